// ghm_stubs.cu -- entry points declared in include/ghm_b200.h whose kernels are not written yet.
// They fail loudly (GHM_EUNSUP); nothing falls back to the CPU.
#include "ghm_common.cuh"

#define NOT_YET(name) return ghm_fail(GHM_EUNSUP, name ": not implemented in this build")

extern "C" int64_t ghm_bp_dns_workspace_bytes(const ghm_model_t*, int64_t) { return 0; }
extern "C" int ghm_bp_dns(const ghm_model_t*, int64_t, const float*, float, const float*, float*, void*, void*) { NOT_YET("ghm_bp_dns"); }
extern "C" int64_t ghm_bp_nwp_workspace_bytes(const ghm_model_t*, int64_t) { return 0; }
extern "C" int ghm_bp_nwp(const ghm_model_t*, int64_t, const void*, int, const float*, float*, void*, void*) { NOT_YET("ghm_bp_nwp"); }
extern "C" int ghm_guides_cls(const ghm_model_t*, int64_t, const void*, int, float* const*, float*, float*, void*) { NOT_YET("ghm_guides_cls"); }
extern "C" int64_t ghm_guides_dns_workspace_bytes(const ghm_model_t*, int64_t) { return 0; }
extern "C" int ghm_guides_dns(const ghm_model_t*, int64_t, const float*, float, const float*, float* const*, float*, void*, void*) { NOT_YET("ghm_guides_dns"); }
extern "C" int64_t ghm_guides_nwp_workspace_bytes(const ghm_model_t*, int64_t) { return 0; }
extern "C" int ghm_guides_nwp(const ghm_model_t*, int64_t, const void*, int, const float*, float* const*, float*, void*, void*) { NOT_YET("ghm_guides_nwp"); }
