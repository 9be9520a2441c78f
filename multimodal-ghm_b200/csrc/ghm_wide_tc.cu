// ghm_wide_tc.cu -- tcgen05 variant of the wide path's batched row-GEMM (placeholder until the UMMA kernel lands).
#include "ghm_wide.cuh"

int ghm_wide_gemm_tc(const ghm_model*, int64_t, int, int, int, const float*, float*, cudaStream_t) { return GHM_EUNSUP; }
