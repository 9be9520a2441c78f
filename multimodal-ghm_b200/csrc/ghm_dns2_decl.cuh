// ghm_dns2_decl.cuh -- what ghm_dns.cu needs of the k_dns2 translation units (ghm_dns2_inst_q*.cu)
#pragma once
#include "ghm_common.cuh"

#define DNS_NT 128

struct DnsArgs {
    int64_t B;
    const float* z;
    float c2;              // -0.5 * log2(e) / sigma^2
    const float* ext;      // [B, q] log-message or null
    float* mean;           // [B, n_L]
    float* scratch;        // [E_int][Q][B]
    float* root_bu;        // [B, q] or null: root_node.hd_message after BP_DNS = shifted root hd + ext (:501-506, aliasing)
};

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// GHM_EUNSUP when the tables do not fit the constant-bank parameter; defined in ghm_dns2_inst_q<Q>.cu
template <int Q, int S>
int launch_dns2(const ghm_model* m, const DnsArgs& a, cudaStream_t st);
