// ghm_vec.cuh -- register-resident q-vector primitives shared by the BP kernels.
//
// One THREAD owns one tree; a message is a float[Q] in registers (Q = q padded to
// {4,8,10,16}, padded entries are exactly 0 in the linear domain and stay 0).  All lanes of
// a warp walk their trees in lock-step, so every transition-table read is a shared-memory
// broadcast (one LDS.128 serves 32 trees) and there are no shuffles and no divergence.
#pragma once
#include "ghm_common.cuh"

// y[a] = sum_b T[a][b] * x[b]      (child -> parent, reference `T @ exp(h)`, data_random_GHM.py:207,487,497)
template <int Q>
__device__ __forceinline__ void ghm_matvec(const float* __restrict__ T, const float (&x)[Q], float (&y)[Q]) {
#pragma unroll
    for (int a = 0; a < Q; ++a) y[a] = 0.f;
    const float4* T4 = reinterpret_cast<const float4*>(T);
#pragma unroll
    for (int i = 0; i < Q * Q / 4; ++i) {
        const float4 t = T4[i];
        y[(4 * i + 0) / Q] = fmaf(t.x, x[(4 * i + 0) % Q], y[(4 * i + 0) / Q]);
        y[(4 * i + 1) / Q] = fmaf(t.y, x[(4 * i + 1) % Q], y[(4 * i + 1) / Q]);
        y[(4 * i + 2) / Q] = fmaf(t.z, x[(4 * i + 2) % Q], y[(4 * i + 2) / Q]);
        y[(4 * i + 3) / Q] = fmaf(t.w, x[(4 * i + 3) % Q], y[(4 * i + 3) / Q]);
    }
}

// y[b] = sum_a T[a][b] * x[a]      (parent -> child, reference `T.T @ exp(diff)`, :513,451,453)
template <int Q>
__device__ __forceinline__ void ghm_matvec_t(const float* __restrict__ T, const float (&x)[Q], float (&y)[Q]) {
#pragma unroll
    for (int b = 0; b < Q; ++b) y[b] = 0.f;
    const float4* T4 = reinterpret_cast<const float4*>(T);
#pragma unroll
    for (int i = 0; i < Q * Q / 4; ++i) {
        const float4 t = T4[i];
        y[(4 * i + 0) % Q] = fmaf(t.x, x[(4 * i + 0) / Q], y[(4 * i + 0) % Q]);
        y[(4 * i + 1) % Q] = fmaf(t.y, x[(4 * i + 1) / Q], y[(4 * i + 1) % Q]);
        y[(4 * i + 2) % Q] = fmaf(t.z, x[(4 * i + 2) / Q], y[(4 * i + 2) % Q]);
        y[(4 * i + 3) % Q] = fmaf(t.w, x[(4 * i + 3) / Q], y[(4 * i + 3) % Q]);
    }
}

// load one table row (Q consecutive 4-byte words, row start 8-byte aligned; 16-byte when Q%4==0)
template <int Q, typename W>
__device__ __forceinline__ void ghm_load_row(const W* __restrict__ row, W (&out)[Q]) {
    static_assert(sizeof(W) == 4, "4-byte words");
    if constexpr (Q % 4 == 0) {
        const uint4* p = reinterpret_cast<const uint4*>(row);
#pragma unroll
        for (int i = 0; i < Q / 4; ++i) {
            uint4 v = p[i];
            out[4 * i + 0] = reinterpret_cast<W&>(v.x);
            out[4 * i + 1] = reinterpret_cast<W&>(v.y);
            out[4 * i + 2] = reinterpret_cast<W&>(v.z);
            out[4 * i + 3] = reinterpret_cast<W&>(v.w);
        }
    } else {
        static_assert(Q % 2 == 0, "even Q");
        const uint2* p = reinterpret_cast<const uint2*>(row);
#pragma unroll
        for (int i = 0; i < Q / 2; ++i) {
            uint2 v = p[i];
            out[2 * i + 0] = reinterpret_cast<W&>(v.x);
            out[2 * i + 1] = reinterpret_cast<W&>(v.y);
        }
    }
}

template <int Q>
__device__ __forceinline__ float ghm_vmax(const float (&x)[Q]) {
    float m = x[0];
#pragma unroll
    for (int k = 1; k < Q; ++k) m = fmaxf(m, x[k]);
    return m;
}

// rescale so the largest entry is 1 (the linear-domain twin of the reference's `h -= max(h)`, :197,208,496)
template <int Q>
__device__ __forceinline__ void ghm_normalize(float (&x)[Q]) {
    const float inv = 1.0f / ghm_vmax<Q>(x);
#pragma unroll
    for (int k = 0; k < Q; ++k) x[k] *= inv;
}

// parity-mode inverse CDF: first k with u < cdf[k], else 0   (reference argmax semantics, :164-165)
__device__ __forceinline__ int ghm_search_f64(const double* __restrict__ row, double u, int q) {
    int lo = 0, hi = q;                                  // cdf is non-decreasing: lower bound of {k : u < cdf[k]}
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (u >= __ldg(row + mid)) lo = mid + 1; else hi = mid;
    }
    return lo == q ? 0 : lo;
}
