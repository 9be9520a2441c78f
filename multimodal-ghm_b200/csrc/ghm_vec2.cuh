// ghm_vec2.cuh -- packed-FP32 (f32x2) q-vector primitives for the two-trees-per-thread kernels.
//
// sm_100a has 64-bit packed single-precision instructions (PTX fma.rn.f32x2 / mul.rn.f32x2, SASS
// FFMA2 / FMUL2, with a scalar-broadcast operand form).  A BP message is kept as Q/2 register PAIRS
// along the state index, so one issue slot does two multiply-adds: the child->parent matvec
// `T @ m` (reference data_random_GHM.py:207) is Q*Q/2 FFMA2 whose table operand comes straight out
// of an LDS.128 of the transposed, 16-byte-row-aligned table TTp, and whose vector operand is the
// broadcast scalar m[b].  Each thread owns TWO trees: the table rows are loaded once and used for
// both, and the two independent dependency chains hide the FMA / LDS latencies.
#pragma once
#include "ghm_vec.cuh"

typedef float2 f2;

__device__ __forceinline__ unsigned long long f2_pack(f2 v) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(v.x), "f"(v.y));
    return r;
}
__device__ __forceinline__ f2 f2_unpack(unsigned long long r) {
    f2 v;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(v.x), "=f"(v.y) : "l"(r));
    return v;
}
__device__ __forceinline__ f2 f2_mul(f2 a, f2 b) {
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_pack(a)), "l"(f2_pack(b)));
    return f2_unpack(d);
}
__device__ __forceinline__ f2 f2_fma(f2 a, f2 b, f2 c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(f2_pack(a)), "l"(f2_pack(b)), "l"(f2_pack(c)));
    return f2_unpack(d);
}
__device__ __forceinline__ f2 f2_muls(f2 a, float s) { return f2_mul(a, make_float2(s, s)); }
__device__ __forceinline__ f2 f2_fmas(f2 a, float s, f2 c) { return f2_fma(a, make_float2(s, s), c); }

// load one TTp row (Q floats, 16-byte aligned) as Q/2 pairs
template <int Q>
__device__ __forceinline__ void f2_load_row(const float* __restrict__ row, f2 (&out)[Q / 2]) {
    static_assert(Q % 2 == 0, "even Q");
    const float4* p4 = reinterpret_cast<const float4*>(row);
#pragma unroll
    for (int i = 0; i < Q / 4; ++i) {
        const float4 v = p4[i];
        out[2 * i] = make_float2(v.x, v.y);
        out[2 * i + 1] = make_float2(v.z, v.w);
    }
    if constexpr (Q % 4 == 2) out[Q / 2 - 1] = *reinterpret_cast<const float2*>(row + Q - 2);
}

// the same row through the read-only data path (ld.global.nc: L1-cacheable gathers of a table that lives in L2)
template <int Q>
__device__ __forceinline__ void f2_ldg_row(const float* __restrict__ row, f2 (&out)[Q / 2]) {
    static_assert(Q % 2 == 0, "even Q");
    const float4* p4 = reinterpret_cast<const float4*>(row);
#pragma unroll
    for (int i = 0; i < Q / 4; ++i) {
        const float4 v = __ldg(p4 + i);
        out[2 * i] = make_float2(v.x, v.y);
        out[2 * i + 1] = make_float2(v.z, v.w);
    }
    if constexpr (Q % 4 == 2) out[Q / 2 - 1] = __ldg(reinterpret_cast<const float2*>(row + Q - 2));
}

template <int Q>
__device__ __forceinline__ float f2_elem(const f2 (&x)[Q / 2], int b) { return (b & 1) ? x[b >> 1].y : x[b >> 1].x; }

// y = T x for two trees at once:  y[a] = sum_b TT[b][a] x[b];  TT rows have stride QS floats
template <int Q, int QS>
__device__ __forceinline__ void f2_matvec_up2(const float* __restrict__ TT, const f2 (&x0)[Q / 2], const f2 (&x1)[Q / 2],
                                              f2 (&y0)[Q / 2], f2 (&y1)[Q / 2]) {
#pragma unroll
    for (int b = 0; b < Q; ++b) {
        f2 row[Q / 2];
        f2_load_row<Q>(TT + b * QS, row);
        const float xb0 = f2_elem<Q>(x0, b), xb1 = f2_elem<Q>(x1, b);
#pragma unroll
        for (int i = 0; i < Q / 2; ++i) {
            y0[i] = b == 0 ? f2_muls(row[i], xb0) : f2_fmas(row[i], xb0, y0[i]);
            y1[i] = b == 0 ? f2_muls(row[i], xb1) : f2_fmas(row[i], xb1, y1[i]);
        }
    }
}

template <int Q>
__device__ __forceinline__ float f2_vmax(const f2 (&x)[Q / 2]) {
    float m = fmaxf(x[0].x, x[0].y);
#pragma unroll
    for (int i = 1; i < Q / 2; ++i) m = fmaxf(m, fmaxf(x[i].x, x[i].y));
    return m;
}

// rescale so the largest entry is ~1 (linear-domain twin of the reference's `h -= max(h)`, :197,208).
// The scale is a per-node constant that cancels in every normalised output, so the approximate
// reciprocal (MUFU.RCP) is exact enough by construction.
template <int Q>
__device__ __forceinline__ void f2_normalize(f2 (&x)[Q / 2]) {
    const float inv = __fdividef(1.0f, f2_vmax<Q>(x));
#pragma unroll
    for (int i = 0; i < Q / 2; ++i) x[i] = f2_muls(x[i], inv);
}

// Philox-mode draw through the Walker alias row of (matrix, parent state): one LDS.32 per draw
__device__ __forceinline__ int ghm_draw_alias(const uint32_t* __restrict__ arow, uint32_t r, int q) {
    const unsigned long long m = (unsigned long long)r * (unsigned)q;
    const int k = (int)(m >> 32);
    const uint32_t e = arow[k];
    return ((uint32_t)m < e) ? k : (int)(e & 255u);
}

// single-tree form of f2_matvec_up2
template <int Q, int QS>
__device__ __forceinline__ void f2_matvec_up1(const float* __restrict__ TT, const f2 (&x0)[Q / 2], f2 (&y0)[Q / 2]) {
#pragma unroll
    for (int b = 0; b < Q; ++b) {
        f2 row[Q / 2];
        f2_load_row<Q>(TT + b * QS, row);
        const float xb0 = f2_elem<Q>(x0, b);
#pragma unroll
        for (int i = 0; i < Q / 2; ++i) y0[i] = b == 0 ? f2_muls(row[i], xb0) : f2_fmas(row[i], xb0, y0[i]);
    }
}

// ---- transition tables as a by-value kernel parameter (constant bank) ------------------------------------------
// [0, n_mat*Q*Q): TlinT (child->parent matvec) | [n_mat*Q*Q, 2*n_mat*Q*Q): Tlin (parent->child matvec).  With a
// warp-uniform matrix index the packed FMAs take their table operand from uniform registers (LDCU), i.e. the matvec
// issues no LSU instruction at all.
template <int NW>
struct __align__(16) DnsTab { float v[NW]; };
#define GHM_TAB_WORDS 6144

// y[i] = sum_r T[r][2i..2i+1] * x[r]  with T rows of Q floats (8-byte aligned), table in the constant bank
template <int Q>
__device__ __forceinline__ void f2_matvec_c(const float* __restrict__ T, const f2 (&x)[Q / 2], f2 (&y)[Q / 2]) {
#pragma unroll
    for (int r = 0; r < Q; ++r) {
        const float xr = f2_elem<Q>(x, r);
#pragma unroll
        for (int i = 0; i < Q / 2; ++i) {
            const f2 t = *reinterpret_cast<const f2*>(T + r * Q + 2 * i);
            y[i] = r == 0 ? f2_muls(t, xr) : f2_fmas(t, xr, y[i]);
        }
    }
}
