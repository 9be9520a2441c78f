// ghm_tree_kernel.cuh -- K1 (sampler) and K2 (root-posterior BP), separately or fused in one pass: the kernel
// template and its launch logic.  Instantiated per (padded q, mode) in ghm_tree_inst_*.cu so the translation
// units compile in parallel; the C entry points live in ghm_tree.cu.
//
// Replaces GHMTree.gen_values (src/ghmclip/data/data_random_GHM.py:145-165) and
// GHMTree.BP_CLS (:185-221) of the reference.
//
// k_tree2 (Philox sampling and/or BP; FP32 CUDA cores, issue-slot bound -- see DESIGN.md):
//   * one THREAD owns TPT (1 or 2) trees and walks them depth-first in lock-step: sampling goes down
//     the current root path, the BP message comes back up the same path, so a leaf state lives in a
//     register between being drawn and being absorbed -- fused mode moves no leaf through HBM twice;
//   * all 32 lanes of a warp are at the same node of their trees: transition-table reads are
//     shared-memory broadcasts (shared by both trees of a thread when TPT = 2), index arithmetic is
//     warp-uniform, and there is no divergence; the per-lane gathers (alias entry of the parent state,
//     T^T row of the leaf state) are the only non-broadcast LDS;
//   * messages are Q/2 packed f32x2 register pairs: the child->parent matvec is Q*Q/2 FFMA2 fed by
//     LDS.128 of the 16-byte-row-aligned transposed table (ghm_vec2.cuh);
//   * BP runs in the LINEAR domain with a max-rescale per node: msg(v) = prod_c (T_c msg(c)) / max.
//     This is the reference's log-space recursion `hd = sum_c log(T_c @ exp(hd_c)) - max`
//     (:207-208) exponentiated -- same rescale points, no exp/log in the inner loop;
//   * Philox-mode draws use a Walker alias table (one 32-bit LDS + one compare per draw instead of a
//     q-term CDF scan); the distribution is the reference's up to the 2^-24 threshold quantisation;
//   * Philox counter layout (restated by oracle/philox.py): the s leaves under depth-(L-1) node j take
//     words 0..s-1 of blocks (level L, j*ceil(s/4) + c/4); when s % 4 != 0 the spare word s % 4 of the
//     last block draws node j itself, so the hot path makes one Philox call per node and every word
//     pick is a compile-time index; shallower levels use word idx & 3 of block (level, idx >> 2);
//   * the accumulator of the depth-(L-2) ancestor stays in registers; shallower ancestors (touched
//     every s-th, s^2-th ... node) and their cached Philox blocks sit in shared memory as
//     [level][..][tree][thread] (conflict-free), so the depth L stays a runtime value;
//   * leaves are staged per warp in shared memory as bytes in exactly the [trees][n_L] order of the
//     global tensor, so the flush is a flat, fully coalesced stream of 16-byte stores
//     (the [B, n_L] int64 API layout is 8*n_L contiguous bytes per tree).
// k_sample_parity: the reference's f64 inverse-CDF on caller-supplied uniforms (bit-exact leaves).
#pragma once
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "ghm_vec2.cuh"

#define T2_NT 128          // threads per CTA
#define T2_WARPS (T2_NT / 32)

enum { MODE_PHILOX = 0, MODE_GIVEN = 2 };

struct TreeArgs {
    int64_t B;
    int root_mode;
    int64_t n_given;       // GHM_ROOT_GIVEN: trees [0, n_given) take root_in, the rest draw uniform roots (ClipSampler image layout)
    const int64_t* root_in;
    uint64_t root_seed;    // GHM_ROOT_SHARED: Philox key of the roots of trees [0, n_given) (the partner modality's seed)
    const double* U;
    uint64_t seed, tree_offset;
    uint32_t blk_len;              // blk_len > 0 (ghm_sample_blocked, B < 2^32): local tree b has the global Philox index
    uint64_t blk_extra;            //   tree_offset + b + (b / blk_len) * blk_extra,  blk_extra = blk_stride - blk_len
    int64_t* root_out;
    void* leaves;          // output (sampling modes) or input (MODE_GIVEN); may be null when sampling
    int leaf_dtype;
    float* post;
    float* root_hd;
    int chunk_j;           // (row-chunk staging) depth-(L-1) nodes per staging chunk
    int stage_stride;      // bytes per staged tree row
    int stage_bytes;       // bytes of staging per warp
    int base0, base1;      // index of the matrices into depth L-1 / L-2 (child 0): (L-2)*s, (L-3)*s
    int acc_stride;        // k_tree_fast: f2 elements of shared memory per warp for the parked accumulators / flush ring
};

// ------------------------------------------------------------------------------------------------
// staging: generic (row-strided chunks, any shape) and flat (stage layout == global layout)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void stage_flush_rows(const uint8_t* st, int stride, int wtrees, void* leaves, int dtype,
                                                 int64_t tree0, int64_t B, int nL, int base, int len, int lane) {
    __syncwarp();
    for (int r = 0; r < wtrees; ++r) {
        const int64_t t = tree0 + r;
        if (t >= B) break;
        const uint8_t* row = st + r * stride;
        if (dtype == GHM_LEAF_I64) {
            int64_t* dst = reinterpret_cast<int64_t*>(leaves) + t * nL + base;
            for (int i = lane; i < len; i += 32) dst[i] = (int64_t)row[i];
        } else {
            uint8_t* dst = reinterpret_cast<uint8_t*>(leaves) + t * nL + base;
            for (int i = lane; i < len; i += 32) dst[i] = row[i];
        }
    }
    __syncwarp();
}

__device__ __forceinline__ void stage_load_rows(uint8_t* st, int stride, int wtrees, const void* leaves, int dtype,
                                                int64_t tree0, int64_t B, int nL, int base, int len, int lane, int q,
                                                int* status) {
    __syncwarp();
    bool bad = false;
    for (int r = 0; r < wtrees; ++r) {
        int64_t t = tree0 + r;
        if (t >= B) t = B - 1;
        uint8_t* row = st + r * stride;
        if (dtype == GHM_LEAF_I64) {
            const int64_t* src = reinterpret_cast<const int64_t*>(leaves) + t * nL + base;
            for (int i = lane; i < len; i += 32) {
                int64_t v = src[i];
                if (v < 0 || v >= q) { bad = true; v = v < 0 ? 0 : q - 1; }
                row[i] = (uint8_t)v;
            }
        } else {
            const uint8_t* src = reinterpret_cast<const uint8_t*>(leaves) + t * nL + base;
            for (int i = lane; i < len; i += 32) {
                int v = src[i];
                if (v >= q) { bad = true; v = q - 1; }
                row[i] = (uint8_t)v;
            }
        }
    }
    if (bad) atomicOr(status, 1);
    __syncwarp();
}

// flat flush: stage bytes [rows][nL] -> global [rows][nL] of int64 / uint8, 16-byte stores
__device__ __forceinline__ void stage_flush_flat(const uint8_t* st, int wtrees, void* leaves, int dtype, int64_t tree0,
                                                 int64_t B, int nL, int lane) {
    __syncwarp();
    const int rows = (int)min((int64_t)wtrees, B - tree0);
    const int n = rows * nL;
    if (dtype == GHM_LEAF_I64) {
        int64_t* dst = reinterpret_cast<int64_t*>(leaves) + tree0 * nL;
        const uint32_t* st32 = reinterpret_cast<const uint32_t*>(st);
        ulonglong2* dst2 = reinterpret_cast<ulonglong2*>(dst);
        const int n4 = n >> 2;
#pragma unroll 4
        for (int g = lane; g < n4; g += 32) {
            const uint32_t w = st32[g];
            dst2[2 * g] = make_ulonglong2(w & 255u, (w >> 8) & 255u);
            dst2[2 * g + 1] = make_ulonglong2((w >> 16) & 255u, w >> 24);
        }
        for (int i = (n4 << 2) + lane; i < n; i += 32) dst[i] = (int64_t)st[i];
    } else {
        uint8_t* dst = reinterpret_cast<uint8_t*>(leaves) + tree0 * nL;
        const uint4* st16 = reinterpret_cast<const uint4*>(st);
        uint4* dst16 = reinterpret_cast<uint4*>(dst);
        const int n16 = n >> 4;
        for (int g = lane; g < n16; g += 32) dst16[g] = st16[g];
        for (int i = (n16 << 4) + lane; i < n; i += 32) dst[i] = st[i];
    }
    __syncwarp();
}

// flat load with range validation; rows past the batch end are filled with state 0
__device__ __forceinline__ void stage_load_flat(uint8_t* st, int wtrees, const void* leaves, int dtype, int64_t tree0,
                                                int64_t B, int nL, int lane, int q, int* status) {
    __syncwarp();
    const int rows = (int)min((int64_t)wtrees, B - tree0);
    const int n = rows * nL, ntot = wtrees * nL;
    bool bad = false;
    if (dtype == GHM_LEAF_I64) {
        const int64_t* src = reinterpret_cast<const int64_t*>(leaves) + tree0 * nL;
        const ulonglong2* src2 = reinterpret_cast<const ulonglong2*>(src);
        uint32_t* st32 = reinterpret_cast<uint32_t*>(st);
        const int n4 = n >> 2;
#pragma unroll 2
        for (int g = lane; g < n4; g += 32) {
            const ulonglong2 a = src2[2 * g], b = src2[2 * g + 1];
            unsigned long long v[4] = {a.x, a.y, b.x, b.y};
            uint32_t w = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (v[k] >= (unsigned long long)q) { bad = true; v[k] = ((long long)v[k] < 0) ? 0 : q - 1; }
                w |= (uint32_t)v[k] << (8 * k);
            }
            st32[g] = w;
        }
        for (int i = (n4 << 2) + lane; i < n; i += 32) {
            int64_t v = src[i];
            if (v < 0 || v >= q) { bad = true; v = v < 0 ? 0 : q - 1; }
            st[i] = (uint8_t)v;
        }
    } else {
        const uint8_t* src = reinterpret_cast<const uint8_t*>(leaves) + tree0 * nL;
        for (int i = lane; i < n; i += 32) {
            int v = src[i];
            if (v >= q) { bad = true; v = q - 1; }
            st[i] = (uint8_t)v;
        }
    }
    for (int i = n + lane; i < ntot; i += 32) st[i] = 0;
    if (bad) atomicOr(status, 1);
    __syncwarp();
}

// ------------------------------------------------------------------------------------------------
// k_tree2
// ------------------------------------------------------------------------------------------------
template <int S>
__device__ __forceinline__ int div_s(int idx, const GhmDev& d) {
    if constexpr (S > 0) return idx / S;
    else return ghm_div_s(idx, d);
}

// Transition tables as a by-value kernel parameter: they live in the constant bank, the matvec reads them
// with LDCU.64 into UNIFORM registers and FFMA2 takes the uniform pair as an operand, so the table side of
// the child->parent matvec costs no shared-memory / LSU-writeback bandwidth at all (a broadcast LDS.128
// still writes 512 B of registers per warp; measured: the LDS form is bound by the 128 B/clk LSU pipe).
template <int NW>
struct __align__(16) TabParam { float v[NW]; };

// TPT trees per thread; S = compile-time branching (0: runtime); FLAT: whole-row staging;
// NW > 4: the T^T tables of every matrix are in the constant-bank parameter `tab`
template <int Q, int S, int TPT, int MODE, bool BP, bool FLAT, bool SMEM_TAB, int NW>
__global__ void __launch_bounds__(T2_NT, TPT == 1 ? 5 : 4)
k_tree2(const __grid_constant__ GhmDev d, const __grid_constant__ TreeArgs a, const __grid_constant__ TabParam<NW> tab) {
    constexpr bool CTAB = NW > 4;
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int NT = T2_NT, H = Q / 2, QS = (Q + 3) / 4 * 4, WTREES = 32 * TPT;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int L = d.L, s = S > 0 ? S : d.s, q = d.q, nL = d.n_leaves;
    const int64_t warp_tree0 = ((int64_t)blockIdx.x * T2_WARPS + warp) * WTREES;
    const bool warp_has_work = warp_tree0 < a.B;
    const int nb = (s + 3) >> 2;                                 // Philox blocks per depth-(L-1) node
    const bool spare = (s & 3) != 0 && L >= 2;                   // node j drawn from the spare word of its last leaf block

    // ---- carve shared memory ------------------------------------------------------------
    size_t off = 0;
    const float* TT = d.TTp;
    const uint32_t* AL = d.alias;
    if (SMEM_TAB) {
        if (BP) {
            // constant-bank tables serve the matvecs; shared memory then only holds the leaf-level matrices
            // (their rows are gathered per lane by leaf state)
            const int first = CTAB ? d.mat_off[L] : 0;
            const int words = (d.n_mat - first) * Q * QS;
            float* s1 = reinterpret_cast<float*>(smem + off); off += (size_t)words * 4;
            const float* src = d.TTp + (size_t)first * Q * QS;
            for (int i = tid; i < words; i += NT) s1[i] = src[i];
            TT = s1 - (size_t)first * Q * QS;
        }
        if (MODE == MODE_PHILOX) {
            const int words = d.n_mat * q * q;
            uint32_t* s2 = reinterpret_cast<uint32_t*>(smem + off); off += ((size_t)words * 4 + 15) / 16 * 16;
            for (int i = tid; i < words; i += NT) s2[i] = d.alias[i];
            AL = s2;
        }
    }
    const int n_deep = L > 2 ? L - 2 : 0;                       // ancestors kept in shared memory: depths 0 .. L-3
    f2* ACC = reinterpret_cast<f2*>(smem + off);                 // [n_deep][H][TPT][NT]
    if (BP) off += (size_t)n_deep * H * TPT * NT * sizeof(f2);
    const int n_rng = n_deep + ((L >= 2 && !spare) ? 1 : 0);     // levels 1 .. L-2 (+ L-1 when it has no spare word)
    uint32_t* RNG = reinterpret_cast<uint32_t*>(smem + off);     // [n_rng][3][TPT][NT]  words 1..3 of the cached Philox blocks
    if (MODE == MODE_PHILOX) off += (size_t)n_rng * 3 * TPT * NT * 4;
    uint8_t* VAL = smem + off;                                   // [n_deep][TPT][NT] states of the path nodes at depths 0 .. L-3
    if (MODE == MODE_PHILOX) off += ((size_t)n_deep * TPT * NT + 15) / 16 * 16;
    const bool use_stage = (a.leaves != nullptr);
    uint8_t* stage = smem + off + (size_t)warp * a.stage_bytes;
    if (SMEM_TAB) __syncthreads();
    if (!warp_has_work) return;

    const int n1 = d.spow[L - 1];
    bool active[TPT];
    uint64_t tree[TPT];
    int64_t bt[TPT];
    int srow[TPT];
#pragma unroll
    for (int t = 0; t < TPT; ++t) {
        bt[t] = warp_tree0 + 32 * t + lane;
        active[t] = bt[t] < a.B;
        const int64_t bc = active[t] ? bt[t] : a.B - 1;          // tail threads shadow the last tree and never write
        tree[t] = a.tree_offset + (uint64_t)bc;
        if (a.blk_len) tree[t] += (uint64_t)((uint32_t)bc / a.blk_len) * a.blk_extra;
        srow[t] = (32 * t + lane) * a.stage_stride;
    }

    if (MODE == MODE_GIVEN && FLAT) stage_load_flat(stage, WTREES, a.leaves, a.leaf_dtype, warp_tree0, a.B, nL, lane, q, d.status);

    // ---- root ---------------------------------------------------------------------------
    int xcur[TPT];                                               // state of the current depth-(L-1) node
    int xpar[TPT];                                               // state of its parent (depth L-2)
#pragma unroll
    for (int t = 0; t < TPT; ++t) { xcur[t] = 0; xpar[t] = 0; }
    if (MODE == MODE_PHILOX) {
#pragma unroll
        for (int t = 0; t < TPT; ++t) {
            int x0;
            if (a.root_mode == GHM_ROOT_GIVEN && (active[t] ? bt[t] : a.B - 1) < a.n_given) {
                int64_t r = a.root_in[active[t] ? bt[t] : a.B - 1];
                if (r < 0 || r >= q) { atomicOr(d.status, 1); r = r < 0 ? 0 : q - 1; }
                x0 = (int)r;
            } else {
                const bool shared = a.root_mode == GHM_ROOT_SHARED && (active[t] ? bt[t] : a.B - 1) < a.n_given;
                const uint4 rb = ghm_rng_block(shared ? a.root_seed : a.seed, tree[t], 0u, 0u, GHM_STREAM_TREE);
                const uint32_t* rc = a.root_mode == GHM_ROOT_PRIOR ? d.root_cdfu_prior : d.root_cdfu_unif;
                int cnt = 0;
                for (int k = 0; k < q - 1; ++k) cnt += (rb.x >= __ldg(rc + k)) ? 1 : 0;
                x0 = cnt;
            }
            xcur[t] = x0;                                        // L == 1: node j = 0 is the root itself
            xpar[t] = x0;                                        // L == 2: the parent of every depth-1 node
            if (L > 2) VAL[t * NT + tid] = (uint8_t)x0;           // depth 0 is read back when depth-1 nodes are redrawn
            if (a.root_out && active[t]) a.root_out[bt[t]] = x0;
        }
    }

    f2 msg[TPT][H], accT[TPT][H];
#pragma unroll
    for (int t = 0; t < TPT; ++t)
#pragma unroll
        for (int i = 0; i < H; ++i) { msg[t][i] = make_float2(0.f, 0.f); accT[t][i] = make_float2(0.f, 0.f); }

    int chunk_base = 0, chunk_left = a.chunk_j;                  // row-chunk staging only
    int tz = L > 2 ? L - 2 : 0;                                  // ancestors to (re)draw before node j: depths L-1-tz .. L-2
    int cj = 0;                                                  // j mod s  (child index of node j under its parent)
    const uint32_t* alias_leaf0 = AL + (size_t)d.mat_off[L] * q * q;
    const float* tt_leaf0 = TT + (size_t)d.mat_off[L] * Q * QS;

    uint4 rbN[TPT];                                              // Philox block of the NEXT node's first leaves: issued ahead of
    if (MODE == MODE_PHILOX) {                                   // the BP climb so its integer chain overlaps the FP matvecs
#pragma unroll
        for (int t = 0; t < TPT; ++t) rbN[t] = ghm_rng_block(a.seed, tree[t], (uint32_t)L, 0u, GHM_STREAM_TREE);
    }
    for (int j = 0; j < n1; ++j) {
        if (MODE == MODE_GIVEN && !FLAT && chunk_left == a.chunk_j) {
            const int len = min(a.chunk_j * s, nL - chunk_base);
            stage_load_rows(stage, a.stage_stride, WTREES, a.leaves, a.leaf_dtype, warp_tree0, a.B, nL, chunk_base, len,
                            lane, q, d.status);
        }
        // ---- (re)draw the ancestors at depths <= L-2 that changed (every s-th node at most) ----
        if (MODE == MODE_PHILOX && L > 2 && cj == 0) {
            int t2 = div_s<S>(j, d);                             // trailing zero base-s digits of j, capped at L-2
            tz = 1;
            while (tz < L - 2) {
                const int tq = div_s<S>(t2, d);
                if (t2 - tq * s != 0) break;
                t2 = tq; ++tz;
            }
            const int lstart = L - 1 - tz;                       // >= 1
            int xp[TPT];
#pragma unroll
            for (int t = 0; t < TPT; ++t) xp[t] = VAL[((lstart - 1) * TPT + t) * NT + tid];
            for (int l = lstart; l <= L - 2; ++l) {
                const int idx = ghm_div_pow(j, L - 1 - l, d);
                const int c = idx - div_s<S>(idx, d) * s;
                const int mi = d.mat_off[l] + (d.ti ? c : idx);
                const uint32_t* arow = AL + (size_t)mi * q * q;
#pragma unroll
                for (int t = 0; t < TPT; ++t) {
                    uint32_t* rl = RNG + (size_t)((l - 1) * 3 * TPT + t) * NT + tid;
                    uint32_t r;
                    if ((idx & 3) == 0) {
                        const uint4 rb = ghm_rng_block(a.seed, tree[t], (uint32_t)l, (uint32_t)(idx >> 2), GHM_STREAM_TREE);
                        rl[0] = rb.y; rl[TPT * NT] = rb.z; rl[2 * TPT * NT] = rb.w;
                        r = rb.x;
                    } else {
                        r = rl[((idx & 3) - 1) * TPT * NT];
                    }
                    const int x = ghm_draw_alias(arow + xp[t] * q, r, q);
                    if (l < L - 2) VAL[(l * TPT + t) * NT + tid] = (uint8_t)x;
                    xp[t] = x;
                }
            }
#pragma unroll
            for (int t = 0; t < TPT; ++t) xpar[t] = xp[t];
        }
        // ---- node j (depth L-1) and the s leaves under it -------------------------------------
        const int mi_j = d.mat_off[L > 1 ? L - 1 : 1] + (d.ti ? cj : j);   // matrix of the edge into node j
        uint4 rbL[TPT];
        if (MODE == MODE_PHILOX && L >= 2 && !spare) {           // s % 4 == 0: node j from the per-level block
#pragma unroll
            for (int t = 0; t < TPT; ++t) {
                uint32_t* rl = RNG + (size_t)((L - 2) * 3 * TPT + t) * NT + tid;
                uint32_t r;
                if ((j & 3) == 0) {
                    const uint4 rb = ghm_rng_block(a.seed, tree[t], (uint32_t)(L - 1), (uint32_t)(j >> 2), GHM_STREAM_TREE);
                    rl[0] = rb.y; rl[TPT * NT] = rb.z; rl[2 * TPT * NT] = rb.w;
                    r = rb.x;
                } else {
                    r = rl[((j & 3) - 1) * TPT * NT];
                }
                xcur[t] = ghm_draw_alias(AL + ((size_t)mi_j * q + xpar[t]) * q, r, q);
            }
        }
        f2 h[TPT][H];
#pragma unroll
        for (int c = 0; c < s; ++c) {
            const int lidx = j * s + c;
            const int moff = d.ti ? c : lidx;
            if (MODE == MODE_PHILOX && (c & 3) == 0) {
#pragma unroll
                for (int t = 0; t < TPT; ++t)
                    rbL[t] = c == 0 ? rbN[t]
                                    : ghm_rng_block(a.seed, tree[t], (uint32_t)L, (uint32_t)(j * nb + (c >> 2)), GHM_STREAM_TREE);
            }
            if (MODE == MODE_PHILOX && spare && c == 0) {        // node j: spare word s % 4 of its LAST leaf block
#pragma unroll
                for (int t = 0; t < TPT; ++t) {
                    uint32_t r;
                    if (s < 4) {
                        r = ghm_pick(rbL[t], s & 3);
                    } else {
                        const uint4 rb = ghm_rng_block(a.seed, tree[t], (uint32_t)L, (uint32_t)(j * nb + nb - 1), GHM_STREAM_TREE);
                        r = ghm_pick(rb, s & 3);
                    }
                    xcur[t] = ghm_draw_alias(AL + ((size_t)mi_j * q + xpar[t]) * q, r, q);
                }
            }
#pragma unroll
            for (int t = 0; t < TPT; ++t) {
                int x;
                if (MODE == MODE_PHILOX) {
                    x = ghm_draw_alias(alias_leaf0 + ((size_t)moff * q + xcur[t]) * q, ghm_pick(rbL[t], c & 3), q);
                    if (use_stage) stage[srow[t] + (lidx - chunk_base)] = (uint8_t)x;
                } else {
                    x = stage[srow[t] + (lidx - chunk_base)];
                }
                if (BP) {
                    f2 row[H];
                    f2_load_row<Q>(tt_leaf0 + ((size_t)moff * Q + x) * QS, row);
#pragma unroll
                    for (int i = 0; i < H; ++i) h[t][i] = c == 0 ? row[i] : f2_mul(h[t][i], row[i]);
                }
            }
        }
        if (MODE == MODE_PHILOX) {                               // prefetch (one block past the end is computed and dropped)
#pragma unroll
            for (int t = 0; t < TPT; ++t)
                rbN[t] = ghm_rng_block(a.seed, tree[t], (uint32_t)L, (uint32_t)((j + 1) * nb), GHM_STREAM_TREE);
        }
        // ---- carry the finished node's message up the path ----------------------------------
        if (BP) {
#pragma unroll
            for (int t = 0; t < TPT; ++t) {
#pragma unroll
                for (int i = 0; i < H; ++i) msg[t][i] = h[t][i];
                f2_normalize<Q>(msg[t]);
            }
            // One climb step: u = T m for every tree of the thread, times the parked product of the earlier siblings;
            // park it again (more siblings to come) or rescale and keep climbing.  A == nullptr: register accumulator.
            auto climb_step = [&](const float* __restrict__ Tm, bool has_prev, bool last, f2* A) -> bool {
                f2 u[TPT][H];
                if constexpr (TPT == 2) f2_matvec_up2<Q, QS>(Tm, msg[0], msg[1], u[0], u[1]);
                else f2_matvec_up1<Q, QS>(Tm, msg[0], u[0]);
                if (has_prev) {
#pragma unroll
                    for (int t = 0; t < TPT; ++t)
#pragma unroll
                        for (int i = 0; i < H; ++i) u[t][i] = f2_mul(u[t][i], A ? A[(i * TPT + t) * NT] : accT[t][i]);
                }
                if (!last) {
#pragma unroll
                    for (int t = 0; t < TPT; ++t)
#pragma unroll
                        for (int i = 0; i < H; ++i) {
                            if (A) A[(i * TPT + t) * NT] = u[t][i]; else accT[t][i] = u[t][i];
                        }
                    return false;
                }
#pragma unroll
                for (int t = 0; t < TPT; ++t) {
#pragma unroll
                    for (int i = 0; i < H; ++i) msg[t][i] = u[t][i];
                    f2_normalize<Q>(msg[t]);
                }
                return true;
            };
            int idx = j;
            for (int l = L - 1; l > 0; --l) {
                const int pidx = div_s<S>(idx, d);
                const int c = idx - pidx * s;
                const int mi = d.mat_off[l] + (d.ti ? c : idx);
                idx = pidx;
                f2* A = l == L - 1 ? nullptr : ACC + (size_t)(l - 1) * H * TPT * NT + tid;
                if (!climb_step(TT + (size_t)mi * Q * QS, c != 0, c == s - 1, A)) break;
            }
        }
        // ---- advance the odometer ---------------------------------------------------------------
        if (++cj == s) cj = 0;
        if (!FLAT && use_stage) {
            if (--chunk_left == 0 || j == n1 - 1) {
                if (MODE == MODE_PHILOX) {
                    const int len = (j + 1) * s - chunk_base;
                    stage_flush_rows(stage, a.stage_stride, WTREES, a.leaves, a.leaf_dtype, warp_tree0, a.B, nL, chunk_base,
                                     len, lane);
                }
                chunk_base = (j + 1) * s;
                chunk_left = a.chunk_j;
            }
        }
    }
    if (MODE == MODE_PHILOX && use_stage && FLAT)
        stage_flush_flat(stage, WTREES, a.leaves, a.leaf_dtype, warp_tree0, a.B, nL, lane);

    // ---- root outputs (reference :213-217; root_node.hd_message is the shifted log-likelihood) --
    if (BP) {
#pragma unroll
        for (int t = 0; t < TPT; ++t) {
            if (!active[t]) continue;
            const int64_t b = bt[t];
            if (a.root_hd) {
#pragma unroll
                for (int k = 0; k < Q; ++k)
                    if (k < q) a.root_hd[b * q + k] = logf(f2_elem<Q>(msg[t], k));
            }
            if (a.post) {
                float w[Q], sum = 0.f;
#pragma unroll
                for (int k = 0; k < Q; ++k) { w[k] = f2_elem<Q>(msg[t], k) * __ldg(d.py + k); sum += w[k]; }
                const float inv = 1.0f / sum;
#pragma unroll
                for (int k = 0; k < Q; ++k)
                    if (k < q) a.post[b * q + k] = w[k] * inv;
            }
        }
    }
}

// ----------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------
template <int Q, int S, int TPT, int MODE, bool BP, bool FLAT, int NW>
static int launch_tree2(const ghm_model* m, const TreeArgs& a0, bool want_smem_tab, cudaStream_t st) {
    const GhmDev& d = m->d;
    TreeArgs a = a0;
    constexpr int QS = (Q + 3) / 4 * 4, WTREES = 32 * TPT;
    const int n1 = d.spow[d.L - 1];
    const int n_deep = d.L > 2 ? d.L - 2 : 0;

    size_t fixed = 0;
    if (BP) fixed += (size_t)n_deep * (Q / 2) * TPT * T2_NT * sizeof(float2);
    if (MODE == MODE_PHILOX) {
        const bool spare = (d.s & 3) != 0 && d.L >= 2;
        const int n_rng = n_deep + ((d.L >= 2 && !spare) ? 1 : 0);
        fixed += (size_t)n_rng * 3 * TPT * T2_NT * 4 + ((size_t)n_deep * TPT * T2_NT + 15) / 16 * 16;
    }
    size_t tab_bytes = 0;
    if (BP) tab_bytes += (size_t)(NW > 4 ? d.n_mat - d.mat_off[d.L] : d.n_mat) * Q * QS * 4;
    if (MODE == MODE_PHILOX) tab_bytes += ((size_t)d.n_mat * d.q * d.q * 4 + 15) / 16 * 16;

    a.chunk_j = n1; a.stage_stride = 0; a.stage_bytes = 0;
    a.base0 = (d.L - 2) * d.s; a.base1 = (d.L - 3) * d.s;
    if (a.leaves) {
        if (FLAT) {
            a.stage_stride = d.n_leaves;
            a.stage_bytes = (int)(((size_t)WTREES * d.n_leaves + 15) / 16 * 16);
        } else {
            int chunk_j = n1;
            if (d.n_leaves > 128) chunk_j = std::max(1, 128 / d.s);
            a.chunk_j = chunk_j;
            int stride = (std::min(chunk_j * d.s, d.n_leaves) + 3) / 4 * 4;
            if (((stride / 4) & 1) == 0) stride += 4;
            a.stage_stride = stride;
            a.stage_bytes = (WTREES * stride + 15) / 16 * 16;
        }
    }
    const size_t dyn = fixed + (size_t)a.stage_bytes * T2_WARPS + (want_smem_tab ? tab_bytes : 0);
    if (dyn > 200 * 1024)
        return ghm_fail(GHM_EUNSUP, "tree kernel needs %zu bytes of shared memory (L=%d s=%d q=%d)", dyn, d.L, d.s, d.q);
    const int64_t trees_per_cta = (int64_t)T2_WARPS * WTREES;
    const unsigned grid = (unsigned)((a.B + trees_per_cta - 1) / trees_per_cta);
    TabParam<NW> tab;                                              // by-value table parameter (copied at launch)
    tab.v[0] = 0.f;
    if (NW > 4 && BP) {
        const size_t words = (size_t)d.n_mat * Q * QS;
        if (words > (size_t)NW) return ghm_fail(GHM_EUNSUP, "internal: constant table overflow");
        memcpy(tab.v, m->h_TTp, words * sizeof(float));
    }
    auto go = [&](auto kern) -> int {
        GHM_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
        kern<<<grid, T2_NT, dyn, st>>>(d, a, tab);
        GHM_CHECK_LAUNCH();
        return GHM_OK;
    };
    if constexpr (S > 0) {
        return go(k_tree2<Q, S, TPT, MODE, BP, FLAT, true, NW>);      // fast variants always stage the tables
    } else {
        return want_smem_tab ? go(k_tree2<Q, S, TPT, MODE, BP, FLAT, true, NW>)
                             : go(k_tree2<Q, S, TPT, MODE, BP, FLAT, false, NW>);
    }
}

#include "ghm_tree_fast.cuh"

// fast variant (k_tree_fast): translation-invariant tables, s in {2,3,4}, L >= 3, T^T tables small enough for the
// constant-bank parameter, alias / leaf tables fit in shared memory, whole rows staged (flat 16-byte streams);
// everything else (any s, shallow trees, per-edge tables, n_L too large to stage whole rows, unaligned leaf pointer)
// takes the generic one-tree-per-thread instantiation of k_tree2 (S = 0, TPT = 1).
template <int Q, int S, int MODE, bool BP>
static int launch_fast(const ghm_model* m, const TreeArgs& a, cudaStream_t st) {
    constexpr int QS = (Q + 3) / 4 * 4;
    if (!BP) return launch_tree_fast<Q, S, MODE, BP, 4>(m, a, st);
    if constexpr (Q < 16) {                                        // small tables: a 6 KB parameter instead of 24 KB per launch
        if ((size_t)m->d.n_mat * Q * QS <= 1536) return launch_tree_fast<Q, S, MODE, BP, 1536>(m, a, st);
    }
    return launch_tree_fast<Q, S, MODE, BP, 6144>(m, a, st);
}

typedef int (*ghm_fast_fn)(const ghm_model*, const TreeArgs&, cudaStream_t);

// fast[k]: the software-pipelined variant for s = 2 + k.  The fused-BP modes compile each of them in a translation
// unit of its own (ghm_tree_fast_q*_s*.cu: they are the long pole of the build), sampling-only instantiates them here.
template <int Q, int MODE, bool BP>
static int dispatch_variant(const ghm_model* m, const TreeArgs& a, cudaStream_t st, const ghm_fast_fn (&fast)[3]) {
    const GhmDev& d = m->d;
    constexpr int QS = (Q + 3) / 4 * 4;
    const size_t ctab_words = (size_t)d.n_mat * Q * QS;
    size_t smem_tab = 0;
    if (BP) smem_tab += (size_t)(d.n_mat - d.mat_off[d.L]) * Q * QS * 4;
    if (MODE == MODE_PHILOX) smem_tab += ((size_t)d.n_mat * d.q * d.q * 4 + 15) / 16 * 16;
    const size_t flat_bytes = a.leaves ? ((size_t)64 * d.n_leaves + 15) / 16 * 16 * T2_WARPS : 0;
    const bool flat_ok = !a.leaves || (flat_bytes <= 48 * 1024 && ((uintptr_t)a.leaves % 16) == 0);
    if (d.ti && d.L >= 3 && ctab_words <= 6144 && smem_tab <= 40 * 1024 && flat_ok && d.s >= 2 && d.s <= 4)
        return fast[d.s - 2](m, a, st);                            // software-pipelined variant (ghm_tree_fast.cuh)
    size_t gen_tab = 0;
    if (BP) gen_tab += (size_t)d.n_mat * Q * QS * 4;
    if (MODE == MODE_PHILOX) gen_tab += ((size_t)d.n_mat * d.q * d.q * 4 + 15) / 16 * 16;
    return launch_tree2<Q, 0, 1, MODE, BP, false, 4>(m, a, gen_tab <= 40 * 1024, st);
}

// one translation unit per (padded q, mode): ghm_tree_inst_*.cu define these
#define GHM_TREE_DECLARE(Q, TAG) int ghm_tree_run_q##Q##_##TAG(const ghm_model* m, const TreeArgs& a, cudaStream_t st);
#define GHM_TREE_DEFINE(Q, TAG, MODE, BP)                                                                          \
    int ghm_tree_run_q##Q##_##TAG(const ghm_model* m, const TreeArgs& a, cudaStream_t st) {                         \
        static const ghm_fast_fn fast[3] = {&launch_fast<Q, 2, MODE, BP>, &launch_fast<Q, 3, MODE, BP>,             \
                                            &launch_fast<Q, 4, MODE, BP>};                                          \
        return dispatch_variant<Q, MODE, BP>(m, a, st, fast);                                                      \
    }
// ... with the fast variants in ghm_tree_fast_q<Q>_<TAG>_s<S>.cu (GHM_TREE_FAST_DEFINE)
#define GHM_TREE_FAST_NAME(Q, TAG, S) ghm_tree_fast_q##Q##_##TAG##_s##S
#define GHM_TREE_FAST_DEFINE(Q, TAG, MODE, BP, S)                                                                  \
    int GHM_TREE_FAST_NAME(Q, TAG, S)(const ghm_model* m, const TreeArgs& a, cudaStream_t st) {                     \
        return launch_fast<Q, S, MODE, BP>(m, a, st);                                                              \
    }
#define GHM_TREE_DEFINE_SPLIT(Q, TAG, MODE, BP)                                                                    \
    int GHM_TREE_FAST_NAME(Q, TAG, 2)(const ghm_model*, const TreeArgs&, cudaStream_t);                             \
    int GHM_TREE_FAST_NAME(Q, TAG, 3)(const ghm_model*, const TreeArgs&, cudaStream_t);                             \
    int GHM_TREE_FAST_NAME(Q, TAG, 4)(const ghm_model*, const TreeArgs&, cudaStream_t);                             \
    int ghm_tree_run_q##Q##_##TAG(const ghm_model* m, const TreeArgs& a, cudaStream_t st) {                         \
        static const ghm_fast_fn fast[3] = {&GHM_TREE_FAST_NAME(Q, TAG, 2), &GHM_TREE_FAST_NAME(Q, TAG, 3),         \
                                            &GHM_TREE_FAST_NAME(Q, TAG, 4)};                                        \
        return dispatch_variant<Q, MODE, BP>(m, a, st, fast);                                                      \
    }
