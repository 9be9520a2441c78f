// ghm_common.cuh -- shared device/host definitions for libghm_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/ghm_b200.h"

#define GHM_MAX_LEVELS 16          // L <= 16
#define GHM_MAX_Q_REG  16          // register-resident kernels: q <= 16 (padded to Q in {4,8,10,16})

// ------------------------------------------------------------------------------------
// Device-visible model descriptor, passed BY VALUE as a kernel parameter.
//
// Tables (one entry per matrix m; translation-invariant models have L*s matrices, per-edge
// models have E).  All rows/columns are padded with zeros from q up to QP = padded q:
//   Tlin  [m][a][b]   f32  P(child=b | parent=a)              (up / down matvecs)
//   TlinT [m][b][a]   f32  same, transposed                   (leaf column gather, linear)
//   TlogT [m][b][a]   f32  natural log of the above           (leaf column gather, log domain)
//   TTp   [m][b][a]   f32  TlinT with the row stride padded to QS = roundup(QP, 4) floats, so every
//                          row starts 16-byte aligned (LDS.128 / packed f32x2 operands)
//   alias [m][a][k]   u32  Walker alias table of row T[a][:], entry = (thr24 << 8) | alias8, stride q:
//                          draw u32 r -> m = r*q; k = m >> 32; frac = (u32)m; child = frac < e ? k : e & 255
//                          (Philox mode: O(1) per draw; oracle/philox.py builds the identical table)
//   cdfd  [m][a][k]   f64  sequential cumsum, unpadded stride q                     (parity mode)
// ------------------------------------------------------------------------------------
struct GhmDev {
    int L, s, q, QP, QS, QW, ti;                // QW = roundup(q, 32): row stride of the wide (q > 16) path, 0 when q <= 16
    int n_mat;
    int n_leaves;                          // s^L
    int n_edges;                           // sum_{l=1..L} s^l
    int mat_off[GHM_MAX_LEVELS + 1];       // mat_off[l]: first matrix of edges INTO depth l (l=1..L)
    int edge_off[GHM_MAX_LEVELS + 1];      // edge_off[l]: BFS edge index of node 0 at depth l (l=1..L)
    int spow[GHM_MAX_LEVELS + 1];          // s^l
    unsigned s_magic;                      // ceil(2^32 / s): idx / s == __umulhi(idx, s_magic) for idx < 2^28
    unsigned pow_magic[GHM_MAX_LEVELS + 1]; // ceil(2^32 / s^k), k>=1: j / s^k == __umulhi(j, .) for j*s^k < 2^32
    const float* Tlin;
    const float* TlinT;
    const float* TlogT;
    const float* TTp;
    const float* Wup;                      // wide path: [m][a][b] f32 = T[a][b], stride QW, zero padded
    const float* Wdn;                      // wide path: [m][b][a] f32 = T[a][b], stride QW, zero padded
    const uint32_t* alias;
    const double* cdfd;
    const float* py;                       // [QP] prior, zero padded
    const uint32_t* root_cdfu_prior;       // [QP]
    const uint32_t* root_cdfu_unif;        // [QP]
    const float* leaf_memo;                // [s][q^s][QS] f32: message of a depth-(L-1) node to its parent as a function of
                                           //   (child number, its s leaf states), see ghm_leaf_memo_rows; null when not built
    int* status;                           // sticky device status word
};

// Leaf memo of the fused sampler + root-posterior kernel (ghm_tree_fast.cuh).  The upward message a depth-(L-1) node
// sends to its parent, u = T_cj normalise(prod_c T_c^T[x_c, :]) (reference :191-208), depends on nothing but the
// node's child number cj and the states (x_0 .. x_{s-1}) of its s leaves, so for translation-invariant tables it is
// tabulated once per table upload (s * q^s rows of QS floats, built on the device by k_build_leaf_memo with the very
// operations the kernel used to perform per node) and the kernel gathers one row instead of s leaf rows + a q x q
// matvec.  Built when the whole table stays small against L2 (<= 1 MB at the padded q).
#define GHM_MEMO_MAX_BYTES (1u << 20)
constexpr bool ghm_memo_ok(int Q, int S) {
#ifdef GHM_NO_LEAF_MEMO                                           // A/B builds only
    return false;
#endif
    if (S < 2 || S > 4) return false;
    size_t rows = (size_t)S;
    for (int i = 0; i < S; ++i) rows *= (size_t)Q;
    return rows * (size_t)((Q + 3) / 4 * 4) * 4 <= GHM_MEMO_MAX_BYTES;
}
static inline size_t ghm_leaf_memo_rows(const GhmDev& d) {      // 0: this model has no memo
    if (!d.ti || d.L < 3 || d.QW || !ghm_memo_ok(d.QP, d.s)) return 0;
    size_t rows = (size_t)d.s;
    for (int i = 0; i < d.s; ++i) rows *= (size_t)d.q;
    return rows;
}

// Source-offset tables of the fused guide kernels (ghm_guides.cu): for every 4- or 8-byte unit of a tree's row in
// every guide tensor, the shared-memory offset (in floats) of the message element it copies.  A pure function of
// (L, s, q), so built once at model creation.
#define GHM_EXP_MAX_T (2 * GHM_MAX_LEVELS + 1)
struct GhmGuideTab {
    int G;                                 // trees per CTA the offsets were built for; 0 = fused kernel not usable
    int n_t;                               // tensors in the guide set
    int W;                                 // floats per unit (2 when q is even)
    int stride;                            // floats per tree per message array in shared memory
    int tab_off[GHM_EXP_MAX_T];            // first table entry of tensor t
    int upt[GHM_EXP_MAX_T];                // units per tree row of tensor t
};

struct ghm_model {
    GhmDev d;
    GhmGuideTab gt_cls, gt_dns;
    uint16_t* guide_tab;                   // device: cls tables followed by dns tables
    int device;
    int gemm_mode;       // GHM_GEMM_*: arithmetic of the wide path's row-GEMMs
    void* h_slab;        // pinned host image of the slab (table derivation target, source of H2D uploads)
    float* h_TTp;        // -> TTp inside h_slab (source of the constant-bank kernel parameter)
    float* h_Tlin;       // -> Tlin / TlinT inside h_slab (constant-bank parameter of k_dns2)
    float* h_TlinT;
    cudaEvent_t upload_done;   // fences the last H2D table upload ONLY (ghm_model_update waits on it before rewriting h_slab)
    cudaEvent_t order_ev;      // plain stream-ordering event of the ghm_host_* entry points
    void* slab;          // device allocation holding every table (the ACTIVE one of slabs[])
    size_t slab_bytes;
    // ghm_model_update double-buffers the tables: update u fills (h_slabs / slabs)[u & 1] and re-points `d`, so kernels
    // launched before it keep reading the other slab (GhmDev travels by value) and consecutive evaluations of a p_flip
    // sweep can overlap on different streams.  slabs[0] is the slab of ghm_model_create and owns the sticky status word.
    void* slabs[2];
    void* h_slabs[2];
    cudaEvent_t upload_evs[2];
    int active;
    cudaStream_t stream; // internal stream for ghm_host_* entry points
    // host scratch for ghm_host_* (lazily sized)
    void* h_scratch; size_t h_scratch_bytes;
    void* d_scratch; size_t d_scratch_bytes;
};

// Makes `dev` (or the device that owns `ptr`) current for the lifetime of the guard: every C-ABI entry point launches
// on the device of its model / buffers regardless of the caller's current device.
struct GhmDeviceGuard {
    int prev = -1;
    explicit GhmDeviceGuard(int dev) { enter(dev); }
    explicit GhmDeviceGuard(const void* ptr) {
        cudaPointerAttributes at{};
        if (cudaPointerGetAttributes(&at, ptr) == cudaSuccess && at.type == cudaMemoryTypeDevice) enter(at.device);
        else cudaGetLastError();
    }
    ~GhmDeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
    GhmDeviceGuard(const GhmDeviceGuard&) = delete;
    GhmDeviceGuard& operator=(const GhmDeviceGuard&) = delete;
private:
    void enter(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
};

int ghm_guides_init(ghm_model* m);        // ghm_guides.cu: builds guide_tab (current device = m->device)
int ghm_build_leaf_memo(const ghm_model* m, cudaStream_t st);   // ghm_tree.cu: fills d.leaf_memo from d.TTp on `st` (no-op without a memo)

// ------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------
void ghm_set_error(const std::string& msg);
int ghm_fail(int code, const char* fmt, ...);

#define GHM_CUDA_TRY(expr)                                                                 \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess)                                                             \
            return ghm_fail(GHM_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                            __FILE__, __LINE__);                                           \
    } while (0)

#define GHM_CHECK_LAUNCH()                                                                 \
    do {                                                                                   \
        cudaError_t _e = cudaGetLastError();                                               \
        if (_e != cudaSuccess)                                                             \
            return ghm_fail(GHM_ECUDA, "kernel launch failed: %s (%s:%d)",                 \
                            cudaGetErrorString(_e), __FILE__, __LINE__);                   \
    } while (0)

// padded q used by the register-resident kernels; 0 when q is outside their range
static inline int ghm_pad_q(int q) {
    if (q <= 4) return 4;
    if (q <= 8) return 8;
    if (q <= 10) return 10;
    if (q <= 16) return 16;
    return 0;
}

#ifdef __CUDACC__
// ------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011), counter = (tree_lo, tree_hi, block, level|stream<<8),
// key = (seed_lo, seed_hi).  One call yields 4 x 32 random bits; word (node & 3) of block
// (node >> 2) of level l is the draw for BFS node `node` at depth l.  Restated in numpy by
// oracle/philox.py so Philox-mode samples are checkable bit-for-bit.
// ------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
        uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += W0;
        k.y += W1;
    }
    return c;
}

#define GHM_STREAM_TREE  0u
#define GHM_STREAM_NOISE 1u

__device__ __forceinline__ uint4 ghm_rng_block(uint64_t seed, uint64_t tree, uint32_t level, uint32_t block,
                                               uint32_t stream) {
    return philox4x32_10(make_uint4((uint32_t)tree, (uint32_t)(tree >> 32), block, level | (stream << 8)),
                         make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
}

__device__ __forceinline__ uint32_t ghm_pick(const uint4& v, int w) {
    return w == 0 ? v.x : (w == 1 ? v.y : (w == 2 ? v.z : v.w));
}

// idx / s for uniform small integers
__device__ __forceinline__ int ghm_div_s(int idx, const GhmDev& d) {
    return d.s == 1 ? idx : (int)__umulhi((unsigned)idx, d.s_magic);
}
// j / s^k (k >= 0)
__device__ __forceinline__ int ghm_div_pow(int j, int k, const GhmDev& d) {
    return (k == 0 || d.s == 1) ? j : (int)__umulhi((unsigned)j, d.pow_magic[k]);
}
#endif  // __CUDACC__
