// ghm_guides.cu -- K5: guide tensors for BP_CLS / BP_DNS, written in the layout the guided
// losses index ([B, n_L, C] float32).
//
// Replaces GHMTree.guided_info (reference src/ghmclip/data/data_random_GHM.py:526-592), which
// in the reference costs 80 % of get_batch(guide=True) (Python list-extend + np.array + two
// transposes).  Two phases:
//   1. a LOG-DOMAIN belief-propagation kernel with the reference's exact shift conventions
//      (SURVEY.md Appendix A "shift points") stores every node message compactly as
//      M[b][node][q] float32 (thread per tree, level-synchronous).  It is a second, independent
//      implementation of BP_CLS (:185-221) / BP_DNS (:467-523) next to the linear-domain DFS
//      kernels of ghm_tree.cu / ghm_dns.cu, and the tests cross-check the two.
//   2. k_expand: a pure streaming kernel that broadcasts each depth-l node message over the
//      s^(L-l) leaves below it and concatenates (hd | qd | bu): out[b][i][j*q+k] =
//      SRC_j[b][node(l, i / s^(L-l))][k].  This is the HBM-write-bound part (13-71 KB per tree).
#include <string.h>

#include <algorithm>
#include <type_traits>
#include <vector>

#include "ghm_vec2.cuh"
#include "ghm_wide_lvl.cuh"

#define GD_NT 128

struct LvlArgs {
    int64_t B;
    const void* leaves; int leaf_dtype;     // cls
    const float* z; float sigma;            // dns
    const float* ext;                       // [B,q] or null
    float* HD; float* QD; float* BU;        // [B][n_nodes][q]
    int n_nodes;                            // nodes stored per tree
    float* post; float* root_hd; float* mean;
};

// compact message rows are q floats, 8-byte aligned when q is even: move them as float2 (half the memory
// transactions of scalar accesses at a 4q-byte stride)
template <int Q>
__device__ __forceinline__ void load_vec(const float* p, float (&v)[Q], int q) {
    if ((q & 1) == 0 && (reinterpret_cast<uintptr_t>(p) & 7) == 0) {
        const float2* p2 = reinterpret_cast<const float2*>(p);
#pragma unroll
        for (int k = 0; k < Q; k += 2) {
            if (k < q) { const float2 t = p2[k >> 1]; v[k] = t.x; v[k + 1] = t.y; }
            else { v[k] = -INFINITY; v[k + 1] = -INFINITY; }
        }
    } else {
#pragma unroll
        for (int k = 0; k < Q; ++k) v[k] = (k < q) ? p[k] : -INFINITY;
    }
}
template <int Q>
__device__ __forceinline__ void store_vec(float* p, const float (&v)[Q], int q) {
    if ((q & 1) == 0 && (reinterpret_cast<uintptr_t>(p) & 7) == 0) {
        float2* p2 = reinterpret_cast<float2*>(p);
#pragma unroll
        for (int k = 0; k < Q; k += 2)
            if (k < q) p2[k >> 1] = make_float2(v[k], v[k + 1]);
    } else {
#pragma unroll
        for (int k = 0; k < Q; ++k)
            if (k < q) p[k] = v[k];
    }
}

// out[a] = log(sum_b T[a][b] exp(h[b] - m)) + m   with m = max h   (== log(T @ exp(h)) without overflow).
// MUFU-based __expf / __logf: arguments are <= 0 / sums in (0, q], where their error (<= 2 ulp resp. 2^-21.4 absolute)
// is far inside the 1e-5 budget; the precise library versions made these kernels instruction bound.
template <int Q>
__device__ __forceinline__ void log_matvec(const float* __restrict__ T, const float (&h)[Q], float (&out)[Q], int q) {
    const float m = ghm_vmax<Q>(h);
    float e[Q], u[Q];
#pragma unroll
    for (int k = 0; k < Q; ++k) e[k] = (k < q) ? __expf(h[k] - m) : 0.f;
    ghm_matvec<Q>(T, e, u);
#pragma unroll
    for (int k = 0; k < Q; ++k) out[k] = (k < q) ? __logf(u[k]) + m : -INFINITY;
}
template <int Q>
__device__ __forceinline__ void log_matvec_t(const float* __restrict__ T, const float (&h)[Q], float (&out)[Q], int q) {
    const float m = ghm_vmax<Q>(h);
    float e[Q], u[Q];
#pragma unroll
    for (int k = 0; k < Q; ++k) e[k] = (k < q) ? __expf(h[k] - m) : 0.f;
    ghm_matvec_t<Q>(T, e, u);
#pragma unroll
    for (int k = 0; k < Q; ++k) out[k] = (k < q) ? __logf(u[k]) + m : -INFINITY;
}

__device__ __forceinline__ int leaf_at(const void* leaves, int dtype, int64_t off, int q, int* status) {
    int64_t v = dtype == GHM_LEAF_I64 ? reinterpret_cast<const int64_t*>(leaves)[off]
                                      : (int64_t) reinterpret_cast<const uint8_t*>(leaves)[off];
    if (v < 0 || v >= q) { atomicOr(status, 1); v = v < 0 ? 0 : q - 1; }
    return (int)v;
}

// node id of (depth l, index idx) when depths 0..L are stored: off_all(l) = (s^l - 1)/(s - 1)
__device__ __forceinline__ int node_off_all(const GhmDev& d, int l) { return l == 0 ? 0 : 1 + d.edge_off[l]; }

// One THREAD per (tree, node) of one tree level, one launch per level: B * s^l threads, consecutive threads own
// consecutive nodes of a tree, so the compact [B][node][q] message rows they read (children) and write are
// adjacent 4q-byte rows (coalesced).  The level-serial, thread-per-tree form of round-1a left the SMs at 20 %
// occupancy behind long dependent chains and was 3x slower than the expansion it feeds.
struct LvlThread { int64_t b; int idx; bool ok; };
__device__ __forceinline__ LvlThread lvl_thread(const GhmDev& d, int64_t B, int l) {
    const int64_t t = (int64_t)blockIdx.x * GD_NT + threadIdx.x;
    const int n = d.spow[l];
    LvlThread r;
    r.b = t / n;
    r.idx = (int)(t - r.b * n);
    r.ok = r.b < B;
    return r;
}
static unsigned lvl_grid(const GhmDev& d, int64_t B, int l) { return (unsigned)((B * d.spow[l] + GD_NT - 1) / GD_NT); }

// ---- BP_CLS, log domain (reference :185-221): nodes of depth l, bottom-up --------------------------
// node (depth l, idx) of tree b; HD = this tree's compact rows [n_int][q] (global or shared memory)
// EX: q == Q, so the padding predicates (k < q) fold away at compile time
template <int Q, bool EX>
__device__ __forceinline__ void cls_node(const GhmDev& d, const LvlArgs& a, int64_t b, int l, int idx, float* HD) {
    const int L = d.L, s = d.s, q = EX ? Q : d.q, nL = d.n_leaves;
    float acc[Q];
#pragma unroll
    for (int k = 0; k < Q; ++k) acc[k] = (k < q) ? 0.f : -INFINITY;
    for (int c = 0; c < s; ++c) {
        const int child = idx * s + c;
        const int mi = d.mat_off[l + 1] + (d.ti ? c : child);
        if (l == L - 1) {
            const int x = leaf_at(a.leaves, a.leaf_dtype, b * nL + child, q, d.status);
            const float* row = d.TlogT + ((size_t)mi * Q + x) * Q;          // log T[:, x]   (:196)
#pragma unroll
            for (int k = 0; k < Q; ++k)
                if (k < q) acc[k] += __ldg(row + k);
        } else {
            float h[Q], m[Q];
            load_vec<Q>(HD + (int64_t)(node_off_all(d, l + 1) + child) * q, h, q);
            log_matvec<Q>(d.Tlin + (size_t)mi * Q * Q, h, m, q);             // log(T @ exp(hd))  (:207)
#pragma unroll
            for (int k = 0; k < Q; ++k)
                if (k < q) acc[k] += m[k];
        }
    }
    const float mx = ghm_vmax<Q>(acc);
#pragma unroll
    for (int k = 0; k < Q; ++k) acc[k] -= mx;                                 // (:197,208)
    store_vec<Q>(HD + (int64_t)(node_off_all(d, l) + idx) * q, acc, q);
    if (l > 0) return;
    // root: acc == hd(root)
    if (a.root_hd) store_vec<Q>(a.root_hd + b * q, acc, q);
    if (a.post) {
        float h0[Q];
#pragma unroll
        for (int k = 0; k < Q; ++k) h0[k] = (k < q) ? acc[k] + logf(__ldg(d.py + k)) : -INFINITY;   // (:213)
        const float m0 = ghm_vmax<Q>(h0);
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < Q; ++k) { h0[k] = (k < q) ? expf(h0[k] - m0) : 0.f; sum += h0[k]; }
        const float inv = 1.0f / sum;
#pragma unroll
        for (int k = 0; k < Q; ++k) h0[k] *= inv;
        store_vec<Q>(a.post + b * q, h0, q);
    }
}

template <int Q>
__global__ void __launch_bounds__(GD_NT) k_lvl_cls(const GhmDev d, const LvlArgs a, int l) {
    const LvlThread th = lvl_thread(d, a.B, l);
    if (!th.ok) return;
    cls_node<Q, false>(d, a, th.b, l, th.idx, a.HD + th.b * (int64_t)a.n_nodes * d.q);
}

// ---- BP_DNS, log domain (reference :467-523): upward pass, nodes of depth l ------------------------
template <int Q, bool EX>
__device__ __forceinline__ void dns_up_node(const GhmDev& d, const LvlArgs& a, int64_t b, int l, int idx, float* HD, float* QD,
                                            float* BU) {
    const int L = d.L, s = d.s, q = EX ? Q : d.q, nL = d.n_leaves;
    const int node = node_off_all(d, l) + idx;
    float acc[Q];
    if (l == L) {                                             // leaves: hd unshifted (:485), qd = log(T @ exp(hd)) (:487)
        const float inv2s2 = 0.5f / (a.sigma * a.sigma);
        const float zi = a.z[b * nL + idx];
#pragma unroll
        for (int k = 0; k < Q; ++k) { const float dlt = zi - (float)k; acc[k] = (k < q) ? -dlt * dlt * inv2s2 : -INFINITY; }
    } else {                                                  // hd = sum qd(children) - max (:494-496)
#pragma unroll
        for (int k = 0; k < Q; ++k) acc[k] = (k < q) ? 0.f : -INFINITY;
        for (int c = 0; c < s; ++c) {
            float m[Q];
            load_vec<Q>(QD + (int64_t)(node_off_all(d, l + 1) + idx * s + c) * q, m, q);
#pragma unroll
            for (int k = 0; k < Q; ++k)
                if (k < q) acc[k] += m[k];
        }
        const float mx = ghm_vmax<Q>(acc);
#pragma unroll
        for (int k = 0; k < Q; ++k) acc[k] -= mx;
    }
    if (l > 0) {
        const int mi = d.mat_off[l] + (d.ti ? idx - ghm_div_s(idx, d) * s : idx);
        float m[Q];
        log_matvec<Q>(d.Tlin + (size_t)mi * Q * Q, acc, m, q);                // qd = log(T @ exp(hd)) (:497)
        store_vec<Q>(HD + (int64_t)node * q, acc, q);
        store_vec<Q>(QD + (int64_t)node * q, m, q);
        return;
    }
    // root: bu aliases hd and gets the external message without a re-shift (:501-506)
    if (a.ext) {
#pragma unroll
        for (int k = 0; k < Q; ++k)
            if (k < q) acc[k] += a.ext[b * q + k];
    }
    store_vec<Q>(HD, acc, q);
    store_vec<Q>(BU, acc, q);
}

template <int Q>
__global__ void __launch_bounds__(GD_NT) k_lvl_dns_up(const GhmDev d, const LvlArgs a, int l) {
    const LvlThread th = lvl_thread(d, a.B, l);
    if (!th.ok) return;
    const int64_t base = th.b * (int64_t)a.n_nodes * d.q;
    dns_up_node<Q, false>(d, a, th.b, l, th.idx, a.HD + base, a.QD + base, a.BU + base);
}

// ---- BP_DNS downward pass: bu = hd + log(T^T @ exp(bu_parent - qd)) - max (:509-514), nodes of depth l >= 1 ----
template <int Q, bool EX>
__device__ __forceinline__ void dns_down_node(const GhmDev& d, const LvlArgs& a, int64_t b, int l, int idx, float* HD, float* QD,
                                              float* BU) {
    const int L = d.L, s = d.s, q = EX ? Q : d.q, nL = d.n_leaves;
    const int node = node_off_all(d, l) + idx;
    const int pidx = ghm_div_s(idx, d);
    const int par = node_off_all(d, l - 1) + pidx;
    float bp[Q], qv[Q], hv[Q], m[Q];
    load_vec<Q>(BU + (int64_t)par * q, bp, q);
    load_vec<Q>(QD + (int64_t)node * q, qv, q);
    load_vec<Q>(HD + (int64_t)node * q, hv, q);
#pragma unroll
    for (int k = 0; k < Q; ++k) bp[k] = (k < q) ? bp[k] - qv[k] : -INFINITY;
    const int mi = d.mat_off[l] + (d.ti ? idx - pidx * s : idx);
    log_matvec_t<Q>(d.Tlin + (size_t)mi * Q * Q, bp, m, q);
#pragma unroll
    for (int k = 0; k < Q; ++k) m[k] = (k < q) ? hv[k] + m[k] : -INFINITY;
    const float mx = ghm_vmax<Q>(m);
#pragma unroll
    for (int k = 0; k < Q; ++k) m[k] -= mx;
    store_vec<Q>(BU + (int64_t)node * q, m, q);
    if (l == L && a.mean) {                                                   // (:516-519)
        float num = 0.f, den = 0.f;
#pragma unroll
        for (int k = 0; k < Q; ++k) { const float e = (k < q) ? expf(m[k]) : 0.f; num += (float)k * e; den += e; }
        a.mean[b * nL + idx] = num / den;
    }
}

template <int Q>
__global__ void __launch_bounds__(GD_NT) k_lvl_dns_down(const GhmDev d, const LvlArgs a, int l) {
    const LvlThread th = lvl_thread(d, a.B, l);
    if (!th.ok) return;
    const int64_t base = th.b * (int64_t)a.n_nodes * d.q;
    dns_down_node<Q, false>(d, a, th.b, l, th.idx, a.HD + base, a.QD + base, a.BU + base);
}

// ---- expansion: node messages -> [B, n_L, C] guide tensors ------------------------------------------
// ONE launch writes every tensor of a guide set (blockIdx.y = tensor).  A CTA streams whole tree rows: the
// row of tree b in tensor t is n_L * C contiguous floats, written as consecutive 8-byte units by consecutive
// lanes (fully coalesced); unit u = (leaf i, pair j) reads the float2 (k, k+1) of part j*2/q (hd | qd | bu)
// of the depth-l ancestor of leaf i from the compact [B][node][q] store (L2 resident, 8-byte aligned for even q).
#define EXP_MAX_T GHM_EXP_MAX_T
struct ExpandT {
    float* out;              // [B][nL][C]
    int level;               // depth of the broadcast nodes
    int nsrc;                // C = nsrc * q
    int src[3];              // index into ExpandAll::src for each part
};
struct ExpandAll {
    int64_t B;
    int n_t, n_nodes, q, nL, L;
    int node_off[GHM_MAX_LEVELS + 1];
    int R[GHM_MAX_LEVELS + 1]; unsigned R_magic[GHM_MAX_LEVELS + 1];      // leaves per depth-l node
    const float* src[3];     // message arrays [B][n_nodes][q]
    ExpandT t[EXP_MAX_T];
};

__device__ __forceinline__ int div_magic(int x, int div, unsigned magic) {
    return div == 1 ? x : (int)__umulhi((unsigned)x, magic);
}
static unsigned magic_of(int d) { return d >= 2 ? (unsigned)((0x100000000ull + (unsigned)d - 1) / (unsigned)d) : 0u; }

template <bool VEC2>
__global__ void __launch_bounds__(256) k_expand_all(const __grid_constant__ ExpandAll a) {
    const ExpandT& t = a.t[blockIdx.y];
    const int q = a.q, W = VEC2 ? 2 : 1;
    const int hq = q / W;                                   // units per part
    const int C2 = t.nsrc * hq;                             // units per leaf cell
    const int upt = a.nL * C2;                              // units per tree row
    const unsigned C2_magic = C2 >= 2 ? (unsigned)((0x100000000ull + (unsigned)C2 - 1) / (unsigned)C2) : 0u;
    const unsigned hq_magic = hq >= 2 ? (unsigned)((0x100000000ull + (unsigned)hq - 1) / (unsigned)hq) : 0u;
    const int R = a.R[t.level]; const unsigned Rm = a.R_magic[t.level];
    const int noff = a.node_off[t.level];
    const float* s0 = a.src[t.src[0]];
    const float* s1 = a.src[t.src[t.nsrc > 1 ? 1 : 0]];
    const float* s2 = a.src[t.src[t.nsrc > 2 ? 2 : 0]];
    for (int64_t b = blockIdx.x; b < a.B; b += gridDim.x) {
        float* out = t.out + b * (int64_t)upt * W;
        const int64_t sbase = b * (int64_t)a.n_nodes * q;
#pragma unroll 4
        for (int u = threadIdx.x; u < upt; u += 256) {
            const int i = div_magic(u, C2, C2_magic);
            const int j = u - i * C2;
            const int part = div_magic(j, hq, hq_magic);
            const int kk = j - part * hq;
            const int node = noff + div_magic(i, R, Rm);
            const float* src = (part == 0 ? s0 : (part == 1 ? s1 : s2)) + sbase + (int64_t)node * q + kk * W;
            if (VEC2) reinterpret_cast<float2*>(out)[u] = __ldg(reinterpret_cast<const float2*>(src));
            else out[u] = __ldg(src);
        }
    }
}

// queue tensors into an ExpandAll, then one launch
static void expand_setup(ExpandAll& a, const GhmDev& d, int64_t B, int n_nodes, const float* s0, const float* s1,
                         const float* s2) {
    memset(&a, 0, sizeof a);
    a.B = B; a.n_nodes = n_nodes; a.q = d.q; a.nL = d.n_leaves; a.L = d.L;
    a.src[0] = s0; a.src[1] = s1; a.src[2] = s2;
    for (int l = 0; l <= d.L; ++l) {
        a.node_off[l] = l == 0 ? 0 : 1 + d.edge_off[l];
        a.R[l] = d.spow[d.L - l];
        a.R_magic[l] = magic_of(a.R[l]);
    }
}
static void expand_add(ExpandAll& a, float* out, int level, int nsrc, int p0, int p1, int p2) {
    ExpandT& t = a.t[a.n_t++];
    t.out = out; t.level = level; t.nsrc = nsrc; t.src[0] = p0; t.src[1] = p1; t.src[2] = p2;
}
static int expand_launch(const ExpandAll& a, cudaStream_t st) {
    if (a.n_t == 0) return GHM_OK;
    if ((int64_t)a.nL * 3 * a.q >= (1ll << 30)) return ghm_fail(GHM_EUNSUP, "guide row too large");
    dim3 grid((unsigned)std::min<int64_t>(a.B, 148 * 32), (unsigned)a.n_t);
    bool vec2 = (a.q % 2) == 0;
    for (int i = 0; i < a.n_t; ++i) vec2 = vec2 && ((uintptr_t)a.t[i].out % 8) == 0;
    for (int i = 0; i < 3; ++i) vec2 = vec2 && ((uintptr_t)a.src[i] % 8) == 0;
    if (vec2) k_expand_all<true><<<grid, 256, 0, st>>>(a); else k_expand_all<false><<<grid, 256, 0, st>>>(a);
    GHM_CHECK_LAUNCH();
    return GHM_OK;
}

// ------------------------------------------------------------------------------------------------
// Tree-tiled fused guide kernels: a CTA owns G consecutive trees, runs the whole log-domain BP level by level on
// compact message rows held in SHARED memory (no HBM round trip of the node messages, no per-level launches), then
// streams every guide tensor of its trees straight from shared memory.  HBM traffic = the inputs + the guide
// tensors, i.e. the algorithmic bytes.
//
// Expansion is table driven: which message element a unit of a tree row copies depends only on (L, s, q), so the
// model carries one uint16 shared-memory offset per unit (ghm_guides_init).  A thread owns unit u of the row, reads
// its offset once and copies that unit for each of the CTA's trees: per store one LDS and one STG, no index
// arithmetic; consecutive lanes write consecutive 8-byte units of a row (fully coalesced).
// ------------------------------------------------------------------------------------------------
// Levels near the root keep only the first warp(s) of a CTA busy, and warp w of every CTA lives on scheduler w % 4:
// rotate the thread -> work-item mapping by (CTA, level) so those single-warp rounds spread over the four schedulers.
__device__ __forceinline__ int rot_tid(int l) {
    return (int)((threadIdx.x + 32u * (blockIdx.x + (unsigned)l)) & (blockDim.x - 1u));     // block size is a power of two
}

#define GF_NT_CLS 128                                        // measured best: 128 threads for the cls set, 256 for the dns set
#define GF_NT_DNS 256
struct FusedExp {
    const uint16_t* tab;
    int n_t, stride;
    int tab_off[GHM_EXP_MAX_T];
    int upt[GHM_EXP_MAX_T];
    float* out[GHM_EXP_MAX_T];
};

#define GF_CH 6                                              // table entries a thread holds per chunk
template <bool VEC2, int NT>
__device__ __forceinline__ void expand_tab(const FusedExp& e, const float* gsm, int64_t tree0, int g) {
    typedef typename std::conditional<VEC2, float2, float>::type U;
    const char* sbase = reinterpret_cast<const char*>(gsm);
    const unsigned tstride = (unsigned)e.stride * 4u;          // bytes per tree per array
    // work list = (tensor, chunk of GF_CH * NT units).  Global loads queue behind the stores already in the memory
    // pipe, so the offsets of the NEXT chunk are fetched before the stores of the current one are issued.  Within a
    // chunk the tree is the outer loop: the GF_CH stores of a tree sit at compile-time distances (k * NT units) from
    // one pointer, so a store costs one shared-memory address add, one LDS and one STG.
    auto fetch = [&](int ti, int base, unsigned (&o)[GF_CH]) {
        const uint16_t* __restrict__ tab = e.tab + e.tab_off[ti] + base + threadIdx.x;
        const int left = e.upt[ti] - base - (int)threadIdx.x;
#pragma unroll
        for (int k = 0; k < GF_CH; ++k) o[k] = (k * NT < left) ? (unsigned)__ldg(tab + k * NT) * 4u : 0u;
    };
    unsigned cur[GF_CH], nxt[GF_CH];
    int ti = 0, base = 0;
    if (e.n_t > 0) fetch(0, 0, cur);
    while (ti < e.n_t) {
        const int upt = e.upt[ti];
        int nti = ti, nbase = base + GF_CH * NT;
        if (nbase >= upt) { ++nti; nbase = 0; }
        if (nti < e.n_t) fetch(nti, nbase, nxt);
        U* __restrict__ out = reinterpret_cast<U*>(e.out[ti]) + tree0 * upt + base + threadIdx.x;
        const int left = upt - base - (int)threadIdx.x;
        const char* sa = sbase;
#pragma unroll 2
        for (int tr = 0; tr < g; ++tr, sa += tstride, out += upt) {
#pragma unroll
            for (int k = 0; k < GF_CH; ++k) {
                if (k * NT < left) {
                    out[k * NT] = *reinterpret_cast<const U*>(sa + cur[k]);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < GF_CH; ++k) cur[k] = nxt[k];
        ti = nti; base = nbase;
    }
}

template <int Q, bool EX, bool VEC2>
__global__ void __launch_bounds__(GF_NT_DNS) k_guides_dns_fused(const GhmDev d, const LvlArgs a, const __grid_constant__ FusedExp e,
                                                            int G) {
    extern __shared__ __align__(16) float gsm[];
    const int L = d.L;
    const int stride = e.stride;                            // floats per tree per array
    float* HD = gsm; float* QD = HD + (size_t)G * stride; float* BU = QD + (size_t)G * stride;
    const int64_t tree0 = (int64_t)blockIdx.x * G;
    const int g = (int)min((int64_t)G, a.B - tree0);
    for (int l = L; l >= 0; --l) {
        const int n = d.spow[l];
        const unsigned nm = d.pow_magic[l];                 // ceil(2^32 / s^l), host-computed
        for (int w = rot_tid(l); w < g * n; w += blockDim.x) {
            const int t = div_magic(w, n, nm), idx = w - t * n;
            dns_up_node<Q, EX>(d, a, tree0 + t, l, idx, HD + t * stride, QD + t * stride, BU + t * stride);
        }
        __syncthreads();
    }
    for (int l = 1; l <= L; ++l) {
        const int n = d.spow[l];
        const unsigned nm = d.pow_magic[l];                 // ceil(2^32 / s^l), host-computed
        for (int w = rot_tid(l); w < g * n; w += blockDim.x) {
            const int t = div_magic(w, n, nm), idx = w - t * n;
            dns_down_node<Q, EX>(d, a, tree0 + t, l, idx, HD + t * stride, QD + t * stride, BU + t * stride);
        }
        __syncthreads();
    }
    expand_tab<VEC2, GF_NT_DNS>(e, gsm, tree0, g);
}

// ---- packed / constant-bank variant of the fused dns kernel (translation-invariant tables, q == Q) ----------------
// One thread per PARENT node, looping over its s children: the matrix of child c is the same for every thread of
// the CTA, so its index is a uniform loop counter and the Q*Q/2 packed FMAs of a matvec read the table from the
// constant bank through uniform registers -- no LSU traffic for the tables, which otherwise queues behind the
// guide-tensor stores of the co-resident CTAs (that queueing made BP and expansion time add up instead of overlap).
struct GuideC { int up_base, dn_off, s_u; };               // uniform-side twins: (L-1)*s, float offset of Tlin, s

template <int H>
__device__ __forceinline__ void f2_ld(const float* p, f2 (&v)[H]) {
#pragma unroll
    for (int i = 0; i < H; ++i) v[i] = reinterpret_cast<const f2*>(p)[i];
}
template <int H>
__device__ __forceinline__ void f2_st(float* p, const f2 (&v)[H]) {
#pragma unroll
    for (int i = 0; i < H; ++i) reinterpret_cast<f2*>(p)[i] = v[i];
}
template <int H>
__device__ __forceinline__ float f2_vmax(const f2 (&v)[H]) {
    float m = fmaxf(v[0].x, v[0].y);
#pragma unroll
    for (int i = 1; i < H; ++i) m = fmaxf(m, fmaxf(v[i].x, v[i].y));
    return m;
}
// out = log(M @ exp(h - max h)) + max h, M = the Q x Q table at T (constant bank, see f2_matvec_c)
template <int Q>
__device__ __forceinline__ void f2_log_matvec_c(const float* __restrict__ T, const f2 (&h)[Q / 2], f2 (&out)[Q / 2]) {
    const float m = f2_vmax<Q / 2>(h);
    f2 e[Q / 2], u[Q / 2];
#pragma unroll
    for (int i = 0; i < Q / 2; ++i) e[i] = make_float2(__expf(h[i].x - m), __expf(h[i].y - m));
    f2_matvec_c<Q>(T, e, u);
#pragma unroll
    for (int i = 0; i < Q / 2; ++i) out[i] = make_float2(__logf(u[i].x) + m, __logf(u[i].y) + m);
}

template <int Q, bool VEC2, int NW>
__global__ void __launch_bounds__(GF_NT_DNS)
k_guides_dns_fused_c(const __grid_constant__ GhmDev d, const __grid_constant__ LvlArgs a, const __grid_constant__ FusedExp e,
                     const __grid_constant__ GuideC gc, int G, const __grid_constant__ DnsTab<NW> tab) {
    extern __shared__ __align__(16) float gsm[];
    constexpr int H = Q / 2, QQ = Q * Q;
    const int L = d.L, s = d.s, nL = d.n_leaves;
    const int stride = e.stride, NT = blockDim.x, tid = threadIdx.x;
    float* HD = gsm; float* QD = HD + (size_t)G * stride; float* BU = QD + (size_t)G * stride;
    const int64_t tree0 = (int64_t)blockIdx.x * G;
    const int g = (int)min((int64_t)G, a.B - tree0);
    const float* Tup = tab.v;
    const float* Tdn = tab.v + gc.dn_off;
    const float inv2s2 = 0.5f / (a.sigma * a.sigma);
    // the CTA's noisy leaves, one coalesced read, parked in the BU region (first written at the root of the up pass,
    // after the leaf level has consumed them)
    float* zs = BU;
    for (int i = tid; i < g * nL; i += NT) zs[i] = a.z[tree0 * nL + i];
    __syncthreads();
    // upward pass (reference :483-506): parents of depth l = L-1 .. 0
    int tmb = gc.up_base;
    for (int l = L - 1; l >= 0; --l, tmb -= gc.s_u) {
        const int n = d.spow[l], items = g * n;
        const unsigned nm = d.pow_magic[l];                 // ceil(2^32 / s^l), host-computed
        const int noff_c = 1 + d.edge_off[l + 1], noff_p = l == 0 ? 0 : 1 + d.edge_off[l];
        const int rt = rot_tid(l);
        for (int w0 = 0; w0 < items; w0 += NT) {                // uniform trip count; the tail is predicated
            const bool act = w0 + rt < items;
            if (!__any_sync(0xffffffffu, act)) continue;        // warp-uniform skip (vote results stay on the uniform side)
            const int w = act ? w0 + rt : items - 1;
            const int t = div_magic(w, n, nm), idx = w - t * n;
            float* hd = HD + t * stride; float* qd = QD + t * stride;
            f2 acc[H];
#pragma unroll
            for (int i = 0; i < H; ++i) acc[i] = make_float2(0.f, 0.f);
            int tm = tmb;
#pragma unroll 1
            for (int c = 0; c < s; ++c, ++tm) {
                const int ci = idx * s + c;
                f2 h[H], m[H];
                if (l == L - 1) {                               // leaf hd, unshifted (:485)
                    const float zi = zs[t * nL + ci];
#pragma unroll
                    for (int i = 0; i < H; ++i) {
                        const float da = zi - (float)(2 * i), db = zi - (float)(2 * i + 1);
                        h[i] = make_float2(-da * da * inv2s2, -db * db * inv2s2);
                    }
                    if (act) f2_st<H>(hd + (noff_c + ci) * Q, h);
                } else {
                    f2_ld<H>(hd + (noff_c + ci) * Q, h);
                }
                // qd = log(T @ exp(hd)) (:487,497) = mh + r with mh = max hd.  The parent sums the SHIFTED parts r: the
                // omitted sum of the mh is constant over the states and cancels in the max-shift below, while leaf hd is
                // unshifted in the reference (:485) and reaches -600 .. -4400 at small sigma, where adding mh first
                // would round r to the float32 grid of |mh| (6e-5 at 600) before the cancellation
                const float mh = f2_vmax<H>(h);
                f2 ex[H], u[H];
#pragma unroll
                for (int i = 0; i < H; ++i) ex[i] = make_float2(__expf(h[i].x - mh), __expf(h[i].y - mh));
                f2_matvec_c<Q>(Tup + tm * QQ, ex, u);
#pragma unroll
                for (int i = 0; i < H; ++i) {
                    const float ra = __logf(u[i].x), rb = __logf(u[i].y);
                    m[i] = make_float2(ra + mh, rb + mh);
                    acc[i].x += ra; acc[i].y += rb;
                }
                if (act) f2_st<H>(qd + (noff_c + ci) * Q, m);
            }
            const float mx = f2_vmax<H>(acc);                   // hd = sum qd(children) - max (:494-496)
#pragma unroll
            for (int i = 0; i < H; ++i) { acc[i].x -= mx; acc[i].y -= mx; }
            if (l > 0) {
                if (act) f2_st<H>(hd + (noff_p + idx) * Q, acc);
            } else {                                            // root: bu aliases hd, external message unshifted (:501-506)
                if (a.ext) {
                    f2 x[H];
                    f2_ld<H>(a.ext + (tree0 + t) * Q, x);
#pragma unroll
                    for (int i = 0; i < H; ++i) { acc[i].x += x[i].x; acc[i].y += x[i].y; }
                }
                if (act) { f2_st<H>(hd, acc); f2_st<H>(BU + t * stride, acc); }
            }
        }
        __syncthreads();
    }
    // downward pass (:509-519): children of the depth-l parents
    int tmd = 0;
    for (int l = 0; l < L; ++l, tmd += gc.s_u) {
        const int n = d.spow[l], items = g * n;
        const unsigned nm = d.pow_magic[l];                 // ceil(2^32 / s^l), host-computed
        const int noff_c = 1 + d.edge_off[l + 1], noff_p = l == 0 ? 0 : 1 + d.edge_off[l];
        const int rt = rot_tid(l + 1);
        for (int w0 = 0; w0 < items; w0 += NT) {
            const bool act = w0 + rt < items;
            if (!__any_sync(0xffffffffu, act)) continue;        // warp-uniform skip (vote results stay on the uniform side)
            const int w = act ? w0 + rt : items - 1;
            const int t = div_magic(w, n, nm), idx = w - t * n;
            float* hd = HD + t * stride; float* qd = QD + t * stride; float* bu = BU + t * stride;
            f2 bp[H];
            f2_ld<H>(bu + (noff_p + idx) * Q, bp);
            int tm = tmd;
#pragma unroll 1
            for (int c = 0; c < s; ++c, ++tm) {
                const int ci = idx * s + c, node = noff_c + ci;
                f2 qv[H], hv[H], m[H];
                f2_ld<H>(qd + node * Q, qv);
                f2_ld<H>(hd + node * Q, hv);
#pragma unroll
                for (int i = 0; i < H; ++i) { qv[i].x = bp[i].x - qv[i].x; qv[i].y = bp[i].y - qv[i].y; }
                f2_log_matvec_c<Q>(Tdn + tm * QQ, qv, m);
#pragma unroll
                for (int i = 0; i < H; ++i) { m[i].x += hv[i].x; m[i].y += hv[i].y; }
                const float mx = f2_vmax<H>(m);
#pragma unroll
                for (int i = 0; i < H; ++i) { m[i].x -= mx; m[i].y -= mx; }
                if (act) f2_st<H>(bu + node * Q, m);
                if (l == L - 1 && a.mean) {
                    float num = 0.f, den = 0.f;
#pragma unroll
                    for (int i = 0; i < H; ++i) {
                        const float ea = __expf(m[i].x), eb = __expf(m[i].y);
                        num += (float)(2 * i) * ea + (float)(2 * i + 1) * eb; den += ea + eb;
                    }
                    if (act) a.mean[(tree0 + t) * nL + ci] = __fdividef(num, den);
                }
            }
        }
        __syncthreads();
    }
    expand_tab<VEC2, GF_NT_DNS>(e, gsm, tree0, g);
}

template <int Q, bool EX, bool VEC2>
__global__ void __launch_bounds__(GF_NT_CLS) k_guides_cls_fused(const GhmDev d, const LvlArgs a, const __grid_constant__ FusedExp e,
                                                            int G) {
    extern __shared__ __align__(16) float gsm[];
    const int L = d.L;
    const int stride = e.stride;
    float* HD = gsm;
    const int64_t tree0 = (int64_t)blockIdx.x * G;
    const int g = (int)min((int64_t)G, a.B - tree0);
    for (int l = L - 1; l >= 0; --l) {
        const int n = d.spow[l];
        const unsigned nm = d.pow_magic[l];                 // ceil(2^32 / s^l), host-computed
        for (int w = rot_tid(l); w < g * n; w += blockDim.x) {
            const int t = div_magic(w, n, nm), idx = w - t * n;
            cls_node<Q, EX>(d, a, tree0 + t, l, idx, HD + t * stride);
        }
        __syncthreads();
    }
    expand_tab<VEC2, GF_NT_CLS>(e, gsm, tree0, g);
}

// ---- packed / constant-bank variant of the fused cls kernel (translation-invariant tables, q == Q) ----------------
// Internal levels as in k_guides_dns_fused_c (uniform matrix index, tables in the constant bank).  The leaf level
// gathers rows log T_c[:, x] (:196) of the s leaf-level matrices from a shared-memory copy: the row index depends on
// the observed leaf, so it cannot come through the uniform path, and a global gather would queue behind the stores.
template <int Q, bool VEC2, int NW>
__global__ void __launch_bounds__(GF_NT_CLS)
k_guides_cls_fused_c(const __grid_constant__ GhmDev d, const __grid_constant__ LvlArgs a, const __grid_constant__ FusedExp e,
                     const __grid_constant__ GuideC gc, int G, const __grid_constant__ DnsTab<NW> tab) {
    extern __shared__ __align__(16) float gsm[];
    constexpr int H = Q / 2, QQ = Q * Q;
    const int L = d.L, s = d.s, nL = d.n_leaves;
    const int stride = e.stride, NT = blockDim.x, tid = threadIdx.x;
    float* HD = gsm;
    float* LT = gsm + (size_t)G * stride;                   // [s][Q][Q] log T^T of the edges into the leaves
    const int64_t tree0 = (int64_t)blockIdx.x * G;
    const int g = (int)min((int64_t)G, a.B - tree0);
    const float* Tup = tab.v;
    {
        const float* src = d.TlogT + (size_t)d.mat_off[L] * QQ;
        for (int i = tid; i < s * QQ; i += NT) LT[i] = __ldg(src + i);
    }
    __syncthreads();
    int tmb = gc.up_base - gc.s_u;                          // matrices of the edges into depth L-1
    for (int l = L - 1; l >= 0; --l) {
        const int n = d.spow[l], items = g * n;
        const unsigned nm = d.pow_magic[l];
        const int noff_c = 1 + d.edge_off[l + 1], noff_p = l == 0 ? 0 : 1 + d.edge_off[l];
        const int rt = rot_tid(l);
        for (int w0 = 0; w0 < items; w0 += NT) {
            const bool act = w0 + rt < items;
            if (!__any_sync(0xffffffffu, act)) continue;
            const int w = act ? w0 + rt : items - 1;
            const int t = div_magic(w, n, nm), idx = w - t * n;
            float* hd = HD + t * stride;
            f2 acc[H];
#pragma unroll
            for (int i = 0; i < H; ++i) acc[i] = make_float2(0.f, 0.f);
            if (l == L - 1) {
                for (int c = 0; c < s; ++c) {
                    const int x = leaf_at(a.leaves, a.leaf_dtype, (tree0 + t) * nL + idx * s + c, Q, d.status);
                    f2 r[H];
                    f2_ld<H>(LT + (c * Q + x) * Q, r);
#pragma unroll
                    for (int i = 0; i < H; ++i) { acc[i].x += r[i].x; acc[i].y += r[i].y; }
                }
            } else {
                int tm = tmb;
#pragma unroll 1
                for (int c = 0; c < s; ++c, ++tm) {
                    f2 h[H], m[H];
                    f2_ld<H>(hd + (noff_c + idx * s + c) * Q, h);
                    f2_log_matvec_c<Q>(Tup + tm * QQ, h, m);    // log(T @ exp(hd))  (:207)
#pragma unroll
                    for (int i = 0; i < H; ++i) { acc[i].x += m[i].x; acc[i].y += m[i].y; }
                }
            }
            const float mx = f2_vmax<H>(acc);                   // (:197,208)
#pragma unroll
            for (int i = 0; i < H; ++i) { acc[i].x -= mx; acc[i].y -= mx; }
            if (act) f2_st<H>(hd + (noff_p + idx) * Q, acc);
            if (l == 0 && act) {
                const int64_t b = tree0 + t;
                if (a.root_hd) f2_st<H>(a.root_hd + b * Q, acc);
                if (a.post) {                                   // posterior = softmax(hd + log p_y)  (:213-217)
                    f2 h0[H];
#pragma unroll
                    for (int i = 0; i < H; ++i)
                        h0[i] = make_float2(acc[i].x + logf(__ldg(d.py + 2 * i)), acc[i].y + logf(__ldg(d.py + 2 * i + 1)));
                    const float m0 = f2_vmax<H>(h0);
                    float sum = 0.f;
#pragma unroll
                    for (int i = 0; i < H; ++i) { h0[i].x = expf(h0[i].x - m0); h0[i].y = expf(h0[i].y - m0); sum += h0[i].x + h0[i].y; }
                    const float inv = 1.0f / sum;
#pragma unroll
                    for (int i = 0; i < H; ++i) { h0[i].x *= inv; h0[i].y *= inv; }
                    f2_st<H>(a.post + b * Q, h0);
                }
            }
        }
        if (l < L - 1) tmb -= gc.s_u;
        __syncthreads();
    }
    expand_tab<VEC2, GF_NT_CLS>(e, gsm, tree0, g);
}

// trees per CTA of the fused kernels for `arrays` compact arrays; 0 -> does not fit, use the level kernels
static int fused_trees_per_cta(int64_t n_nodes, int q, int arrays) {
    const size_t per_tree = (size_t)arrays * n_nodes * q * sizeof(float);
    const size_t budget = 73 * 1024;                        // 3 CTAs per SM (227 KB: 3 x (73 + 1 KB reserved))
    int G = (int)(budget / per_tree);
    return G > 16 ? 16 : G;
}

static int64_t n_nodes_all(const GhmDev& d) { return 1 + (int64_t)d.n_edges; }
static int64_t n_nodes_int(const GhmDev& d) { return 1 + (int64_t)d.edge_off[d.L]; }   // depths 0..L-1

// offsets of one tensor: unit u = (leaf i, part, k) -> float offset of SRC_part[node(level, i / s^(L-level))][k] in a CTA's
// shared memory, where array `a` of tree 0 starts at a * G * stride
static void tab_add(std::vector<uint16_t>& tab, GhmGuideTab& gt, const GhmDev& d, int level, int nsrc, int p0, int p1, int p2) {
    const int W = gt.W, hq = d.q / W, C2 = nsrc * hq, nL = d.n_leaves;
    const int R = d.spow[d.L - level], noff = level == 0 ? 0 : 1 + d.edge_off[level];
    const int part_arr[3] = {p0, p1, p2};
    gt.tab_off[gt.n_t] = (int)tab.size();
    gt.upt[gt.n_t] = nL * C2;
    ++gt.n_t;
    for (int i = 0; i < nL; ++i)
        for (int j = 0; j < C2; ++j) {
            const int part = j / hq, kk = j - part * hq;
            tab.push_back((uint16_t)((size_t)part_arr[part] * gt.G * gt.stride + (size_t)(noff + i / R) * d.q + kk * W));
        }
}

int ghm_guides_init(ghm_model* m) {
    const GhmDev& d = m->d;
    m->gt_cls = GhmGuideTab{}; m->gt_dns = GhmGuideTab{};
    if (d.q > GHM_MAX_Q_REG || d.s < 2) return GHM_OK;
    std::vector<uint16_t> tab;
    const int L = d.L, W = (d.q % 2 == 0) ? 2 : 1;
    {   // cls set: guides[j] <- hd of depth L-1-j     (reference :533-549)
        GhmGuideTab& gt = m->gt_cls;
        gt.W = W; gt.stride = (int)(n_nodes_int(d) * d.q);
        gt.G = fused_trees_per_cta(n_nodes_int(d), d.q, 1);
        if (gt.G >= 1)
            for (int j = 0; j < L; ++j) tab_add(tab, gt, d, L - 1 - j, 1, 0, 0, 0);
    }
    {   // dns set: (hd|qd) depth L..1 ; root (hd|bu) ; (hd|qd|bu) depth 1..L      (reference :554-590)
        GhmGuideTab& gt = m->gt_dns;
        gt.W = W; gt.stride = (int)(n_nodes_all(d) * d.q);
        gt.G = fused_trees_per_cta(n_nodes_all(d), d.q, 3);
        if (gt.G >= 1) {
            for (int j = 0; j < L; ++j) tab_add(tab, gt, d, L - j, 2, 0, 1, 0);
            tab_add(tab, gt, d, 0, 2, 0, 2, 0);
            for (int j = 1; j <= L; ++j) tab_add(tab, gt, d, j, 3, 0, 1, 2);
        }
    }
    if (tab.empty()) return GHM_OK;
    int prev = 0;
    cudaGetDevice(&prev);
    if (prev != m->device) cudaSetDevice(m->device);
    cudaError_t ce = cudaMalloc(&m->guide_tab, tab.size() * sizeof(uint16_t));
    if (ce == cudaSuccess) ce = cudaMemcpy(m->guide_tab, tab.data(), tab.size() * sizeof(uint16_t), cudaMemcpyHostToDevice);
    if (prev != m->device) cudaSetDevice(prev);
    if (ce != cudaSuccess) return ghm_fail(GHM_ECUDA, "guide offset tables: %s", cudaGetErrorString(ce));
    return GHM_OK;
}

// fused launch descriptor for one call; false when an output pointer rules the fused kernel out
static bool fused_setup(FusedExp& e, const ghm_model* m, const GhmGuideTab& gt, float* const* guides) {
    if (gt.G < 1 || !m->guide_tab) return false;
    e.tab = m->guide_tab; e.n_t = gt.n_t; e.stride = gt.stride;
    for (int i = 0; i < gt.n_t; ++i) {
        e.tab_off[i] = gt.tab_off[i]; e.upt[i] = gt.upt[i]; e.out[i] = guides[i];
        if (!guides[i] || (gt.W == 2 && ((uintptr_t)guides[i] % 8) != 0)) return false;
    }
    return true;
}

template <typename F>
static int dispatch_q(int q, F&& f) {
    switch (ghm_pad_q(q)) {
        case 4: return f(std::integral_constant<int, 4>{});
        case 8: return f(std::integral_constant<int, 8>{});
        case 10: return f(std::integral_constant<int, 10>{});
        case 16: return f(std::integral_constant<int, 16>{});
        default:
            return ghm_fail(GHM_EUNSUP, "variable_type=%d: register-resident kernels cover q <= %d in this build", q,
                            GHM_MAX_Q_REG);
    }
}

struct DevGuard {
    int prev;
    explicit DevGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DevGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// ---- cls guides ------------------------------------------------------------------------------------
// workspace: HD [B][n_int][q] f32 -- carried in the LAST guide tensor's tail?  No: guides[L-1] is the
// root guide [B, n_L, q] and n_L*q >= n_int*q always holds (n_int = (n_L-1)/(s-1)+... <= n_L for s >= 2),
// so the compact messages are staged inside guides[L-1] and expanded in place last.  For s == 1 the
// model constructor's limits make n_int = L <= n_L*... false; handled by requiring s >= 2 here.
extern "C" int ghm_guides_cls(const ghm_model_t* m, int64_t B, const void* leaves, int leaf_dtype,
                              float* const* guides, float* post, float* root_hd, void* stream) {
    if (!m || !leaves || !guides) return ghm_fail(GHM_EINVAL, "ghm_guides_cls: null argument");
    if (B <= 0) return GHM_OK;
    const GhmDev& d = m->d;
    if (d.s < 2) return ghm_fail(GHM_EUNSUP, "ghm_guides_cls needs n_child >= 2");
    DevGuard g(m->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int L = d.L;
    const int64_t n_int = n_nodes_int(d);
    // compact store lives at the END of the root guide tensor so the in-place expansion of level 0
    // (which only reads node 0 of each tree) cannot overwrite data it still needs -- see below.
    // Simpler and always safe: use guides[0] (depth L-1 tensor, written FIRST by expansion) is not
    // possible either; so allocate nothing and stage in the root guide, expanding levels L-1..1 first
    // and the root level through a tiny copy of the root messages kept in root_hd/post scratch.
    if (!root_hd) return ghm_fail(GHM_EINVAL, "ghm_guides_cls: root_hd output is required");
    LvlArgs a{};
    a.B = B; a.leaves = leaves; a.leaf_dtype = leaf_dtype; a.n_nodes = (int)n_int;
    a.post = post; a.root_hd = root_hd;
    FusedExp fe;
    if (fused_setup(fe, m, m->gt_cls, guides)) {               // tree-tiled fused kernel: messages never leave shared memory
        const int G = m->gt_cls.G;
        const size_t dyn = (size_t)G * n_int * d.q * sizeof(float);
        const unsigned grid = (unsigned)((B + G - 1) / G);
        const bool v2 = m->gt_cls.W == 2;
        return dispatch_q(d.q, [&](auto Qc) -> int {
            constexpr int Q = decltype(Qc)::value;
            auto go = [&](auto kern) -> int {
                GHM_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
                kern<<<grid, GF_NT_CLS, dyn, st>>>(d, a, fe, G);
                GHM_CHECK_LAUNCH();
                return GHM_OK;
            };
            const size_t words = (size_t)d.n_mat * Q * Q;
            const size_t dyn_c = dyn + (size_t)d.s * Q * Q * sizeof(float);
            if (d.ti && d.q == Q && 2 * words <= (size_t)GHM_TAB_WORDS && ((uintptr_t)root_hd % 8) == 0 &&
                (!post || ((uintptr_t)post % 8) == 0)) {
                GuideC gc{(d.L - 1) * d.s, (int)words, d.s};
                DnsTab<GHM_TAB_WORDS> tab;
                tab.v[0] = 0.f;
                memcpy(tab.v, m->h_TlinT, words * sizeof(float));
                auto goc = [&](auto kern) -> int {
                    GHM_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_c));
                    kern<<<grid, GF_NT_CLS, dyn_c, st>>>(d, a, fe, gc, G, tab);
                    GHM_CHECK_LAUNCH();
                    return GHM_OK;
                };
                return v2 ? goc(k_guides_cls_fused_c<Q, true, GHM_TAB_WORDS>) : goc(k_guides_cls_fused_c<Q, false, GHM_TAB_WORDS>);
            }
            if (d.q == Q) return v2 ? go(k_guides_cls_fused<Q, true, true>) : go(k_guides_cls_fused<Q, true, false>);
            return v2 ? go(k_guides_cls_fused<Q, false, true>) : go(k_guides_cls_fused<Q, false, false>);
        });
    }
    // large trees: level kernels on a compact store staged inside the root guide tensor, expanded last
    float* HD = guides[L - 1];
    if (n_int > (int64_t)d.n_leaves) return ghm_fail(GHM_EUNSUP, "unexpected tree shape");
    a.HD = HD;
    int rc;
    if (d.QW) {                                                // 16 < q <= 256: one warp per (tree, node) row (ghm_wide_lvl.cuh)
        WLvlArgs wa{};
        wa.B = B; wa.leaves = leaves; wa.leaf_dtype = leaf_dtype; wa.HD = HD; wa.n_nodes = (int)n_int;
        wa.post = post; wa.root_hd = root_hd;
        rc = GHM_OK;
        for (int l = d.L - 1; l >= 0; --l) {
            k_wl_cls<<<wl_grid(B * d.spow[l]), WL_NT, (size_t)(WL_NT / 32) * d.QW * sizeof(float), st>>>(d, wa, l);
            GHM_CHECK_LAUNCH();
        }
    } else {
        rc = dispatch_q(d.q, [&](auto Qc) -> int {
            constexpr int Q = decltype(Qc)::value;
            for (int l = d.L - 1; l >= 0; --l) {
                k_lvl_cls<Q><<<lvl_grid(d, B, l), GD_NT, 0, st>>>(d, a, l);
                GHM_CHECK_LAUNCH();
            }
            return GHM_OK;
        });
    }
    if (rc) return rc;
    // guides[j] <- depth L-1-j, j = 0..L-2, read from HD (inside guides[L-1]); then the root level from root_hd
    {
        ExpandAll e;
        expand_setup(e, d, B, (int)n_int, HD, nullptr, nullptr);
        for (int j = 0; j < L - 1; ++j) expand_add(e, guides[j], L - 1 - j, 1, 0, 0, 0);
        rc = expand_launch(e, st);
        if (rc) return rc;
    }
    {   // root level: source = root_hd viewed as [B][1 node][q]
        ExpandAll e;
        expand_setup(e, d, B, 1, root_hd, nullptr, nullptr);
        expand_add(e, guides[L - 1], 0, 1, 0, 0, 0);
        rc = expand_launch(e, st);
        if (rc) return rc;
    }
    return GHM_OK;
}

// ---- dns guides ------------------------------------------------------------------------------------
extern "C" int64_t ghm_guides_dns_workspace_bytes(const ghm_model_t* m, int64_t B) {
    if (!m || B <= 0) return 0;
    // tree-tiled fused kernel (messages in shared memory): no workspace; it needs non-null, 8-byte aligned guide
    // pointers -- ghm_guides_dns fails with GHM_EINVAL if they are not and no workspace was supplied
    if (m->gt_dns.G >= 1 && m->guide_tab) return 0;
    return 3 * B * n_nodes_all(m->d) * m->d.q * (int64_t)sizeof(float);
}

extern "C" int ghm_guides_dns(const ghm_model_t* m, int64_t B, const float* z, float sigma, const float* ext,
                              float* const* guides, float* mean, void* workspace, void* stream) {
    if (!m || !z) return ghm_fail(GHM_EINVAL, "ghm_guides_dns: null argument");
    if (B <= 0) return GHM_OK;
    if (!(sigma > 0.f)) return ghm_fail(GHM_EINVAL, "ghm_guides_dns: sigma must be positive");
    const GhmDev& d = m->d;
    DevGuard g(m->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int L = d.L;
    const int64_t nn = n_nodes_all(d);
    float* HD = (float*)workspace;
    float* QD = HD + B * nn * d.q;
    float* BU = QD + B * nn * d.q;
    LvlArgs a{};
    a.B = B; a.z = z; a.sigma = sigma; a.ext = ext; a.HD = HD; a.QD = QD; a.BU = BU; a.n_nodes = (int)nn; a.mean = mean;
    FusedExp fe;
    if (guides && fused_setup(fe, m, m->gt_dns, guides)) {     // tree-tiled fused kernel: messages never leave shared memory
        const int G = m->gt_dns.G;
        const size_t dyn = (size_t)3 * G * nn * d.q * sizeof(float);
        const unsigned grid = (unsigned)((B + G - 1) / G);
        const bool v2 = m->gt_dns.W == 2;
        return dispatch_q(d.q, [&](auto Qc) -> int {
            constexpr int Q = decltype(Qc)::value;
            const size_t words = (size_t)d.n_mat * Q * Q;
            if (d.ti && d.q == Q && 2 * words <= (size_t)GHM_TAB_WORDS) {
                GuideC gc{(d.L - 1) * d.s, (int)words, d.s};
                DnsTab<GHM_TAB_WORDS> tab;
                tab.v[0] = 0.f;
                memcpy(tab.v, m->h_TlinT, words * sizeof(float));
                memcpy(tab.v + words, m->h_Tlin, words * sizeof(float));
                auto goc = [&](auto kern) -> int {
                    GHM_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
                    kern<<<grid, GF_NT_DNS, dyn, st>>>(d, a, fe, gc, G, tab);
                    GHM_CHECK_LAUNCH();
                    return GHM_OK;
                };
                return v2 ? goc(k_guides_dns_fused_c<Q, true, GHM_TAB_WORDS>) : goc(k_guides_dns_fused_c<Q, false, GHM_TAB_WORDS>);
            }
            auto go = [&](auto kern) -> int {
                GHM_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
                kern<<<grid, GF_NT_DNS, dyn, st>>>(d, a, fe, G);
                GHM_CHECK_LAUNCH();
                return GHM_OK;
            };
            if (d.q == Q) return v2 ? go(k_guides_dns_fused<Q, true, true>) : go(k_guides_dns_fused<Q, true, false>);
            return v2 ? go(k_guides_dns_fused<Q, false, true>) : go(k_guides_dns_fused<Q, false, false>);
        });
    }
    if (!workspace)
        return ghm_fail(GHM_EINVAL, "ghm_guides_dns: the level-synchronous path needs a workspace of 3*B*nodes*q floats "
                                    "(guide pointers null / not 8-byte aligned, or the tree does not fit shared memory)");
    int rc;
    if (d.QW) {                                                // 16 < q <= 256: one warp per (tree, node) row (ghm_wide_lvl.cuh)
        WLvlArgs wa{};
        wa.B = B; wa.z = z; wa.sigma = sigma; wa.ext = ext; wa.HD = HD; wa.QD = QD; wa.BU = BU; wa.n_nodes = (int)nn;
        wa.mean = mean;
        const size_t dyn = (size_t)(WL_NT / 32) * d.QW * sizeof(float);
        rc = GHM_OK;
        for (int l = d.L; l >= 0; --l) {
            k_wl_dns_up<<<wl_grid(B * d.spow[l]), WL_NT, dyn, st>>>(d, wa, l);
            GHM_CHECK_LAUNCH();
        }
        for (int l = 1; l <= d.L; ++l) {
            k_wl_dns_down<<<wl_grid(B * d.spow[l]), WL_NT, dyn, st>>>(d, wa, l);
            GHM_CHECK_LAUNCH();
        }
    } else {
        rc = dispatch_q(d.q, [&](auto Qc) -> int {
            constexpr int Q = decltype(Qc)::value;
            for (int l = d.L; l >= 0; --l) {
                k_lvl_dns_up<Q><<<lvl_grid(d, B, l), GD_NT, 0, st>>>(d, a, l);
                GHM_CHECK_LAUNCH();
            }
            for (int l = 1; l <= d.L; ++l) {
                k_lvl_dns_down<Q><<<lvl_grid(d, B, l), GD_NT, 0, st>>>(d, a, l);
                GHM_CHECK_LAUNCH();
            }
            return GHM_OK;
        });
    }
    if (rc || !guides) return rc;
    // (hd|qd) depth L..1 ; root (hd|bu) ; (hd|qd|bu) depth 1..L      (reference :554-590) -- one launch
    ExpandAll e;
    expand_setup(e, d, B, (int)nn, HD, QD, BU);
    for (int j = 0; j < L; ++j) expand_add(e, guides[j], L - j, 2, 0, 1, 0);
    expand_add(e, guides[L], 0, 2, 0, 2, 0);
    for (int j = 1; j <= L; ++j) expand_add(e, guides[L + j], j, 3, 0, 1, 2);
    return expand_launch(e, st);
}
