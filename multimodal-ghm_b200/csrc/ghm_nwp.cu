// ghm_nwp.cu -- K4: next-token posteriors p(leaf_{t+1} | leaves_{<=t}, ext) for every prefix.
//
// Replaces GHMTree.BP_NWP_autoregressive (reference src/ghmclip/data/data_random_GHM.py:336-463),
// including its per-position guide tensors (:357-364,382,405-406,438-439,459).
//
// The reference walks the positions in order and keeps the messages of finished subtrees on its Node
// objects; that loop is serial in the position.  Here the work is one THREAD per (tree, position):
//   1. k_nwp_full: one full upward pass (level-synchronous, one thread per (tree, node)) stores the
//      message qd_full(v) of every internal node v given ALL leaves under it.  A subtree the reference
//      has finished by position t holds exactly this message (`parent.children[c].qd_message` for
//      c < current, :394-396), with the NWP shift convention (every hd and qd max-shifted, :397-399).
//   2. k_nwp_pos: position t climbs the root path of leaf t -- at each ancestor the product of the stored
//      messages of the children LEFT of the path times the partial message coming up the path (:389-415),
//      root belief with the external message (:420-435) -- then walks down the path of leaf t+1: a
//      cavity update while the ancestors are shared, a plain push-down after they split (:443-454).
// Consecutive threads are consecutive positions of one tree, so posteriors and guide tensors
// ([B, n_L-1, C]) are written as contiguous rows (coalesced), finished-subtree messages are warp-wide
// broadcast loads, and the launch has B*(n_L-1) threads instead of B.  LINEAR domain with a max-rescale
// wherever the reference shifts; guide tensors are the natural logs of those rescaled vectors, i.e. the
// reference's shifted log-messages (including the aliased root halves, :425-439).
#include <string.h>

#include <algorithm>

#include "ghm_vec2.cuh"
#include "ghm_wide_lvl.cuh"

#define NWP_NT 128
#define NWP_MAX_GUIDES (2 * GHM_MAX_LEVELS + 1)

struct NwpArgs {
    int64_t B;
    const void* leaves; int leaf_dtype;
    const float* ext;
    float* pp;                         // [B][nL-1][q]
    float* full;                       // [B][n_int][q] finished-subtree messages (workspace)
    int n_int;                         // internal nodes per tree (depths 0 .. L-1; slot 0 unused)
    float* guides[NWP_MAX_GUIDES];     // only when GUIDE
};

// rows are q floats, 8-byte aligned for even q: float2 stores (half the transactions of scalar stores at a 4q-byte stride)
template <int Q>
__device__ __forceinline__ void store_row(float* dst, const float (&v)[Q], int q) {
    if ((q & 1) == 0 && (reinterpret_cast<uintptr_t>(dst) & 7) == 0) {
        float2* d2 = reinterpret_cast<float2*>(dst);
#pragma unroll
        for (int k = 0; k < Q; k += 2)
            if (k < q) d2[k >> 1] = make_float2(v[k], v[k + 1]);
    } else {
#pragma unroll
        for (int k = 0; k < Q; ++k)
            if (k < q) dst[k] = v[k];
    }
}
template <int Q>
__device__ __forceinline__ void store_log(float* dst, const float (&v)[Q], int q) {
    float lg[Q];
#pragma unroll
    for (int k = 0; k < Q; ++k) lg[k] = (k < q) ? __logf(v[k]) : 0.f;      // v in (0, 1]: MUFU.LG2 error << 1e-5 budget
    store_row<Q>(dst, lg, q);
}

__device__ __forceinline__ int nwp_leaf(const NwpArgs& a, const GhmDev& d, int64_t off) {
    int64_t xv = a.leaf_dtype == GHM_LEAF_I64 ? reinterpret_cast<const int64_t*>(a.leaves)[off]
                                              : (int64_t) reinterpret_cast<const uint8_t*>(a.leaves)[off];
    if (xv < 0 || xv >= d.q) { atomicOr(d.status, 1); xv = xv < 0 ? 0 : d.q - 1; }
    return (int)xv;
}
__device__ __forceinline__ int nwp_node(const GhmDev& d, int l) { return l == 0 ? 0 : 1 + d.edge_off[l]; }

// finished message of leaf `leaf`: T[:, x] rescaled (:374-375)
template <int Q>
__device__ __forceinline__ void nwp_leaf_msg(const NwpArgs& a, const GhmDev& d, int64_t b, int leaf, float (&m)[Q]) {
    const int x = nwp_leaf(a, d, b * d.n_leaves + leaf);
    const int c = leaf - ghm_div_s(leaf, d) * d.s;
    const int mi = d.mat_off[d.L] + (d.ti ? c : leaf);
    ghm_load_row<Q, float>(d.TlinT + ((size_t)mi * Q + x) * Q, m);
    ghm_normalize<Q>(m);
}

// ---- 1. full upward pass: qd_full of the depth-l nodes (1 <= l <= L-1) --------------------------------
template <int Q>
__global__ void __launch_bounds__(NWP_NT) k_nwp_full(const GhmDev d, const NwpArgs a, int l) {
    const int64_t t = (int64_t)blockIdx.x * NWP_NT + threadIdx.x;
    const int n = d.spow[l];
    const int64_t b = t / n;
    if (b >= a.B) return;
    const int idx = (int)(t - b * n);
    const int L = d.L, s = d.s, q = d.q;
    float* F = a.full + b * (int64_t)a.n_int * q;
    float h[Q];
#pragma unroll
    for (int k = 0; k < Q; ++k) h[k] = (k < q) ? 1.f : 0.f;
    for (int c = 0; c < s; ++c) {
        float m[Q];
        if (l == L - 1) nwp_leaf_msg<Q>(a, d, b, idx * s + c, m);
        else {
            const float* src = F + (int64_t)(nwp_node(d, l + 1) + idx * s + c) * q;
#pragma unroll
            for (int k = 0; k < Q; ++k) m[k] = (k < q) ? src[k] : 0.f;
        }
#pragma unroll
        for (int k = 0; k < Q; ++k) h[k] *= m[k];
    }
    ghm_normalize<Q>(h);
    const int mi = d.mat_off[l] + (d.ti ? idx - ghm_div_s(idx, d) * s : idx);
    float u[Q];
    ghm_matvec<Q>(d.Tlin + (size_t)mi * Q * Q, h, u);
    ghm_normalize<Q>(u);
    float* dst = F + (int64_t)(nwp_node(d, l) + idx) * q;
#pragma unroll
    for (int k = 0; k < Q; ++k)
        if (k < q) dst[k] = u[k];
}

// ---- 2. one thread per (tree, position) ---------------------------------------------------------------
template <int Q, bool GUIDE, bool SMEM_TAB>
__global__ void __launch_bounds__(NWP_NT) k_nwp_pos(const GhmDev d, const NwpArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int NT = NWP_NT;
    const int tid = threadIdx.x;
    const int L = d.L, s = d.s, q = d.q, nL = d.n_leaves;
    const int npos = nL - 1;

    size_t off = 0;
    const float* Tlin = d.Tlin;
    if (SMEM_TAB) {
        const int tab_words = d.n_mat * Q * Q;
        float* s1 = reinterpret_cast<float*>(smem);
        off += (size_t)tab_words * 4;
        for (int i = tid; i < tab_words; i += NT) s1[i] = d.Tlin[i];
        Tlin = s1;
    }
    const int nlv = L > 1 ? L - 1 : 1;                       // path nodes at depth 1 .. L-1: one slot each (a fourth, unused slot
    float* HD = reinterpret_cast<float*>(smem + off); off += (size_t)nlv * Q * NT * 4;   // cost 10 KB: 4 instead of 6 CTAs/SM)
    float* QD = reinterpret_cast<float*>(smem + off);                                    // [L][Q][NT]
    if (SMEM_TAB) __syncthreads();

    const int64_t row = (int64_t)blockIdx.x * NT + tid;                  // = b * npos + t
    if (row >= a.B * (int64_t)npos) return;
    const int64_t b = row / npos;
    const int t = (int)(row - b * npos);
    const float* F = a.full + b * (int64_t)a.n_int * q;

    // ---- observed leaf t: qd = log T[:, x_t], shifted (:374-375) -------------------------------------------
    float m[Q];
    nwp_leaf_msg<Q>(a, d, b, t, m);
    if (GUIDE) store_log<Q>(a.guides[0] + row * q, m, q);
    // ---- up the path: depth L-1 .. 1 (:389-415) -----------------------------------------------------------
    int idx = t;                                       // index of the node whose message `m` is, at depth l+1
    for (int l = L - 1; l >= 1; --l) {
        const int pidx = ghm_div_s(idx, d);            // ancestor at depth l
        const int c = idx - pidx * s;                  // which child of it the path goes through
        float h[Q];
#pragma unroll
        for (int k = 0; k < Q; ++k) h[k] = m[k];
        for (int cc = 0; cc < c; ++cc) {               // finished children left of the path
            float f[Q];
            if (l == L - 1) nwp_leaf_msg<Q>(a, d, b, pidx * s + cc, f);
            else {
                const float* src = F + (int64_t)(nwp_node(d, l + 1) + pidx * s + cc) * q;
#pragma unroll
                for (int k = 0; k < Q; ++k) f[k] = (k < q) ? src[k] : 0.f;
            }
#pragma unroll
            for (int k = 0; k < Q; ++k) h[k] *= f[k];
        }
        ghm_normalize<Q>(h);
        const int pc = pidx - ghm_div_s(pidx, d) * s;
        const int mi = d.mat_off[l] + (d.ti ? pc : pidx);
        ghm_matvec<Q>(Tlin + (size_t)mi * Q * Q, h, m);
        ghm_normalize<Q>(m);
        float* Hl = HD + (size_t)(l - 1) * Q * NT + tid;
        float* Ql = QD + (size_t)(l - 1) * Q * NT + tid;
#pragma unroll
        for (int k = 0; k < Q; ++k) { Hl[k * NT] = h[k]; Ql[k * NT] = m[k]; }
        if (GUIDE) {
            float* g = a.guides[L - l] + row * 2 * q;
            store_log<Q>(g, h, q);
            store_log<Q>(g + q, m, q);
        }
        idx = pidx;
    }
    // ---- root (:420-439): idx is now the depth-1 node on the path -----------------------------------------
    float bel[Q];
#pragma unroll
    for (int k = 0; k < Q; ++k) bel[k] = m[k];
    for (int cc = 0; cc < idx; ++cc) {
        float f[Q];
        if (L == 1) nwp_leaf_msg<Q>(a, d, b, cc, f);
        else {
            const float* src = F + (int64_t)(nwp_node(d, 1) + cc) * q;
#pragma unroll
            for (int k = 0; k < Q; ++k) f[k] = (k < q) ? src[k] : 0.f;
        }
#pragma unroll
        for (int k = 0; k < Q; ++k) bel[k] *= f[k];
    }
    ghm_normalize<Q>(bel);
    if (a.ext) {                                       // external root message in the linear domain, rescaled (:425-435)
        float x[Q];
#pragma unroll
        for (int k = 0; k < Q; ++k) x[k] = (k < q) ? a.ext[b * q + k] : -INFINITY;
        const float mx = ghm_vmax<Q>(x);
#pragma unroll
        for (int k = 0; k < Q; ++k) bel[k] *= (k < q) ? __expf(x[k] - mx) : 0.f;
        ghm_normalize<Q>(bel);
    }
    if (GUIDE) {
        float* g = a.guides[L] + row * 2 * q;
        store_log<Q>(g, bel, q);
        store_log<Q>(g + q, bel, q);
    }
    // ---- down the path of leaf t+1 (:443-459) ---------------------------------------------------------------
    for (int l = 1; l <= L; ++l) {
        const int g = ghm_div_pow(t + 1, L - l, d);            // goal-path node at depth l
        const int a_l = ghm_div_pow(t, L - l, d);              // observed-path node at depth l
        const int pg = ghm_div_s(g, d);
        const int cg = g - pg * s;
        const int mi = d.mat_off[l] + (d.ti ? cg : g);
        float w[Q], tt[Q];
        if (g == a_l) {                                         // shared ancestor: cavity update (l <= L-1 here)
            const float* Hl = HD + (size_t)(l - 1) * Q * NT + tid;
            const float* Ql = QD + (size_t)(l - 1) * Q * NT + tid;
#pragma unroll
            for (int k = 0; k < Q; ++k) { const float qv = Ql[k * NT]; w[k] = qv > 0.f ? __fdividef(bel[k], qv) : 0.f; }
            ghm_matvec_t<Q>(Tlin + (size_t)mi * Q * Q, w, tt);
#pragma unroll
            for (int k = 0; k < Q; ++k) bel[k] = Hl[k * NT] * tt[k];
        } else {
            ghm_matvec_t<Q>(Tlin + (size_t)mi * Q * Q, bel, tt);
#pragma unroll
            for (int k = 0; k < Q; ++k) bel[k] = tt[k];
        }
        ghm_normalize<Q>(bel);
        if (GUIDE) store_log<Q>(a.guides[L + l] + row * q, bel, q);
    }
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < Q; ++k) sum += bel[k];
    const float inv = 1.0f / sum;
#pragma unroll
    for (int k = 0; k < Q; ++k) bel[k] *= inv;
    store_row<Q>(a.pp + row * q, bel, q);
}

// ------------------------------------------------------------------------------------------------
// Uniform variant (translation-invariant tables, q == Q even): one thread per TREE, the position is a property of
// the CTA (blockIdx.y strides over the positions), so every path index -- ancestors, child numbers, matrices,
// the depth at which the goal path splits off -- is the same for all threads: the 2L-1 matvecs of a position take
// their table from the constant bank through uniform registers (no per-lane table rows through the LSU, which is
// what bounds k_nwp_pos), and the finished-subtree messages live transposed, FT[node][state pair][tree], so that
// consecutive trees read them coalesced.
// ------------------------------------------------------------------------------------------------
struct NwpU { int up0, dn_off, s_u, pg; };                   // (L-2)*s, float offset of Tlin in the table parameter, s, gridDim.y

template <int H>
__device__ __forceinline__ void f2_mul_into(f2 (&h)[H], const f2 (&f)[H]) {
#pragma unroll
    for (int i = 0; i < H; ++i) h[i] = f2_mul(h[i], f[i]);
}
// rescaled leaf message T_c[:, x] from the shared-memory copy of the leaf-level matrices (:374-375)
template <int Q>
__device__ __forceinline__ void nwp_leaf_msg_s(const float* LT, int c, int x, f2 (&m)[Q / 2]) {
    const f2* r = reinterpret_cast<const f2*>(LT + (c * Q + x) * Q);
#pragma unroll
    for (int i = 0; i < Q / 2; ++i) m[i] = r[i];
    f2_normalize<Q>(m);
}

template <int Q, int NW>
__global__ void __launch_bounds__(NWP_NT)
k_nwp_full_u(const __grid_constant__ GhmDev d, const __grid_constant__ NwpArgs a, const __grid_constant__ NwpU u,
             const __grid_constant__ DnsTab<NW> tab) {
    extern __shared__ __align__(16) float nsm[];
    constexpr int H = Q / 2, QQ = Q * Q;
    const int L = d.L, s = d.s, nL = d.n_leaves, tid = threadIdx.x;
    float* LT = nsm;
    {
        const float* src = d.TlinT + (size_t)d.mat_off[L] * QQ;
        for (int i = tid; i < s * QQ; i += NWP_NT) LT[i] = __ldg(src + i);
    }
    // The CTA's leaves as bytes in shared memory, [tree][n_L], read from the global [B, n_L] tensor as ONE contiguous span
    // with consecutive threads on consecutive elements.  Each thread reading its own row (n_L int64 at an 8 n_L-byte stride
    // between lanes) left this kernel waiting on those loads for 69 % of its samples (profiles/r02_ncu_full_k_nwp_full_u.csv).
    uint8_t* LV = reinterpret_cast<uint8_t*>(nsm) + (((size_t)s * QQ * sizeof(float) + 15) / 16 * 16);
    {
        const int64_t t0 = (int64_t)blockIdx.x * NWP_NT;
        const int64_t cnt = min((int64_t)NWP_NT, a.B - t0) * nL;
        bool bad = false;
        if (a.leaf_dtype == GHM_LEAF_I64) {
            const int64_t* src = reinterpret_cast<const int64_t*>(a.leaves) + t0 * nL;
            for (int64_t i = tid; i < cnt; i += NWP_NT) {
                int64_t v = src[i];
                if (v < 0 || v >= d.q) { bad = true; v = v < 0 ? 0 : d.q - 1; }
                LV[i] = (uint8_t)v;
            }
        } else {
            const uint8_t* src = reinterpret_cast<const uint8_t*>(a.leaves) + t0 * nL;
            for (int64_t i = tid; i < cnt; i += NWP_NT) {
                int v = src[i];
                if (v >= d.q) { bad = true; v = d.q - 1; }
                LV[i] = (uint8_t)v;
            }
        }
        if (bad) atomicOr(d.status, 1);
    }
    __syncthreads();
    const int64_t b0 = (int64_t)blockIdx.x * NWP_NT + tid;
    const bool act = b0 < a.B;
    const int64_t b = act ? b0 : a.B - 1;
    const uint8_t* lrow = LV + (size_t)(act ? tid : (int)(a.B - 1 - (int64_t)blockIdx.x * NWP_NT)) * nL;
    f2* FT = reinterpret_cast<f2*>(a.full);
    const int64_t B = a.B;
    int lbase = u.up0;                                        // matrices of the edges into depth l
    for (int l = L - 1; l >= 1; --l, lbase -= u.s_u) {
        const int n = d.spow[l];
        const int noff = nwp_node(d, l), noff_c = nwp_node(d, l + 1);
        int toff = lbase;
        for (int idx = 0; idx < n; ++idx) {
            f2 h[H];
#pragma unroll
            for (int i = 0; i < H; ++i) h[i] = make_float2(1.f, 1.f);
            for (int cc = 0; cc < s; ++cc) {
                f2 f[H];
                if (l == L - 1) nwp_leaf_msg_s<Q>(LT, cc, (int)lrow[idx * s + cc], f);
                else {
                    const f2* src = FT + (int64_t)(noff_c + idx * s + cc) * H * B + b;
#pragma unroll
                    for (int i = 0; i < H; ++i) f[i] = src[(int64_t)i * B];
                }
                f2_mul_into<H>(h, f);
            }
            f2_normalize<Q>(h);
            f2 v[H];
            f2_matvec_c<Q>(tab.v + toff * QQ, h, v);
            f2_normalize<Q>(v);
            if (act) {
                f2* dst = FT + (int64_t)(noff + idx) * H * B + b;
#pragma unroll
                for (int i = 0; i < H; ++i) dst[(int64_t)i * B] = v[i];
            }
            if (++toff == lbase + u.s_u) toff = lbase;
        }
    }
}

#define NWP_PC 2                                             // consecutive positions per chunk (their posterior rows are flushed as
                                                             // one 8q-byte span per tree; 4 measured the same at B = 65536 and
                                                             // slower at B = 10000: fewer chunks, fewer resident CTAs)
template <int Q, int NW>
__global__ void __launch_bounds__(NWP_NT)
k_nwp_pos_u(const __grid_constant__ GhmDev d, const __grid_constant__ NwpArgs a, const __grid_constant__ NwpU u,
            const __grid_constant__ DnsTab<NW> tab) {
    extern __shared__ __align__(16) float nsm[];
    constexpr int H = Q / 2, QQ = Q * Q, NT = NWP_NT, PC = NWP_PC;
    const int L = d.L, s = d.s, nL = d.n_leaves, tid = threadIdx.x;
    const int npos = nL - 1;
    float* LT = nsm;                                         // [s][Q][Q]
    f2* HS = reinterpret_cast<f2*>(nsm + (size_t)((s * QQ + 3) / 4) * 4);   // [L-1][H][NT]  path hd of depth 1..L-1
    f2* QS = HS + (size_t)(L - 1) * H * NT;                  // [L-1][H][NT]  path qd
    f2* ST = QS + (size_t)(L - 1) * H * NT;                        // [NT][PC][H] posterior rows of the chunk, flushed coalesced
    uint8_t* LV = reinterpret_cast<uint8_t*>(ST + (size_t)NT * PC * H);     // [PC + s][NT] observed leaves of the chunk
    {
        const float* src = d.TlinT + (size_t)d.mat_off[L] * QQ;
        for (int i = tid; i < s * QQ; i += NT) LT[i] = __ldg(src + i);
    }
    __syncthreads();
    const int64_t cta0 = (int64_t)blockIdx.x * NT;
    const int64_t b0 = cta0 + tid;
    const bool act = b0 < a.B;
    const int64_t b = act ? b0 : a.B - 1;
    const int64_t B = a.B;
    const f2* FT = reinterpret_cast<const f2*>(a.full);
    const float* Tup = tab.v;
    const float* Tdn = tab.v + u.dn_off;
    f2 ex[H];                                                // external root message, linear domain (:425-435)
    if (a.ext) {
        const f2* xe = reinterpret_cast<const f2*>(a.ext + b * Q);
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < H; ++i) { ex[i] = xe[i]; mx = fmaxf(mx, fmaxf(ex[i].x, ex[i].y)); }
#pragma unroll
        for (int i = 0; i < H; ++i) ex[i] = make_float2(__expf(ex[i].x - mx), __expf(ex[i].y - mx));
    }
    for (int t0 = blockIdx.y * PC; t0 < npos; t0 += u.pg * PC) {     // uniform: the positions are the same for the whole CTA
        const int nv = min(PC, npos - t0);
        // the leaves this chunk conditions on beyond the finished subtrees: from the first sibling of leaf t0 to leaf
        // t0+nv-1, fetched with independent loads (one memory latency per chunk instead of one per use)
        const int lbase = ghm_div_s(t0, d) * s;
        const int nlv = t0 + nv - lbase;
        for (int j = 0; j < nlv; ++j) LV[j * NT + tid] = (uint8_t)nwp_leaf(a, d, b * nL + lbase + j);
        for (int p = 0; p < nv; ++p) {
            const int t = t0 + p;
            f2 m[H];
            {
                const int pt = ghm_div_s(t, d);
                nwp_leaf_msg_s<Q>(LT, t - pt * s, LV[(t - lbase) * NT + tid], m);
            }
            // ---- up the path: depth L-1 .. 1 (:389-415) ----
            int idx = t;
            for (int l = L - 1; l >= 1; --l) {
                const int pidx = ghm_div_s(idx, d);
                const int c = idx - pidx * s;
                f2 h[H];
#pragma unroll
                for (int i = 0; i < H; ++i) h[i] = m[i];
                for (int cc = 0; cc < c; ++cc) {             // finished children left of the path
                    f2 f[H];
                    if (l == L - 1) nwp_leaf_msg_s<Q>(LT, cc, LV[(pidx * s + cc - lbase) * NT + tid], f);
                    else {
                        const f2* src = FT + (int64_t)(nwp_node(d, l + 1) + pidx * s + cc) * H * B + b;
#pragma unroll
                        for (int i = 0; i < H; ++i) f[i] = src[(int64_t)i * B];
                    }
                    f2_mul_into<H>(h, f);
                }
                f2_normalize<Q>(h);
                const int pc = pidx - ghm_div_s(pidx, d) * s;
                f2_matvec_c<Q>(Tup + ((l - 1) * s + pc) * QQ, h, m);
                f2_normalize<Q>(m);
                f2* Hl = HS + (size_t)(l - 1) * H * NT + tid;
                f2* Ql = QS + (size_t)(l - 1) * H * NT + tid;
#pragma unroll
                for (int i = 0; i < H; ++i) { Hl[i * NT] = h[i]; Ql[i * NT] = m[i]; }
                idx = pidx;
            }
            // ---- root (:420-439): idx is the depth-1 node on the path ----
            f2 bel[H];
#pragma unroll
            for (int i = 0; i < H; ++i) bel[i] = m[i];
            for (int cc = 0; cc < idx; ++cc) {
                f2 f[H];
                const f2* src = FT + (int64_t)(nwp_node(d, 1) + cc) * H * B + b;
#pragma unroll
                for (int i = 0; i < H; ++i) f[i] = src[(int64_t)i * B];
                f2_mul_into<H>(bel, f);
            }
            f2_normalize<Q>(bel);
            if (a.ext) { f2_mul_into<H>(bel, ex); f2_normalize<Q>(bel); }
            // ---- down the path of leaf t+1 (:443-459) ----
            for (int l = 1; l <= L; ++l) {
                const int gl = ghm_div_pow(t + 1, L - l, d);  // goal-path node at depth l
                const int al = ghm_div_pow(t, L - l, d);      // observed-path node at depth l
                const int cg = gl - ghm_div_s(gl, d) * s;
                const float* Tm = Tdn + ((l - 1) * s + cg) * QQ;
                f2 tt[H];
                if (gl == al) {                               // shared ancestor: cavity update (l <= L-1 here)
                    const f2* Hl = HS + (size_t)(l - 1) * H * NT + tid;
                    const f2* Ql = QS + (size_t)(l - 1) * H * NT + tid;
                    f2 w[H];
#pragma unroll
                    for (int i = 0; i < H; ++i) {
                        const f2 qv = Ql[i * NT];
                        w[i].x = qv.x > 0.f ? __fdividef(bel[i].x, qv.x) : 0.f;
                        w[i].y = qv.y > 0.f ? __fdividef(bel[i].y, qv.y) : 0.f;
                    }
                    f2_matvec_c<Q>(Tm, w, tt);
#pragma unroll
                    for (int i = 0; i < H; ++i) bel[i] = f2_mul(Hl[i * NT], tt[i]);
                } else {
                    f2_matvec_c<Q>(Tm, bel, tt);
#pragma unroll
                    for (int i = 0; i < H; ++i) bel[i] = tt[i];
                }
                f2_normalize<Q>(bel);
            }
            float sum = 0.f;
#pragma unroll
            for (int i = 0; i < H; ++i) sum += bel[i].x + bel[i].y;
            const float inv = 1.0f / sum;
            f2* srow = ST + ((size_t)tid * PC + p) * H;
#pragma unroll
            for (int i = 0; i < H; ++i) srow[i] = make_float2(bel[i].x * inv, bel[i].y * inv);
        }
        __syncthreads();
        // flush: the nv rows of a tree are nv*q contiguous floats of pp; consecutive lanes write consecutive 8-byte units
        const int upt = nv * H;
        const int ntree = (int)min((int64_t)NT, B - cta0);
        if (nv == PC) {                                       // compile-time divisor on the common path
            for (int v = tid; v < ntree * (PC * H); v += NT) {
                const int tr = v / (PC * H), k = v - tr * (PC * H);
                reinterpret_cast<f2*>(a.pp + ((cta0 + tr) * npos + t0) * Q)[k] = ST[(size_t)tr * PC * H + k];
            }
        } else {
            for (int v = tid; v < ntree * upt; v += NT) {
                const int tr = v / upt, k = v - tr * upt;
                reinterpret_cast<f2*>(a.pp + ((cta0 + tr) * npos + t0) * Q)[k] = ST[(size_t)tr * PC * H + k];
            }
        }
        __syncthreads();
    }
}

// GHM_EUNSUP -> the caller takes the generic kernels
template <int Q, bool GUIDE>
static int launch_nwp_u(const ghm_model* m, const NwpArgs& a, cudaStream_t st) {
    const GhmDev& d = m->d;
    constexpr int NW = GHM_TAB_WORDS;
    const size_t words = (size_t)d.n_mat * Q * Q;
    if (!d.ti || d.q != Q || d.s < 2 || d.s > 16 || d.L < 2 || 2 * words > (size_t)NW) return GHM_EUNSUP;
    if (((uintptr_t)a.pp % 8) || ((uintptr_t)a.full % 8) || (a.ext && ((uintptr_t)a.ext % 8))) return GHM_EUNSUP;
    if (GUIDE) return GHM_EUNSUP;                             // guide rows of one position sit 4q(n_L-1) bytes apart between
                                                             // trees: the (tree, position)-per-thread kernel writes them better
    const size_t lt = ((size_t)d.s * Q * Q + 3) / 4 * 4 * sizeof(float);
    const size_t lt_full = (lt + 15) / 16 * 16 + (size_t)NWP_NT * d.n_leaves;     // k_nwp_full_u: + the CTA's leaves as bytes
    if (lt_full > 100 * 1024) return GHM_EUNSUP;
    const size_t dyn_pos = lt + ((size_t)2 * (d.L - 1) + NWP_PC) * (Q / 2) * NWP_NT * sizeof(float2) + (size_t)(NWP_PC + d.s) * NWP_NT;
    if (dyn_pos > 100 * 1024) return GHM_EUNSUP;
    const int npos = d.n_leaves - 1;
    const unsigned gx = (unsigned)((a.B + NWP_NT - 1) / NWP_NT);
    // enough position groups to fill the GPU a few times over
    const int nchunk = (npos + NWP_PC - 1) / NWP_PC;
    int pg = (int)std::min<int64_t>(nchunk, std::max<int64_t>(1, (148 * 16 + gx - 1) / gx));
    if (pg > 65535) pg = 65535;
    NwpU u{(d.L - 2) * d.s, (int)words, d.s, pg};
    DnsTab<NW> tab;
    tab.v[0] = 0.f;
    memcpy(tab.v, m->h_TlinT, words * sizeof(float));
    memcpy(tab.v + words, m->h_Tlin, words * sizeof(float));
    GHM_CUDA_TRY(cudaFuncSetAttribute(k_nwp_full_u<Q, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lt_full));
    k_nwp_full_u<Q, NW><<<gx, NWP_NT, lt_full, st>>>(d, a, u, tab);
    GHM_CHECK_LAUNCH();
    GHM_CUDA_TRY(cudaFuncSetAttribute(k_nwp_pos_u<Q, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_pos));
    k_nwp_pos_u<Q, NW><<<dim3(gx, (unsigned)pg), NWP_NT, dyn_pos, st>>>(d, a, u, tab);
    GHM_CHECK_LAUNCH();
    return GHM_OK;
}

// ----------------------------------------------------------------------------------------
template <int Q, bool GUIDE>
static int launch_nwp(const ghm_model* m, const NwpArgs& a, cudaStream_t st) {
    const GhmDev& d = m->d;
    {
        const int rc = launch_nwp_u<Q, GUIDE>(m, a, st);
        if (rc != GHM_EUNSUP) return rc;
    }
    for (int l = d.L - 1; l >= 1; --l) {
        const int64_t nthreads = a.B * d.spow[l];
        k_nwp_full<Q><<<(unsigned)((nthreads + NWP_NT - 1) / NWP_NT), NWP_NT, 0, st>>>(d, a, l);
        GHM_CHECK_LAUNCH();
    }
    const size_t tab_bytes = (size_t)d.n_mat * Q * Q * 4;
    size_t dyn = (size_t)2 * (d.L > 1 ? d.L - 1 : 1) * Q * NWP_NT * 4;
    const bool smem_tab = tab_bytes + dyn <= 100 * 1024;
    if (smem_tab) dyn += tab_bytes;
    if (dyn > 200 * 1024)
        return ghm_fail(GHM_EUNSUP, "nwp kernel needs %zu bytes of shared memory (L=%d q=%d)", dyn, d.L, d.q);
    const int64_t rows = a.B * (int64_t)(d.n_leaves - 1);
    const unsigned grid = (unsigned)((rows + NWP_NT - 1) / NWP_NT);
    auto go = [&](auto kern) -> int {
        GHM_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
        kern<<<grid, NWP_NT, dyn, st>>>(d, a);
        GHM_CHECK_LAUNCH();
        return GHM_OK;
    };
    return smem_tab ? go(k_nwp_pos<Q, GUIDE, true>) : go(k_nwp_pos<Q, GUIDE, false>);
}

template <bool GUIDE>
static int dispatch_nwp(const ghm_model* m, const NwpArgs& a, cudaStream_t st) {
    switch (ghm_pad_q(m->d.q)) {
        case 4: return launch_nwp<4, GUIDE>(m, a, st);
        case 8: return launch_nwp<8, GUIDE>(m, a, st);
        case 10: return launch_nwp<10, GUIDE>(m, a, st);
        case 16: return launch_nwp<16, GUIDE>(m, a, st);
        default:
            return ghm_fail(GHM_EUNSUP, "variable_type=%d: register-resident kernels cover q <= %d in this build", m->d.q,
                            GHM_MAX_Q_REG);
    }
}

// wide q (ghm_wide_lvl.cuh): log-domain, warp per row; finished-subtree messages in the same [B][n_int][q] workspace
static int launch_nwp_wide(const ghm_model* m, const NwpArgs& a, float* const* guides, cudaStream_t st) {
    const GhmDev& d = m->d;
    WNwpArgs w{};
    w.B = a.B; w.leaves = a.leaves; w.leaf_dtype = a.leaf_dtype; w.ext = a.ext; w.pp = a.pp; w.full = a.full; w.n_int = a.n_int;
    w.guide = guides ? 1 : 0;
    if (guides)
        for (int i = 0; i < 2 * d.L + 1; ++i) w.guides[i] = guides[i];
    const int warps = WL_NT / 32;
    for (int l = d.L - 1; l >= 1; --l) {
        k_wl_nwp_full<<<wl_grid(a.B * d.spow[l]), WL_NT, (size_t)warps * d.QW * sizeof(float), st>>>(d, w, l);
        GHM_CHECK_LAUNCH();
    }
    const size_t dyn = (size_t)warps * (size_t)(d.QW + 2 * (d.L - 1) * d.QW) * sizeof(float);
    if (dyn > 200 * 1024)
        return ghm_fail(GHM_EUNSUP, "wide nwp kernel needs %zu bytes of shared memory (L=%d q=%d)", dyn, d.L, d.q);
    GHM_CUDA_TRY(cudaFuncSetAttribute(k_wl_nwp_pos, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    k_wl_nwp_pos<<<wl_grid(a.B * (int64_t)(d.n_leaves - 1)), WL_NT, dyn, st>>>(d, w);
    GHM_CHECK_LAUNCH();
    return GHM_OK;
}

static int64_t nwp_ws_bytes(const ghm_model_t* m, int64_t B) {
    if (!m || B <= 0) return 16;
    return std::max<int64_t>(16, B * (int64_t)(1 + m->d.edge_off[m->d.L]) * m->d.q * (int64_t)sizeof(float));
}

static int nwp_common(const ghm_model_t* m, int64_t B, const void* leaves, int leaf_dtype, const float* ext,
                      float* const* guides, float* pp, void* workspace, void* stream) {
    if (!m || !leaves || !pp || !workspace) return ghm_fail(GHM_EINVAL, "ghm_bp_nwp: null argument");
    if (B <= 0) return B == 0 ? GHM_OK : ghm_fail(GHM_EINVAL, "ghm_bp_nwp: negative batch");
    if (leaf_dtype != GHM_LEAF_I64 && leaf_dtype != GHM_LEAF_U8) return ghm_fail(GHM_EINVAL, "bad leaf_dtype %d", leaf_dtype);
    if (m->d.n_leaves < 2) return GHM_OK;
    int prev = 0;
    cudaGetDevice(&prev);
    if (prev != m->device) cudaSetDevice(m->device);
    NwpArgs a{};
    a.B = B; a.leaves = leaves; a.leaf_dtype = leaf_dtype; a.ext = ext; a.pp = pp;
    a.full = (float*)workspace; a.n_int = 1 + m->d.edge_off[m->d.L];
    int rc;
    if (m->d.QW) {                                             // 16 < q <= 256: one warp per (tree, node / position) row
        rc = launch_nwp_wide(m, a, guides, (cudaStream_t)stream);
    } else if (guides) {
        for (int i = 0; i < 2 * m->d.L + 1; ++i) a.guides[i] = guides[i];
        rc = dispatch_nwp<true>(m, a, (cudaStream_t)stream);
    } else {
        rc = dispatch_nwp<false>(m, a, (cudaStream_t)stream);
    }
    if (prev != m->device) cudaSetDevice(prev);
    return rc;
}

extern "C" int64_t ghm_bp_nwp_workspace_bytes(const ghm_model_t* m, int64_t B) { return nwp_ws_bytes(m, B); }
extern "C" int ghm_bp_nwp(const ghm_model_t* m, int64_t B, const void* leaves, int leaf_dtype, const float* ext,
                          float* pp, void* workspace, void* stream) {
    return nwp_common(m, B, leaves, leaf_dtype, ext, nullptr, pp, workspace, stream);
}
extern "C" int64_t ghm_guides_nwp_workspace_bytes(const ghm_model_t* m, int64_t B) { return nwp_ws_bytes(m, B); }
extern "C" int ghm_guides_nwp(const ghm_model_t* m, int64_t B, const void* leaves, int leaf_dtype, const float* ext,
                              float* const* guides, float* pp, void* workspace, void* stream) {
    if (!guides) return ghm_fail(GHM_EINVAL, "ghm_guides_nwp: null guides");
    return nwp_common(m, B, leaves, leaf_dtype, ext, guides, pp, workspace, stream);
}
