// ghm_wide_lvl.cuh -- log-domain belief propagation for 16 < q <= 256 with one WARP per message row.
//
// The register-resident kernels (one thread per tree / node, q <= 16) cannot hold a q-vector per thread beyond
// q = 16.  Here a message is spread over the 32 lanes of a warp (lane owns states lane, lane + 32, ...), the
// child -> parent product `log(T @ exp(h))` (reference src/ghmclip/data/data_random_GHM.py:207,497) streams the
// table rows coalesced through L1/L2 with the exponentiated vector broadcast from shared memory, and max / sum
// reductions are warp shuffles.  Same shift conventions as the reference, so these kernels
// serve the outputs that expose log-messages:
//   * the compact [B][node][q] message stores behind the guide tensors of BP_CLS / BP_DNS (guided_info, :526-592),
//   * BP_NWP_autoregressive (:336-463) with and without its per-position guide tensors.
// The throughput path for wide q (posteriors only) is the batched row-GEMM form of ghm_wide.cu / ghm_wide_tc.cu.
#pragma once
#include "ghm_common.cuh"

#define WL_NT 128                      // 4 warps = 4 rows per CTA
#define WL_NV 8                        // q <= 256: at most 8 states per lane
// compile-time trip count (the per-lane vectors stay in registers), runtime guard
#define WL_FOR(i, nv) _Pragma("unroll") for (int i = 0; i < WL_NV; ++i) if (i < (nv))

__device__ __forceinline__ float wl_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float wl_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// a q-vector spread over a warp: v[i] is state lane + 32 i; entries past q are -inf (log domain)
struct WVec { float v[WL_NV]; };

__device__ __forceinline__ float wv_max(const WVec& x, int nv) {
    float m = -INFINITY;
    WL_FOR(i, nv) m = fmaxf(m, x.v[i]);
    return wl_max(m);
}
__device__ __forceinline__ void wv_shift(WVec& x, int nv) {                 // x -= max(x)   (:197,208,...)
    const float m = wv_max(x, nv);
    WL_FOR(i, nv) x.v[i] -= m;
}
__device__ __forceinline__ void wv_load(WVec& x, const float* p, int q, int nv, int lane) {
    WL_FOR(i, nv) { const int k = lane + 32 * i; x.v[i] = k < q ? p[k] : -INFINITY; }
}
__device__ __forceinline__ void wv_store(const WVec& x, float* p, int q, int nv, int lane) {
    WL_FOR(i, nv) { const int k = lane + 32 * i; if (k < q) p[k] = x.v[i]; }
}
__device__ __forceinline__ void wv_add(WVec& x, const WVec& y, int nv) {
    WL_FOR(i, nv) x.v[i] += y.v[i];
}

// out = log(W^T-contract: sum_k W[k][n] exp(h[k] - m)) + m  with W rows of stride QW (zero padded):
// W = d.Wdn + mi*QW*QW gives log(T @ exp(h)) (child -> parent), W = d.Wup + ... gives log(T^T @ exp(h)).
// `ebuf` is this warp's shared-memory scratch of QW floats.
__device__ __forceinline__ void wl_log_matvec(const float* __restrict__ W, int QW, int q, int nv, int lane, const WVec& h,
                                              WVec& out, float* ebuf) {
    const float m = wv_max(h, nv);
    __syncwarp();
    WL_FOR(i, nv) { const int k = lane + 32 * i; ebuf[k] = k < q ? __expf(h.v[i] - m) : 0.f; }
    __syncwarp();
    float acc[WL_NV];
    for (int i = 0; i < WL_NV; ++i) acc[i] = 0.f;
    for (int k = 0; k < q; ++k) {
        const float e = ebuf[k];
        const float* row = W + (size_t)k * QW + lane;
#pragma unroll
        for (int i = 0; i < WL_NV; ++i)
            if (i < nv) acc[i] = fmaf(e, __ldg(row + 32 * i), acc[i]);
    }
    WL_FOR(i, nv) { const int k = lane + 32 * i; out.v[i] = k < q ? __logf(acc[i]) + m : -INFINITY; }
}

__device__ __forceinline__ int wl_leaf(const void* leaves, int dtype, int64_t off, int q, int* status, int lane) {
    int64_t v = dtype == GHM_LEAF_I64 ? reinterpret_cast<const int64_t*>(leaves)[off]
                                      : (int64_t) reinterpret_cast<const uint8_t*>(leaves)[off];
    if (v < 0 || v >= q) { if (lane == 0) atomicOr(status, 1); v = v < 0 ? 0 : q - 1; }
    return (int)v;
}
__device__ __forceinline__ int wl_node_off(const GhmDev& d, int l) { return l == 0 ? 0 : 1 + d.edge_off[l]; }
__device__ __forceinline__ int wl_mat(const GhmDev& d, int l, int idx) {     // matrix of the edge INTO node (l, idx)
    return d.mat_off[l] + (d.ti ? idx - ghm_div_s(idx, d) * d.s : idx);
}
// log T[:, x] of the edge into leaf `leaf` (TlogT is [m][b][a] with stride QP)
__device__ __forceinline__ void wl_leaf_col(const GhmDev& d, int leaf, int x, WVec& out, int nv, int lane) {
    const float* row = d.TlogT + ((size_t)wl_mat(d, d.L, leaf) * d.QP + x) * d.QP;
    WL_FOR(i, nv) { const int k = lane + 32 * i; out.v[i] = k < d.q ? __ldg(row + k) : -INFINITY; }
}

struct WLvlArgs {
    int64_t B;
    const void* leaves; int leaf_dtype;     // cls / nwp
    const float* z; float sigma;            // dns
    const float* ext;                       // [B,q] or null
    float* HD; float* QD; float* BU;        // [B][n_nodes][q] compact stores
    int n_nodes;
    float* post; float* root_hd; float* mean;
};

// row -> (tree b, node idx of depth l); one warp per row
struct WRow { int64_t b; int idx; bool ok; int lane, warp; };
__device__ __forceinline__ WRow wl_row(const GhmDev& d, int64_t B, int n_per_tree) {
    WRow r;
    r.lane = threadIdx.x & 31; r.warp = threadIdx.x >> 5;
    const int64_t row = (int64_t)blockIdx.x * (WL_NT / 32) + r.warp;
    r.b = row / n_per_tree;
    r.idx = (int)(row - r.b * n_per_tree);
    r.ok = r.b < B;
    return r;
}
static inline unsigned wl_grid(int64_t rows) { return (unsigned)((rows + WL_NT / 32 - 1) / (WL_NT / 32)); }

// ---- BP_CLS, log domain (reference :185-221): nodes of depth l, bottom-up ------------------------------------
static __global__ void __launch_bounds__(WL_NT) k_wl_cls(const GhmDev d, const WLvlArgs a, int l) {
    extern __shared__ float wl_smem[];
    const WRow r = wl_row(d, a.B, d.spow[l]);
    if (!r.ok) return;
    const int L = d.L, s = d.s, q = d.q, QW = d.QW, nv = QW / 32, lane = r.lane;
    float* ebuf = wl_smem + (size_t)r.warp * QW;
    float* HD = a.HD + r.b * (int64_t)a.n_nodes * q;
    WVec acc;
    WL_FOR(i, nv) acc.v[i] = (lane + 32 * i) < q ? 0.f : -INFINITY;
    for (int c = 0; c < s; ++c) {
        const int child = r.idx * s + c;
        WVec m;
        if (l == L - 1) {
            const int x = wl_leaf(a.leaves, a.leaf_dtype, r.b * d.n_leaves + child, q, d.status, lane);
            wl_leaf_col(d, child, x, m, nv, lane);                                   // log T[:, x]   (:196)
        } else {
            WVec h;
            wv_load(h, HD + (int64_t)(wl_node_off(d, l + 1) + child) * q, q, nv, lane);
            wl_log_matvec(d.Wdn + (size_t)wl_mat(d, l + 1, child) * QW * QW, QW, q, nv, lane, h, m, ebuf);   // (:207)
        }
        wv_add(acc, m, nv);
    }
    wv_shift(acc, nv);                                                               // (:197,208)
    wv_store(acc, HD + (int64_t)(wl_node_off(d, l) + r.idx) * q, q, nv, lane);
    if (l > 0) return;
    if (a.root_hd) wv_store(acc, a.root_hd + r.b * q, q, nv, lane);
    if (a.post) {                                                                    // (:213-217)
        WVec h0;
        WL_FOR(i, nv) { const int k = lane + 32 * i; h0.v[i] = k < q ? acc.v[i] + logf(__ldg(d.py + k)) : -INFINITY; }
        const float m0 = wv_max(h0, nv);
        float sum = 0.f;
        WL_FOR(i, nv) { h0.v[i] = (lane + 32 * i) < q ? expf(h0.v[i] - m0) : 0.f; sum += h0.v[i]; }
        const float inv = 1.0f / wl_sum(sum);
        WL_FOR(i, nv) h0.v[i] *= inv;
        wv_store(h0, a.post + r.b * q, q, nv, lane);
    }
}

// ---- BP_DNS upward pass (reference :483-506): nodes of depth l -------------------------------------------------
static __global__ void __launch_bounds__(WL_NT) k_wl_dns_up(const GhmDev d, const WLvlArgs a, int l) {
    extern __shared__ float wl_smem[];
    const WRow r = wl_row(d, a.B, d.spow[l]);
    if (!r.ok) return;
    const int L = d.L, s = d.s, q = d.q, QW = d.QW, nv = QW / 32, lane = r.lane;
    float* ebuf = wl_smem + (size_t)r.warp * QW;
    const int64_t base = r.b * (int64_t)a.n_nodes * q;
    float *HD = a.HD + base, *QD = a.QD + base, *BU = a.BU + base;
    const int node = wl_node_off(d, l) + r.idx;
    WVec acc;
    if (l == L) {                                             // leaves: hd unshifted (:485)
        const float inv2s2 = 0.5f / (a.sigma * a.sigma);
        const float zi = a.z[r.b * d.n_leaves + r.idx];
        WL_FOR(i, nv) {
            const int k = lane + 32 * i;
            const float dlt = zi - (float)k;
            acc.v[i] = k < q ? -dlt * dlt * inv2s2 : -INFINITY;
        }
    } else {                                                  // hd = sum qd(children) - max (:494-496)
        WL_FOR(i, nv) acc.v[i] = (lane + 32 * i) < q ? 0.f : -INFINITY;
        for (int c = 0; c < s; ++c) {
            WVec m;
            wv_load(m, QD + (int64_t)(wl_node_off(d, l + 1) + r.idx * s + c) * q, q, nv, lane);
            wv_add(acc, m, nv);
        }
        wv_shift(acc, nv);
    }
    if (l > 0) {
        WVec m;
        wl_log_matvec(d.Wdn + (size_t)wl_mat(d, l, r.idx) * QW * QW, QW, q, nv, lane, acc, m, ebuf);   // qd (:487,497)
        wv_store(acc, HD + (int64_t)node * q, q, nv, lane);
        wv_store(m, QD + (int64_t)node * q, q, nv, lane);
        return;
    }
    if (a.ext) {                                              // root: bu aliases hd, + ext without a re-shift (:501-506)
        WL_FOR(i, nv) { const int k = lane + 32 * i; if (k < q) acc.v[i] += a.ext[r.b * q + k]; }
    }
    wv_store(acc, HD, q, nv, lane);
    wv_store(acc, BU, q, nv, lane);
}

// ---- BP_DNS downward pass (reference :509-519): nodes of depth l >= 1 --------------------------------------------
static __global__ void __launch_bounds__(WL_NT) k_wl_dns_down(const GhmDev d, const WLvlArgs a, int l) {
    extern __shared__ float wl_smem[];
    const WRow r = wl_row(d, a.B, d.spow[l]);
    if (!r.ok) return;
    const int L = d.L, q = d.q, QW = d.QW, nv = QW / 32, lane = r.lane;
    float* ebuf = wl_smem + (size_t)r.warp * QW;
    const int64_t base = r.b * (int64_t)a.n_nodes * q;
    float *HD = a.HD + base, *QD = a.QD + base, *BU = a.BU + base;
    const int node = wl_node_off(d, l) + r.idx;
    const int par = wl_node_off(d, l - 1) + ghm_div_s(r.idx, d);
    WVec bp, qv, hv, m;
    wv_load(bp, BU + (int64_t)par * q, q, nv, lane);
    wv_load(qv, QD + (int64_t)node * q, q, nv, lane);
    wv_load(hv, HD + (int64_t)node * q, q, nv, lane);
    WL_FOR(i, nv) bp.v[i] = (lane + 32 * i) < q ? bp.v[i] - qv.v[i] : -INFINITY;
    wl_log_matvec(d.Wup + (size_t)wl_mat(d, l, r.idx) * QW * QW, QW, q, nv, lane, bp, m, ebuf);
    WL_FOR(i, nv) m.v[i] = (lane + 32 * i) < q ? hv.v[i] + m.v[i] : -INFINITY;
    wv_shift(m, nv);
    wv_store(m, BU + (int64_t)node * q, q, nv, lane);
    if (l == L && a.mean) {                                   // (:516-519)
        float num = 0.f, den = 0.f;
        WL_FOR(i, nv) {
            const int k = lane + 32 * i;
            const float e = k < q ? expf(m.v[i]) : 0.f;
            num += (float)k * e; den += e;
        }
        num = wl_sum(num); den = wl_sum(den);
        if (lane == 0) a.mean[r.b * d.n_leaves + r.idx] = num / den;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// BP_NWP_autoregressive (reference :336-463), parallel in the position like ghm_nwp.cu:
//   k_wl_nwp_full  one warp per (tree, internal node): the shifted message qd_full(v) of the subtree under v given
//                  ALL its leaves -- what the reference's finished subtrees hold (:394-399);
//   k_wl_nwp_pos   one warp per (tree, position t): up the root path of leaf t, root belief with the external
//                  message, down the path of leaf t+1 (cavity while shared, push-down after the split), posterior
//                  of leaf t+1 and -- when asked -- the 2L+1 guide tensors with the reference's shift points.
// ------------------------------------------------------------------------------------------------------------------
struct WNwpArgs {
    int64_t B;
    const void* leaves; int leaf_dtype;
    const float* ext;
    float* pp;                         // [B][nL-1][q]
    float* full;                       // [B][n_int][q] finished-subtree messages (workspace), log domain
    int n_int;
    float* guides[2 * GHM_MAX_LEVELS + 1];
    int guide;
};

// shifted leaf message log T[:, x] - max (:374-375)
__device__ __forceinline__ void wl_nwp_leaf(const GhmDev& d, const WNwpArgs& a, int64_t b, int leaf, WVec& m, int nv, int lane) {
    const int x = wl_leaf(a.leaves, a.leaf_dtype, b * d.n_leaves + leaf, d.q, d.status, lane);
    wl_leaf_col(d, leaf, x, m, nv, lane);
    wv_shift(m, nv);
}

static __global__ void __launch_bounds__(WL_NT) k_wl_nwp_full(const GhmDev d, const WNwpArgs a, int l) {
    extern __shared__ float wl_smem[];
    const WRow r = wl_row(d, a.B, d.spow[l]);
    if (!r.ok) return;
    const int L = d.L, s = d.s, q = d.q, QW = d.QW, nv = QW / 32, lane = r.lane;
    float* ebuf = wl_smem + (size_t)r.warp * QW;
    float* F = a.full + r.b * (int64_t)a.n_int * q;
    WVec h;
    WL_FOR(i, nv) h.v[i] = (lane + 32 * i) < q ? 0.f : -INFINITY;
    for (int c = 0; c < s; ++c) {
        WVec m;
        if (l == L - 1) wl_nwp_leaf(d, a, r.b, r.idx * s + c, m, nv, lane);
        else wv_load(m, F + (int64_t)(wl_node_off(d, l + 1) + r.idx * s + c) * q, q, nv, lane);
        wv_add(h, m, nv);
    }
    wv_shift(h, nv);                                                                 // (:397)
    WVec u;
    wl_log_matvec(d.Wdn + (size_t)wl_mat(d, l, r.idx) * QW * QW, QW, q, nv, lane, h, u, ebuf);
    wv_shift(u, nv);                                                                 // (:399)
    wv_store(u, F + (int64_t)(wl_node_off(d, l) + r.idx) * q, q, nv, lane);
}

// shared memory: per warp QW floats of matvec scratch + 2 (L-1) q-vectors (hd, qd of the path nodes at depth 1..L-1)
static __global__ void __launch_bounds__(WL_NT) k_wl_nwp_pos(const GhmDev d, const WNwpArgs a) {
    extern __shared__ float wl_smem[];
    const int L = d.L, s = d.s, q = d.q, QW = d.QW, nv = QW / 32, nL = d.n_leaves, npos = nL - 1;
    const WRow r = wl_row(d, a.B, npos);
    if (!r.ok) return;
    const int lane = r.lane, t = r.idx;
    const int64_t b = r.b, row = b * npos + t;
    float* wbase = wl_smem + (size_t)r.warp * (size_t)(QW + 2 * (L - 1) * QW);
    float* ebuf = wbase;
    float* HDs = wbase + QW;                                  // [L-1][QW]  depth l at (l-1)
    float* QDs = HDs + (size_t)(L - 1) * QW;
    const float* F = a.full + b * (int64_t)a.n_int * q;

    WVec m;
    wl_nwp_leaf(d, a, b, t, m, nv, lane);                     // observed leaf t (:374-375)
    if (a.guide) wv_store(m, a.guides[0] + row * q, q, nv, lane);
    // ---- up the path: depth L-1 .. 1 (:389-415) ----
    int idx = t;
    for (int l = L - 1; l >= 1; --l) {
        const int pidx = ghm_div_s(idx, d);
        const int c = idx - pidx * s;
        WVec h = m;
        for (int cc = 0; cc < c; ++cc) {                      // finished children left of the path
            WVec f;
            if (l == L - 1) wl_nwp_leaf(d, a, b, pidx * s + cc, f, nv, lane);
            else wv_load(f, F + (int64_t)(wl_node_off(d, l + 1) + pidx * s + cc) * q, q, nv, lane);
            wv_add(h, f, nv);
        }
        wv_shift(h, nv);
        wl_log_matvec(d.Wdn + (size_t)wl_mat(d, l, pidx) * QW * QW, QW, q, nv, lane, h, m, ebuf);
        wv_shift(m, nv);
        WL_FOR(i, nv) { HDs[(l - 1) * QW + lane + 32 * i] = h.v[i]; QDs[(l - 1) * QW + lane + 32 * i] = m.v[i]; }
        if (a.guide) {
            float* g = a.guides[L - l] + row * 2 * q;
            wv_store(h, g, q, nv, lane);
            wv_store(m, g + q, q, nv, lane);
        }
        idx = pidx;
    }
    // ---- root (:420-439): idx is the depth-1 node on the path ----
    WVec bel = m;
    for (int cc = 0; cc < idx; ++cc) {
        WVec f;
        if (L == 1) wl_nwp_leaf(d, a, b, cc, f, nv, lane);
        else wv_load(f, F + (int64_t)(wl_node_off(d, 1) + cc) * q, q, nv, lane);
        wv_add(bel, f, nv);
    }
    wv_shift(bel, nv);
    if (a.ext) {
        WL_FOR(i, nv) { const int k = lane + 32 * i; if (k < q) bel.v[i] += a.ext[b * q + k]; }
        wv_shift(bel, nv);
    }
    if (a.guide) {
        float* g = a.guides[L] + row * 2 * q;
        wv_store(bel, g, q, nv, lane);
        wv_store(bel, g + q, q, nv, lane);
    }
    // ---- down the path of leaf t+1 (:443-459) ----
    for (int l = 1; l <= L; ++l) {
        const int g = ghm_div_pow(t + 1, L - l, d);            // goal-path node at depth l
        const int a_l = ghm_div_pow(t, L - l, d);              // observed-path node at depth l
        const float* W = d.Wup + (size_t)wl_mat(d, l, g) * QW * QW;
        WVec tt;
        if (g == a_l) {                                         // shared ancestor: cavity update (l <= L-1 here)
            WVec w;
            WL_FOR(i, nv) w.v[i] = (lane + 32 * i) < q ? bel.v[i] - QDs[(l - 1) * QW + lane + 32 * i] : -INFINITY;
            wl_log_matvec(W, QW, q, nv, lane, w, tt, ebuf);
            WL_FOR(i, nv) bel.v[i] = (lane + 32 * i) < q ? HDs[(l - 1) * QW + lane + 32 * i] + tt.v[i] : -INFINITY;
        } else {
            wl_log_matvec(W, QW, q, nv, lane, bel, tt, ebuf);
            bel = tt;
        }
        wv_shift(bel, nv);
        if (a.guide) wv_store(bel, a.guides[L + l] + row * q, q, nv, lane);
    }
    float sum = 0.f;
    WVec p;
    WL_FOR(i, nv) { p.v[i] = (lane + 32 * i) < q ? __expf(bel.v[i]) : 0.f; sum += p.v[i]; }
    const float inv = 1.0f / wl_sum(sum);
    WL_FOR(i, nv) p.v[i] *= inv;
    wv_store(p, a.pp + row * q, q, nv, lane);
}
