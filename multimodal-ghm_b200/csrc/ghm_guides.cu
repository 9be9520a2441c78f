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
#include <algorithm>

#include "ghm_vec.cuh"

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

template <int Q>
__device__ __forceinline__ void load_vec(const float* p, float (&v)[Q], int q) {
#pragma unroll
    for (int k = 0; k < Q; ++k) v[k] = (k < q) ? p[k] : -INFINITY;
}
template <int Q>
__device__ __forceinline__ void store_vec(float* p, const float (&v)[Q], int q) {
#pragma unroll
    for (int k = 0; k < Q; ++k)
        if (k < q) p[k] = v[k];
}

// out[a] = log(sum_b T[a][b] exp(h[b] - m)) + m   with m = max h   (== log(T @ exp(h)) without overflow)
template <int Q>
__device__ __forceinline__ void log_matvec(const float* __restrict__ T, const float (&h)[Q], float (&out)[Q], int q) {
    const float m = ghm_vmax<Q>(h);
    float e[Q], u[Q];
#pragma unroll
    for (int k = 0; k < Q; ++k) e[k] = (k < q) ? expf(h[k] - m) : 0.f;
    ghm_matvec<Q>(T, e, u);
#pragma unroll
    for (int k = 0; k < Q; ++k) out[k] = (k < q) ? logf(u[k]) + m : -INFINITY;
}
template <int Q>
__device__ __forceinline__ void log_matvec_t(const float* __restrict__ T, const float (&h)[Q], float (&out)[Q], int q) {
    const float m = ghm_vmax<Q>(h);
    float e[Q], u[Q];
#pragma unroll
    for (int k = 0; k < Q; ++k) e[k] = (k < q) ? expf(h[k] - m) : 0.f;
    ghm_matvec_t<Q>(T, e, u);
#pragma unroll
    for (int k = 0; k < Q; ++k) out[k] = (k < q) ? logf(u[k]) + m : -INFINITY;
}

__device__ __forceinline__ int leaf_at(const void* leaves, int dtype, int64_t off, int q, int* status) {
    int64_t v = dtype == GHM_LEAF_I64 ? reinterpret_cast<const int64_t*>(leaves)[off]
                                      : (int64_t) reinterpret_cast<const uint8_t*>(leaves)[off];
    if (v < 0 || v >= q) { atomicOr(status, 1); v = v < 0 ? 0 : q - 1; }
    return (int)v;
}

// node id of (depth l, index idx) when depths 0..L are stored: off_all(l) = (s^l - 1)/(s - 1)
__device__ __forceinline__ int node_off_all(const GhmDev& d, int l) { return l == 0 ? 0 : 1 + d.edge_off[l]; }

// ---- BP_CLS, log domain (reference :185-221) ---------------------------------------------------
template <int Q>
__global__ void __launch_bounds__(GD_NT) k_lvl_cls(const GhmDev d, const LvlArgs a) {
    const int64_t b = (int64_t)blockIdx.x * GD_NT + threadIdx.x;
    if (b >= a.B) return;
    const int L = d.L, s = d.s, q = d.q, nL = d.n_leaves;
    float* HD = a.HD + b * (int64_t)a.n_nodes * q;
    float acc[Q];
    for (int l = L - 1; l >= 0; --l) {
        const int n = d.spow[l];
        for (int idx = 0; idx < n; ++idx) {
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
        }
    }
    // root: acc == hd(root)
    if (a.root_hd) store_vec<Q>(a.root_hd + b * q, acc, q);
    if (a.post) {
        float h0[Q];
#pragma unroll
        for (int k = 0; k < Q; ++k) h0[k] = (k < q) ? acc[k] + logf(__ldg(d.py + k)) : -INFINITY;   // (:213)
        const float mx = ghm_vmax<Q>(h0);
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < Q; ++k) { h0[k] = (k < q) ? expf(h0[k] - mx) : 0.f; sum += h0[k]; }
        const float inv = 1.0f / sum;
#pragma unroll
        for (int k = 0; k < Q; ++k) h0[k] *= inv;
        store_vec<Q>(a.post + b * q, h0, q);
    }
}

// ---- BP_DNS, log domain (reference :467-523) ---------------------------------------------------
template <int Q>
__global__ void __launch_bounds__(GD_NT) k_lvl_dns(const GhmDev d, const LvlArgs a) {
    const int64_t b = (int64_t)blockIdx.x * GD_NT + threadIdx.x;
    if (b >= a.B) return;
    const int L = d.L, s = d.s, q = d.q, nL = d.n_leaves;
    const int64_t base = b * (int64_t)a.n_nodes * q;
    float* HD = a.HD + base;
    float* QD = a.QD + base;
    float* BU = a.BU + base;
    const float inv2s2 = 0.5f / (a.sigma * a.sigma);
    // leaves: hd unshifted (:485), qd = log(T @ exp(hd)) (:487)
    {
        const int o = node_off_all(d, L);
        for (int i = 0; i < nL; ++i) {
            const float zi = a.z[b * nL + i];
            float h[Q], m[Q];
#pragma unroll
            for (int k = 0; k < Q; ++k) { const float dlt = zi - (float)k; h[k] = (k < q) ? -dlt * dlt * inv2s2 : -INFINITY; }
            const int mi = d.mat_off[L] + (d.ti ? i % s : i);
            log_matvec<Q>(d.Tlin + (size_t)mi * Q * Q, h, m, q);
            store_vec<Q>(HD + (int64_t)(o + i) * q, h, q);
            store_vec<Q>(QD + (int64_t)(o + i) * q, m, q);
        }
    }
    // internal nodes bottom-up: hd = sum qd(children) - max (:494-496), qd = log(T @ exp(hd)) (:497)
    float acc[Q];
    for (int l = L - 1; l >= 0; --l) {
        const int n = d.spow[l];
        for (int idx = 0; idx < n; ++idx) {
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
            if (l > 0) {
                const int mi = d.mat_off[l] + (d.ti ? idx % s : idx);
                float m[Q];
                log_matvec<Q>(d.Tlin + (size_t)mi * Q * Q, acc, m, q);
                store_vec<Q>(HD + (int64_t)(node_off_all(d, l) + idx) * q, acc, q);
                store_vec<Q>(QD + (int64_t)(node_off_all(d, l) + idx) * q, m, q);
            }
        }
    }
    // root: bu aliases hd and gets the external message without a re-shift (:501-506)
    if (a.ext) {
#pragma unroll
        for (int k = 0; k < Q; ++k)
            if (k < q) acc[k] += a.ext[b * q + k];
    }
    store_vec<Q>(HD, acc, q);
    store_vec<Q>(BU, acc, q);
    // top-down: bu = hd + log(T^T @ exp(bu_parent - qd)) - max (:509-514)
    for (int l = 1; l <= L; ++l) {
        const int n = d.spow[l];
        for (int idx = 0; idx < n; ++idx) {
            const int node = node_off_all(d, l) + idx;
            const int par = node_off_all(d, l - 1) + ghm_div_s(idx, d);
            float bp[Q], qv[Q], hv[Q], m[Q];
            load_vec<Q>(BU + (int64_t)par * q, bp, q);
            load_vec<Q>(QD + (int64_t)node * q, qv, q);
            load_vec<Q>(HD + (int64_t)node * q, hv, q);
#pragma unroll
            for (int k = 0; k < Q; ++k) bp[k] = (k < q) ? bp[k] - qv[k] : -INFINITY;
            const int mi = d.mat_off[l] + (d.ti ? idx % s : idx);
            log_matvec_t<Q>(d.Tlin + (size_t)mi * Q * Q, bp, m, q);
#pragma unroll
            for (int k = 0; k < Q; ++k) m[k] = (k < q) ? hv[k] + m[k] : -INFINITY;
            const float mx = ghm_vmax<Q>(m);
#pragma unroll
            for (int k = 0; k < Q; ++k) m[k] -= mx;
            store_vec<Q>(BU + (int64_t)node * q, m, q);
            if (l == L && a.mean) {                                               // (:516-519)
                float num = 0.f, den = 0.f;
#pragma unroll
                for (int k = 0; k < Q; ++k) { const float e = (k < q) ? expf(m[k]) : 0.f; num += (float)k * e; den += e; }
                a.mean[b * nL + idx] = num / den;
            }
        }
    }
}

// ---- expansion: node messages -> [B, n_L, C] guide tensor ------------------------------------------
struct ExpandArgs {
    int64_t B;
    const float* src[3];     // up to 3 message arrays [B][n_nodes][q]
    int nsrc;
    int n_nodes, q, nL;
    int node_off;            // first node of the level inside a tree
    unsigned R_magic; int R; // leaves per node at this level: node = i / R
    unsigned C_magic; int C; // C = nsrc * q
    unsigned q_magic;
    float* out;              // [B][nL][C]
};

__device__ __forceinline__ int div_magic(int x, int div, unsigned magic) {
    return div == 1 ? x : (int)__umulhi((unsigned)x, magic);
}

__global__ void __launch_bounds__(256) k_expand(const ExpandArgs a) {
    const int row = a.nL * a.C;                          // floats per tree
    for (int64_t b = blockIdx.y; b < a.B; b += gridDim.y) {
        float* out = a.out + b * row;
        for (int e = blockIdx.x * 256 + threadIdx.x; e < row; e += gridDim.x * 256) {
            const int i = div_magic(e, a.C, a.C_magic);
            const int c = e - i * a.C;
            const int j = div_magic(c, a.q, a.q_magic);
            const int k = c - j * a.q;
            const int node = a.node_off + div_magic(i, a.R, a.R_magic);
            const float* src = j == 0 ? a.src[0] : (j == 1 ? a.src[1] : a.src[2]);
            out[e] = __ldg(src + (b * a.n_nodes + node) * a.q + k);
        }
    }
}

static unsigned magic_of(int d) { return d >= 2 ? (unsigned)((0x100000000ull + (unsigned)d - 1) / (unsigned)d) : 0u; }

static int launch_expand(const GhmDev& d, int64_t B, int level, int n_nodes, bool all_levels, const float* s0,
                         const float* s1, const float* s2, int nsrc, float* out, cudaStream_t st) {
    ExpandArgs a{};
    a.B = B; a.src[0] = s0; a.src[1] = s1; a.src[2] = s2; a.nsrc = nsrc;
    a.n_nodes = n_nodes; a.q = d.q; a.nL = d.n_leaves;
    a.node_off = level == 0 ? 0 : 1 + d.edge_off[level];
    (void)all_levels;
    a.R = d.spow[d.L - level]; a.R_magic = magic_of(a.R);
    a.C = nsrc * d.q; a.C_magic = magic_of(a.C);
    a.q_magic = magic_of(d.q);
    a.out = out;
    const int row = a.nL * a.C;
    if ((int64_t)row * 1 >= (1ll << 31)) return ghm_fail(GHM_EUNSUP, "guide row too large");
    dim3 grid((unsigned)std::min((row + 255) / 256, 64), (unsigned)std::min<int64_t>(B, 65535));
    k_expand<<<grid, 256, 0, st>>>(a);
    GHM_CHECK_LAUNCH();
    return GHM_OK;
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

static int64_t n_nodes_all(const GhmDev& d) { return 1 + (int64_t)d.n_edges; }
static int64_t n_nodes_int(const GhmDev& d) { return 1 + (int64_t)d.edge_off[d.L]; }   // depths 0..L-1

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
    float* HD = guides[L - 1];
    if (n_int > (int64_t)d.n_leaves) return ghm_fail(GHM_EUNSUP, "unexpected tree shape");
    LvlArgs a{};
    a.B = B; a.leaves = leaves; a.leaf_dtype = leaf_dtype; a.HD = HD; a.n_nodes = (int)n_int;
    a.post = post; a.root_hd = root_hd;
    // the root guide needs hd(root) after HD's storage is overwritten: keep it in root_hd (caller buffer)
    if (!root_hd) return ghm_fail(GHM_EINVAL, "ghm_guides_cls: root_hd output is required");
    int rc = dispatch_q(d.q, [&](auto Qc) -> int {
        constexpr int Q = decltype(Qc)::value;
        k_lvl_cls<Q><<<(unsigned)((B + GD_NT - 1) / GD_NT), GD_NT, 0, st>>>(d, a);
        GHM_CHECK_LAUNCH();
        return GHM_OK;
    });
    if (rc) return rc;
    // guides[j] <- depth L-1-j, j = 0..L-2 read from HD (inside guides[L-1]); then the root level from root_hd
    for (int j = 0; j < L - 1; ++j) {
        rc = launch_expand(d, B, L - 1 - j, (int)n_int, false, HD, nullptr, nullptr, 1, guides[j], st);
        if (rc) return rc;
    }
    // root level: source = root_hd viewed as [B][1 node][q]
    {
        ExpandArgs e{};
        e.B = B; e.src[0] = root_hd; e.nsrc = 1; e.n_nodes = 1; e.q = d.q; e.nL = d.n_leaves; e.node_off = 0;
        e.R = d.n_leaves; e.R_magic = magic_of(e.R); e.C = d.q; e.C_magic = magic_of(e.C); e.q_magic = magic_of(d.q);
        e.out = guides[L - 1];
        const int row = e.nL * e.C;
        dim3 grid((unsigned)std::min((row + 255) / 256, 64), (unsigned)std::min<int64_t>(B, 65535));
        k_expand<<<grid, 256, 0, st>>>(e);
        GHM_CHECK_LAUNCH();
    }
    return GHM_OK;
}

// ---- dns guides ------------------------------------------------------------------------------------
extern "C" int64_t ghm_guides_dns_workspace_bytes(const ghm_model_t* m, int64_t B) {
    if (!m || B <= 0) return 0;
    return 3 * B * n_nodes_all(m->d) * m->d.q * (int64_t)sizeof(float);
}

extern "C" int ghm_guides_dns(const ghm_model_t* m, int64_t B, const float* z, float sigma, const float* ext,
                              float* const* guides, float* mean, void* workspace, void* stream) {
    if (!m || !z || !workspace) return ghm_fail(GHM_EINVAL, "ghm_guides_dns: null argument");
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
    int rc = dispatch_q(d.q, [&](auto Qc) -> int {
        constexpr int Q = decltype(Qc)::value;
        k_lvl_dns<Q><<<(unsigned)((B + GD_NT - 1) / GD_NT), GD_NT, 0, st>>>(d, a);
        GHM_CHECK_LAUNCH();
        return GHM_OK;
    });
    if (rc || !guides) return rc;
    // (hd|qd) depth L..1 ; root (hd|bu) ; (hd|qd|bu) depth 1..L      (reference :554-590)
    for (int j = 0; j < L; ++j) {
        rc = launch_expand(d, B, L - j, (int)nn, true, HD, QD, nullptr, 2, guides[j], st);
        if (rc) return rc;
    }
    rc = launch_expand(d, B, 0, (int)nn, true, HD, BU, nullptr, 2, guides[L], st);
    if (rc) return rc;
    for (int j = 1; j <= L; ++j) {
        rc = launch_expand(d, B, j, (int)nn, true, HD, QD, BU, 3, guides[L + j], st);
        if (rc) return rc;
    }
    return GHM_OK;
}
