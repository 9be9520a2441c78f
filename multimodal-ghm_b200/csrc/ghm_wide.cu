// ghm_wide.cu -- belief propagation for 16 < q <= 256: level-synchronous, batched row-GEMM form.
//
// Replaces GHMTree.BP_CLS (reference src/ghmclip/data/data_random_GHM.py:185-221) and GHMTree.BP_DNS
// (:467-523) when the q-vector no longer fits a thread's registers.  Here the [B*nodes, q] x [q, q]
// product of the north star is a real dense contraction:
//
//   layout    messages of one tree level live NODE-major: M[node][tree][QW] f32 (QW = q rounded up to
//             32, zero padded), so "every tree's message at node v" is one contiguous [B, QW] matrix and
//             the child->parent step `T_v @ m` (:207,497) for all trees is the GEMM  U_v = M_v * T_v^T;
//             nodes of a level are the GEMM batch (blockIdx.z), the weight is picked per node
//             (translation-invariant: (level, child index); per-edge: the edge) -- no gather, no scatter.
//   combine   parent p: h_p = prod_c U_{p*s+c} / max  -- one warp per (node, tree) row, lanes over
//             states; this is the reference's `sum_c log(.) - max` (:207-208) in the linear domain.
//   GEMM      k_wide_sgemm: FP32 CUDA cores (128x64 CTA tile, 8x4 register tile, packed f32x2 FMAs);
//             ghm_wide_tc.cu provides the tcgen05 TF32 / BF16 variant of the same contract.
//   BP_DNS    upward as above with the Gaussian leaf likelihoods (:485) as the leaf-level GEMM input,
//             then the mirrored downward pass: cavity w = b_parent / u_v (:513), GEMM with T (not
//             transposed), belief b_v = h_v * (T^T w) / max (:512-514), posterior mean at the leaves (:516-519).
// Rescales happen where the reference subtracts the max; per-node constants cancel in every normalised
// output (posterior, posterior mean); root_hd is the log of the max-rescaled root message exactly like
// the reference's shifted `hd_message`.
#include <algorithm>

#include "ghm_wide.cuh"

#define WR_NT 256                   // row kernels: 8 warps = 8 rows per CTA

// weight of the edge INTO node (level l >= 1, BFS index idx)
__device__ __forceinline__ int wide_mat(const GhmDev& d, int l, int idx) {
    return d.mat_off[l] + (d.ti ? idx - ghm_div_s(idx, d) * d.s : idx);
}

template <int NV>
__device__ __forceinline__ float warp_max_nv(const float (&v)[NV]) {
    float m = v[0];
#pragma unroll
    for (int i = 1; i < NV; ++i) m = fmaxf(m, v[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    return m;
}
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- root outputs from the rescaled root message (reference :213-217) ---------------------------
template <int NV>
__device__ __forceinline__ void wide_root_out(const GhmDev& d, const float (&v)[NV], int lane, int64_t b, float* post,
                                              float* root_hd) {
    const int q = d.q;
    if (root_hd) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int k = lane + 32 * i;
            if (k < q) root_hd[b * q + k] = logf(v[i]);
        }
    }
    if (post) {
        float w[NV], sum = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) { w[i] = v[i] * __ldg(d.py + lane + 32 * i); sum += w[i]; }
        sum = warp_sum_f(sum);
        const float inv = 1.0f / sum;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int k = lane + 32 * i;
            if (k < q) post[b * q + k] = w[i] * inv;
        }
    }
}

// ---- BP_CLS leaf level: h_p = prod_c T_c[:, x_c] / max for the depth-(L-1) nodes (:191-197) --------
template <int NV>
__global__ void __launch_bounds__(WR_NT) k_wide_leaf_cls(const GhmDev d, int64_t B, const void* leaves, int leaf_dtype,
                                                         float* H, float* post, float* root_hd) {
    const int lane = threadIdx.x & 31;
    const int QW = d.QW, L = d.L, s = d.s, nL = d.n_leaves;
    const int n1 = d.spow[L - 1];
    const int64_t row = (int64_t)blockIdx.x * (WR_NT / 32) + (threadIdx.x >> 5);
    if (row >= (int64_t)n1 * B) return;
    const int node = (int)(row / B);
    const int64_t b = row - (int64_t)node * B;
    float v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = 1.f;
    for (int c = 0; c < s; ++c) {
        const int leaf = node * s + c;
        int64_t x = leaf_dtype == GHM_LEAF_I64 ? reinterpret_cast<const int64_t*>(leaves)[b * nL + leaf]
                                               : (int64_t) reinterpret_cast<const uint8_t*>(leaves)[b * nL + leaf];
        if (x < 0 || x >= d.q) { if (lane == 0) atomicOr(d.status, 1); x = x < 0 ? 0 : d.q - 1; }
        const float* trow = d.Wdn + ((size_t)wide_mat(d, L, leaf) * QW + (int)x) * QW;      // T[:, x]
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] *= __ldg(trow + lane + 32 * i);
    }
    const float inv = __fdividef(1.f, warp_max_nv<NV>(v));
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] *= inv;
    if (L == 1) { wide_root_out<NV>(d, v, lane, b, post, root_hd); return; }
    float* out = H + row * QW;
#pragma unroll
    for (int i = 0; i < NV; ++i) out[lane + 32 * i] = v[i];
}

// ---- combine: h_p = prod_c U_{p*s+c} / max for the n_par nodes of a level ---------------------------
// U rows: (child node index within its level)*B + b; H rows: p*B + b.  ext != null (root of BP_DNS): the
// external root message is multiplied in (:505-506).  is_root: write post / root_hd instead of H when asked.
template <int NV>
__global__ void __launch_bounds__(WR_NT) k_wide_combine(const GhmDev d, int64_t B, int n_par, const float* __restrict__ U,
                                                        float* __restrict__ H, const float* __restrict__ ext, int is_root,
                                                        float* post, float* root_hd) {
    const int lane = threadIdx.x & 31;
    const int QW = d.QW, s = d.s, q = d.q;
    const int64_t row = (int64_t)blockIdx.x * (WR_NT / 32) + (threadIdx.x >> 5);
    if (row >= (int64_t)n_par * B) return;
    const int p = (int)(row / B);
    const int64_t b = row - (int64_t)p * B;
    float v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = 1.f;
    for (int c = 0; c < s; ++c) {
        const float* u = U + ((int64_t)(p * s + c) * B + b) * QW;
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] *= u[lane + 32 * i];
    }
    float inv = __fdividef(1.f, warp_max_nv<NV>(v));
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] *= inv;
    if (H) {                                                   // BP_DNS keeps hd(root) before the external message
        float* out = H + row * QW;
#pragma unroll
        for (int i = 0; i < NV; ++i) out[lane + 32 * i] = v[i];
    }
    if (is_root && (post || root_hd)) wide_root_out<NV>(d, v, lane, b, post, root_hd);
    (void)ext; (void)q;
}

// ---- BP_DNS leaf likelihoods e[k] = exp(-(z-k)^2 / 2 sigma^2) / max (:485) ---------------------------
template <int NV>
__global__ void __launch_bounds__(WR_NT) k_wide_leaf_like(const GhmDev d, int64_t B, const float* __restrict__ z, float c2,
                                                          float* __restrict__ E) {
    const int lane = threadIdx.x & 31;
    const int QW = d.QW, q = d.q, nL = d.n_leaves;
    const int64_t row = (int64_t)blockIdx.x * (WR_NT / 32) + (threadIdx.x >> 5);
    if (row >= (int64_t)nL * B) return;
    const int leaf = (int)(row / B);
    const int64_t b = row - (int64_t)leaf * B;
    const float zi = z[b * nL + leaf];
    const float kstar = fminf(fmaxf(rintf(zi), 0.f), (float)(q - 1));
    const float d0 = (zi - kstar) * (zi - kstar);
    float* out = E + row * QW;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int k = lane + 32 * i;
        const float dk = zi - (float)k;
        out[k] = k < q ? exp2f(c2 * (dk * dk - d0)) : 0.f;
    }
}

// ---- BP_DNS root belief: b_0 = h_0 * exp(ext - max) / max (:501-506) ---------------------------------
template <int NV>
__global__ void __launch_bounds__(WR_NT) k_wide_root_belief(const GhmDev d, int64_t B, const float* __restrict__ H0,
                                                            const float* __restrict__ ext, float* __restrict__ BU0,
                                                            float* __restrict__ root_bu) {
    const int lane = threadIdx.x & 31;
    const int QW = d.QW, q = d.q;
    const int64_t b = (int64_t)blockIdx.x * (WR_NT / 32) + (threadIdx.x >> 5);
    if (b >= B) return;
    float v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = H0[b * QW + lane + 32 * i];
    if (root_bu) {                                                 // root_node.hd_message after BP_DNS: log h_0 + ext (unshifted)
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int k = lane + 32 * i;
            if (k < q) root_bu[b * q + k] = __logf(v[i]) + (ext ? ext[b * q + k] : 0.f);
        }
    }
    if (ext) {
        float x[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) { const int k = lane + 32 * i; x[i] = k < q ? ext[b * q + k] : -INFINITY; }
        const float mx = warp_max_nv<NV>(x);
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] *= (lane + 32 * i < q) ? __expf(x[i] - mx) : 0.f;
        const float inv = __fdividef(1.f, warp_max_nv<NV>(v));
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] *= inv;
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) BU0[b * QW + lane + 32 * i] = v[i];
}

// ---- BP_DNS cavity: w_v = b_parent / u_v for the n nodes of a level (:513) -----------------------------
template <int NV>
__global__ void __launch_bounds__(WR_NT) k_wide_cavity(const GhmDev d, int64_t B, int n, const float* __restrict__ BUpar,
                                                       const float* __restrict__ U, float* __restrict__ Wt) {
    const int lane = threadIdx.x & 31;
    const int QW = d.QW;
    const int64_t row = (int64_t)blockIdx.x * (WR_NT / 32) + (threadIdx.x >> 5);
    if (row >= (int64_t)n * B) return;
    const int node = (int)(row / B);
    const int64_t b = row - (int64_t)node * B;
    const float* bp = BUpar + ((int64_t)ghm_div_s(node, d) * B + b) * QW;
    const float* u = U + row * QW;
    float* out = Wt + row * QW;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const float uv = u[lane + 32 * i];
        out[lane + 32 * i] = uv > 0.f ? __fdividef(bp[lane + 32 * i], uv) : 0.f;
    }
}

// ---- BP_DNS belief: b_v = h_v * tt_v / max (:512-514); leaves: h_v = e_v recomputed from z, output the mean ----
template <int NV>
__global__ void __launch_bounds__(WR_NT) k_wide_belief(const GhmDev d, int64_t B, int n, const float* __restrict__ Hlev,
                                                       const float* __restrict__ TT, float* __restrict__ BUlev,
                                                       const float* __restrict__ z, float c2, float* __restrict__ mean) {
    const int lane = threadIdx.x & 31;
    const int QW = d.QW, q = d.q, nL = d.n_leaves;
    const int64_t row = (int64_t)blockIdx.x * (WR_NT / 32) + (threadIdx.x >> 5);
    if (row >= (int64_t)n * B) return;
    const int node = (int)(row / B);
    const int64_t b = row - (int64_t)node * B;
    const float* tt = TT + row * QW;
    float v[NV];
    if (mean) {                                                // leaf level
        const float zi = z[b * nL + node];
        const float kstar = fminf(fmaxf(rintf(zi), 0.f), (float)(q - 1));
        const float d0 = (zi - kstar) * (zi - kstar);
        float num = 0.f, den = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int k = lane + 32 * i;
            const float dk = zi - (float)k;
            const float bl = k < q ? exp2f(c2 * (dk * dk - d0)) * tt[k] : 0.f;
            num = fmaf((float)k, bl, num);
            den += bl;
        }
        num = warp_sum_f(num); den = warp_sum_f(den);
        if (lane == 0) mean[b * nL + node] = num / den;
        return;
    }
    const float* h = Hlev + row * QW;
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = h[lane + 32 * i] * tt[lane + 32 * i];
    const float inv = __fdividef(1.f, warp_max_nv<NV>(v));
    float* out = BUlev + row * QW;
#pragma unroll
    for (int i = 0; i < NV; ++i) out[lane + 32 * i] = v[i] * inv;
}

// ------------------------------------------------------------------------------------------------
// FP32 batched row-GEMM on CUDA cores:  Y[node][b][n] = sum_k X[node][b][k] * W[mat(node)][k][n]
// CTA tile 128 (trees) x 64 (or 32) columns, K step 16, 256 threads, thread tile 8 x 4 (8 x 2), double-buffered smem.
// ------------------------------------------------------------------------------------------------
#define SG_BM 128
#define SG_BK 16
template <int BN>                                               // 64, or 32 when QW is an odd multiple of 32
__global__ void __launch_bounds__(256) k_wide_sgemm(const GhmDev d, int64_t B, int level, int down, const float* __restrict__ X,
                                                    float* __restrict__ Y) {
    constexpr int TN = BN / 16;                                 // columns per thread: 4 or 2
    __shared__ __align__(16) float As[2][SG_BK][SG_BM + 4];     // [k][m] (transposed on load)
    __shared__ __align__(16) float Bs[2][SG_BK][BN];            // [k][n]
    const int QW = d.QW, tid = threadIdx.x;
    const int node = blockIdx.z;
    const int64_t m0 = (int64_t)blockIdx.x * SG_BM;
    const int n0 = blockIdx.y * BN;
    const float* W = (down ? d.Wup : d.Wdn) + (size_t)wide_mat(d, level, node) * QW * QW;   // W[k][n]
    const float* Xn = X + (int64_t)node * B * QW;
    float* Yn = Y + (int64_t)node * B * QW;
    const int tm = (tid >> 4) * 8, tn = (tid & 15) * TN;
    float acc[8][TN];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
    // load mapping: A tile 128 x 16 floats = 512 float4 (2 per thread); B tile 16 x BN = 4*BN float4
    auto load_tiles = [&](int buf, int k0) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int f = tid + r * 256;            // 0..511
            const int m = f >> 2, kq = (f & 3) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m0 + m < B) v = *reinterpret_cast<const float4*>(Xn + (m0 + m) * QW + k0 + kq);
            As[buf][kq + 0][m] = v.x; As[buf][kq + 1][m] = v.y; As[buf][kq + 2][m] = v.z; As[buf][kq + 3][m] = v.w;
        }
        if (tid < 4 * BN) {
            const int k = tid / (BN / 4), nq = (tid % (BN / 4)) * 4;
            *reinterpret_cast<float4*>(&Bs[buf][k][nq]) = *reinterpret_cast<const float4*>(W + (size_t)(k0 + k) * QW + n0 + nq);
        }
    };
    load_tiles(0, 0);
    __syncthreads();
    const int nk = QW / SG_BK;
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) load_tiles(buf ^ 1, (kt + 1) * SG_BK);
#pragma unroll
        for (int k = 0; k < SG_BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][tm]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][tm + 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float bv[TN];
#pragma unroll
            for (int j = 0; j < TN; ++j) bv[j] = Bs[buf][k][tn + j];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t m = m0 + tm + i;
        if (m < B) {
            float* dst = Yn + m * QW + n0 + tn;
            if constexpr (TN == 4) *reinterpret_cast<float4*>(dst) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
            else *reinterpret_cast<float2*>(dst) = make_float2(acc[i][0], acc[i][1]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
int ghm_wide_gemm(const ghm_model* m, int64_t B, int level, int n_nodes, int down, const float* X, float* Y,
                  cudaStream_t st) {
    const GhmDev& d = m->d;
    if (m->gemm_mode != GHM_GEMM_F32) {
        int rc = ghm_wide_gemm_tc(m, B, level, n_nodes, down, X, Y, st);
        if (rc != GHM_EUNSUP) return rc;                       // shapes the tensor variant does not cover run on CUDA cores
    }
    if (n_nodes > 65535) return ghm_fail(GHM_EUNSUP, "level with more than 65535 nodes");
    if (d.QW % 64 == 0) {
        dim3 grid((unsigned)((B + SG_BM - 1) / SG_BM), (unsigned)(d.QW / 64), (unsigned)n_nodes);
        k_wide_sgemm<64><<<grid, 256, 0, st>>>(d, B, level, down, X, Y);
    } else {
        dim3 grid((unsigned)((B + SG_BM - 1) / SG_BM), (unsigned)(d.QW / 32), (unsigned)n_nodes);
        k_wide_sgemm<32><<<grid, 256, 0, st>>>(d, B, level, down, X, Y);
    }
    GHM_CHECK_LAUNCH();
    return GHM_OK;
}

template <typename F>
static int dispatch_nv(int QW, F&& f) {
    switch (QW / 32) {
        case 1: return f(std::integral_constant<int, 1>{});
        case 2: return f(std::integral_constant<int, 2>{});
        case 3: return f(std::integral_constant<int, 3>{});
        case 4: return f(std::integral_constant<int, 4>{});
        case 5: return f(std::integral_constant<int, 5>{});
        case 6: return f(std::integral_constant<int, 6>{});
        case 7: return f(std::integral_constant<int, 7>{});
        case 8: return f(std::integral_constant<int, 8>{});
        default: return ghm_fail(GHM_EUNSUP, "wide path covers q <= 256");
    }
}

static unsigned row_grid(int64_t rows) { return (unsigned)((rows + WR_NT / 32 - 1) / (WR_NT / 32)); }

int64_t ghm_wide_cls_workspace_bytes(const ghm_model* m, int64_t B) {
    const GhmDev& d = m->d;
    const int64_t n1 = d.spow[d.L > 1 ? d.L - 1 : 0];
    return 2 * n1 * B * d.QW * (int64_t)sizeof(float) + 256;
}

int ghm_wide_bp_cls(const ghm_model* m, int64_t B, const void* leaves, int leaf_dtype, float* post, float* root_hd,
                    void* workspace, cudaStream_t st) {
    const GhmDev& d = m->d;
    if (!workspace) return ghm_fail(GHM_EINVAL, "ghm_bp_cls: q=%d needs a workspace of ghm_bp_cls_workspace_bytes()", d.q);
    const int64_t n1 = d.spow[d.L - 1];
    float* H = (float*)workspace;
    float* U = H + n1 * B * d.QW;
    return dispatch_nv(d.QW, [&](auto nvc) -> int {
        constexpr int NV = decltype(nvc)::value;
        // TF32: the depth-(L-1) messages are formed inside the first GEMM (ghm_wide_tc.cu, LF_CLS)
        int fused = d.L >= 2 ? ghm_wide_cls_leaf_fused(m, B, leaves, leaf_dtype, U, st) : GHM_EUNSUP;
        if (fused != GHM_OK && fused != GHM_EUNSUP) return fused;
        if (fused == GHM_EUNSUP) {
            k_wide_leaf_cls<NV><<<row_grid(n1 * B), WR_NT, 0, st>>>(d, B, leaves, leaf_dtype, H, post, root_hd);
            GHM_CHECK_LAUNCH();
        }
        for (int l = d.L - 1; l >= 1; --l) {                   // H holds the depth-l messages
            const int n = d.spow[l];
            int rc = (fused == GHM_OK && l == d.L - 1) ? GHM_OK : ghm_wide_gemm(m, B, l, n, 0, H, U, st);
            if (rc) return rc;
            const int np = d.spow[l - 1];
            k_wide_combine<NV><<<row_grid((int64_t)np * B), WR_NT, 0, st>>>(d, B, np, U, l == 1 ? nullptr : H, nullptr,
                                                                             l == 1, post, root_hd);
            GHM_CHECK_LAUNCH();
        }
        return GHM_OK;
    });
}

// BP_DNS workspace: Hd [N_int][B][QW] | Ud [1 + E][B][QW] (slot 0 unused) | BU [N_int][B][QW] | tmpA, tmpB [n_L][B][QW]
static int64_t n_internal(const GhmDev& d) { return 1 + (int64_t)d.edge_off[d.L]; }
int64_t ghm_wide_dns_workspace_bytes(const ghm_model* m, int64_t B) {
    const GhmDev& d = m->d;
    const int64_t rows = 2 * n_internal(d) + (1 + (int64_t)d.n_edges) + 2 * (int64_t)d.n_leaves;
    return rows * B * d.QW * (int64_t)sizeof(float) + 256;
}

int ghm_wide_bp_dns(const ghm_model* m, int64_t B, const float* z, float sigma, const float* ext, float* mean, float* root_bu,
                    void* workspace, cudaStream_t st) {
    const GhmDev& d = m->d;
    const int L = d.L, nL = d.n_leaves;
    const int64_t RW = B * d.QW;                               // floats per node
    const int64_t Nint = n_internal(d);
    float* Hd = (float*)workspace;                             // node id of (depth l, idx) = off(l) + idx, off(0) = 0
    float* Ud = Hd + Nint * RW;
    float* BU = Ud + (1 + (int64_t)d.n_edges) * RW;
    float* tA = BU + Nint * RW;
    float* tB = tA + (int64_t)nL * RW;
    auto off = [&](int l) -> int64_t { return l == 0 ? 0 : 1 + (int64_t)d.edge_off[l]; };
    const float c2 = -0.5f * 1.4426950408889634f / (sigma * sigma);
    return dispatch_nv(d.QW, [&](auto nvc) -> int {
        constexpr int NV = decltype(nvc)::value;
        int rc;
        // ---- upward ----
        rc = ghm_wide_leaf_up_fused(m, B, z, c2, Ud + off(L) * RW, st);      // TF32: likelihoods generated inside the GEMM
        if (rc == GHM_EUNSUP) {
            k_wide_leaf_like<NV><<<row_grid((int64_t)nL * B), WR_NT, 0, st>>>(d, B, z, c2, tA);
            GHM_CHECK_LAUNCH();
            rc = ghm_wide_gemm(m, B, L, nL, 0, tA, Ud + off(L) * RW, st);
        }
        if (rc) return rc;
        for (int l = L - 1; l >= 0; --l) {
            const int n = d.spow[l];
            k_wide_combine<NV><<<row_grid((int64_t)n * B), WR_NT, 0, st>>>(d, B, n, Ud + off(l + 1) * RW, Hd + off(l) * RW,
                                                                            nullptr, 0, nullptr, nullptr);
            GHM_CHECK_LAUNCH();
            if (l >= 1 && (rc = ghm_wide_gemm(m, B, l, n, 0, Hd + off(l) * RW, Ud + off(l) * RW, st))) return rc;
        }
        k_wide_root_belief<NV><<<row_grid(B), WR_NT, 0, st>>>(d, B, Hd, ext, BU, root_bu);
        GHM_CHECK_LAUNCH();
        // ---- downward ----
        for (int l = 1; l <= L; ++l) {
            const int n = d.spow[l];
            if (l == L) {                                     // TF32: cavity + GEMM + posterior mean in one kernel
                rc = ghm_wide_leaf_down_fused(m, B, z, c2, BU + off(L - 1) * RW, Ud + off(L) * RW, mean, st);
                if (rc != GHM_EUNSUP) return rc;
            } else {                                          // TF32: cavity + GEMM + belief in one kernel
                rc = ghm_wide_down_fused(m, B, l, BU + off(l - 1) * RW, Ud + off(l) * RW, Hd + off(l) * RW, BU + off(l) * RW, st);
                if (rc == GHM_OK) continue;
                if (rc != GHM_EUNSUP) return rc;
            }
            k_wide_cavity<NV><<<row_grid((int64_t)n * B), WR_NT, 0, st>>>(d, B, n, BU + off(l - 1) * RW, Ud + off(l) * RW, tA);
            GHM_CHECK_LAUNCH();
            if ((rc = ghm_wide_gemm(m, B, l, n, 1, tA, tB, st))) return rc;
            k_wide_belief<NV><<<row_grid((int64_t)n * B), WR_NT, 0, st>>>(d, B, n, l < L ? Hd + off(l) * RW : nullptr, tB,
                                                                           l < L ? BU + off(l) * RW : nullptr, z, c2,
                                                                           l == L ? mean : nullptr);
            GHM_CHECK_LAUNCH();
        }
        return GHM_OK;
    });
}
