// ghm_risk.cu -- K6: risk reductions, and Gaussian observation noise.
//
// Replaces PPCLIPLoss / ClipSampler.get_Bayes / clip_loss_compute (reference
// src/ghmclip/data/data_random_GHM.py:13-41, :794-817, :819-844), the loss tails of
// ConditionalDenoiseSampler.get_Bayes (:886-894), NextWordPredictSampler.get_Bayes (:931-942),
// ClassificationSampler.get_Bayes (:707-720) and the noise draw of :733 / :867.
//
// Every kernel reduces per-item losses to {sum, sum of squares, count} in float64 and adds
// them atomically into a 3-double device buffer: that buffer is what crosses NVLink (one
// all-reduce of 24 bytes per risk evaluation) when the batch is sharded over GPUs.  The
// reference's dense ((K-1)n x n) 0/1 matmul (:26-27, O(n^2) memory) is a K-1 term segmented sum here.
#include <algorithm>

#include "ghm_vec.cuh"

#define RISK_NT 256

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-reduce (sum, sumsq, count) and add to sums[0..2]
__device__ __forceinline__ void block_accumulate(double s1, double s2, double cnt, double* sums) {
    __shared__ double sh[3][RISK_NT / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    s1 = warp_sum(s1); s2 = warp_sum(s2); cnt = warp_sum(cnt);
    if (lane == 0) { sh[0][warp] = s1; sh[1][warp] = s2; sh[2][warp] = cnt; }
    __syncthreads();
    if (warp == 0) {
        s1 = lane < RISK_NT / 32 ? sh[0][lane] : 0.0;
        s2 = lane < RISK_NT / 32 ? sh[1][lane] : 0.0;
        cnt = lane < RISK_NT / 32 ? sh[2][lane] : 0.0;
        s1 = warp_sum(s1); s2 = warp_sum(s2); cnt = warp_sum(cnt);
        if (lane == 0 && cnt > 0.0) {
            atomicAdd(sums + 0, s1);
            atomicAdd(sums + 1, s2);
            atomicAdd(sums + 2, cnt);
        }
    }
}

// ---- CLIP: symmetric K-way contrastive Bayes loss ------------------------------------------
__global__ void __launch_bounds__(RISK_NT) k_risk_clip(const float* __restrict__ t_pp, const float* __restrict__ i_pp,
                                                       int64_t n, int K, int q, int64_t lo, int64_t hi, double* sums) {
    const int64_t i = lo + (int64_t)blockIdx.x * RISK_NT + threadIdx.x;
    double loss = 0.0, cnt = 0.0;
    if (i < hi) {
        auto dot = [&](const float* a, const float* b) {
            double acc = 0.0;
            for (int y = 0; y < q; ++y) acc += (double)a[y] * (double)b[y];
            return acc * (double)q;
        };
        // direction 1: text negatives against image match i      (:19-28)
        const float* tm = t_pp + i * q;
        const float* im = i_pp + i * q;
        double sm = dot(tm, im), sn = 0.0;
        for (int k = 0; k < K - 1; ++k) sn += dot(t_pp + (2 * n + (int64_t)k * n + i) * q, im);
        loss = -log(sm / (sn + sm));
        // direction 2: image negatives against text match n+i    (:31-39)
        tm = t_pp + (n + i) * q;
        im = i_pp + (n + i) * q;
        sm = dot(tm, im); sn = 0.0;
        for (int k = 0; k < K - 1; ++k) sn += dot(i_pp + (2 * n + (int64_t)k * n + i) * q, tm);
        loss += -log(sm / (sn + sm));
        cnt = 1.0;
    }
    block_accumulate(loss, loss * loss, cnt, sums);
}

extern "C" int ghm_risk_clip(const float* t_pp, const float* i_pp, int64_t n, int K, int q, int64_t pair_lo,
                             int64_t pair_hi, double* sums, void* stream) {
    if (!t_pp || !i_pp || !sums) return ghm_fail(GHM_EINVAL, "ghm_risk_clip: null argument");
    if (K < 2 || q < 2 || n <= 0 || pair_lo < 0 || pair_hi > n || pair_lo > pair_hi)
        return ghm_fail(GHM_EINVAL, "ghm_risk_clip: bad sizes n=%lld K=%d q=%d pairs=[%lld,%lld)", (long long)n, K, q,
                        (long long)pair_lo, (long long)pair_hi);
    const int64_t cnt = pair_hi - pair_lo;
    if (cnt == 0) return GHM_OK;
    GhmDeviceGuard guard(sums);
    k_risk_clip<<<(unsigned)((cnt + RISK_NT - 1) / RISK_NT), RISK_NT, 0, (cudaStream_t)stream>>>(t_pp, i_pp, n, K, q,
                                                                                                 pair_lo, pair_hi, sums);
    GHM_CHECK_LAUNCH();
    return GHM_OK;
}

// ---- CDM: per-tree squared error of the posterior mean -------------------------------------
template <typename LeafT>
__global__ void __launch_bounds__(RISK_NT) k_risk_cdm(const float* __restrict__ mean, const LeafT* __restrict__ leaves,
                                                      int64_t B, int64_t nL, double* sums) {
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = ((int64_t)blockIdx.x * RISK_NT + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * RISK_NT) >> 5;
    double s1 = 0.0, s2 = 0.0, cnt = 0.0;
    for (int64_t b = warp_global; b < B; b += n_warps) {
        double acc = 0.0;
        for (int64_t i = lane; i < nL; i += 32) {
            const double dlt = (double)mean[b * nL + i] - (double)leaves[b * nL + i];
            acc += dlt * dlt;
        }
        acc = warp_sum(acc);
        if (lane == 0) { s1 += acc; s2 += acc * acc; cnt += 1.0; }
    }
    block_accumulate(s1, s2, cnt, sums);
}

extern "C" int ghm_risk_cdm(const float* mean, const void* leaves, int leaf_dtype, int64_t B, int64_t nL, double* sums,
                            void* stream) {
    if (!mean || !leaves || !sums) return ghm_fail(GHM_EINVAL, "ghm_risk_cdm: null argument");
    if (B <= 0) return GHM_OK;
    GhmDeviceGuard guard(sums);
    const int64_t warps = B;
    unsigned grid = (unsigned)std::min<int64_t>((warps * 32 + RISK_NT - 1) / RISK_NT, 148 * 8);
    if (leaf_dtype == GHM_LEAF_I64)
        k_risk_cdm<int64_t><<<grid, RISK_NT, 0, (cudaStream_t)stream>>>(mean, (const int64_t*)leaves, B, nL, sums);
    else if (leaf_dtype == GHM_LEAF_U8)
        k_risk_cdm<uint8_t><<<grid, RISK_NT, 0, (cudaStream_t)stream>>>(mean, (const uint8_t*)leaves, B, nL, sums);
    else
        return ghm_fail(GHM_EINVAL, "bad leaf_dtype %d", leaf_dtype);
    GHM_CHECK_LAUNCH();
    return GHM_OK;
}

// ---- cross-entropy of a categorical posterior against integer targets -----------------------
template <typename LeafT>
__global__ void __launch_bounds__(RISK_NT) k_risk_ce(const float* __restrict__ pp, const LeafT* __restrict__ tgt,
                                                     int64_t rows, int q, int64_t t_stride, int64_t t_off,
                                                     int64_t row_group, double* sums) {
    double s1 = 0.0, s2 = 0.0, cnt = 0.0;
    for (int64_t r = (int64_t)blockIdx.x * RISK_NT + threadIdx.x; r < rows; r += (int64_t)gridDim.x * RISK_NT) {
        const int64_t b = r / row_group, t = r - b * row_group;
        int64_t y = (int64_t)tgt[b * t_stride + t_off + t];
        y = y < 0 ? 0 : (y >= q ? q - 1 : y);
        const double l = -(double)logf(pp[r * q + y]);     // float32 log like the reference's torch path
        s1 += l; s2 += l * l; cnt += 1.0;
    }
    block_accumulate(s1, s2, cnt, sums);
}

extern "C" int ghm_risk_ce(const float* pp, const void* target, int leaf_dtype, int64_t rows, int q, int64_t t_stride,
                           int64_t t_off, int64_t row_group, double* sums, void* stream) {
    if (!pp || !target || !sums) return ghm_fail(GHM_EINVAL, "ghm_risk_ce: null argument");
    if (rows <= 0) return GHM_OK;
    if (row_group <= 0) return ghm_fail(GHM_EINVAL, "ghm_risk_ce: row_group must be positive");
    GhmDeviceGuard guard(sums);
    unsigned grid = (unsigned)std::min<int64_t>((rows + RISK_NT - 1) / RISK_NT, 148 * 8);
    if (leaf_dtype == GHM_LEAF_I64)
        k_risk_ce<int64_t><<<grid, RISK_NT, 0, (cudaStream_t)stream>>>(pp, (const int64_t*)target, rows, q, t_stride,
                                                                       t_off, row_group, sums);
    else if (leaf_dtype == GHM_LEAF_U8)
        k_risk_ce<uint8_t><<<grid, RISK_NT, 0, (cudaStream_t)stream>>>(pp, (const uint8_t*)target, rows, q, t_stride,
                                                                       t_off, row_group, sums);
    else
        return ghm_fail(GHM_EINVAL, "bad leaf_dtype %d", leaf_dtype);
    GHM_CHECK_LAUNCH();
    return GHM_OK;
}

// ---- zero-shot classification risk (figures/eval-zsc-risk.py:66-83, eval-zsc-ood.py:23-28) ---------------
// The image-side root posterior is pushed down the LEFTMOST path of the text tree, x <- x @ T_l[0] for every level,
// which is the Bayes predictive distribution of the first text leaf given the image; loss = -log x[first text leaf].
template <int Q, typename LeafT>
__global__ void __launch_bounds__(RISK_NT) k_risk_zsc(const GhmDev d, const float* __restrict__ i_pp,
                                                      const LeafT* __restrict__ t_leaves, int64_t B, double* sums) {
    const int q = d.q;
    double s1 = 0.0, s2 = 0.0, cnt = 0.0;
    for (int64_t b = (int64_t)blockIdx.x * RISK_NT + threadIdx.x; b < B; b += (int64_t)gridDim.x * RISK_NT) {
        float v[Q], w[Q];
#pragma unroll
        for (int k = 0; k < Q; ++k) v[k] = (k < q) ? i_pp[b * q + k] : 0.f;
        for (int l = 1; l <= d.L; ++l) {
            ghm_matvec_t<Q>(d.Tlin + (size_t)d.mat_off[l] * Q * Q, v, w);       // edge into node 0 of depth l
#pragma unroll
            for (int k = 0; k < Q; ++k) v[k] = w[k];
        }
        int64_t y = (int64_t)t_leaves[b * d.n_leaves];
        y = y < 0 ? 0 : (y >= q ? q - 1 : y);
        float vy = 0.f;
#pragma unroll
        for (int k = 0; k < Q; ++k) vy = (k == (int)y) ? v[k] : vy;
        const double l = -(double)logf(vy);                  // float32 log like the reference's torch path
        s1 += l; s2 += l * l; cnt += 1.0;
    }
    block_accumulate(s1, s2, cnt, sums);
}

extern "C" int ghm_risk_zsc(const ghm_model_t* text, int64_t B, const float* i_pp, const void* t_leaves, int leaf_dtype,
                            double* sums, void* stream) {
    if (!text || !i_pp || !t_leaves || !sums) return ghm_fail(GHM_EINVAL, "ghm_risk_zsc: null argument");
    if (B <= 0) return GHM_OK;
    if (leaf_dtype != GHM_LEAF_I64 && leaf_dtype != GHM_LEAF_U8) return ghm_fail(GHM_EINVAL, "bad leaf_dtype %d", leaf_dtype);
    const GhmDev& d = text->d;
    GhmDeviceGuard guard(text->device);
    const unsigned grid = (unsigned)std::min<int64_t>((B + RISK_NT - 1) / RISK_NT, 148 * 8);
    cudaStream_t st = (cudaStream_t)stream;
#define ZSC_GO(Q)                                                                                           \
    do {                                                                                                    \
        if (leaf_dtype == GHM_LEAF_I64)                                                                     \
            k_risk_zsc<Q, int64_t><<<grid, RISK_NT, 0, st>>>(d, i_pp, (const int64_t*)t_leaves, B, sums);   \
        else                                                                                                \
            k_risk_zsc<Q, uint8_t><<<grid, RISK_NT, 0, st>>>(d, i_pp, (const uint8_t*)t_leaves, B, sums);   \
    } while (0)
    switch (ghm_pad_q(d.q)) {
        case 4: ZSC_GO(4); break;
        case 8: ZSC_GO(8); break;
        case 10: ZSC_GO(10); break;
        case 16: ZSC_GO(16); break;
        default: return ghm_fail(GHM_EUNSUP, "ghm_risk_zsc covers q <= %d in this build", GHM_MAX_Q_REG);
    }
#undef ZSC_GO
    GHM_CHECK_LAUNCH();
    return GHM_OK;
}

// ---- z = x + sigma * N(0,1)  (Philox stream 1, Box-Muller; two words per leaf) ---------------
template <typename LeafT>
__global__ void __launch_bounds__(RISK_NT) k_gauss_noise(const LeafT* __restrict__ leaves, int64_t B, int nL,
                                                         float sigma, uint64_t seed, uint64_t tree_offset,
                                                         float* __restrict__ z) {
    // one thread per PAIR of leaves (one Philox block = 4 words = 2 normals)
    const int64_t pairs_per_tree = (nL + 1) / 2;
    const int64_t total = B * pairs_per_tree;
    for (int64_t p = (int64_t)blockIdx.x * RISK_NT + threadIdx.x; p < total; p += (int64_t)gridDim.x * RISK_NT) {
        const int64_t b = p / pairs_per_tree;
        const int pi = (int)(p - b * pairs_per_tree);
        const uint4 w = ghm_rng_block(seed, tree_offset + (uint64_t)b, 0u, (uint32_t)pi, GHM_STREAM_NOISE);
        const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int leaf = 2 * pi + h;
            if (leaf < nL) {
                const float u1 = ((float)ws[2 * h] + 0.5f) * 2.3283064365386963e-10f;      // all 32 bits: tails to 6.7 sigma
                const float u2 = ((float)(ws[2 * h + 1] >> 8) + 0.5f) * 5.9604644775390625e-08f;
                const float g = sqrtf(-2.0f * logf(u1)) * cosf(6.2831853071795864f * u2);
                z[b * nL + leaf] = (float)leaves[b * nL + leaf] + sigma * g;
            }
        }
    }
}

extern "C" int ghm_gauss_noise(const ghm_model_t* m, int64_t B, const void* leaves, int leaf_dtype, float sigma,
                               uint64_t seed, uint64_t tree_offset, float* z, void* stream) {
    if (!m || !leaves || !z) return ghm_fail(GHM_EINVAL, "ghm_gauss_noise: null argument");
    if (B <= 0) return GHM_OK;
    GhmDeviceGuard guard(m->device);
    const int nL = m->d.n_leaves;
    const int64_t total = B * ((nL + 1) / 2);
    unsigned grid = (unsigned)std::min<int64_t>((total + RISK_NT - 1) / RISK_NT, 148 * 16);
    if (leaf_dtype == GHM_LEAF_I64)
        k_gauss_noise<int64_t><<<grid, RISK_NT, 0, (cudaStream_t)stream>>>((const int64_t*)leaves, B, nL, sigma, seed,
                                                                           tree_offset, z);
    else if (leaf_dtype == GHM_LEAF_U8)
        k_gauss_noise<uint8_t><<<grid, RISK_NT, 0, (cudaStream_t)stream>>>((const uint8_t*)leaves, B, nL, sigma, seed,
                                                                           tree_offset, z);
    else
        return ghm_fail(GHM_EINVAL, "bad leaf_dtype %d", leaf_dtype);
    GHM_CHECK_LAUNCH();
    return GHM_OK;
}
