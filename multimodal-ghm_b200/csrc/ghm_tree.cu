// ghm_tree.cu -- C entry points of K1 (sampler) and K2 (root-posterior BP) + the parity-mode sampler kernel.
//
// Replaces GHMTree.gen_values (src/ghmclip/data/data_random_GHM.py:145-165) and GHMTree.BP_CLS (:185-221) of the
// reference.  The Philox sampler / BP kernel template is in ghm_tree_kernel.cuh (design notes there).
#include "ghm_tree_kernel.cuh"
#include "ghm_wide.cuh"

// ------------------------------------------------------------------------------------------------
// k_sample_parity: reference uniforms, f64 compare, one thread per tree, level by level down each
// root path (depth-first so only the L states of the current path are live).
// ------------------------------------------------------------------------------------------------
#define PAR_NT 128
__global__ void __launch_bounds__(PAR_NT) k_sample_parity(const GhmDev d, const TreeArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x;
    const int L = d.L, s = d.s, q = d.q, nL = d.n_leaves;
    const int64_t b = (int64_t)blockIdx.x * PAR_NT + tid;
    if (b >= a.B) return;
    int* val = reinterpret_cast<int*>(smem);                    // [L][NT]
    int64_t r = a.root_in[b];
    if (r < 0 || r >= q) { atomicOr(d.status, 1); r = r < 0 ? 0 : q - 1; }
    val[tid] = (int)r;
    if (a.root_out) a.root_out[b] = r;
    const int n1 = d.spow[L - 1];
    for (int j = 0; j < n1; ++j) {
        int tz = 0, t = j;
        while (tz < L - 1) {
            const int tq = ghm_div_s(t, d);
            if (t - tq * s != 0) break;
            t = tq; ++tz;
        }
        for (int l = max(1, L - 1 - tz); l < L; ++l) {
            const int idx = ghm_div_pow(j, L - 1 - l, d);
            const int pidx = ghm_div_s(idx, d);
            const int c = idx - pidx * s;
            const int mi = d.mat_off[l] + (d.ti ? c : idx);
            const int xp = val[(l - 1) * PAR_NT + tid];
            const double u = a.U[(size_t)(d.edge_off[l] + idx) * a.B + b];
            val[l * PAR_NT + tid] = ghm_search_f64(d.cdfd + ((size_t)mi * q + xp) * q, u, q);
        }
        const int xj = val[(L - 1) * PAR_NT + tid];
        for (int c = 0; c < s; ++c) {
            const int lidx = j * s + c;
            const int mi = d.mat_off[L] + (d.ti ? c : lidx);
            const double u = a.U[(size_t)(d.edge_off[L] + lidx) * a.B + b];
            const int x = ghm_search_f64(d.cdfd + ((size_t)mi * q + xj) * q, u, q);
            if (a.leaves) {
                if (a.leaf_dtype == GHM_LEAF_I64) reinterpret_cast<int64_t*>(a.leaves)[b * nL + lidx] = x;
                else reinterpret_cast<uint8_t*>(a.leaves)[b * nL + lidx] = (uint8_t)x;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// k_build_leaf_memo: one thread per (child number cj, leaf states x_0 .. x_{s-1}) tabulates the message a
// depth-(L-1) node sends to its parent (ghm_common.cuh, GhmDev::leaf_memo).  The row is produced with the
// operations, in the order, of the per-node code it replaces in k_tree_fast / k_tree2 (product of the leaf rows
// c = 0 .. s-1, max-rescale through the approximate reciprocal, packed FMA matvec over b = 0 .. Q-1), so the
// memoised kernel returns bit-identical posteriors.  Reference: GHMTree.BP_CLS :191-208.
// ------------------------------------------------------------------------------------------------
template <int Q>
__global__ void __launch_bounds__(128) k_build_leaf_memo(const GhmDev d, float* __restrict__ out, int rows) {
    // one thread per (row, state pair): the rescaled leaf-row product is recomputed by the Q/2 threads of a row (cheap),
    // each then forms its own pair of the matvec -- independent loads, two dependent global round trips in all
    constexpr int H = Q / 2, QS = (Q + 3) / 4 * 4;
    const int idx = blockIdx.x * 128 + threadIdx.x;
    const int r = idx / H, i = idx - r * H;
    if (r >= rows) return;
    const int q = d.q, s = d.s, L = d.L;
    int xs[4] = {0, 0, 0, 0}, rem = r;                           // row = ((cj*q + x_0)*q + x_1)*q + ...
#pragma unroll
    for (int c = 3; c >= 0; --c)
        if (c < s) { xs[c] = rem % q; rem /= q; }
    const int cj = rem;
    f2 h[H];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        if (c < s) {
            f2 row[H];
            f2_load_row<Q>(d.TTp + ((size_t)(d.mat_off[L] + c) * Q + xs[c]) * QS, row);
#pragma unroll
            for (int k = 0; k < H; ++k) h[k] = c == 0 ? row[k] : f2_mul(h[k], row[k]);
        }
    }
    f2_normalize<Q>(h);
    const float* TT = d.TTp + (size_t)(d.mat_off[L - 1] + cj) * Q * QS + 2 * i;
    f2 u = make_float2(0.f, 0.f);
#pragma unroll
    for (int b = 0; b < Q; ++b) {                                // the pair-i column of f2_matvec_up1, same operation order
        const f2 t = *reinterpret_cast<const f2*>(TT + b * QS);
        const float xb = f2_elem<Q>(h, b);
        u = b == 0 ? f2_muls(t, xb) : f2_fmas(t, xb, u);
    }
    float* o = out + (size_t)r * QS;
    *reinterpret_cast<f2*>(o + 2 * i) = u;
    if (i == 0) {
#pragma unroll
        for (int k = Q; k < QS; ++k) o[k] = 0.f;
    }
}

int ghm_build_leaf_memo(const ghm_model* m, cudaStream_t st) {
    const GhmDev& d = m->d;
    const size_t rows = ghm_leaf_memo_rows(d);
    if (!rows || !d.leaf_memo) return GHM_OK;
    float* out = const_cast<float*>(d.leaf_memo);
    const unsigned grid = (unsigned)((rows * (size_t)(d.QP / 2) + 127) / 128);
    switch (d.QP) {
        case 4: k_build_leaf_memo<4><<<grid, 128, 0, st>>>(d, out, (int)rows); break;
        case 8: k_build_leaf_memo<8><<<grid, 128, 0, st>>>(d, out, (int)rows); break;
        case 10: k_build_leaf_memo<10><<<grid, 128, 0, st>>>(d, out, (int)rows); break;
        case 16: k_build_leaf_memo<16><<<grid, 128, 0, st>>>(d, out, (int)rows); break;
        default: return ghm_fail(GHM_EUNSUP, "internal: leaf memo for padded q = %d", d.QP);
    }
    GHM_CHECK_LAUNCH();
    return GHM_OK;
}

#ifdef GHM_TREE_PROBE   // development aid: compile one instantiation only (nvcc -DGHM_TREE_PROBE=2 -cubin)
#ifndef GHM_PROBE_Q
#define GHM_PROBE_Q 10
#define GHM_PROBE_S 3
#endif
template __global__ void k_tree_fast<GHM_PROBE_Q, GHM_PROBE_S, MODE_PHILOX, true, 6144, false>(
    const __grid_constant__ GhmDev, const __grid_constant__ TreeArgs, const __grid_constant__ TabParam<6144>);
#endif

#ifndef GHM_TREE_PROBE
#define GHM_TREE_DECLARE_ALL(Q) GHM_TREE_DECLARE(Q, s) GHM_TREE_DECLARE(Q, sb) GHM_TREE_DECLARE(Q, g)
GHM_TREE_DECLARE_ALL(4)
GHM_TREE_DECLARE_ALL(8)
GHM_TREE_DECLARE_ALL(10)
GHM_TREE_DECLARE_ALL(16)

// TAG: s = Philox sampling only, sb = Philox sampling + fused BP, g = BP on given leaves
#define GHM_TREE_SWITCH(TAG)                                                                                     \
    switch (ghm_pad_q(m->d.q)) {                                                                                 \
        case 4: return ghm_tree_run_q4_##TAG(m, a, st);                                                           \
        case 8: return ghm_tree_run_q8_##TAG(m, a, st);                                                           \
        case 10: return ghm_tree_run_q10_##TAG(m, a, st);                                                         \
        case 16: return ghm_tree_run_q16_##TAG(m, a, st);                                                         \
        default:                                                                                                 \
            return ghm_fail(GHM_EUNSUP, "variable_type=%d: register-resident kernels cover q <= %d in this build", \
                            m->d.q, GHM_MAX_Q_REG);                                                               \
    }
static int run_sample(const ghm_model* m, const TreeArgs& a, cudaStream_t st) {
    if (m->d.q > GHM_MAX_Q_REG) return ghm_tree_run_q4_s(m, a, st);   // sampling never touches a q-vector: any instantiation serves
    GHM_TREE_SWITCH(s)
}
static int run_sample_bp(const ghm_model* m, const TreeArgs& a, cudaStream_t st) { GHM_TREE_SWITCH(sb) }
static int run_given_bp(const ghm_model* m, const TreeArgs& a, cudaStream_t st) { GHM_TREE_SWITCH(g) }

struct DeviceGuard {
    int prev;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

static int sample_common(const ghm_model_t* m, int64_t B, int root_mode, int64_t n_given, const int64_t* root_in,
                         uint64_t root_seed, const double* U, uint64_t seed, uint64_t tree_offset, int64_t* root_out, void* leaves_out,
                         int leaf_dtype, float* post_out, float* root_hd_out, void* stream, int64_t blk_len = 0,
                         int64_t blk_stride = 0) {
    if (!m) return ghm_fail(GHM_EINVAL, "ghm_sample: null model");
    if (B <= 0) return B == 0 ? GHM_OK : ghm_fail(GHM_EINVAL, "ghm_sample: negative batch");
    if (root_mode < 0 || root_mode > 3) return ghm_fail(GHM_EINVAL, "ghm_sample: bad root_mode %d", root_mode);
    if (root_mode == GHM_ROOT_GIVEN && n_given > 0 && !root_in) return ghm_fail(GHM_EINVAL, "ghm_sample: root_in is null");
    if (n_given < 0 || n_given > B) return ghm_fail(GHM_EINVAL, "ghm_sample: n_given outside [0, B]");
    if (U && (root_mode != GHM_ROOT_GIVEN || n_given != B))
        return ghm_fail(GHM_EINVAL, "ghm_sample: parity mode (U given) needs host-drawn roots (GHM_ROOT_GIVEN)");
    if (leaf_dtype != GHM_LEAF_I64 && leaf_dtype != GHM_LEAF_U8) return ghm_fail(GHM_EINVAL, "bad leaf_dtype %d", leaf_dtype);
    DeviceGuard g(m->device);
    TreeArgs a{};
    a.B = B; a.root_mode = root_mode; a.n_given = n_given; a.root_in = root_in; a.root_seed = root_seed; a.U = U;
    a.seed = seed;
    a.tree_offset = tree_offset;
    if (blk_len > 0 && (B >= (1ll << 32) || blk_len >= (1ll << 32)))
        return ghm_fail(GHM_EUNSUP, "ghm_sample_blocked: batches of 2^32 trees or more are not supported");
    a.blk_len = (uint32_t)blk_len; a.blk_extra = (uint64_t)(blk_stride - blk_len);
    a.root_out = root_out; a.leaves = leaves_out; a.leaf_dtype = leaf_dtype; a.post = post_out; a.root_hd = root_hd_out;
    const bool bp = post_out || root_hd_out;
    if (bp && m->d.QW)
        return ghm_fail(GHM_EUNSUP, "ghm_sample: fused BP covers q <= %d; for q = %d sample, then call ghm_bp_cls", GHM_MAX_Q_REG,
                        m->d.q);
    cudaStream_t st = (cudaStream_t)stream;
    if (U) {
        if (bp) return ghm_fail(GHM_EINVAL, "ghm_sample: fused BP is a Philox-mode feature; in parity mode call ghm_bp_cls");
        const size_t dyn = (size_t)m->d.L * PAR_NT * sizeof(int);
        k_sample_parity<<<(unsigned)((B + PAR_NT - 1) / PAR_NT), PAR_NT, dyn, st>>>(m->d, a);
        GHM_CHECK_LAUNCH();
        return GHM_OK;
    }
    return bp ? run_sample_bp(m, a, st) : run_sample(m, a, st);
}

extern "C" int ghm_sample(const ghm_model_t* m, int64_t B, int root_mode, const int64_t* root_in, const double* U,
                          uint64_t seed, uint64_t tree_offset, int64_t* root_out, void* leaves_out, int leaf_dtype,
                          float* post_out, float* root_hd_out, void* stream) {
    if (root_mode == GHM_ROOT_SHARED) return ghm_fail(GHM_EINVAL, "ghm_sample: GHM_ROOT_SHARED is ghm_sample_paired's mode");
    return sample_common(m, B, root_mode, root_mode == GHM_ROOT_GIVEN ? B : 0, root_in, 0, U, seed, tree_offset, root_out,
                         leaves_out, leaf_dtype, post_out, root_hd_out, stream);
}

extern "C" int ghm_sample_mixed(const ghm_model_t* m, int64_t B, int64_t n_given, const int64_t* root_in, uint64_t seed,
                                uint64_t tree_offset, int64_t* root_out, void* leaves_out, int leaf_dtype,
                                float* post_out, float* root_hd_out, void* stream) {
    return sample_common(m, B, GHM_ROOT_GIVEN, n_given, root_in, 0, nullptr, seed, tree_offset, root_out, leaves_out,
                         leaf_dtype, post_out, root_hd_out, stream);
}

extern "C" int ghm_sample_paired(const ghm_model_t* m, int64_t B, int64_t n_shared, uint64_t root_seed, uint64_t seed,
                                 uint64_t tree_offset, int64_t* root_out, void* leaves_out, int leaf_dtype,
                                 float* post_out, float* root_hd_out, void* stream) {
    return sample_common(m, B, GHM_ROOT_SHARED, n_shared, nullptr, root_seed, nullptr, seed, tree_offset, root_out,
                         leaves_out, leaf_dtype, post_out, root_hd_out, stream);
}

extern "C" int ghm_sample_blocked(const ghm_model_t* m, int64_t B, int64_t blk_len, int64_t blk_stride, int root_mode,
                                  int64_t n_given, const int64_t* root_in, uint64_t root_seed, uint64_t seed, uint64_t tree_offset,
                                  int64_t* root_out, void* leaves_out, int leaf_dtype, float* post_out, float* root_hd_out,
                                  void* stream) {
    if (blk_len <= 0 || blk_stride < blk_len) return ghm_fail(GHM_EINVAL, "ghm_sample_blocked: need 0 < blk_len <= blk_stride");
    if (root_mode == GHM_ROOT_GIVEN && n_given != B && n_given != 0)
        return ghm_fail(GHM_EINVAL, "ghm_sample_blocked: GHM_ROOT_GIVEN takes all roots (n_given = B) or none");
    return sample_common(m, B, root_mode, root_mode == GHM_ROOT_GIVEN ? B : n_given, root_in, root_seed, nullptr, seed, tree_offset,
                         root_out, leaves_out, leaf_dtype, post_out, root_hd_out, stream, blk_len, blk_stride);
}

extern "C" int64_t ghm_bp_cls_workspace_bytes(const ghm_model_t* m, int64_t B) {
    if (!m || B <= 0 || m->d.QW == 0) return 0;
    return ghm_wide_cls_workspace_bytes(m, B);
}

extern "C" int ghm_bp_cls(const ghm_model_t* m, int64_t B, const void* leaves, int leaf_dtype, float* post,
                          float* root_hd, void* workspace, void* stream) {
    if (!m || !leaves) return ghm_fail(GHM_EINVAL, "ghm_bp_cls: null argument");
    if (B <= 0) return B == 0 ? GHM_OK : ghm_fail(GHM_EINVAL, "ghm_bp_cls: negative batch");
    if (leaf_dtype != GHM_LEAF_I64 && leaf_dtype != GHM_LEAF_U8) return ghm_fail(GHM_EINVAL, "bad leaf_dtype %d", leaf_dtype);
    DeviceGuard g(m->device);
    if (m->d.QW) return ghm_wide_bp_cls(m, B, leaves, leaf_dtype, post, root_hd, workspace, (cudaStream_t)stream);
    TreeArgs a{};
    a.B = B; a.leaves = const_cast<void*>(leaves); a.leaf_dtype = leaf_dtype; a.post = post; a.root_hd = root_hd;
    return run_given_bp(m, a, (cudaStream_t)stream);
}
#endif  // GHM_TREE_PROBE
