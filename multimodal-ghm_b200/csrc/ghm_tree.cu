// ghm_tree.cu -- K1 (sampler) and K2 (root-posterior BP), separately or fused in one pass.
//
// Replaces GHMTree.gen_values (src/ghmclip/data/data_random_GHM.py:145-165) and
// GHMTree.BP_CLS (:185-221) of the reference.
//
// Design (B200, FP32 CUDA cores; HBM-write bound when leaves are materialised, FP32-issue
// bound when BP is fused):
//   * one THREAD per tree walks its tree depth-first: sampling goes down the current root
//     path, the BP message comes back up the same path, so a leaf state lives in a register
//     between being drawn and being absorbed -- fused mode moves no leaf through HBM twice;
//   * all 32 lanes of a warp are at the same node of their 32 trees, so transition-table reads
//     are shared-memory broadcasts and there is no divergence; per-lane gathers (CDF row of the
//     parent state, T^T row of the leaf state) are the only non-broadcast LDS;
//   * BP runs in the LINEAR domain with a max-rescale per node: msg(v) = prod_c (T_c msg(c)) /
//     max.  This is the reference's log-space recursion `hd = sum_c log(T_c @ exp(hd_c)) - max`
//     (:207-208) exponentiated -- same rescale points, no exp/log in the inner loop, so the
//     kernel is bound by FFMA issue rather than by the MUFU pipe;
//   * per-level accumulators (the partial products of the ancestors on the current path) sit in
//     shared memory as [level][state][thread] (conflict-free), so the depth L is a runtime value;
//   * leaves are staged per warp in shared memory as bytes and written/read with coalesced
//     row-contiguous transactions (the [B, n_L] int64 API layout is 8*n_L contiguous bytes per tree).
#include <algorithm>

#include "ghm_vec.cuh"

#define GHM_NT 128   // threads (= trees) per CTA

enum { MODE_PHILOX = 0, MODE_PARITY = 1, MODE_GIVEN = 2 };

struct TreeArgs {
    int64_t B;
    int root_mode;
    const int64_t* root_in;
    const double* U;
    uint64_t seed, tree_offset;
    int64_t* root_out;
    void* leaves;          // output (sampling modes) or input (MODE_GIVEN); may be null when sampling
    int leaf_dtype;
    float* post;
    float* root_hd;
    int chunk_j;           // depth-(L-1) nodes per staging chunk
    int stage_stride;      // bytes per staged tree row
};

__device__ __forceinline__ void stage_flush(const uint8_t* st, int stride, void* leaves, int dtype, int64_t tree0,
                                            int64_t B, int nL, int base, int len, int lane) {
    __syncwarp();
    for (int r = 0; r < 32; ++r) {
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

__device__ __forceinline__ void stage_load(uint8_t* st, int stride, const void* leaves, int dtype, int64_t tree0,
                                           int64_t B, int nL, int base, int len, int lane, int q, int* status) {
    __syncwarp();
    bool bad = false;
    for (int r = 0; r < 32; ++r) {
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

template <int Q, int MODE, bool BP, bool SMEM_TAB>
__global__ void __launch_bounds__(GHM_NT) k_tree(const GhmDev d, const TreeArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int NT = GHM_NT;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int L = d.L, s = d.s, q = d.q;
    const int64_t b = (int64_t)blockIdx.x * NT + tid;
    const bool active = b < a.B;
    const int64_t bc = active ? b : a.B - 1;   // tail threads shadow the last tree and never write
    const uint64_t tree = a.tree_offset + (uint64_t)bc;

    // ---- carve shared memory ------------------------------------------------------------
    size_t off = 0;
    const int tab_words = d.n_mat * Q * Q;
    const float* Tlin = d.Tlin;
    const float* TlinT = d.TlinT;
    const uint32_t* cdfu = d.cdfu;
    if (SMEM_TAB) {
        if (BP) {
            float* s1 = reinterpret_cast<float*>(smem + off); off += (size_t)tab_words * 4;
            float* s2 = reinterpret_cast<float*>(smem + off); off += (size_t)tab_words * 4;
            for (int i = tid; i < tab_words; i += NT) { s1[i] = d.Tlin[i]; s2[i] = d.TlinT[i]; }
            Tlin = s1; TlinT = s2;
        }
        if (MODE == MODE_PHILOX) {
            uint32_t* s3 = reinterpret_cast<uint32_t*>(smem + off); off += (size_t)tab_words * 4;
            for (int i = tid; i < tab_words; i += NT) s3[i] = d.cdfu[i];
            cdfu = s3;
        }
    }
    float* acc = reinterpret_cast<float*>(smem + off);          // [L-1][Q][NT]
    if (BP) off += (size_t)(L > 1 ? L - 1 : 0) * Q * NT * 4;
    int* val = reinterpret_cast<int*>(smem + off);              // [L][NT] states on the current path
    if (MODE != MODE_GIVEN) off += (size_t)L * NT * 4;
    uint32_t* rng = reinterpret_cast<uint32_t*>(smem + off);    // [L][4][NT] Philox block per level
    if (MODE == MODE_PHILOX) off += (size_t)L * 4 * NT * 4;
    const bool use_stage = (a.leaves != nullptr);
    uint8_t* stage = smem + off + (size_t)warp * 32 * a.stage_stride;
    if (SMEM_TAB) __syncthreads();

    const int nL = d.n_leaves;
    const int n1 = d.spow[L - 1];
    const int64_t warp_tree0 = (int64_t)blockIdx.x * NT + warp * 32;
    const bool warp_has_work = warp_tree0 < a.B;

    // ---- root ---------------------------------------------------------------------------
    if (MODE != MODE_GIVEN) {
        int x0;
        if (a.root_mode == GHM_ROOT_GIVEN) {
            int64_t r = a.root_in[bc];
            if (r < 0 || r >= q) { atomicOr(d.status, 1); r = r < 0 ? 0 : q - 1; }
            x0 = (int)r;
        } else {
            const uint4 rb = ghm_rng_block(a.seed, tree, 0u, 0u, GHM_STREAM_TREE);
            const uint32_t* rc = a.root_mode == GHM_ROOT_PRIOR ? d.root_cdfu_prior : d.root_cdfu_unif;
            int cnt = 0;
            for (int k = 0; k < q - 1; ++k) cnt += (rb.x >= __ldg(rc + k)) ? 1 : 0;
            x0 = cnt;
        }
        val[tid] = x0;
        if (a.root_out && active) a.root_out[b] = x0;
    }

    float msg[Q];
#pragma unroll
    for (int k = 0; k < Q; ++k) msg[k] = 0.f;
    uint4 leaf_rb = make_uint4(0, 0, 0, 0);
    int leaf_blk = -1;

    for (int j = 0; j < n1; ++j) {
        const int chunk_base = (j / a.chunk_j) * a.chunk_j * s;   // leaf index of the first staged leaf
        if (MODE == MODE_GIVEN && (j % a.chunk_j) == 0) {   // idle tail warps shadow tree B-1 (clamped inside)
            const int len = min(a.chunk_j * s, nL - chunk_base);
            stage_load(stage, a.stage_stride, a.leaves, a.leaf_dtype, warp_tree0, a.B, nL, chunk_base, len, lane, q,
                       d.status);
        }
        // ---- (re)draw the internal nodes of the root path that changed ----------------------
        if (MODE != MODE_GIVEN) {
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
                const int xp = val[(l - 1) * NT + tid];
                int x;
                if (MODE == MODE_PHILOX) {
                    uint32_t* rl = rng + (size_t)l * 4 * NT + tid;
                    if ((idx & 3) == 0) {
                        const uint4 rb = ghm_rng_block(a.seed, tree, (uint32_t)l, (uint32_t)(idx >> 2), GHM_STREAM_TREE);
                        rl[0] = rb.x; rl[NT] = rb.y; rl[2 * NT] = rb.z; rl[3 * NT] = rb.w;
                    }
                    const uint32_t r = rl[(idx & 3) * NT];
                    x = ghm_search_u32<Q>(cdfu + ((size_t)mi * Q + xp) * Q, r, q);
                } else {
                    const double u = a.U[(size_t)(d.edge_off[l] + idx) * a.B + bc];
                    x = ghm_search_f64(d.cdfd + ((size_t)mi * q + xp) * q, u, q);
                }
                val[l * NT + tid] = x;
            }
        }
        // ---- the s leaves under depth-(L-1) node j --------------------------------------------
        const int xj = (MODE != MODE_GIVEN) ? val[(L - 1) * NT + tid] : 0;
        float h[Q];
#pragma unroll
        for (int k = 0; k < Q; ++k) h[k] = 1.f;
        for (int c = 0; c < s; ++c) {
            const int lidx = j * s + c;
            const int mi = d.mat_off[L] + (d.ti ? c : lidx);
            int x;
            if (MODE == MODE_PHILOX) {
                if ((lidx >> 2) != leaf_blk) {
                    leaf_blk = lidx >> 2;
                    leaf_rb = ghm_rng_block(a.seed, tree, (uint32_t)L, (uint32_t)leaf_blk, GHM_STREAM_TREE);
                }
                x = ghm_search_u32<Q>(cdfu + ((size_t)mi * Q + xj) * Q, ghm_pick(leaf_rb, lidx & 3), q);
            } else if (MODE == MODE_PARITY) {
                const double u = a.U[(size_t)(d.edge_off[L] + lidx) * a.B + bc];
                x = ghm_search_f64(d.cdfd + ((size_t)mi * q + xj) * q, u, q);
            } else {
                x = stage[lane * a.stage_stride + (lidx - chunk_base)];
            }
            if (MODE != MODE_GIVEN && use_stage) stage[lane * a.stage_stride + (lidx - chunk_base)] = (uint8_t)x;
            if (BP) {
                float row[Q];
                ghm_load_row<Q, float>(TlinT + ((size_t)mi * Q + x) * Q, row);
#pragma unroll
                for (int k = 0; k < Q; ++k) h[k] *= row[k];
            }
        }
        // ---- carry the finished node's message up the path ----------------------------------
        if (BP) {
#pragma unroll
            for (int k = 0; k < Q; ++k) msg[k] = h[k];
            ghm_normalize<Q>(msg);
            int l = L - 1, idx = j;
            while (l > 0) {
                const int pidx = ghm_div_s(idx, d);
                const int c = idx - pidx * s;
                const int mi = d.mat_off[l] + (d.ti ? c : idx);
                float u[Q];
                ghm_matvec<Q>(Tlin + (size_t)mi * Q * Q, msg, u);
                float* A = acc + (size_t)(l - 1) * Q * NT + tid;
                if (c != 0) {
#pragma unroll
                    for (int k = 0; k < Q; ++k) u[k] *= A[k * NT];
                }
                if (c != s - 1) {
#pragma unroll
                    for (int k = 0; k < Q; ++k) A[k * NT] = u[k];
                    break;
                }
#pragma unroll
                for (int k = 0; k < Q; ++k) msg[k] = u[k];
                ghm_normalize<Q>(msg);
                --l;
                idx = pidx;
            }
        }
        if (MODE != MODE_GIVEN && use_stage && (((j + 1) % a.chunk_j) == 0 || j == n1 - 1) && warp_has_work) {
            const int len = (j + 1) * s - chunk_base;
            stage_flush(stage, a.stage_stride, a.leaves, a.leaf_dtype, warp_tree0, a.B, nL, chunk_base, len, lane);
        }
    }

    // ---- root outputs (reference :213-217; root_node.hd_message is the shifted log-likelihood) --
    if (BP && active) {
        if (a.root_hd) {
#pragma unroll
            for (int k = 0; k < Q; ++k)
                if (k < q) a.root_hd[b * q + k] = logf(msg[k]);
        }
        if (a.post) {
            float w[Q], sum = 0.f;
#pragma unroll
            for (int k = 0; k < Q; ++k) { w[k] = msg[k] * __ldg(d.py + k); sum += w[k]; }
            const float inv = 1.0f / sum;
#pragma unroll
            for (int k = 0; k < Q; ++k)
                if (k < q) a.post[b * q + k] = w[k] * inv;
        }
    }
}

// ----------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------
template <int Q, int MODE, bool BP>
static int launch_tree(const ghm_model* m, const TreeArgs& a0, cudaStream_t st) {
    const GhmDev& d = m->d;
    TreeArgs a = a0;
    const int n1 = d.spow[d.L - 1];
    int chunk_j = n1;
    if (d.n_leaves > 256) chunk_j = 256 / d.s > 0 ? 256 / d.s : 1;
    a.chunk_j = chunk_j;
    int stride = (int)((std::min(chunk_j * d.s, d.n_leaves) + 3) / 4 * 4);
    if (((stride / 4) & 1) == 0) stride += 4;
    a.stage_stride = stride;

    const size_t tab_bytes = (size_t)d.n_mat * Q * Q * 4 * ((BP ? 2 : 0) + (MODE == MODE_PHILOX ? 1 : 0));
    size_t dyn = 0;
    if (BP) dyn += (size_t)(d.L > 1 ? d.L - 1 : 0) * Q * GHM_NT * 4;
    if (MODE != MODE_GIVEN) dyn += (size_t)d.L * GHM_NT * 4;
    if (MODE == MODE_PHILOX) dyn += (size_t)d.L * 4 * GHM_NT * 4;
    if (a.leaves) dyn += (size_t)(GHM_NT / 32) * 32 * stride;
    const bool smem_tab = tab_bytes > 0 && tab_bytes + dyn <= 100 * 1024;
    if (smem_tab) dyn += tab_bytes;
    if (dyn > 200 * 1024)
        return ghm_fail(GHM_EUNSUP, "tree kernel needs %zu bytes of shared memory (L=%d s=%d q=%d)", dyn, d.L, d.s, d.q);
    const unsigned grid = (unsigned)((a.B + GHM_NT - 1) / GHM_NT);
    auto go = [&](auto kern) -> int {
        GHM_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
        kern<<<grid, GHM_NT, dyn, st>>>(d, a);
        GHM_CHECK_LAUNCH();
        return GHM_OK;
    };
    return smem_tab ? go(k_tree<Q, MODE, BP, true>) : go(k_tree<Q, MODE, BP, false>);
}

template <int MODE, bool BP>
static int dispatch_tree(const ghm_model* m, const TreeArgs& a, cudaStream_t st) {
    switch (ghm_pad_q(m->d.q)) {
        case 4: return launch_tree<4, MODE, BP>(m, a, st);
        case 8: return launch_tree<8, MODE, BP>(m, a, st);
        case 10: return launch_tree<10, MODE, BP>(m, a, st);
        case 16: return launch_tree<16, MODE, BP>(m, a, st);
        default:
            return ghm_fail(GHM_EUNSUP, "variable_type=%d: register-resident kernels cover q <= %d in this build",
                            m->d.q, GHM_MAX_Q_REG);
    }
}

struct DeviceGuard {
    int prev;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

extern "C" int ghm_sample(const ghm_model_t* m, int64_t B, int root_mode, const int64_t* root_in, const double* U,
                          uint64_t seed, uint64_t tree_offset, int64_t* root_out, void* leaves_out, int leaf_dtype,
                          float* post_out, float* root_hd_out, void* stream) {
    if (!m) return ghm_fail(GHM_EINVAL, "ghm_sample: null model");
    if (B <= 0) return B == 0 ? GHM_OK : ghm_fail(GHM_EINVAL, "ghm_sample: negative batch");
    if (root_mode < 0 || root_mode > 2) return ghm_fail(GHM_EINVAL, "ghm_sample: bad root_mode %d", root_mode);
    if (root_mode == GHM_ROOT_GIVEN && !root_in) return ghm_fail(GHM_EINVAL, "ghm_sample: root_in is null");
    if (U && root_mode != GHM_ROOT_GIVEN)
        return ghm_fail(GHM_EINVAL, "ghm_sample: parity mode (U given) needs host-drawn roots (GHM_ROOT_GIVEN)");
    if (leaf_dtype != GHM_LEAF_I64 && leaf_dtype != GHM_LEAF_U8) return ghm_fail(GHM_EINVAL, "bad leaf_dtype %d", leaf_dtype);
    DeviceGuard g(m->device);
    TreeArgs a{};
    a.B = B; a.root_mode = root_mode; a.root_in = root_in; a.U = U; a.seed = seed; a.tree_offset = tree_offset;
    a.root_out = root_out; a.leaves = leaves_out; a.leaf_dtype = leaf_dtype; a.post = post_out; a.root_hd = root_hd_out;
    const bool bp = post_out || root_hd_out;
    cudaStream_t st = (cudaStream_t)stream;
    if (U) {
        if (bp) return ghm_fail(GHM_EINVAL, "ghm_sample: fused BP is a Philox-mode feature; in parity mode call ghm_bp_cls");
        return dispatch_tree<MODE_PARITY, false>(m, a, st);
    }
    return bp ? dispatch_tree<MODE_PHILOX, true>(m, a, st) : dispatch_tree<MODE_PHILOX, false>(m, a, st);
}

extern "C" int ghm_bp_cls(const ghm_model_t* m, int64_t B, const void* leaves, int leaf_dtype, float* post,
                          float* root_hd, void* stream) {
    if (!m || !leaves) return ghm_fail(GHM_EINVAL, "ghm_bp_cls: null argument");
    if (B <= 0) return B == 0 ? GHM_OK : ghm_fail(GHM_EINVAL, "ghm_bp_cls: negative batch");
    if (leaf_dtype != GHM_LEAF_I64 && leaf_dtype != GHM_LEAF_U8) return ghm_fail(GHM_EINVAL, "bad leaf_dtype %d", leaf_dtype);
    DeviceGuard g(m->device);
    TreeArgs a{};
    a.B = B; a.leaves = const_cast<void*>(leaves); a.leaf_dtype = leaf_dtype; a.post = post; a.root_hd = root_hd;
    return dispatch_tree<MODE_GIVEN, true>(m, a, (cudaStream_t)stream);
}
