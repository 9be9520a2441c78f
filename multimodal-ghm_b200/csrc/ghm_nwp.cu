// ghm_nwp.cu -- K4: next-token posteriors p(leaf_{t+1} | leaves_{<=t}, ext) for every prefix.
//
// Replaces GHMTree.BP_NWP_autoregressive (reference src/ghmclip/data/data_random_GHM.py:336-463),
// including its per-position guide tensors (:357-364,382,405-406,438-439,459).
//
// One thread per tree walks the positions t = 0 .. n_L-2 in order, exactly like the reference's
// stateful loop, but the state is O(L q) per tree instead of a Python Node graph:
//   S[l]   product of the (max-rescaled) messages of the COMPLETED children of the depth-l ancestor
//          of leaf t; this is what `parent.children[c].qd_message` holds for c < current (:394-396)
//   hd[l], qd[l]  the partial messages of the depth-l ancestor given leaves <= t (:397-399)
// Up the path: L-1 matvecs (`T @ exp(hd)`), root belief with the external message (:420-435),
// down the path of leaf t+1: a cavity update while the ancestors are shared, a plain push-down
// after they split (:443-454).  LINEAR domain with a max-rescale wherever the reference shifts;
// the guide tensors are the natural logs of those rescaled vectors, i.e. the reference's shifted
// log-messages (including the aliased root halves, :425-439).
#include <algorithm>

#include "ghm_vec.cuh"

#define NWP_NT 128
#define NWP_MAX_GUIDES (2 * GHM_MAX_LEVELS + 1)

struct NwpArgs {
    int64_t B;
    const void* leaves; int leaf_dtype;
    const float* ext;
    float* pp;                         // [B][nL-1][q]
    float* guides[NWP_MAX_GUIDES];     // only when GUIDE
};

template <int Q>
__device__ __forceinline__ void store_log(float* dst, const float (&v)[Q], int q) {
#pragma unroll
    for (int k = 0; k < Q; ++k)
        if (k < q) dst[k] = logf(v[k]);
}

template <int Q, bool GUIDE, bool SMEM_TAB>
__global__ void __launch_bounds__(NWP_NT) k_nwp(const GhmDev d, const NwpArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int NT = NWP_NT;
    const int tid = threadIdx.x;
    const int L = d.L, s = d.s, q = d.q, nL = d.n_leaves;
    const int64_t b = (int64_t)blockIdx.x * NT + tid;
    const bool active = b < a.B;
    const int64_t bc = active ? b : a.B - 1;

    size_t off = 0;
    const float* Tlin = d.Tlin;
    const float* TlinT = d.TlinT;
    if (SMEM_TAB) {
        const int tab_words = d.n_mat * Q * Q;
        float* s1 = reinterpret_cast<float*>(smem);
        float* s2 = s1 + tab_words;
        off += (size_t)tab_words * 8;
        for (int i = tid; i < tab_words; i += NT) { s1[i] = d.Tlin[i]; s2[i] = d.TlinT[i]; }
        Tlin = s1; TlinT = s2;
    }
    float* S = reinterpret_cast<float*>(smem + off);  off += (size_t)L * Q * NT * 4;   // [L][Q][NT]
    float* HD = reinterpret_cast<float*>(smem + off); off += (size_t)L * Q * NT * 4;   // [L][Q][NT] (depth 1..L-1 used)
    float* QD = reinterpret_cast<float*>(smem + off);                                    // [L][Q][NT]
    if (SMEM_TAB) __syncthreads();

    float ex[Q];          // exp(ext - max), the external root message in the linear domain
    if (a.ext) {
        float x[Q];
#pragma unroll
        for (int k = 0; k < Q; ++k) x[k] = (k < q) ? a.ext[bc * q + k] : -INFINITY;
        const float mx = ghm_vmax<Q>(x);
#pragma unroll
        for (int k = 0; k < Q; ++k) ex[k] = (k < q) ? __expf(x[k] - mx) : 0.f;
    } else {
#pragma unroll
        for (int k = 0; k < Q; ++k) ex[k] = (k < q) ? 1.f : 0.f;
    }

    const int64_t row = bc * (int64_t)(nL - 1);
    for (int t = 0; t < nL - 1; ++t) {
        // ---- observed leaf t: qd = log T[:, x_t], shifted (:374-375) -------------------------------
        int64_t xv = a.leaf_dtype == GHM_LEAF_I64 ? reinterpret_cast<const int64_t*>(a.leaves)[bc * nL + t]
                                                   : (int64_t) reinterpret_cast<const uint8_t*>(a.leaves)[bc * nL + t];
        if (xv < 0 || xv >= q) { atomicOr(d.status, 1); xv = xv < 0 ? 0 : q - 1; }
        float m[Q];
        {
            const int pidx = ghm_div_s(t, d);
            const int c = t - pidx * s;
            const int mi = d.mat_off[L] + (d.ti ? c : t);
            ghm_load_row<Q, float>(TlinT + ((size_t)mi * Q + (int)xv) * Q, m);
            ghm_normalize<Q>(m);
        }
        if (GUIDE && active) store_log<Q>(a.guides[0] + (row + t) * q, m, q);
        // ---- up the path: depth L-1 .. 1 (:389-415) -------------------------------------------------
        int idx = t;                                   // index of the node whose message `m` is, at depth l+1
        for (int l = L - 1; l >= 1; --l) {
            const int pidx = ghm_div_s(idx, d);        // ancestor at depth l
            const int c = idx - pidx * s;              // which child of it the path goes through
            float* Sl = S + (size_t)l * Q * NT + tid;
            float h[Q];
            if (c == 0) {
#pragma unroll
                for (int k = 0; k < Q; ++k) h[k] = m[k];
            } else {
#pragma unroll
                for (int k = 0; k < Q; ++k) h[k] = Sl[k * NT] * m[k];
            }
            // the child (depth l+1, idx) is complete iff leaf t+1 lies under a different depth-(l+1) node
            const bool child_done = ghm_div_pow(t + 1, L - 1 - l, d) != idx;
            ghm_normalize<Q>(h);
            if (child_done) {
#pragma unroll
                for (int k = 0; k < Q; ++k) Sl[k * NT] = h[k];
            }
            const int ppidx = ghm_div_s(pidx, d);
            const int pc = pidx - ppidx * s;
            const int mi = d.mat_off[l] + (d.ti ? pc : pidx);
            ghm_matvec<Q>(Tlin + (size_t)mi * Q * Q, h, m);
            ghm_normalize<Q>(m);
            float* Hl = HD + (size_t)l * Q * NT + tid;
            float* Ql = QD + (size_t)l * Q * NT + tid;
#pragma unroll
            for (int k = 0; k < Q; ++k) { Hl[k * NT] = h[k]; Ql[k * NT] = m[k]; }
            if (GUIDE && active) {
                float* g = a.guides[L - l] + (row + t) * 2 * q;
                store_log<Q>(g, h, q);
                store_log<Q>(g + q, m, q);
            }
            idx = pidx;
        }
        // ---- root (:420-439): idx is now the depth-1 node on the path --------------------------------
        float bel[Q];
        {
            float* S0 = S + tid;
            if (idx == 0) {
#pragma unroll
                for (int k = 0; k < Q; ++k) bel[k] = m[k];
            } else {
#pragma unroll
                for (int k = 0; k < Q; ++k) bel[k] = S0[k * NT] * m[k];
            }
            ghm_normalize<Q>(bel);
            const bool child_done = ghm_div_pow(t + 1, L - 1, d) != idx;
            if (child_done) {
#pragma unroll
                for (int k = 0; k < Q; ++k) S0[k * NT] = bel[k];
            }
#pragma unroll
            for (int k = 0; k < Q; ++k) bel[k] *= ex[k];
            ghm_normalize<Q>(bel);
            if (GUIDE && active) {
                float* g = a.guides[L] + (row + t) * 2 * q;
                store_log<Q>(g, bel, q);
                store_log<Q>(g + q, bel, q);
            }
        }
        // ---- down the path of leaf t+1 (:443-459) ------------------------------------------------------
        for (int l = 1; l <= L; ++l) {
            const int g = ghm_div_pow(t + 1, L - l, d);            // goal-path node at depth l
            const int a_l = ghm_div_pow(t, L - l, d);              // observed-path node at depth l
            const int pg = ghm_div_s(g, d);
            const int cg = g - pg * s;
            const int mi = d.mat_off[l] + (d.ti ? cg : g);
            float w[Q], tt[Q];
            if (g == a_l) {                                         // shared ancestor: cavity update (l <= L-1 here)
                const float* Hl = HD + (size_t)l * Q * NT + tid;
                const float* Ql = QD + (size_t)l * Q * NT + tid;
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
            if (GUIDE && active) store_log<Q>(a.guides[L + l] + (row + t) * q, bel, q);
        }
        if (active) {
            float sum = 0.f;
#pragma unroll
            for (int k = 0; k < Q; ++k) sum += bel[k];
            const float inv = 1.0f / sum;
            float* o = a.pp + (row + t) * q;
#pragma unroll
            for (int k = 0; k < Q; ++k)
                if (k < q) o[k] = bel[k] * inv;
        }
    }
}

// ----------------------------------------------------------------------------------------
template <int Q, bool GUIDE>
static int launch_nwp(const ghm_model* m, const NwpArgs& a, cudaStream_t st) {
    const GhmDev& d = m->d;
    const size_t tab_bytes = (size_t)d.n_mat * Q * Q * 8;
    size_t dyn = (size_t)3 * d.L * Q * NWP_NT * 4;
    const bool smem_tab = tab_bytes + dyn <= 100 * 1024;
    if (smem_tab) dyn += tab_bytes;
    if (dyn > 200 * 1024)
        return ghm_fail(GHM_EUNSUP, "nwp kernel needs %zu bytes of shared memory (L=%d q=%d)", dyn, d.L, d.q);
    const unsigned grid = (unsigned)((a.B + NWP_NT - 1) / NWP_NT);
    auto go = [&](auto kern) -> int {
        GHM_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
        kern<<<grid, NWP_NT, dyn, st>>>(d, a);
        GHM_CHECK_LAUNCH();
        return GHM_OK;
    };
    return smem_tab ? go(k_nwp<Q, GUIDE, true>) : go(k_nwp<Q, GUIDE, false>);
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

static int nwp_common(const ghm_model_t* m, int64_t B, const void* leaves, int leaf_dtype, const float* ext,
                      float* const* guides, float* pp, void* stream) {
    if (!m || !leaves || !pp) return ghm_fail(GHM_EINVAL, "ghm_bp_nwp: null argument");
    if (B <= 0) return B == 0 ? GHM_OK : ghm_fail(GHM_EINVAL, "ghm_bp_nwp: negative batch");
    if (leaf_dtype != GHM_LEAF_I64 && leaf_dtype != GHM_LEAF_U8) return ghm_fail(GHM_EINVAL, "bad leaf_dtype %d", leaf_dtype);
    if (m->d.n_leaves < 2) return GHM_OK;
    int prev = 0;
    cudaGetDevice(&prev);
    if (prev != m->device) cudaSetDevice(m->device);
    NwpArgs a{};
    a.B = B; a.leaves = leaves; a.leaf_dtype = leaf_dtype; a.ext = ext; a.pp = pp;
    int rc;
    if (guides) {
        for (int i = 0; i < 2 * m->d.L + 1; ++i) a.guides[i] = guides[i];
        rc = dispatch_nwp<true>(m, a, (cudaStream_t)stream);
    } else {
        rc = dispatch_nwp<false>(m, a, (cudaStream_t)stream);
    }
    if (prev != m->device) cudaSetDevice(prev);
    return rc;
}

extern "C" int64_t ghm_bp_nwp_workspace_bytes(const ghm_model_t*, int64_t) { return 16; }
extern "C" int ghm_bp_nwp(const ghm_model_t* m, int64_t B, const void* leaves, int leaf_dtype, const float* ext,
                          float* pp, void*, void* stream) {
    return nwp_common(m, B, leaves, leaf_dtype, ext, nullptr, pp, stream);
}
extern "C" int64_t ghm_guides_nwp_workspace_bytes(const ghm_model_t*, int64_t) { return 16; }
extern "C" int ghm_guides_nwp(const ghm_model_t* m, int64_t B, const void* leaves, int leaf_dtype, const float* ext,
                              float* const* guides, float* pp, void*, void* stream) {
    if (!guides) return ghm_fail(GHM_EINVAL, "ghm_guides_nwp: null guides");
    return nwp_common(m, B, leaves, leaf_dtype, ext, guides, pp, stream);
}
