// ghm_dns.cu -- K3: Gaussian-denoiser belief propagation (posterior mean of every leaf).
//
// Replaces GHMTree.BP_DNS (reference src/ghmclip/data/data_random_GHM.py:467-523).
//
// Same thread-per-tree depth-first walk as ghm_tree.cu, two passes, LINEAR domain:
//   up    e_i[k]  = exp(-(z_i-k)^2 / 2 sigma^2) / max          leaf likelihood            (:485)
//         u_v     = T_v h_v,  h_v = prod_c u_c / max           message to the parent      (:487,494-497)
//   root  b_0     = h_0 * exp(ext)                             external root message      (:501-506)
//   down  b_v     = h_v * T_v^T (b_parent / u_v) / max         cavity rule                (:509-514)
//   out   mean_i  = sum_k k b_i[k] / sum_k b_i[k]                                         (:516-519)
// These are the reference's log-space recursions exponentiated; every rescale only changes a
// per-node constant that cancels in the normalised leaf marginals, so the result is the same
// posterior mean (tests: <= 1e-5 relative against the float64 oracle and the reference fixtures).
// The division replaces exp(bu - qd): one MUFU.RCP per state instead of LG2 + EX2, which moves the
// kernel from the MUFU pipe to the FFMA pipe (2E matvecs of q^2 FMAs dominate).
//
// State: the upward messages u_v of the internal nodes (E_int * q floats per tree) are parked in a
// global scratch laid out [node][state][tree] (coalesced, L2-resident for the in-flight trees) and
// read back by the downward pass; leaf messages are recomputed rather than stored (FFMA-only in the
// linear domain), per-level beliefs of the current root path live in shared memory.
#include <string.h>

#include <algorithm>

#include "ghm_vec.cuh"
#include "ghm_wide.cuh"


#include "ghm_dns2_decl.cuh"

// leaf likelihood vector, rescaled so that its largest entry is 1
template <int Q>
__device__ __forceinline__ void leaf_like(float z, float c2, int q, float (&e)[Q]) {
    float kstar = rintf(z);
    kstar = fminf(fmaxf(kstar, 0.f), (float)(q - 1));
    const float d0 = (z - kstar) * (z - kstar);
#pragma unroll
    for (int k = 0; k < Q; ++k) {
        const float dk = z - (float)k;
        e[k] = (k < q) ? ex2_approx(c2 * (dk * dk - d0)) : 0.f;
    }
}

template <int Q>
__device__ __forceinline__ void cavity(const float (&b)[Q], const float (&u)[Q], float (&w)[Q]) {
#pragma unroll
    for (int k = 0; k < Q; ++k) w[k] = u[k] > 0.f ? __fdividef(b[k], u[k]) : 0.f;
}

template <int Q, bool SMEM_TAB>
__global__ void __launch_bounds__(DNS_NT) k_dns(const GhmDev d, const DnsArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int NT = DNS_NT;
    const int tid = threadIdx.x;
    const int L = d.L, s = d.s, q = d.q, nL = d.n_leaves;
    const int64_t b = (int64_t)blockIdx.x * NT + tid;
    const bool active = b < a.B;
    const int64_t bc = active ? b : a.B - 1;
    const int64_t B = a.B;

    size_t off = 0;
    const float* Tlin = d.Tlin;
    if (SMEM_TAB) {
        const int tab_words = d.n_mat * Q * Q;
        float* s1 = reinterpret_cast<float*>(smem);
        off += (size_t)tab_words * 4;
        for (int i = tid; i < tab_words; i += NT) s1[i] = d.Tlin[i];
        Tlin = s1;
    }
    float* stack = reinterpret_cast<float*>(smem + off);          // [L][Q][NT] accumulators (up) / beliefs (down)
    off += (size_t)L * Q * NT * 4;
    float* leafu = reinterpret_cast<float*>(smem + off);          // [s][Q][NT]
    if (SMEM_TAB) __syncthreads();

    const int n1 = d.spow[L - 1];
    const float* zrow = a.z + bc * nL;
    float msg[Q];

    // =============================== upward pass ===============================================
    for (int j = 0; j < n1; ++j) {
        float h[Q];
#pragma unroll
        for (int k = 0; k < Q; ++k) h[k] = 1.f;
        for (int c = 0; c < s; ++c) {
            const int i = j * s + c;
            const int mi = d.mat_off[L] + (d.ti ? c : i);
            float e[Q], u[Q];
            leaf_like<Q>(zrow[i], a.c2, q, e);
            ghm_matvec<Q>(Tlin + (size_t)mi * Q * Q, e, u);
#pragma unroll
            for (int k = 0; k < Q; ++k) h[k] *= u[k];
        }
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
            if (active) {
                float* U = a.scratch + ((size_t)(d.edge_off[l] + idx) * Q) * B + b;
#pragma unroll
                for (int k = 0; k < Q; ++k) U[(size_t)k * B] = u[k];
            }
            float* A = stack + (size_t)(l - 1) * Q * NT + tid;
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
    // =============================== root belief ===============================================
    if (a.root_bu && active) {                                     // log of the max-rescaled root message (+ ext, unshifted)
#pragma unroll
        for (int k = 0; k < Q; ++k)
            if (k < q) a.root_bu[b * q + k] = __logf(msg[k]) + (a.ext ? a.ext[b * q + k] : 0.f);
    }
    if (a.ext) {
        float x[Q];
#pragma unroll
        for (int k = 0; k < Q; ++k) x[k] = (k < q) ? a.ext[bc * q + k] : -INFINITY;
        const float mx = ghm_vmax<Q>(x);
#pragma unroll
        for (int k = 0; k < Q; ++k) msg[k] *= (k < q) ? ex2_approx((x[k] - mx) * 1.4426950408889634f) : 0.f;
        ghm_normalize<Q>(msg);
    }
#pragma unroll
    for (int k = 0; k < Q; ++k) stack[k * NT + tid] = msg[k];
    __syncwarp();

    // =============================== downward pass =============================================
    float* mrow = a.mean + bc * nL;
    for (int j = 0; j < n1; ++j) {
        int tz = 0, t = j;
        while (tz < L - 1) {
            const int tq = ghm_div_s(t, d);
            if (t - tq * s != 0) break;
            t = tq; ++tz;
        }
        // beliefs of the internal nodes of the root path that changed, depths lstart .. L-2
        for (int l = max(1, L - 1 - tz); l <= L - 2; ++l) {
            const int idx = ghm_div_pow(j, L - 1 - l, d);
            const int pidx = ghm_div_s(idx, d);
            const int c = idx - pidx * s;
            const int mi = d.mat_off[l] + (d.ti ? c : idx);
            float uv[Q], hv[Q], bp[Q], w[Q], tt[Q];
            const float* U = a.scratch + ((size_t)(d.edge_off[l] + idx) * Q) * B + bc;
#pragma unroll
            for (int k = 0; k < Q; ++k) uv[k] = U[(size_t)k * B];
#pragma unroll
            for (int k = 0; k < Q; ++k) hv[k] = 1.f;
            for (int cc = 0; cc < s; ++cc) {
                const float* Uc = a.scratch + ((size_t)(d.edge_off[l + 1] + idx * s + cc) * Q) * B + bc;
#pragma unroll
                for (int k = 0; k < Q; ++k) hv[k] *= Uc[(size_t)k * B];
            }
            const float* P = stack + (size_t)(l - 1) * Q * NT + tid;
#pragma unroll
            for (int k = 0; k < Q; ++k) bp[k] = P[k * NT];
            cavity<Q>(bp, uv, w);
            ghm_matvec_t<Q>(Tlin + (size_t)mi * Q * Q, w, tt);
#pragma unroll
            for (int k = 0; k < Q; ++k) hv[k] *= tt[k];
            ghm_normalize<Q>(hv);
            float* S = stack + (size_t)l * Q * NT + tid;
#pragma unroll
            for (int k = 0; k < Q; ++k) S[k * NT] = hv[k];
        }
        // depth L-1 node j: leaf messages (recomputed), its belief, then the s leaf marginals
        float h[Q];
#pragma unroll
        for (int k = 0; k < Q; ++k) h[k] = 1.f;
        for (int c = 0; c < s; ++c) {
            const int i = j * s + c;
            const int mi = d.mat_off[L] + (d.ti ? c : i);
            float e[Q], u[Q];
            leaf_like<Q>(zrow[i], a.c2, q, e);
            ghm_matvec<Q>(Tlin + (size_t)mi * Q * Q, e, u);
            float* LU = leafu + (size_t)c * Q * NT + tid;
#pragma unroll
            for (int k = 0; k < Q; ++k) { h[k] *= u[k]; LU[k * NT] = u[k]; }
        }
        float bj[Q];
        if (L == 1) {
#pragma unroll
            for (int k = 0; k < Q; ++k) bj[k] = stack[k * NT + tid];
        } else {
            const int pidx = ghm_div_s(j, d);
            const int c = j - pidx * s;
            const int mi = d.mat_off[L - 1] + (d.ti ? c : j);
            float uv[Q], bp[Q], w[Q], tt[Q];
            const float* U = a.scratch + ((size_t)(d.edge_off[L - 1] + j) * Q) * B + bc;
#pragma unroll
            for (int k = 0; k < Q; ++k) uv[k] = U[(size_t)k * B];
            const float* P = stack + (size_t)(L - 2) * Q * NT + tid;
#pragma unroll
            for (int k = 0; k < Q; ++k) bp[k] = P[k * NT];
            cavity<Q>(bp, uv, w);
            ghm_matvec_t<Q>(Tlin + (size_t)mi * Q * Q, w, tt);
            ghm_normalize<Q>(h);
#pragma unroll
            for (int k = 0; k < Q; ++k) bj[k] = h[k] * tt[k];
            ghm_normalize<Q>(bj);
        }
        for (int c = 0; c < s; ++c) {
            const int i = j * s + c;
            const int mi = d.mat_off[L] + (d.ti ? c : i);
            float e[Q], u[Q], w[Q], tt[Q];
            const float* LU = leafu + (size_t)c * Q * NT + tid;
#pragma unroll
            for (int k = 0; k < Q; ++k) u[k] = LU[k * NT];
            cavity<Q>(bj, u, w);
            ghm_matvec_t<Q>(Tlin + (size_t)mi * Q * Q, w, tt);
            leaf_like<Q>(zrow[i], a.c2, q, e);
            float num = 0.f, den = 0.f;
#pragma unroll
            for (int k = 0; k < Q; ++k) {
                const float bl = e[k] * tt[k];
                num = fmaf((float)k, bl, num);
                den += bl;
            }
            if (active) mrow[i] = num / den;
        }
    }
}

// ----------------------------------------------------------------------------------------
extern "C" int64_t ghm_bp_dns_workspace_bytes(const ghm_model_t* m, int64_t B) {
    if (!m || B <= 0) return 0;
    if (m->d.QW) return ghm_wide_dns_workspace_bytes(m, B);
    const int Q = ghm_pad_q(m->d.q);
    return std::max<int64_t>(16, (int64_t)m->d.edge_off[m->d.L] * Q * B * (int64_t)sizeof(float));
}

template <int Q>
static int launch_dns(const ghm_model* m, const DnsArgs& a, cudaStream_t st) {
    const GhmDev& d = m->d;
    const size_t tab_bytes = (size_t)d.n_mat * Q * Q * 4;
    size_t dyn = (size_t)d.L * Q * DNS_NT * 4 + (size_t)d.s * Q * DNS_NT * 4;
    const bool smem_tab = tab_bytes + dyn <= 100 * 1024;
    if (smem_tab) dyn += tab_bytes;
    if (dyn > 200 * 1024)
        return ghm_fail(GHM_EUNSUP, "dns kernel needs %zu bytes of shared memory (L=%d s=%d q=%d)", dyn, d.L, d.s, d.q);
    const unsigned grid = (unsigned)((a.B + DNS_NT - 1) / DNS_NT);
    auto go = [&](auto kern) -> int {
        GHM_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
        kern<<<grid, DNS_NT, dyn, st>>>(d, a);
        GHM_CHECK_LAUNCH();
        return GHM_OK;
    };
    return smem_tab ? go(k_dns<Q, true>) : go(k_dns<Q, false>);
}


// GHM_EUNSUP -> the caller falls back to the generic k_dns
static int launch_dns2_any(const ghm_model* m, const DnsArgs& a, cudaStream_t st) {
    const GhmDev& d = m->d;
    if (!d.ti || d.s < 2 || d.s > 4 || d.L > 15) return GHM_EUNSUP;
#define DNS2_CASE(Q)                                                   \
    switch (d.s) {                                                     \
        case 2: return launch_dns2<Q, 2>(m, a, st);                    \
        case 3: return launch_dns2<Q, 3>(m, a, st);                    \
        default: return launch_dns2<Q, 4>(m, a, st);                   \
    }
    switch (ghm_pad_q(d.q)) {
        case 4: DNS2_CASE(4)
        case 8: DNS2_CASE(8)
        case 10: DNS2_CASE(10)
        case 16: DNS2_CASE(16)
        default: return GHM_EUNSUP;
    }
#undef DNS2_CASE
}

extern "C" int ghm_bp_dns(const ghm_model_t* m, int64_t B, const float* z, float sigma, const float* ext, float* mean,
                          float* root_bu, void* workspace, void* stream) {
    if (!m || !z || !mean || !workspace) return ghm_fail(GHM_EINVAL, "ghm_bp_dns: null argument");
    if (B <= 0) return B == 0 ? GHM_OK : ghm_fail(GHM_EINVAL, "ghm_bp_dns: negative batch");
    if (!(sigma > 0.f)) return ghm_fail(GHM_EINVAL, "ghm_bp_dns: sigma must be positive");
    int prev = 0;
    cudaGetDevice(&prev);
    if (prev != m->device) cudaSetDevice(m->device);
    DnsArgs a{};
    a.B = B; a.z = z; a.c2 = -0.5f * 1.4426950408889634f / (sigma * sigma); a.ext = ext; a.mean = mean;
    a.scratch = (float*)workspace; a.root_bu = root_bu;
    int rc;
    if (m->d.QW) {
        rc = ghm_wide_bp_dns(m, B, z, sigma, ext, mean, root_bu, workspace, (cudaStream_t)stream);
        if (prev != m->device) cudaSetDevice(prev);
        return rc;
    }
    rc = launch_dns2_any(m, a, (cudaStream_t)stream);
    if (rc != GHM_EUNSUP) {
        if (prev != m->device) cudaSetDevice(prev);
        return rc;
    }
    switch (ghm_pad_q(m->d.q)) {
        case 4: rc = launch_dns<4>(m, a, (cudaStream_t)stream); break;
        case 8: rc = launch_dns<8>(m, a, (cudaStream_t)stream); break;
        case 10: rc = launch_dns<10>(m, a, (cudaStream_t)stream); break;
        case 16: rc = launch_dns<16>(m, a, (cudaStream_t)stream); break;
        default:
            rc = ghm_fail(GHM_EUNSUP, "variable_type=%d: register-resident kernels cover q <= %d in this build", m->d.q,
                          GHM_MAX_Q_REG);
    }
    if (prev != m->device) cudaSetDevice(prev);
    return rc;
}
