// ghm_dns2_kernel.cuh -- the packed / constant-bank Gaussian-denoiser BP kernel (k_dns2) and its launcher, kept in a
// header so that each padded q gets its own translation unit (ghm_dns2_inst_q*.cu): the kernel inlines ~13 unrolled
// q x q matvecs per (q, s) instance and one file with all twelve instances took nine minutes to compile.
#pragma once
#include <string.h>

#include "ghm_dns2_decl.cuh"
#include "ghm_vec2.cuh"

// ------------------------------------------------------------------------------------------------
// k_dns2: the fast variant (translation-invariant tables, s in {2,3,4}).  Same two depth-first passes and the
// same scratch as k_dns, but
//   * messages are packed f32x2 register pairs (FFMA2 / FMUL2, ghm_vec2.cuh);
//   * both table orientations (T^T rows for the child->parent matvec, T rows for parent->child) are a by-value
//     kernel parameter: constant bank -> LDCU.64 -> uniform-register operands of FFMA2.  k_dns's LDS.128
//     broadcasts ran the 128 B/clk LSU pipe 4x longer than the FMAs they fed (321 matvecs per tree);
//   * the table address of every hot matvec (leaf level: compile-time child index; node j: the running index
//     toff0; its parent: toff1) is built only from host-provided bases and loop counters so that it stays in the
//     uniform datapath (see ghm_tree_kernel.cuh); ancestors at depth <= L-3 (every s^2-th node) use per-thread LDC;
//   * leaf likelihoods e_c and leaf messages u_c of the current node stay in registers between the node's belief
//     and its leaves' marginals instead of being recomputed / bounced through shared memory.
// ------------------------------------------------------------------------------------------------

struct Dns2Args {
    DnsArgs a;
    int base0, base1, base_leaf;                       // matrix index of child 0 of: edges into depth L-1, L-2, L
    int dn_off;                                        // float offset of the Tlin block inside the table parameter
};

__device__ __forceinline__ f2 f2_add(f2 a, f2 b) {
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_pack(a)), "l"(f2_pack(b)));
    return f2_unpack(d);
}

// Gaussian leaf likelihoods e[k] = exp2(c2 ((z - k)^2 - d0)), shifted so that the nearest state has e = 1 (:485).
// Packed f32x2 arithmetic: three packed instructions + two MUFU.EX2 per state pair (the scalar form was 14 and, with
// 162 calls per tree, 14 % of k_dns2's instructions).  EX: q == Q, no padding states to mask.
template <int Q, bool EX>
__device__ __forceinline__ void f2_leaf_like(float z, float c2, int q, f2 (&e)[Q / 2]) {
    float kstar = rintf(z);
    kstar = fminf(fmaxf(kstar, 0.f), (float)(q - 1));
    const float d0 = (z - kstar) * (z - kstar);
    const f2 zz = make_float2(z, z), cc = make_float2(c2, c2), dd = make_float2(-d0, -d0);
#pragma unroll
    for (int i = 0; i < Q / 2; ++i) {
        const f2 da = f2_add(zz, make_float2(-(float)(2 * i), -(float)(2 * i + 1)));
        const f2 arg = f2_mul(f2_fma(da, da, dd), cc);
        e[i].x = (EX || 2 * i < q) ? ex2_approx(arg.x) : 0.f;
        e[i].y = (EX || 2 * i + 1 < q) ? ex2_approx(arg.y) : 0.f;
    }
}

template <int Q>
__device__ __forceinline__ void f2_cavity(const f2 (&b)[Q / 2], const f2 (&u)[Q / 2], f2 (&w)[Q / 2]) {
#pragma unroll
    for (int i = 0; i < Q / 2; ++i) {
        w[i].x = u[i].x > 0.f ? __fdividef(b[i].x, u[i].x) : 0.f;
        w[i].y = u[i].y > 0.f ? __fdividef(b[i].y, u[i].y) : 0.f;
    }
}

template <int Q, int S, int NW>
__global__ void __launch_bounds__(DNS_NT, 4)
k_dns2(const __grid_constant__ GhmDev d, const __grid_constant__ Dns2Args aa, const __grid_constant__ DnsTab<NW> tab) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int NT = DNS_NT, H = Q / 2, QQ = Q * Q;
    const DnsArgs& a = aa.a;
    const int tid = threadIdx.x;
    const int L = d.L, s = S, q = d.q, nL = d.n_leaves;
    const int64_t b = (int64_t)blockIdx.x * NT + tid;
    const bool active = b < a.B;
    const int64_t bc = active ? b : a.B - 1;
    const int64_t B = a.B;
    f2* stack = reinterpret_cast<f2*>(smem);                       // [L][H][NT] accumulators (up) / beliefs (down)
    const int n1 = d.spow[L - 1];
    const float* zrow = a.z + bc * nL;
    const float* Tup = tab.v;                                      // T^T rows: child -> parent
    const float* Tdn = tab.v + aa.dn_off;                          // T rows:   parent -> child

    auto sload = [&](int lvl, f2 (&v)[H]) {
#pragma unroll
        for (int i = 0; i < H; ++i) v[i] = stack[((size_t)lvl * H + i) * NT + tid];
    };
    auto sstore = [&](int lvl, const f2 (&v)[H]) {
#pragma unroll
        for (int i = 0; i < H; ++i) stack[((size_t)lvl * H + i) * NT + tid] = v[i];
    };
    auto uload = [&](int edge, f2 (&v)[H]) {                      // scratch [edge][k][B], coalesced across lanes
        const float* U = a.scratch + (size_t)edge * Q * B + bc;
#pragma unroll
        for (int i = 0; i < H; ++i) { v[i].x = U[(size_t)(2 * i) * B]; v[i].y = U[(size_t)(2 * i + 1) * B]; }
    };
    auto ustore = [&](int edge, const f2 (&v)[H]) {
        if (!active) return;
        float* U = a.scratch + (size_t)edge * Q * B + b;
#pragma unroll
        for (int i = 0; i < H; ++i) { U[(size_t)(2 * i) * B] = v[i].x; U[(size_t)(2 * i + 1) * B] = v[i].y; }
    };

    // =============================== upward pass ===============================================
    f2 msg[H], accT[H];
#pragma unroll
    for (int i = 0; i < H; ++i) { msg[i] = make_float2(0.f, 0.f); accT[i] = make_float2(0.f, 0.f); }
    int cj = 0;
    const int base0 = aa.base0, base1 = aa.base1;
    int toff0 = base0, toff1 = base1;                              // uniform-side twins: matrix index of the two hottest steps
    auto advance = [&]() {
        if (++cj == s) cj = 0;
        toff0 += 1;
        if (toff0 == base0 + s) {
            toff0 = base0;
            toff1 += 1;
            if (toff1 == base1 + s) toff1 = base1;
        }
    };
    for (int j = 0; j < n1; ++j) {
        f2 h[H];
#pragma unroll
        for (int c = 0; c < S; ++c) {
            f2 e[H], u[H];
            f2_leaf_like<Q, false>(zrow[j * S + c], a.c2, q, e);
            f2_matvec_c<Q>(Tup + (aa.base_leaf + c) * QQ, e, u);
#pragma unroll
            for (int i = 0; i < H; ++i) h[i] = c == 0 ? u[i] : f2_mul(h[i], u[i]);
        }
#pragma unroll
        for (int i = 0; i < H; ++i) msg[i] = h[i];
        f2_normalize<Q>(msg);
        // one climb step: u = T msg -> scratch; times the parked product of the earlier siblings; park or continue
        auto climb = [&](const float* __restrict__ Tm, int edge, bool has_prev, bool last, f2* A) -> bool {
            f2 u[H];
            f2_matvec_c<Q>(Tm, msg, u);
            ustore(edge, u);
            if (has_prev) {
#pragma unroll
                for (int i = 0; i < H; ++i) u[i] = f2_mul(u[i], A ? A[(size_t)i * NT] : accT[i]);
            }
            if (!last) {
#pragma unroll
                for (int i = 0; i < H; ++i) { if (A) A[(size_t)i * NT] = u[i]; else accT[i] = u[i]; }
                return false;
            }
#pragma unroll
            for (int i = 0; i < H; ++i) msg[i] = u[i];
            f2_normalize<Q>(msg);
            return true;
        };
        bool up = L >= 2;
        int idx = j;                                               // BFS index of the node whose message is `msg`
        if (up) up = climb(Tup + toff0 * QQ, d.edge_off[L - 1] + idx, toff0 != base0, toff0 == base0 + (s - 1), nullptr);
        idx = idx / S;
        if (up && L >= 3) {
            up = climb(Tup + toff1 * QQ, d.edge_off[L - 2] + idx, toff1 != base1, toff1 == base1 + (s - 1),
                       stack + (size_t)(L - 3) * H * NT + tid);
            idx = idx / S;
        }
        if (up && L >= 4) {
            for (int l = L - 3; l > 0; --l) {
                const int pl = ghm_div_pow(j, L - 1 - l, d);               // path node at depth l (uniform datapath)
                const int c = pl - (pl / S) * S;
                if (!climb(Tup + ((l - 1) * s + c) * QQ, d.edge_off[l] + idx, c != 0, c == s - 1, stack + (size_t)(l - 1) * H * NT + tid))
                    break;
                idx = idx / S;
            }
        }
        advance();
    }
    // =============================== root belief ===============================================
    if (a.root_bu && active) {                                     // log of the max-rescaled root message (+ ext, unshifted)
#pragma unroll
        for (int k = 0; k < Q; ++k)
            if (k < q) a.root_bu[b * q + k] = __logf(f2_elem<Q>(msg, k)) + (a.ext ? a.ext[b * q + k] : 0.f);
    }
    if (a.ext) {
        float x[Q];
#pragma unroll
        for (int k = 0; k < Q; ++k) x[k] = (k < q) ? a.ext[bc * q + k] : -INFINITY;
        const float mx = ghm_vmax<Q>(x);
#pragma unroll
        for (int i = 0; i < H; ++i) {
            msg[i].x *= (2 * i < q) ? ex2_approx((x[2 * i] - mx) * 1.4426950408889634f) : 0.f;
            msg[i].y *= (2 * i + 1 < q) ? ex2_approx((x[2 * i + 1] - mx) * 1.4426950408889634f) : 0.f;
        }
        f2_normalize<Q>(msg);
    }
    __syncwarp();
    sstore(0, msg);                                                // beliefs per depth on the current path: stack[depth]

    // =============================== downward pass =============================================
    cj = 0; toff0 = base0; toff1 = base1;
    float* mrow = a.mean + bc * nL;
    int tz = L > 2 ? L - 2 : 0;                                    // ancestors (depth 1..L-2) to refresh before node j
    for (int j = 0; j < n1; ++j) {
        if (L > 2 && cj == 0) {
            // beliefs of the internal nodes of the root path that changed: depths L-1-tz .. L-2 (every s-th node at most)
            for (int l = L - 1 - tz; l <= L - 2; ++l) {
                const int idx = ghm_div_pow(j, L - 1 - l, d);
                const int c = idx - (idx / S) * S;                        // child digit from the uniform index, not the vector odometer
                f2 uv[H], hv[H], bp[H], w[H], tt[H];
                uload(d.edge_off[l] + idx, uv);
                for (int cc = 0; cc < S; ++cc) {
                    f2 uc[H];
                    uload(d.edge_off[l + 1] + idx * S + cc, uc);
#pragma unroll
                    for (int i = 0; i < H; ++i) hv[i] = cc == 0 ? uc[i] : f2_mul(hv[i], uc[i]);
                }
                sload(l - 1, bp);
                f2_cavity<Q>(bp, uv, w);
                f2_matvec_c<Q>(Tdn + ((l - 1) * s + c) * QQ, w, tt);
#pragma unroll
                for (int i = 0; i < H; ++i) hv[i] = f2_mul(hv[i], tt[i]);
                f2_normalize<Q>(hv);
                sstore(l, hv);
            }
        }
        // depth L-1 node j: leaf likelihoods / messages (kept in registers), its belief, then the s leaf marginals
        f2 e[S][H], u[S][H], h[H];
#pragma unroll
        for (int c = 0; c < S; ++c) {
            f2_leaf_like<Q, false>(zrow[j * S + c], a.c2, q, e[c]);
            f2_matvec_c<Q>(Tup + (aa.base_leaf + c) * QQ, e[c], u[c]);
#pragma unroll
            for (int i = 0; i < H; ++i) h[i] = c == 0 ? u[c][i] : f2_mul(h[i], u[c][i]);
        }
        f2 bj[H];
        if (L == 1) {
            sload(0, bj);
        } else {
            f2 uv[H], bp[H], w[H], tt[H];
            uload(d.edge_off[L - 1] + j, uv);
            sload(L - 2, bp);
            f2_cavity<Q>(bp, uv, w);
            f2_matvec_c<Q>(Tdn + toff0 * QQ, w, tt);
            f2_normalize<Q>(h);
#pragma unroll
            for (int i = 0; i < H; ++i) bj[i] = f2_mul(h[i], tt[i]);
            f2_normalize<Q>(bj);
        }
#pragma unroll
        for (int c = 0; c < S; ++c) {
            f2 w[H], tt[H];
            f2_cavity<Q>(bj, u[c], w);
            f2_matvec_c<Q>(Tdn + (aa.base_leaf + c) * QQ, w, tt);
            float num = 0.f, den = 0.f;
#pragma unroll
            for (int i = 0; i < H; ++i) {
                const float b0 = e[c][i].x * tt[i].x, b1 = e[c][i].y * tt[i].y;
                num = fmaf((float)(2 * i), b0, num);
                num = fmaf((float)(2 * i + 1), b1, num);
                den += b0 + b1;
            }
            if (active) mrow[j * S + c] = num / den;
        }
        // advance the odometer; tz = number of depths <= L-2 that change before the next node
        advance();
        tz = 0;
        if (cj == 0) {                                             // depth l changes before node j+1 iff s^(L-1-l) divides j+1
            const int jn = j + 1;
            for (int l = L - 2; l >= 1; --l) {
                if (ghm_div_pow(jn, L - 1 - l, d) * d.spow[L - 1 - l] != jn) break;
                ++tz;
            }
        }
    }
}

template <int Q, int S>
int launch_dns2(const ghm_model* m, const DnsArgs& a, cudaStream_t st) {
    const GhmDev& d = m->d;
    constexpr int NW = 6144;
    const size_t words = (size_t)d.n_mat * Q * Q;
    if (2 * words > (size_t)NW) return GHM_EUNSUP;
    const size_t dyn = (size_t)d.L * (Q / 2) * DNS_NT * sizeof(float2);
    if (dyn > 48 * 1024) return GHM_EUNSUP;
    Dns2Args aa{};
    aa.a = a;
    aa.base0 = (d.L - 2) * d.s; aa.base1 = (d.L - 3) * d.s; aa.base_leaf = (d.L - 1) * d.s; aa.dn_off = (int)words;
    DnsTab<NW> tab;
    tab.v[0] = 0.f;
    memcpy(tab.v, m->h_TlinT, words * sizeof(float));
    memcpy(tab.v + words, m->h_Tlin, words * sizeof(float));
    const unsigned grid = (unsigned)((a.B + DNS_NT - 1) / DNS_NT);
    k_dns2<Q, S, NW><<<grid, DNS_NT, dyn, st>>>(d, aa, tab);
    GHM_CHECK_LAUNCH();
    return GHM_OK;
}

#define GHM_DNS2_DEFINE(Q)                                                                  \
    template int launch_dns2<Q, 2>(const ghm_model*, const DnsArgs&, cudaStream_t);         \
    template int launch_dns2<Q, 3>(const ghm_model*, const DnsArgs&, cudaStream_t);         \
    template int launch_dns2<Q, 4>(const ghm_model*, const DnsArgs&, cudaStream_t);
#define GHM_DNS2_DEFINE_S(Q, S) template int launch_dns2<Q, S>(const ghm_model*, const DnsArgs&, cudaStream_t);
