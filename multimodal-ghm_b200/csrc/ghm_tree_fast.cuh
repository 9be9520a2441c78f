// ghm_tree_fast.cuh -- the software-pipelined fast variant of the fused sampler + root-posterior kernel.
//
// Same contract, same Philox counter layout and same arithmetic as k_tree2 (ghm_tree_kernel.cuh; reference
// GHMTree.gen_values :145-165 and GHMTree.BP_CLS :185-221 of src/ghmclip/data/data_random_GHM.py), specialised
// for what the hot configurations are: translation-invariant tables, s in {2,3,4}, L >= 3, two trees per thread,
// whole leaf rows staged flat, T^T tables in the constant bank.
//
// What is different from k_tree2 (ncu, profiles/r02b_*: the old loop was LATENCY bound -- 3.5 warps per scheduler,
// issue slots 63 % busy, FMA-heavy pipe 54 %, the rest "short scoreboard" / "wait" stalls on the dependent chain
// alias LDS -> alias LDS -> T^T row LDS -> FMUL2 of every node):
//   * the loop is software pipelined over the depth-(L-1) nodes: one iteration holds three independent chains
//     -- the Philox block of node j+2 (IMAD.WIDE chain, FMA-heavy pipe), the draws + leaf-row products of node
//     j+1 (LDS chain, ALU pipe), and the BP climb of node j (packed FFMA2 with constant-bank operands) -- so
//     the leaf-row product of the next node is in registers before the climb that needs it starts;
//   * no odometer: child digits of the rarely taken deep steps come from `j / s^k` by the host's multiply-high
//     magics (uniform datapath), the two hot steps use running counters;
//   * LEAF MEMO (sampling + BP, where ghm_memo_ok(Q, S): the table stays <= 1 MB): the message of a depth-(L-1) node to
//     its parent is a function of (child number, s leaf states) only, tabulated per table upload (GhmDev::leaf_memo, built
//     on the device by k_build_leaf_memo with the same operations, so the posteriors are bit-identical).  The node then
//     costs ONE gathered row (L1 / L2 resident, fetched a full iteration ahead into the other half of a register
//     ping-pong, loop unrolled by two) instead of s shared-memory row gathers, the row product, a rescale and a q x q
//     matvec: 27 of the 40 matvecs of an L = 4, s = 3 tree disappear.  These variants run ONE tree per thread at 7 (q <= 10)
//     or 6 CTAs per SM (FastCfg) and flush int64 leaves through cp.async.bulk from a ring of 1 KB buffers carved from the
//     warp's accumulator span.  Given-leaves BP keeps the row products (no Philox work to hide the gathers behind).
//     Measurements and what was tried on the way: DESIGN.md 3.1, round 2b.
#pragma once

// Flush of one warp's staged leaf bytes st[0 .. n) (n a multiple of 32) as n contiguous int64 at dst (16-byte aligned)
// through a ring of nb (1 .. 4) staging buffers of 1 KB carved from `buf` (16-byte aligned, owned by this warp).  A round
// moves 128 leaves: every lane expands two byte pairs with one conflict-free STS.128 each, lane 0 hands the buffer to
// the bulk-copy engine and then waits until the ring has a free buffer again, so nb - 1 copies stay in flight.
__device__ __forceinline__ void bulk_wait_read(int pending) {    // cp.async.bulk.wait_group.read takes an immediate
    if (pending <= 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    else if (pending == 1) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    else if (pending == 2) asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
    else asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
}
__device__ __forceinline__ void stage_flush_bulk_i64(const uint8_t* st, int n, int64_t* dst, unsigned char* buf, int nb, int lane) {
    const uint32_t sbuf = (uint32_t)__cvta_generic_to_shared(buf);
    const uint16_t* src = reinterpret_cast<const uint16_t*>(st) + lane;
    unsigned char* my = buf + 16 * lane;
    const int rounds = n >> 7, tail = n & 127;                   // tail: 0, 32, 64 or 96 leaves
    int b = 0;
    __syncwarp();
    for (int r = 0; r < rounds; ++r) {
        const uint32_t w0 = src[0], w1 = src[32];
        *reinterpret_cast<uint4*>(my + b * 1024) = make_uint4(w0 & 255u, 0u, w0 >> 8, 0u);
        *reinterpret_cast<uint4*>(my + b * 1024 + 512) = make_uint4(w1 & 255u, 0u, w1 >> 8, 0u);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 1024;\n\tcp.async.bulk.commit_group;"
                         :: "l"(dst), "r"(sbuf + (uint32_t)b * 1024u) : "memory");
            bulk_wait_read(nb - 1);
        }
        __syncwarp();
        src += 64; dst += 128;
        b = b + 1 == nb ? 0 : b + 1;
    }
    if (tail) {
        if (2 * lane < tail) {
            const uint32_t w0 = src[0];
            *reinterpret_cast<uint4*>(my + b * 1024) = make_uint4(w0 & 255u, 0u, w0 >> 8, 0u);
        }
        if (2 * (lane + 32) < tail) {
            const uint32_t w1 = src[32];
            *reinterpret_cast<uint4*>(my + b * 1024 + 512) = make_uint4(w1 & 255u, 0u, w1 >> 8, 0u);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0)
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n\tcp.async.bulk.commit_group;"
                         :: "l"(dst), "r"(sbuf + (uint32_t)b * 1024u), "r"(tail * 8) : "memory");
    }
    if (lane == 0) bulk_wait_read(0);                            // the buffers are shared memory of a CTA about to exit
    __syncwarp();
}

// Shape of a variant.  With the leaf memo the kernel is short of warps, not of issue slots (ncu r02z: 4 warps per
// scheduler, issue 0.48, every pipe under 0.36), and what two trees per thread used to share -- the constant-bank table
// loads of the leaf-level matvec -- is gone: ONE tree per thread in 70-80 registers and half the shared memory puts 7
// (q <= 10) or 6 CTAs on an SM instead of 4 (measured, L4 s3 q10: 0.1524 -> 0.1390 ms).  Without the memo the two-tree
// form stays (one tree per thread measured equal there, and 20 % slower for sampling only: DESIGN.md).
template <int Q, int S, int MODE, bool BP>
struct FastCfg {
    static constexpr bool MEMO = BP && MODE == MODE_PHILOX && ghm_memo_ok(Q, S);
    static constexpr bool RING = !BP && MODE == MODE_PHILOX;      // sampling only: same form, the "accumulator span" is just the flush ring
    static constexpr int TPT = MEMO ? 1 : 2;
    static constexpr int CTAS = MEMO ? (Q <= 10 ? 7 : 6) : 4;
};

template <int Q, int S, int MODE, bool BP, int NW, bool BLK>
__global__ void __launch_bounds__(T2_NT, (FastCfg<Q, S, MODE, BP>::CTAS))
k_tree_fast(const __grid_constant__ GhmDev d, const __grid_constant__ TreeArgs a, const __grid_constant__ TabParam<NW> tab) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int TPT = FastCfg<Q, S, MODE, BP>::TPT, NT = T2_NT, H = Q / 2, QS = (Q + 3) / 4 * 4, WTREES = 32 * TPT;
    constexpr bool SPARE = (S & 3) != 0;                         // node j is drawn from the spare word of its leaf block
    constexpr bool PHILOX = MODE == MODE_PHILOX;
    constexpr bool MEMO = FastCfg<Q, S, MODE, BP>::MEMO, RING = FastCfg<Q, S, MODE, BP>::RING;
    // Shared-memory stride of the gathered leaf rows T_c^T[x, :].  Each lane reads the row of ITS leaf state, so the
    // loads are true gathers: with 48-byte rows read as LDS.128 + LDS.128 + LDS.64 the ten rows of q = 10 fall on eight
    // 16-byte bank groups (rows 0/8 and 1/9 collide: 0.5 extra wavefronts per load, ncu r02k); with 40-byte rows read
    // as five LDS.64 the ten rows start on ten different 8-byte bank groups (10 x mod 32 = 0,10,20,30,8,18,28,6,16,26).
    // Measured (B = 327 680, L4 s3 q10): BP on given leaves 0.122 -> 0.117 ms, but the fused Philox variant 0.166 -> 0.175 ms
    // (the extra LDS.64 issue slots and 3 more registers cost more there than the conflicts), so only MODE_GIVEN uses it.
    constexpr int LS = (Q % 4 == 2 && MODE == MODE_GIVEN) ? Q : QS;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int L = d.L, q = d.q, nL = d.n_leaves;                 // L >= 3 (host)
    const int64_t warp_tree0 = ((int64_t)blockIdx.x * T2_WARPS + warp) * WTREES;
    const bool warp_has_work = warp_tree0 < a.B;

    // ---- carve shared memory (same layout as k_tree2) -------------------------------------
    size_t off = 0;
    const float* tt_leaf0 = nullptr;
    const uint32_t* AL = d.alias;
    if (BP && !MEMO) {                                           // leaf-level T^T rows are gathered per lane by leaf state
        const int first = d.mat_off[L];
        const int words = S * Q * LS;
        float* s1 = reinterpret_cast<float*>(smem + off); off += ((size_t)words * 4 + 15) / 16 * 16;
        const float* src = d.TTp + (size_t)first * Q * QS;
        for (int i = tid; i < words; i += NT) s1[i] = src[(i / LS) * QS + (i % LS)];
        tt_leaf0 = s1;
    }
    if (PHILOX) {
        const int words = d.n_mat * q * q;
        uint32_t* s2 = reinterpret_cast<uint32_t*>(smem + off); off += ((size_t)words * 4 + 15) / 16 * 16;
        for (int i = tid; i < words; i += NT) s2[i] = d.alias[i];
        AL = s2;
    }
    const int n_deep = L - 2;                                    // ancestors kept in shared memory: depths 0 .. L-3
    // Parked accumulators.  Memoised variants: [warp][n_deep][H][TPT][32], a warp's accumulators are one contiguous span
    // (host: padded to whole 1 KB buffers) that doubles as the ring of the bulk-store flush once the last climb is done.
    // The others keep [n_deep][H][TPT][NT] carved from n_deep: with the per-warp form (or merely a carve that depends on a
    // kernel argument) ptxas moved one of their two hot matvecs from LDCU.64 / uniform registers to LDC.64 and the
    // given-leaves BP lost 12 % (0.115 -> 0.128 ms; SASS: 213 -> 113 LDCU.64).
    constexpr int AST = MEMO ? 32 : NT;                          // stride between the (i, t) rows of one level
    const int acc_warp = (MEMO || RING) ? a.acc_stride : 0;      // f2 elements per warp
    f2* ACC = reinterpret_cast<f2*>(smem + off);                 // (+ acc_tid at the use sites, AFTER the uniform offsets:
    const int acc_tid = MEMO ? warp * acc_warp + lane : tid;     //  the order decides LDCU vs LDC as well)
    if constexpr (MEMO || RING) off += (size_t)acc_warp * T2_WARPS * sizeof(f2);
    else if (BP) off += (size_t)n_deep * H * TPT * NT * sizeof(f2);
    const int n_rng = n_deep + (SPARE ? 0 : 1);                  // levels 1 .. L-2 (+ L-1 when it has no spare word)
    uint32_t* RNG = reinterpret_cast<uint32_t*>(smem + off);     // [n_rng][3][TPT][NT]  words 1..3 of the cached Philox blocks
    if (PHILOX) off += (size_t)n_rng * 3 * TPT * NT * 4;
    uint8_t* VAL = smem + off;                                   // [n_deep][TPT][NT] states of the path nodes at depths 0 .. L-3
    if (PHILOX) off += ((size_t)n_deep * TPT * NT + 15) / 16 * 16;
    const bool use_stage = (a.leaves != nullptr);
    uint8_t* stage = smem + off + (size_t)warp * a.stage_bytes;
    __syncthreads();
    if (!warp_has_work) return;

    const int n1 = d.spow[L - 1];
    bool active[TPT];
    uint64_t tree[TPT];
    int64_t bt[TPT];
    int srow[TPT];
#pragma unroll
    for (int t = 0; t < TPT; ++t) {
        bt[t] = warp_tree0 + 32 * t + lane;
        active[t] = bt[t] < a.B;
        const int64_t bc = active[t] ? bt[t] : a.B - 1;          // tail threads shadow the last tree and never write
        tree[t] = a.tree_offset + (uint64_t)bc;
        // (a template flag, not a runtime branch: the branch alone cost 3 registers and 3-5 % of the unblocked launch)
        if constexpr (BLK) tree[t] += (uint64_t)((uint32_t)bc / a.blk_len) * a.blk_extra;
        srow[t] = (32 * t + lane) * nL;
    }
    if (!PHILOX) stage_load_flat(stage, WTREES, a.leaves, a.leaf_dtype, warp_tree0, a.B, nL, lane, q, d.status);

    // ---- root ---------------------------------------------------------------------------
    int xpar[TPT];                                               // state of the depth-(L-2) parent of the node being drawn
#pragma unroll
    for (int t = 0; t < TPT; ++t) xpar[t] = 0;
    if (PHILOX) {
#pragma unroll
        for (int t = 0; t < TPT; ++t) {
            const int64_t bc = active[t] ? bt[t] : a.B - 1;
            int x0;
            if (a.root_mode == GHM_ROOT_GIVEN && bc < a.n_given) {
                int64_t r = a.root_in[bc];
                if (r < 0 || r >= q) { atomicOr(d.status, 1); r = r < 0 ? 0 : q - 1; }
                x0 = (int)r;
            } else {
                const bool shared = a.root_mode == GHM_ROOT_SHARED && bc < a.n_given;
                const uint4 rb = ghm_rng_block(shared ? a.root_seed : a.seed, tree[t], 0u, 0u, GHM_STREAM_TREE);
                const uint32_t* rc = a.root_mode == GHM_ROOT_PRIOR ? d.root_cdfu_prior : d.root_cdfu_unif;
                int cnt = 0;
                for (int k = 0; k < q - 1; ++k) cnt += (rb.x >= __ldg(rc + k)) ? 1 : 0;
                x0 = cnt;
            }
            VAL[t * NT + tid] = (uint8_t)x0;
            if (a.root_out && active[t]) a.root_out[bt[t]] = x0;
        }
    }

    const uint32_t* alias_leaf0 = AL + (size_t)d.mat_off[L] * q * q;
    const int base0 = a.base0, base1 = a.base1;                  // (L-2)*s, (L-3)*s: constant-bank matrix index of the two hottest climb steps

    // (re)draw the ancestors of node jn at depths 1 .. L-2 that changed: depth l changes iff s^(L-1-l) divides jn
    auto redraw_ancestors = [&](int jn) {
        for (int l = 1; l <= L - 2; ++l) {
            const int idx = ghm_div_pow(jn, L - 1 - l, d);
            if (idx * d.spow[L - 1 - l] != jn) continue;
            const int c = idx - (idx / S) * S;
            const uint32_t* arow = AL + (size_t)((l - 1) * S + c) * q * q;
#pragma unroll
            for (int t = 0; t < TPT; ++t) {
                uint32_t* rl = RNG + (size_t)((l - 1) * 3 * TPT + t) * NT + tid;
                uint32_t r;
                if ((idx & 3) == 0) {
                    const uint4 rb = ghm_rng_block(a.seed, tree[t], (uint32_t)l, (uint32_t)(idx >> 2), GHM_STREAM_TREE);
                    rl[0] = rb.y; rl[TPT * NT] = rb.z; rl[2 * TPT * NT] = rb.w;
                    r = rb.x;
                } else {
                    r = rl[((idx & 3) - 1) * TPT * NT];
                }
                const int xp = VAL[((l - 1) * TPT + t) * NT + tid];
                const int x = ghm_draw_alias(arow + xp * q, r, q);
                if (l < L - 2) VAL[(l * TPT + t) * NT + tid] = (uint8_t)x; else xpar[t] = x;
            }
        }
    };

    // draw node jn (child cjn of its parent) and its S leaves, stage the leaves, and form the product of the leaf
    // rows hout = prod_c T_c^T[x_c, :]  (reference :191-196 in the linear domain)
    auto sample_node = [&](int jn, int cjn, const uint4 (&rb)[TPT], f2 (&hout)[TPT][H]) {
        int xc[TPT];
        if (PHILOX) {
            const uint32_t* arow = AL + (size_t)(base0 + cjn) * q * q;
#pragma unroll
            for (int t = 0; t < TPT; ++t) {
                uint32_t r;
                if (SPARE) {
                    r = ghm_pick(rb[t], S & 3);
                } else {                                         // s == 4: word jn & 3 of block (level L-1, jn >> 2)
                    uint32_t* rl = RNG + (size_t)((L - 2) * 3 * TPT + t) * NT + tid;
                    if ((jn & 3) == 0) {
                        const uint4 rj = ghm_rng_block(a.seed, tree[t], (uint32_t)(L - 1), (uint32_t)(jn >> 2), GHM_STREAM_TREE);
                        rl[0] = rj.y; rl[TPT * NT] = rj.z; rl[2 * TPT * NT] = rj.w;
                        r = rj.x;
                    } else {
                        r = rl[((jn & 3) - 1) * TPT * NT];
                    }
                }
                xc[t] = ghm_draw_alias(arow + xpar[t] * q, r, q);
            }
        }
        int mrow[TPT];                                           // MEMO: ((cjn * q + x_0) * q + x_1) * q + ...
#pragma unroll
        for (int t = 0; t < TPT; ++t) mrow[t] = cjn;
#pragma unroll
        for (int c = 0; c < S; ++c) {
#pragma unroll
            for (int t = 0; t < TPT; ++t) {
                int x;
                if (PHILOX) {
                    x = ghm_draw_alias(alias_leaf0 + (c * q + xc[t]) * q, ghm_pick(rb[t], c), q);
                    if (use_stage) stage[srow[t] + jn * S + c] = (uint8_t)x;
                } else {
                    x = stage[srow[t] + jn * S + c];
                }
                if constexpr (MEMO) {
                    mrow[t] = mrow[t] * q + x;
                } else if (BP) {
                    f2 row[H];
                    if constexpr (LS == Q) {
                        const f2* rp = reinterpret_cast<const f2*>(tt_leaf0 + (c * Q + x) * LS);
#pragma unroll
                        for (int i = 0; i < H; ++i) row[i] = rp[i];
                    } else {
                        f2_load_row<Q>(tt_leaf0 + (c * Q + x) * QS, row);
                    }
#pragma unroll
                    for (int i = 0; i < H; ++i) hout[t][i] = c == 0 ? row[i] : f2_mul(hout[t][i], row[i]);
                }
            }
        }
        if constexpr (MEMO) {                                    // one row of the memo per tree: first used one iteration later
#pragma unroll
            for (int t = 0; t < TPT; ++t) f2_ldg_row<Q>(d.leaf_memo + (size_t)(unsigned)mrow[t] * QS, hout[t]);
        }
    };

    f2 msg[TPT][H], accT[TPT][H];
#pragma unroll
    for (int t = 0; t < TPT; ++t)
#pragma unroll
        for (int i = 0; i < H; ++i) { msg[t][i] = make_float2(0.f, 0.f); accT[t][i] = make_float2(0.f, 0.f); }

    // One climb step: u = T m for both trees, times the parked product of the earlier siblings; park it again (more
    // siblings to come) or rescale and keep climbing.  A == nullptr: register accumulator.
    auto climb_step = [&](const float* __restrict__ Tm, bool has_prev, bool last, f2* A) -> bool {
        f2 u[TPT][H];
        if constexpr (TPT == 2) f2_matvec_up2<Q, QS>(Tm, msg[0], msg[1], u[0], u[1]);
        else f2_matvec_up1<Q, QS>(Tm, msg[0], u[0]);
        if (has_prev) {
#pragma unroll
            for (int t = 0; t < TPT; ++t)
#pragma unroll
                for (int i = 0; i < H; ++i) u[t][i] = f2_mul(u[t][i], A ? A[(i * TPT + t) * AST] : accT[t][i]);
        }
        if (!last) {
#pragma unroll
            for (int t = 0; t < TPT; ++t)
#pragma unroll
                for (int i = 0; i < H; ++i) {
                    if (A) A[(i * TPT + t) * AST] = u[t][i]; else accT[t][i] = u[t][i];
                }
            return false;
        }
#pragma unroll
        for (int t = 0; t < TPT; ++t) {
#pragma unroll
            for (int i = 0; i < H; ++i) msg[t][i] = u[t][i];
            f2_normalize<Q>(msg[t]);
        }
        return true;
    };

    // the BP climb of node j, whose leaf-row product is h
    auto climb = [&](int j, int cj, int c1, const f2 (&h)[TPT][H]) {
        bool up;
        if constexpr (MEMO) {                                    // h IS the message to the parent: fold it into the parked product
            up = cj == S - 1;                                    // (uniform branches: selects cost 2 S q issue slots per node)
            if (cj == 0) {
#pragma unroll
                for (int t = 0; t < TPT; ++t)
#pragma unroll
                    for (int i = 0; i < H; ++i) accT[t][i] = h[t][i];
            } else if (!up) {
#pragma unroll
                for (int t = 0; t < TPT; ++t)
#pragma unroll
                    for (int i = 0; i < H; ++i) accT[t][i] = f2_mul(h[t][i], accT[t][i]);
            } else {
#pragma unroll
                for (int t = 0; t < TPT; ++t) {
#pragma unroll
                    for (int i = 0; i < H; ++i) msg[t][i] = f2_mul(h[t][i], accT[t][i]);
                    f2_normalize<Q>(msg[t]);
                }
            }
        } else {
#pragma unroll
            for (int t = 0; t < TPT; ++t) {
#pragma unroll
                for (int i = 0; i < H; ++i) msg[t][i] = h[t][i];
                f2_normalize<Q>(msg[t]);
            }
            // the two hottest steps (every node / every s-th node) read their table through running offsets that feed
            // nothing but the constant-bank address, so they stay in UNIFORM registers (LDCU.64 + FFMA2 with a UR operand)
            up = climb_step(tab.v + (base0 + cj) * (Q * QS), cj != 0, cj == S - 1, nullptr);
        }
        if (up) up = climb_step(tab.v + (base1 + c1) * (Q * QS), c1 != 0, c1 == S - 1, ACC + (size_t)(L - 3) * H * TPT * AST + acc_tid);
        if (up && L >= 4) {
            f2* A = ACC + (size_t)(L - 4) * H * TPT * AST + acc_tid;
            for (int l = L - 3; l > 0; --l) {
                const int idx = ghm_div_pow(j, L - 1 - l, d);    // index of the path node at depth l
                const int c = idx - (idx / S) * S;
                if (!climb_step(tab.v + ((l - 1) * S + c) * (Q * QS), c != 0, c == S - 1, A)) break;
                A -= H * TPT * AST;
            }
        }
    };

    // ---- prologue: ancestors + node 0 ---------------------------------------------------
    uint4 rbA[TPT], rbB[TPT];                                    // Philox blocks of the leaves of nodes j+1 / j+2
    f2 hcur[TPT][H], hnext[TPT][H];
    if (PHILOX) {
        redraw_ancestors(0);
#pragma unroll
        for (int t = 0; t < TPT; ++t) {
            rbB[t] = ghm_rng_block(a.seed, tree[t], (uint32_t)L, 0u, GHM_STREAM_TREE);
            rbA[t] = ghm_rng_block(a.seed, tree[t], (uint32_t)L, 1u, GHM_STREAM_TREE);   // block of node 1
        }
    }
    sample_node(0, 0, rbB, hcur);

    int cj = 0, c1 = 0;                                          // j mod s, (j / s) mod s
    // One straight-line block: Philox of node j+2 (-> rbN), draws + leaf rows of node j+1 (rbU -> hL), BP climb of node j (hU).
    // (measured: guarding the sampling half with `jn < n1` instead of peeling the last climb costs 4 %)
    if constexpr (MEMO) {                                        // unrolled by two: the buffers swap roles, no register copies
        auto body = [&](int j, int cjn, uint4 (&rbU)[TPT], uint4 (&rbN)[TPT], f2 (&hU)[TPT][H], f2 (&hL)[TPT][H]) {
            const int jn = j + 1;
            if (cjn == 0) redraw_ancestors(jn);
#pragma unroll
            for (int t = 0; t < TPT; ++t)                        // (one block past the end is computed and dropped)
                rbN[t] = ghm_rng_block(a.seed, tree[t], (uint32_t)L, (uint32_t)(jn + 1), GHM_STREAM_TREE);
            // consume the prefetched row BEFORE the next gathers are issued: a wait placed behind them shares their
            // scoreboard and sits out a full L2 round trip (ncu r02x: 29 % of the stall samples on the first use)
            climb(j, cj, c1, hU);
            sample_node(jn, cjn, rbU, hL);
        };
        auto advance = [&](int cjn) {
            cj = cjn;
            if (cjn == 0) c1 = c1 + 1 == S ? 0 : c1 + 1;
        };
        int j = 0;
        for (; j + 2 <= n1 - 1; j += 2) {
            int cjn = cj + 1 == S ? 0 : cj + 1;
            body(j, cjn, rbA, rbB, hcur, hnext);
            advance(cjn);
            cjn = cj + 1 == S ? 0 : cj + 1;
            body(j + 1, cjn, rbB, rbA, hnext, hcur);
            advance(cjn);
        }
        if constexpr (S % 2 == 0) {                              // n1 = s^(L-1) even: one node pair is left (odd s: none)
            const int cjn = cj + 1 == S ? 0 : cj + 1;
            body(j, cjn, rbA, rbB, hcur, hnext);
            advance(cjn);
            climb(n1 - 1, cj, c1, hnext);
        } else {
            climb(n1 - 1, cj, c1, hcur);
        }
    } else {
        for (int j = 0; j < n1 - 1; ++j) {
            const int jn = j + 1;
            const int cjn = cj + 1 == S ? 0 : cj + 1;
            if (PHILOX && cjn == 0) redraw_ancestors(jn);
            if (PHILOX) {
#pragma unroll
                for (int t = 0; t < TPT; ++t)                    // (one block past the end is computed and dropped)
                    rbB[t] = ghm_rng_block(a.seed, tree[t], (uint32_t)L, (uint32_t)(jn + 1), GHM_STREAM_TREE);
            }
            sample_node(jn, cjn, rbA, hnext);
            if (BP) climb(j, cj, c1, hcur);
#pragma unroll
            for (int t = 0; t < TPT; ++t) {
                rbA[t] = rbB[t];
#pragma unroll
                for (int i = 0; i < H; ++i) hcur[t][i] = hnext[t][i];
            }
            cj = cjn;
            if (cjn == 0) c1 = c1 + 1 == S ? 0 : c1 + 1;
        }
        if (BP) climb(n1 - 1, cj, c1, hcur);
    }

    if (PHILOX && use_stage) {
        // int64 leaves of a full tile leave through the bulk-copy engine: the warp expands its staged bytes into a ring of
        // 1 KB buffers (memoised variants: its accumulator span, dead by now; sampling only: 3 KB carved for the purpose)
        // with conflict-free STS.128 and one lane hands each filled buffer to cp.async.bulk.  The STG.128 loop it replaces ran one store at a time -- the next store's data
        // registers were the previous store's, and each wait was a round trip of the busy LSU queue (ncu r02y: 28 % of
        // the kernel's stall samples) -- and every CTA of a wave reached it at the same moment.
        if ((MEMO || RING) && a.leaf_dtype == GHM_LEAF_I64 && warp_tree0 + WTREES <= a.B && acc_warp * (int)sizeof(f2) >= 1024)
            stage_flush_bulk_i64(stage, WTREES * nL, reinterpret_cast<int64_t*>(a.leaves) + warp_tree0 * nL,
                                 reinterpret_cast<unsigned char*>(ACC + warp * acc_warp), min(acc_warp * (int)sizeof(f2) / 1024, 4), lane);
        else
            stage_flush_flat(stage, WTREES, a.leaves, a.leaf_dtype, warp_tree0, a.B, nL, lane);
    }

    // ---- root outputs (reference :213-217; root_node.hd_message is the shifted log-likelihood) --
    if (BP) {
#pragma unroll
        for (int t = 0; t < TPT; ++t) {
            if (!active[t]) continue;
            const int64_t b = bt[t];
            if (a.root_hd) {
#pragma unroll
                for (int k = 0; k < Q; ++k)
                    if (k < q) a.root_hd[b * q + k] = logf(f2_elem<Q>(msg[t], k));
            }
            if (a.post) {
                float w[Q], sum = 0.f;
#pragma unroll
                for (int k = 0; k < Q; ++k) { w[k] = f2_elem<Q>(msg[t], k) * __ldg(d.py + k); sum += w[k]; }
                const float inv = 1.0f / sum;
#pragma unroll
                for (int k = 0; k < Q; ++k)
                    if (k < q) a.post[b * q + k] = w[k] * inv;
            }
        }
    }
}

// host side: same shared-memory carve as the kernel above
template <int Q, int S, int MODE, bool BP, int NW>
static int launch_tree_fast(const ghm_model* m, const TreeArgs& a0, cudaStream_t st) {
    const GhmDev& d = m->d;
    TreeArgs a = a0;
    constexpr int QS = (Q + 3) / 4 * 4, TPT = FastCfg<Q, S, MODE, BP>::TPT, WTREES = 32 * TPT;
    const int n_deep = d.L - 2;
    size_t dyn = 0;
    constexpr bool MEMO = FastCfg<Q, S, MODE, BP>::MEMO;
    if (MEMO && !d.leaf_memo) return ghm_fail(GHM_EUNSUP, "internal: leaf memo missing (L=%d s=%d q=%d)", d.L, d.s, d.q);
    if (BP && !MEMO) dyn += ((size_t)S * Q * ((Q % 4 == 2 && MODE == MODE_GIVEN) ? Q : QS) * 4 + 15) / 16 * 16;
    if (MODE == MODE_PHILOX) dyn += ((size_t)d.n_mat * d.q * d.q * 4 + 15) / 16 * 16;
    constexpr bool RING = FastCfg<Q, S, MODE, BP>::RING;
    a.acc_stride = BP ? n_deep * (Q / 2) * TPT * 32 : 0;          // f2 per warp; memoised int64 sampling: whole 1 KB flush buffers (<= 4)
    if (MEMO && a.leaves && a.leaf_dtype == GHM_LEAF_I64 && a.acc_stride * 8 < 4096)
        a.acc_stride = (a.acc_stride * 8 + 1023) / 1024 * 128;
    if (RING && a.leaves && a.leaf_dtype == GHM_LEAF_I64) a.acc_stride = 3 * 128;   // sampling only: a ring of three buffers
    if (BP || RING) dyn += (size_t)a.acc_stride * T2_WARPS * sizeof(float2);
    if (MODE == MODE_PHILOX) {
        const int n_rng = n_deep + ((S & 3) != 0 ? 0 : 1);
        dyn += (size_t)n_rng * 3 * TPT * T2_NT * 4 + ((size_t)n_deep * TPT * T2_NT + 15) / 16 * 16;
    }
    a.chunk_j = d.spow[d.L - 1]; a.stage_stride = 0; a.stage_bytes = 0;
    a.base0 = (d.L - 2) * d.s; a.base1 = (d.L - 3) * d.s;
    if (a.leaves) {
        a.stage_stride = d.n_leaves;
        a.stage_bytes = (int)(((size_t)WTREES * a.stage_stride + 15) / 16 * 16);
    }
    dyn += (size_t)a.stage_bytes * T2_WARPS;
    if (dyn > 200 * 1024)
        return ghm_fail(GHM_EUNSUP, "tree kernel needs %zu bytes of shared memory (L=%d s=%d q=%d)", dyn, d.L, d.s, d.q);
    const int64_t trees_per_cta = (int64_t)T2_WARPS * WTREES;
    const unsigned grid = (unsigned)((a.B + trees_per_cta - 1) / trees_per_cta);
    TabParam<NW> tab;                                              // by-value table parameter (copied at launch)
    tab.v[0] = 0.f;
    if (BP) {
        const size_t words = (size_t)d.n_mat * Q * QS;
        if (words > (size_t)NW) return ghm_fail(GHM_EUNSUP, "internal: constant table overflow");
        memcpy(tab.v, m->h_TTp, words * sizeof(float));
    }
    auto go = [&](auto kern) -> int {
        GHM_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
        kern<<<grid, T2_NT, dyn, st>>>(d, a, tab);
        GHM_CHECK_LAUNCH();
        return GHM_OK;
    };
    if constexpr (MODE == MODE_PHILOX) {
        if (a.blk_len) return go(k_tree_fast<Q, S, MODE, BP, NW, true>);      // a shard of a block-structured batch
    }
    return go(k_tree_fast<Q, S, MODE, BP, NW, false>);
}
