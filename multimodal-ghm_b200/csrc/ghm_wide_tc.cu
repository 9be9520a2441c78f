// ghm_wide_tc.cu -- tcgen05 (5th-gen tensor core) variant of the wide path's batched row-GEMM.
//
//   Y[node][b][n] = sum_k X[node][b][k] * Wk[mat(node)][n][k]        (both operands K-major)
//
// i.e. the [B*nodes, q] x [q, q] contraction of the north star, used only when q is large (QW in
// {64, 128, 192, 256}) and the model's gemm mode asks for it (GHM_GEMM_TF32 / GHM_GEMM_BF16; the default is
// the FP32 CUDA-core kernel in ghm_wide.cu).  Child -> parent (`T @ m`, reference :207,497): Wk = Wup
// (row a, contiguous b); parent -> child (`T.T @ m`, :513): Wk = Wdn.
//
// One CTA = one 128-tree x QW output tile of one node:
//   * operands are staged in shared memory in the canonical K-major SWIZZLE_128B layout (rows of 128 bytes,
//     8-row / 1024-byte atoms, 16-byte chunk index XOR row%8) by all 128 threads -- TF32 takes the FP32
//     messages as they are, BF16 converts on the fly -- double/triple buffered over 128-byte K chunks;
//   * one elected thread issues tcgen05.mma (M = 128, N = QW, K = 8 (tf32) / 16 (bf16) per instruction, four
//     per chunk, descriptor start address advanced by 32 bytes inside the swizzle atom), accumulating FP32 in
//     TMEM; tcgen05.commit arrives on the stage's mbarrier so the stage can be refilled while later MMAs run;
//   * epilogue: each warp reads its 32 TMEM lanes with tcgen05.ld.32x32b.x32 and writes full 128-byte row
//     segments of Y.
// k_wide_gemm_tma (TF32): the same tile computed by a warp-specialised CTA fed by TMA -- warp 0 issues
// cp.async.bulk.tensor (SWIZZLE_128B tensor maps over the message matrix and the weight table, 128-byte K chunks)
// into a 2-4 stage ring guarded by full/empty mbarriers, warp 1 issues the MMAs and frees stages with tcgen05.commit,
// warps 2-5 drain TMEM.  No thread touches the operands: the FP32 bits go HBM/L2 -> shared memory -> tensor core.
// k_wide_gemm_tc (BF16, and the TF32 fallback) stages operands with ordinary loads because BF16 needs the FP32 ->
// BF16 conversion in flight.
// The combine / cavity / belief row kernels around the GEMM are shared with the FP32 path.
#include "ghm_wide.cuh"

#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

#define TC_THREADS 128
#define TC_M 128
#define TC_MAX_STAGES 3

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra.uni WAIT_DONE;\n"
        "bra.uni WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 format: version 1, SBO = 1024 B between 8-row atoms)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    uint64_t desc = 0;
    desc |= (uint64_t)((saddr >> 4) & 0x3FFF);            // start address, bits [0,14)
    desc |= (uint64_t)0 << 16;                             // leading byte offset (unused: K extent <= one swizzle row)
    desc |= (uint64_t)(1024 >> 4) << 32;                   // stride byte offset, bits [32,46)
    desc |= (uint64_t)1 << 46;                             // descriptor version (Blackwell)
    desc |= (uint64_t)2 << 61;                             // layout type SWIZZLE_128B
    return desc;
}

template <int KIND>
__device__ __forceinline__ void umma_issue(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if constexpr (KIND == GHM_GEMM_TF32) {
        asm volatile(
            "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    } else {
        asm volatile(
            "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    }
}

// store one 128-byte operand row (eight 16-byte chunks) at row r of a SWIZZLE_128B tile
__device__ __forceinline__ void st_row_sw128(unsigned char* tile, int r, const uint4 (&c)[8]) {
    unsigned char* row = tile + (size_t)r * 128;
    const int x = r & 7;
#pragma unroll
    for (int j = 0; j < 8; ++j) *reinterpret_cast<uint4*>(row + ((j ^ x) << 4)) = c[j];
}

// fetch the 128-byte operand row for K chunk kc from an FP32 source row (null -> zeros)
template <int KIND>
__device__ __forceinline__ void load_row(const float* __restrict__ src, int kc, uint4 (&c)[8]) {
    if (!src) {
#pragma unroll
        for (int j = 0; j < 8; ++j) c[j] = make_uint4(0, 0, 0, 0);
        return;
    }
    if constexpr (KIND == GHM_GEMM_TF32) {                  // 32 floats, used as they are (tf32 = the top 19 bits)
        const uint4* p = reinterpret_cast<const uint4*>(src + kc * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) c[j] = __ldg(p + j);
    } else {                                                // 64 floats -> 64 bf16
        const float4* p = reinterpret_cast<const float4*>(src + kc * 64);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 a = __ldg(p + 2 * j), b = __ldg(p + 2 * j + 1);
            __nv_bfloat162 v0 = __floats2bfloat162_rn(a.x, a.y), v1 = __floats2bfloat162_rn(a.z, a.w);
            __nv_bfloat162 v2 = __floats2bfloat162_rn(b.x, b.y), v3 = __floats2bfloat162_rn(b.z, b.w);
            c[j] = make_uint4(*reinterpret_cast<uint32_t*>(&v0), *reinterpret_cast<uint32_t*>(&v1),
                              *reinterpret_cast<uint32_t*>(&v2), *reinterpret_cast<uint32_t*>(&v3));
        }
    }
}

template <int KIND>
__global__ void __launch_bounds__(TC_THREADS) k_wide_gemm_tc(const GhmDev d, int64_t B, int level, int down, int stages,
                                                             const float* __restrict__ X, float* __restrict__ Y) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bars[TC_MAX_STAGES];
    __shared__ uint32_t tmem_base_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N = d.QW;                                     // 64 / 128 / 192 / 256
    const int node = blockIdx.y;
    const int64_t m0 = (int64_t)blockIdx.x * TC_M;
    const int mi = d.mat_off[level] + (d.ti ? node - ghm_div_s(node, d) * d.s : node);
    const float* Wk = (down ? d.Wdn : d.Wup) + (size_t)mi * N * N;            // [n][k]
    const float* Xn = X + (int64_t)node * B * N;
    float* Yn = Y + (int64_t)node * B * N;

    unsigned char* tiles = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int a_bytes = TC_M * 128, b_bytes = N * 128, stage_bytes = a_bytes + b_bytes;
    const int kchunk = KIND == GHM_GEMM_TF32 ? 32 : 64;
    const int nchunks = N / kchunk;
    const uint32_t tmem_cols = N <= 64 ? 64 : (N <= 128 ? 128 : 256);

    if (tid == 0) {
        for (int i = 0; i < TC_MAX_STAGES; ++i) mbar_init(&bars[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                     "r"(tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_slot;

    // instruction descriptor: D = F32, A/B = TF32 or BF16, both K-major, N, M = 128
    const uint32_t fmt = KIND == GHM_GEMM_TF32 ? 2u : 1u;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);

    const int64_t arow = m0 + tid;                          // thread t stages A row t and B rows t, t+128
    const float* asrc = arow < B ? Xn + arow * N : nullptr;

    for (int kc = 0; kc < nchunks; ++kc) {
        const int st = kc % stages;
        unsigned char* At = tiles + (size_t)st * stage_bytes;
        unsigned char* Bt = At + a_bytes;
        if (kc >= stages) mbar_wait(&bars[st], ((kc / stages) - 1) & 1);      // the MMAs that read this stage are done
        uint4 c[8];
        load_row<KIND>(asrc, kc, c);
        st_row_sw128(At, tid, c);
        for (int n = tid; n < N; n += TC_THREADS) {
            load_row<KIND>(Wk + (size_t)n * N, kc, c);
            st_row_sw128(Bt, n, c);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // generic-proxy writes -> async proxy (UMMA)
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint64_t adesc = umma_desc_sw128(smem_u32(At)), bdesc = umma_desc_sw128(smem_u32(Bt));
#pragma unroll
            for (int k = 0; k < 4; ++k)                                         // 4 x 32 bytes of K per 128-byte chunk
                umma_issue<KIND>(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kc | k) != 0);
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[st]))
                         : "memory");
        }
    }
    // wait for the last commit of every stage in flight (commits complete in issue order: the last one suffices)
    {
        const int kc = nchunks - 1, st = kc % stages;
        mbar_wait(&bars[st], (kc / stages) & 1);
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // ---- epilogue: TMEM -> registers -> global (row m0 + 32*warp + lane, 32 columns at a time) ----
    const int64_t row = m0 + warp * 32 + lane;
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t r[32];
        const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
              "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr)
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (row < B) {
            uint4* dst = reinterpret_cast<uint4*>(Yn + row * N + c0);
#pragma unroll
            for (int j = 0; j < 8; ++j) dst[j] = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(tmem_cols) : "memory");
}

// ------------------------------------------------------------------------------------------------
// TMA-fed, warp-specialised TF32 variant
// ------------------------------------------------------------------------------------------------
#define TMA_THREADS 192

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Epilogue of the warp-specialised kernels: TMEM -> registers (thread = row) -> a 32 x 32 transpose tile in shared memory
// -> global with every store instruction covering four full 128-byte row segments.  Writing straight from the
// thread-per-row registers made each STG.128 touch 32 different lines (16 bytes each at a 4N-byte stride): 8192
// half-sector write transactions per 128 x 256 tile kept L1 busy 4 us per tile -- twice the MMA time (ncu: L1 52 %,
// tensor pipe 22 %).
#define EPI_PITCH 36                                           // floats per staged row: 16-byte aligned, conflict-free for LDS/STS.128
// 32 consecutive TMEM columns of this thread's lane (= tile row) <-> 32 registers
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
        "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
        "r"(r[31])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
// one 32-row x 32-column block: registers (thread = row) -> transpose tile -> global, 4 rows x 128 bytes per store
__device__ __forceinline__ void tile_store_block(float* tile, int lane, const uint32_t (&r)[32], float scale, float* __restrict__ Yn,
                                                 int N, int64_t row0, int64_t B, int c0) {
    __syncwarp();                                              // earlier reads of the tile are done
    uint4* srow = reinterpret_cast<uint4*>(tile + (size_t)lane * EPI_PITCH);
#pragma unroll
    for (int j = 0; j < 8; ++j)
        srow[j] = make_uint4(__float_as_uint(__uint_as_float(r[4 * j]) * scale), __float_as_uint(__uint_as_float(r[4 * j + 1]) * scale),
                             __float_as_uint(__uint_as_float(r[4 * j + 2]) * scale), __float_as_uint(__uint_as_float(r[4 * j + 3]) * scale));
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int rr = (lane >> 3) + 4 * i;
        const uint4 v = *reinterpret_cast<const uint4*>(tile + (size_t)rr * EPI_PITCH + 4 * (lane & 7));
        if (row0 + rr < B) *reinterpret_cast<uint4*>(Yn + (row0 + rr) * N + c0 + 4 * (lane & 7)) = v;
    }
}
// the reverse: global block (coalesced) -> transpose tile -> 32 registers of this thread's row
__device__ __forceinline__ void tile_load_block(float* tile, int lane, float (&h)[32], const float* __restrict__ Xn, int N, int64_t row0,
                                                int64_t B, int c0) {
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int rr = (lane >> 3) + 4 * i;
        const int64_t rc = min(row0 + rr, B - 1);
        *reinterpret_cast<uint4*>(tile + (size_t)rr * EPI_PITCH + 4 * (lane & 7)) =
            __ldg(reinterpret_cast<const uint4*>(Xn + rc * N + c0 + 4 * (lane & 7)));
    }
    __syncwarp();
    const float4* srow = reinterpret_cast<const float4*>(tile + (size_t)lane * EPI_PITCH);
#pragma unroll
    for (int j = 0; j < 8; ++j) { const float4 v = srow[j]; h[4 * j] = v.x; h[4 * j + 1] = v.y; h[4 * j + 2] = v.z; h[4 * j + 3] = v.w; }
}

// columns c_first, c_first + c_stride, ... of this warp's TMEM lane quarter -> global rows (tile_id picks the transpose tile)
__device__ __forceinline__ void epilogue_store_rows(uint32_t tmem_d, int quarter, int lane, int N, int64_t m0, int64_t B,
                                                    float* __restrict__ Yn, float* __restrict__ stage, int tile_id = -1,
                                                    int c_first = 0, int c_stride = 32) {
    float* tile = stage + (size_t)(tile_id < 0 ? quarter : tile_id) * 32 * EPI_PITCH;
    const int64_t row0 = m0 + quarter * 32;
    for (int c0 = c_first; c0 < N; c0 += c_stride) {
        uint32_t r[32];
        tmem_ld32(tmem_d + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, r);
        tile_store_block(tile, lane, r, 1.0f, Yn, N, row0, B, c0);
    }
}

__global__ void __launch_bounds__(TMA_THREADS) k_wide_gemm_tma(const __grid_constant__ CUtensorMap mapX,
                                                               const __grid_constant__ CUtensorMap mapW, const GhmDev d,
                                                               int64_t B, int level, int stages, float* __restrict__ Y) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[4], empty_bar[4], tmem_full_bar;
    __shared__ uint32_t tmem_base_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N = d.QW;
    const int node = blockIdx.y;
    const int64_t m0 = (int64_t)blockIdx.x * TC_M;
    const int mi = d.mat_off[level] + (d.ti ? node - ghm_div_s(node, d) * d.s : node);
    unsigned char* tiles = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int a_bytes = TC_M * 128, b_bytes = N * 128, stage_bytes = a_bytes + b_bytes;
    const int nchunks = N / 32;
    const uint32_t tmem_cols = N <= 64 ? 64 : (N <= 128 ? 128 : 256);

    if (tid == 0) {
        for (int i = 0; i < 4; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        mbar_init(&tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                     "r"(tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_slot;

    if (warp == 0) {
        if (lane == 0) {                                      // ===== TMA producer =====
            for (int kc = 0; kc < nchunks; ++kc) {
                const int st = kc % stages;
                if (kc >= stages) mbar_wait(&empty_bar[st], ((kc / stages) - 1) & 1);
                unsigned char* At = tiles + (size_t)st * stage_bytes;
                mbar_expect_tx(&full_bar[st], (uint32_t)stage_bytes);
                tma_load_2d(At, &mapX, &full_bar[st], kc * 32, (int)((int64_t)node * B + m0));
                tma_load_2d(At + a_bytes, &mapW, &full_bar[st], kc * 32, mi * N);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {                                      // ===== MMA issuer =====
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
            for (int kc = 0; kc < nchunks; ++kc) {
                const int st = kc % stages;
                mbar_wait(&full_bar[st], (kc / stages) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                unsigned char* At = tiles + (size_t)st * stage_bytes;
                const uint64_t adesc = umma_desc_sw128(smem_u32(At)), bdesc = umma_desc_sw128(smem_u32(At + a_bytes));
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_issue<GHM_GEMM_TF32>(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kc | k) != 0);
                umma_commit(&empty_bar[st]);                  // stage free once these MMAs have read it
            }
            umma_commit(&tmem_full_bar);                      // accumulator complete
        }
    } else {                                                  // ===== epilogue warps 2..5 =====
        const int quarter = warp & 3;                         // TMEM lane quarter this warp may access
        mbar_wait(&tmem_full_bar, 0);                         // every MMA has completed: the operand stages are free
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        epilogue_store_rows(tmem_d, quarter, lane, N, m0, B, Y + (int64_t)node * B * N, reinterpret_cast<float*>(tiles));
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(tmem_cols) : "memory");
}

// ------------------------------------------------------------------------------------------------
// Cluster variant: CL row tiles of ONE node form a thread-block cluster and share the weight operand.  What bounds
// k_wide_gemm_tma is the L2 -> shared-memory traffic (profiles/r01_ncu_full_k_wide_gemm_tma.csv: L2 56 %, tensor pipe
// 22 %): every 128-row tile pulls its own copy of the N x N weight (256 KB at N = 256) next to 128 KB of messages.
// Here CTA r of the cluster loads rows [r N/CL, (r+1) N/CL) of each weight chunk and MULTICASTS them into the same
// stage of all CL CTAs (cp.async.bulk.tensor ... .multicast::cluster), so a weight byte crosses L2 -> SM once per
// cluster: (128 + 256 / CL) KB per tile instead of 384 KB.  A stage of a CTA is written by every CTA of the cluster,
// so its EMPTY barrier counts CL arrivals: each CTA's tcgen05.commit arrives on that barrier in all CL CTAs
// (.multicast::cluster).  Cluster barriers fence the mbarrier initialisation and the exit.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], "
        "[%2], %5;" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}

template <int CL>
__global__ void __launch_bounds__(TMA_THREADS) k_wide_gemm_tma_mc(const __grid_constant__ CUtensorMap mapX,
                                                                  const __grid_constant__ CUtensorMap mapWs, const GhmDev d,
                                                                  int64_t B, int level, int stages, float* __restrict__ Y) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[4], empty_bar[4], tmem_full_bar;
    __shared__ uint32_t tmem_base_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N = d.QW;
    const int node = blockIdx.y;
    const int64_t m0 = (int64_t)blockIdx.x * TC_M;
    const int mi = d.mat_off[level] + (d.ti ? node - ghm_div_s(node, d) * d.s : node);
    const uint32_t rank = cluster_ctarank();
    const uint16_t mask = (uint16_t)((1u << CL) - 1u);
    unsigned char* tiles = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int a_bytes = TC_M * 128, b_bytes = N * 128, stage_bytes = a_bytes + b_bytes;
    const int slice_rows = N / CL, slice_bytes = slice_rows * 128;
    const int nchunks = N / 32;
    const uint32_t tmem_cols = N <= 64 ? 64 : (N <= 128 ? 128 : 256);

    if (tid == 0) {
        for (int i = 0; i < 4; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], CL); }
        mbar_init(&tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                     "r"(tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();                                       // every CTA's barriers are initialised before any peer signals them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_slot;

    if (warp == 0) {
        if (lane == 0) {                                      // ===== TMA producer =====
            for (int kc = 0; kc < nchunks; ++kc) {
                const int st = kc % stages;
                if (kc >= stages) mbar_wait(&empty_bar[st], ((kc / stages) - 1) & 1);   // all CL consumers released the stage
                unsigned char* At = tiles + (size_t)st * stage_bytes;
                mbar_expect_tx(&full_bar[st], (uint32_t)stage_bytes);
                tma_load_2d(At, &mapX, &full_bar[st], kc * 32, (int)((int64_t)node * B + m0));
                tma_load_2d_mc(At + a_bytes + (size_t)rank * slice_bytes, &mapWs, &full_bar[st], kc * 32,
                               mi * N + (int)rank * slice_rows, mask);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {                                      // ===== MMA issuer =====
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
            for (int kc = 0; kc < nchunks; ++kc) {
                const int st = kc % stages;
                mbar_wait(&full_bar[st], (kc / stages) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                unsigned char* At = tiles + (size_t)st * stage_bytes;
                const uint64_t adesc = umma_desc_sw128(smem_u32(At)), bdesc = umma_desc_sw128(smem_u32(At + a_bytes));
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_issue<GHM_GEMM_TF32>(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kc | k) != 0);
                umma_commit_mc(&empty_bar[st], mask);         // this CTA is done with the stage: tell every producer of the cluster
            }
            umma_commit(&tmem_full_bar);                      // accumulator complete
        }
    } else {                                                  // ===== epilogue warps 2..5 =====
        const int quarter = warp & 3;                         // TMEM lane quarter this warp may access
        mbar_wait(&tmem_full_bar, 0);                         // every MMA has completed: the operand stages are free
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        epilogue_store_rows(tmem_d, quarter, lane, N, m0, B, Y + (int64_t)node * B * N, reinterpret_cast<float*>(tiles));
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();                                       // no CTA leaves while a peer may still signal its barriers
    if (warp == 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(tmem_cols) : "memory");
}

// ------------------------------------------------------------------------------------------------
// BP_DNS GEMMs with the row kernels folded in (TF32).  The per-launch ncu list of the unfused pass
// (profiles/r02_launches_wide_dns_unfused.csv, q = 256, B = 16384) spends 2.1 of 3.8 ms in HBM-bound row kernels that only
// reshuffle GEMM operands:
//   leaves, up   : k_wide_leaf_like writes the Gaussian likelihoods E (1.3 GB) that the GEMM reads straight back;
//   every level, down : k_wide_cavity writes w = b_parent / u, the GEMM reads it and writes T^T w, k_wide_belief reads that
//                       (and hd) to form the belief -- or, at the leaves, one float per row.
// Here the A operand is COMPUTED into the swizzled shared-memory tile by eight producer warps (coalesced: per instruction
// a warp covers 4 tile rows x 128 bytes and writes each 16-byte piece to its SWIZZLE_128B slot):
//   LF_UP       : A[r][k] = exp2(c2 ((z_r - k)^2 - d0_r))  from one float per row   (reference :485)
//   LF_DOWN(_INT): A[r][k] = b_parent[r][k] / u_v[r][k]                              (reference :513)
//   LF_CLS      : A[r][k] = prod_c T_c[k][x_c(r)]  (BP_CLS, depth L-1: the leaf rows gathered from the L2-resident table,
//                 :191-196; left unnormalised -- the scale is removed by the combine that follows the GEMM)
// the weight still arrives by TMA, and the same eight warps run the epilogue on the TMEM accumulator:
//   LF_UP       : store u = T e (transposed to coalesced rows);
//   LF_DOWN     : posterior mean sum_k k e_k tt_k / sum_k e_k tt_k reduced in registers (:516-519);
//   LF_DOWN_INT : belief b_v = hd_v * tt / max (:512-514): product written back to TMEM, row max exchanged between the
//                 two warps of a lane quarter, second pass rescales and stores.
// Nothing but the inputs and the outputs of a level crosses HBM.
// ------------------------------------------------------------------------------------------------
enum { LF_UP = 0, LF_DOWN = 1, LF_DOWN_INT = 2, LF_CLS = 3 };
#define FUSED_PROD_WARPS 8
#define FUSED_THREADS (64 + 32 * FUSED_PROD_WARPS)
struct FusedArgs {
    int level;               // depth of the nodes this launch covers (blockIdx.y = node index within the level)
    const float* z;          // [B][nL]                     (leaf modes)
    float c2;                // -0.5 log2(e) / sigma^2
    const float* BUpar;      // [n_par][B][N] beliefs of the parents          (LF_DOWN, LF_DOWN_INT)
    const float* U;          // [n][B][N]     upward messages u = T h         (LF_DOWN, LF_DOWN_INT)
    const float* H;          // [n][B][N]     hd of the nodes                 (LF_DOWN_INT)
    float* out;              // [n][B][N]     LF_UP: u;  LF_DOWN_INT: beliefs
    float* mean;             // [B][nL]       (LF_DOWN)
    const void* leaves;      // [B][nL] int64 / uint8          (LF_CLS)
    int leaf_dtype;
};

__device__ __forceinline__ float ex2_fast(float x) {          // arguments <= 0: MUFU.EX2, flushes to 0 far below 2^-126
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void producers_sync() {            // named barrier 1: the eight producer / epilogue warps only
    asm volatile("bar.sync 1, %0;" ::"r"(32 * FUSED_PROD_WARPS) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(FUSED_THREADS) k_wide_fused(const __grid_constant__ CUtensorMap mapW, const GhmDev d, int64_t B,
                                                              int stages, const FusedArgs a) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[4], empty_bar[4], tmem_full_bar;
    __shared__ uint32_t tmem_base_slot;
    __shared__ float2 red[TC_M];                              // per-row exchange between the two warps of a lane quarter
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N = d.QW, q = d.q, nL = d.n_leaves;
    const int node = blockIdx.y;
    const int64_t m0 = (int64_t)blockIdx.x * TC_M;
    const int mi = d.mat_off[a.level] + (d.ti ? node - ghm_div_s(node, d) * d.s : node);
    unsigned char* tiles = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int a_bytes = TC_M * 128, b_bytes = N * 128, stage_bytes = a_bytes + b_bytes;
    const int nchunks = N / 32;
    const uint32_t tmem_cols = N <= 64 ? 64 : (N <= 128 ? 128 : 256);

    if (tid == 0) {
        for (int i = 0; i < 4; ++i) { mbar_init(&full_bar[i], 1 + FUSED_PROD_WARPS); mbar_init(&empty_bar[i], 1); }
        mbar_init(&tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                     "r"(tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_slot;

    if (warp == 0) {
        if (lane == 0) {                                      // ===== TMA producer: weight chunks =====
            for (int kc = 0; kc < nchunks; ++kc) {
                const int st = kc % stages;
                if (kc >= stages) mbar_wait(&empty_bar[st], ((kc / stages) - 1) & 1);
                unsigned char* At = tiles + (size_t)st * stage_bytes;
                mbar_expect_tx(&full_bar[st], (uint32_t)b_bytes);
                tma_load_2d(At + a_bytes, &mapW, &full_bar[st], kc * 32, mi * N);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {                                      // ===== MMA issuer =====
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
            for (int kc = 0; kc < nchunks; ++kc) {
                const int st = kc % stages;
                mbar_wait(&full_bar[st], (kc / stages) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                unsigned char* At = tiles + (size_t)st * stage_bytes;
                const uint64_t adesc = umma_desc_sw128(smem_u32(At)), bdesc = umma_desc_sw128(smem_u32(At + a_bytes));
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_issue<GHM_GEMM_TF32>(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kc | k) != 0);
                umma_commit(&empty_bar[st]);
            }
            umma_commit(&tmem_full_bar);
        }
    } else {                                                  // ===== warps 2..9: A producers, then epilogue =====
        const int quarter = warp & 3;                         // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;                     // two warps share a quarter: rows 16 half .. +15 / alternate column blocks
        const int pid = warp - 2;
        const int r = quarter * 32 + lane;                    // tile row owned in the thread = row phases
        const int64_t b = m0 + r;
        const bool ok = b < B;
        float zi = 0.f, d0 = 0.f;
        if (MODE == LF_UP || MODE == LF_DOWN) {
            zi = ok ? a.z[b * nL + node] : 0.f;
            const float kstar = fminf(fmaxf(rintf(zi), 0.f), (float)(q - 1));
            d0 = (zi - kstar) * (zi - kstar);
        }
        const int pc = lane & 7;
        const int par = ghm_div_s(node, d);
        // LF_CLS: table row (matrix of leaf edge c, leaf state x_c) of each of the 4 tile rows this lane produces, c < s
        const float* crow[4][3];
        int s_cls = 0;
        if (MODE == LF_CLS) {
            s_cls = d.s;                                      // host: s <= 3 for this mode
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int rr = quarter * 32 + 16 * half + 4 * i + (lane >> 3);
                const int64_t bc = min(m0 + rr, B - 1);
                for (int c = 0; c < 3; ++c) {
                    crow[i][c] = nullptr;
                    if (c < s_cls) {
                        const int leaf = node * s_cls + c;
                        int64_t x = a.leaf_dtype == GHM_LEAF_I64 ? reinterpret_cast<const int64_t*>(a.leaves)[bc * nL + leaf]
                                                                 : (int64_t) reinterpret_cast<const uint8_t*>(a.leaves)[bc * nL + leaf];
                        if (x < 0 || x >= q) { atomicOr(d.status, 1); x = x < 0 ? 0 : q - 1; }
                        const int ml = d.mat_off[d.L] + (d.ti ? c : leaf);
                        crow[i][c] = d.Wdn + ((size_t)ml * N + (int)x) * N;          // T[:, x] of that edge
                    }
                }
            }
        }
        for (int kc = 0; kc < nchunks; ++kc) {
            const int st = kc % stages;
            unsigned char* At = tiles + (size_t)st * stage_bytes;
            // The chunk's global loads are issued BEFORE waiting for its stage to be free: they only need registers, and their
            // latency then overlaps the MMAs that still read the stage (2.72 -> 2.39 ms per BP_DNS at q = 256).  A second
            // register buffer that fetches chunk kc+1 ahead of chunk kc's conversion costs 127 registers x 320 threads = one
            // CTA per SM instead of two and measured 2.77 ms.
            float4 pbv[4], puv[4];
            if (MODE == LF_CLS) {                                        // gathered rows: pbv = row of child 0 (x child 1), puv = child 2
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    pbv[i] = __ldg(reinterpret_cast<const float4*>(crow[i][0] + kc * 32) + pc);
                    if (s_cls > 1) {
                        const float4 t1 = __ldg(reinterpret_cast<const float4*>(crow[i][1] + kc * 32) + pc);
                        pbv[i].x *= t1.x; pbv[i].y *= t1.y; pbv[i].z *= t1.z; pbv[i].w *= t1.w;
                    }
                    puv[i] = s_cls > 2 ? __ldg(reinterpret_cast<const float4*>(crow[i][2] + kc * 32) + pc) : make_float4(1.f, 1.f, 1.f, 1.f);
                }
            } else if (MODE != LF_UP) {                                  // all 8 loads of the chunk in flight before any use
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int rr = quarter * 32 + 16 * half + 4 * i + (lane >> 3);
                    const int64_t bc = min(m0 + rr, B - 1);              // rows past the batch re-read the last row; zeroed below
                    pbv[i] = __ldg(reinterpret_cast<const float4*>(a.BUpar + ((int64_t)par * B + bc) * N + kc * 32) + pc);
                    puv[i] = __ldg(reinterpret_cast<const float4*>(a.U + ((int64_t)node * B + bc) * N + kc * 32) + pc);
                }
            }
            if (kc >= stages) mbar_wait(&empty_bar[st], ((kc / stages) - 1) & 1);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int rl = 16 * half + 4 * i + (lane >> 3);          // row within the quarter
                const int rr = quarter * 32 + rl;                        // tile row
                const bool in = m0 + rr < B;
                uint4 v;
                if (MODE == LF_UP) {
                    const float zr = __shfl_sync(0xffffffffu, zi, rl), dr = __shfl_sync(0xffffffffu, d0, rl);
                    float e[4];
#pragma unroll
                    for (int x = 0; x < 4; ++x) {
                        const int k = kc * 32 + 4 * pc + x;
                        const float dk = zr - (float)k;
                        e[x] = (in && k < q) ? ex2_fast(a.c2 * (dk * dk - dr)) : 0.f;
                    }
                    v = make_uint4(__float_as_uint(e[0]), __float_as_uint(e[1]), __float_as_uint(e[2]), __float_as_uint(e[3]));
                } else if (MODE == LF_CLS) {
                    const float4 pb = pbv[i], pu = puv[i];
                    v = make_uint4(__float_as_uint(in ? pb.x * pu.x : 0.f), __float_as_uint(in ? pb.y * pu.y : 0.f),
                                   __float_as_uint(in ? pb.z * pu.z : 0.f), __float_as_uint(in ? pb.w * pu.w : 0.f));
                } else {
                    const float4 pb = pbv[i], pu = puv[i];
                    v = make_uint4(__float_as_uint(in && pu.x > 0.f ? __fdividef(pb.x, pu.x) : 0.f),
                                   __float_as_uint(in && pu.y > 0.f ? __fdividef(pb.y, pu.y) : 0.f),
                                   __float_as_uint(in && pu.z > 0.f ? __fdividef(pb.z, pu.z) : 0.f),
                                   __float_as_uint(in && pu.w > 0.f ? __fdividef(pb.w, pu.w) : 0.f));
                }
                *reinterpret_cast<uint4*>(At + (size_t)rr * 128 + ((pc ^ (rr & 7)) << 4)) = v;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // generic-proxy writes -> async proxy (UMMA)
            __syncwarp();
            if (lane == 0) mbar_arrive(&full_bar[st]);
        }
        mbar_wait(&tmem_full_bar, 0);                         // every MMA has completed: the operand stages are free
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        float* stage_f = reinterpret_cast<float*>(tiles);
        const uint32_t tlane = tmem_d + ((uint32_t)(quarter * 32) << 16);
        if (MODE == LF_UP || MODE == LF_CLS) {
            epilogue_store_rows(tmem_d, quarter, lane, N, m0, B, a.out + (int64_t)node * B * N, stage_f, pid, 32 * half, 64);
        } else if (MODE == LF_DOWN) {
            float num = 0.f, den = 0.f;
            for (int c0 = 32 * half; c0 < N; c0 += 64) {
                uint32_t t[32];
                tmem_ld32(tlane + (uint32_t)c0, t);
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int k = c0 + j;
                    const float dk = zi - (float)k;
                    const float bl = k < q ? ex2_fast(a.c2 * (dk * dk - d0)) * __uint_as_float(t[j]) : 0.f;
                    num = fmaf((float)k, bl, num);
                    den += bl;
                }
            }
            if (half == 1) red[r] = make_float2(num, den);
            producers_sync();
            if (half == 0 && ok) {
                const float2 o = red[r];
                a.mean[b * nL + node] = (num + o.x) / (den + o.y);
            }
        } else {
            // pass 1: product hd * tt back into TMEM, row max over this warp's column blocks
            float* tile = stage_f + (size_t)pid * 32 * EPI_PITCH;
            const float* Hn = a.H + (int64_t)node * B * N;
            const int64_t row0 = m0 + quarter * 32;
            float mx = 0.f;
            for (int c0 = 32 * half; c0 < N; c0 += 64) {
                uint32_t t[32];
                float h[32];
                tmem_ld32(tlane + (uint32_t)c0, t);
                tile_load_block(tile, lane, h, Hn, N, row0, B, c0);
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float p = h[j] * __uint_as_float(t[j]);
                    mx = fmaxf(mx, p);
                    t[j] = __float_as_uint(p);
                }
                tmem_st32(tlane + (uint32_t)c0, t);
            }
            if (half == 1) red[r].y = mx; else red[r].x = mx;  // the two warps of a quarter own the same rows
            producers_sync();
            const float2 both = red[r];
            const float inv = __fdividef(1.f, fmaxf(both.x, both.y));
            // pass 2: rescale and store
            float* BUn = a.out + (int64_t)node * B * N;
            for (int c0 = 32 * half; c0 < N; c0 += 64) {
                uint32_t t[32];
                tmem_ld32(tlane + (uint32_t)c0, t);
                tile_store_block(tile, lane, t, inv, BUn, N, row0, B, c0);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(tmem_cols) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
        else
            cudaGetLastError();
    }
    return fn;
}

// 2-D FP32 row-major [rows, N] tensor map with a [box_rows x 32] box (128-byte rows, SWIZZLE_128B)
static bool make_map(CUtensorMap* map, const float* base, uint64_t rows, int N, int box_rows) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return false;
    const cuuint64_t gdim[2] = {(cuuint64_t)N, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)N * sizeof(float)};
    const cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1u, 1u};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// returns GHM_EUNSUP when TMA cannot be used (no driver entry point, misaligned pointer): caller falls back
static int launch_gemm_tma(const ghm_model* m, int64_t B, int level, int n_nodes, int down, const float* X, float* Y,
                           cudaStream_t st) {
    const GhmDev& d = m->d;
    const int N = d.QW;
    if (((uintptr_t)X % 16) != 0 || ((uintptr_t)Y % 16) != 0) return GHM_EUNSUP;
    if ((int64_t)n_nodes * B >= (1ll << 31)) return GHM_EUNSUP;
    CUtensorMap mapX, mapW;
    if (!make_map(&mapX, X, (uint64_t)n_nodes * (uint64_t)B, N, TC_M)) return GHM_EUNSUP;
    if (!make_map(&mapW, down ? d.Wdn : d.Wup, (uint64_t)d.n_mat * (uint64_t)N, N, N)) return GHM_EUNSUP;
    const int stage_bytes = TC_M * 128 + N * 128;
    const int nchunks = N / 32;
    int stages = std::min(4, nchunks);
    while (stages > 2 && (size_t)stages * stage_bytes + 1024 > 110 * 1024) --stages;     // two CTAs per SM
    const size_t dyn = (size_t)stages * stage_bytes + 1024;
    const unsigned tiles = (unsigned)((B + TC_M - 1) / TC_M);
    // cluster of CL row tiles sharing the multicast weight operand (GHM_WIDE_CLUSTER=1 keeps the single-CTA kernel)
    static const int cl_env = getenv("GHM_WIDE_CLUSTER") ? atoi(getenv("GHM_WIDE_CLUSTER")) : 1;
    int CL = tiles >= 4 && cl_env >= 4 ? 4 : (tiles >= 2 && cl_env >= 2 ? 2 : 1);
    if (CL > 1) {
        CUtensorMap mapWs;
        if (!make_map(&mapWs, down ? d.Wdn : d.Wup, (uint64_t)d.n_mat * (uint64_t)N, N, N / CL)) return GHM_EUNSUP;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((tiles + CL - 1) / CL * CL, (unsigned)n_nodes, 1);
        cfg.blockDim = dim3(TMA_THREADS, 1, 1);
        cfg.dynamicSmemBytes = dyn;
        cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = (unsigned)CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        if (CL == 4) {
            GHM_CUDA_TRY(cudaFuncSetAttribute(k_wide_gemm_tma_mc<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
            GHM_CUDA_TRY(cudaLaunchKernelEx(&cfg, k_wide_gemm_tma_mc<4>, mapX, mapWs, d, B, level, stages, Y));
        } else {
            GHM_CUDA_TRY(cudaFuncSetAttribute(k_wide_gemm_tma_mc<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
            GHM_CUDA_TRY(cudaLaunchKernelEx(&cfg, k_wide_gemm_tma_mc<2>, mapX, mapWs, d, B, level, stages, Y));
        }
        return GHM_OK;
    }
    GHM_CUDA_TRY(cudaFuncSetAttribute(k_wide_gemm_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    dim3 grid(tiles, (unsigned)n_nodes);
    k_wide_gemm_tma<<<grid, TMA_THREADS, dyn, st>>>(mapX, mapW, d, B, level, stages, Y);
    GHM_CHECK_LAUNCH();
    return GHM_OK;
}

int ghm_wide_gemm_tc(const ghm_model* m, int64_t B, int level, int n_nodes, int down, const float* X, float* Y, cudaStream_t st) {
    const GhmDev& d = m->d;
    const int N = d.QW;                                        // a multiple of 32 in [32, 256]: UMMA M = 128 takes N % 16 == 0
    const int kind = m->gemm_mode;
    if (N < 32 || N > 256 || (N % 32) != 0) return GHM_EUNSUP;
    if (kind == GHM_GEMM_BF16 && (N % 64) != 0) return GHM_EUNSUP;    // BF16 stages 64-element (128-byte) K chunks
    if (n_nodes > 65535) return GHM_EUNSUP;
    if (kind == GHM_GEMM_TF32) {
        const int rc = launch_gemm_tma(m, B, level, n_nodes, down, X, Y, st);
        if (rc != GHM_EUNSUP) return rc;
    }
    const int kchunk = kind == GHM_GEMM_TF32 ? 32 : 64;
    const int nchunks = N / kchunk;
    const int stage_bytes = TC_M * 128 + N * 128;
    int stages = std::min(TC_MAX_STAGES, nchunks);
    while (stages > 1 && (size_t)stages * stage_bytes + 1024 > 200 * 1024) --stages;
    const size_t dyn = (size_t)stages * stage_bytes + 1024;
    dim3 grid((unsigned)((B + TC_M - 1) / TC_M), (unsigned)n_nodes);
    if (kind == GHM_GEMM_TF32) {
        GHM_CUDA_TRY(cudaFuncSetAttribute(k_wide_gemm_tc<GHM_GEMM_TF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
        k_wide_gemm_tc<GHM_GEMM_TF32><<<grid, TC_THREADS, dyn, st>>>(d, B, level, down, stages, X, Y);
    } else if (kind == GHM_GEMM_BF16) {
        GHM_CUDA_TRY(cudaFuncSetAttribute(k_wide_gemm_tc<GHM_GEMM_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
        k_wide_gemm_tc<GHM_GEMM_BF16><<<grid, TC_THREADS, dyn, st>>>(d, B, level, down, stages, X, Y);
    } else {
        return GHM_EUNSUP;
    }
    GHM_CHECK_LAUNCH();
    return GHM_OK;
}


// BP_DNS levels with the row kernels folded into the TF32 GEMM (see k_wide_fused).  GHM_EUNSUP when the shape / mode /
// driver does not allow it: the caller runs the unfused kernels.
static int launch_fused(const ghm_model* m, int64_t B, int mode, int n_nodes, const FusedArgs& a, cudaStream_t st) {
    const GhmDev& d = m->d;
    const int N = d.QW;
    if (m->gemm_mode != GHM_GEMM_TF32) return GHM_EUNSUP;
    if (const char* e = getenv("GHM_WIDE_UNFUSED")) {          // development switch: "1" = all, "up" / "down" / "int" = one mode
        if (e[0] == '1' || (e[0] == 'u' && mode == LF_UP) || (e[0] == 'd' && mode == LF_DOWN) || (e[0] == 'i' && mode == LF_DOWN_INT) ||
            (e[0] == 'c' && mode == LF_CLS))
            return GHM_EUNSUP;
    }
    if (N < 32 || N > 256 || (N % 32) != 0) return GHM_EUNSUP;
    if (n_nodes > 65535) return GHM_EUNSUP;
    CUtensorMap mapW;
    if (!make_map(&mapW, (mode == LF_UP || mode == LF_CLS) ? d.Wup : d.Wdn, (uint64_t)d.n_mat * (uint64_t)N, N, N)) return GHM_EUNSUP;
    const int stage_bytes = TC_M * 128 + N * 128;
    int stages = std::min(4, N / 32);
    while (stages > 2 && (size_t)stages * stage_bytes + 1024 > 110 * 1024) --stages;     // two CTAs per SM
    // the epilogue reuses the operand stages for its eight 32 x 32 transpose tiles
    const size_t dyn = std::max((size_t)stages * stage_bytes, (size_t)FUSED_PROD_WARPS * 32 * EPI_PITCH * sizeof(float)) + 1024;
    dim3 grid((unsigned)((B + TC_M - 1) / TC_M), (unsigned)n_nodes);
    auto go = [&](auto kern) -> int {
        GHM_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
        kern<<<grid, FUSED_THREADS, dyn, st>>>(mapW, d, B, stages, a);
        GHM_CHECK_LAUNCH();
        return GHM_OK;
    };
    if (mode == LF_UP) return go(k_wide_fused<LF_UP>);
    if (mode == LF_DOWN) return go(k_wide_fused<LF_DOWN>);
    if (mode == LF_CLS) return go(k_wide_fused<LF_CLS>);
    return go(k_wide_fused<LF_DOWN_INT>);
}

int ghm_wide_leaf_up_fused(const ghm_model* m, int64_t B, const float* z, float c2, float* Uout, cudaStream_t st) {
    FusedArgs a{};
    a.level = m->d.L; a.z = z; a.c2 = c2; a.out = Uout;
    return launch_fused(m, B, LF_UP, m->d.n_leaves, a, st);
}

int ghm_wide_leaf_down_fused(const ghm_model* m, int64_t B, const float* z, float c2, const float* BUpar, const float* U,
                             float* mean, cudaStream_t st) {
    FusedArgs a{};
    a.level = m->d.L; a.z = z; a.c2 = c2; a.BUpar = BUpar; a.U = U; a.mean = mean;
    return launch_fused(m, B, LF_DOWN, m->d.n_leaves, a, st);
}

int ghm_wide_down_fused(const ghm_model* m, int64_t B, int level, const float* BUpar, const float* U, const float* H, float* BU,
                        cudaStream_t st) {
    FusedArgs a{};
    a.level = level; a.BUpar = BUpar; a.U = U; a.H = H; a.out = BU;
    return launch_fused(m, B, LF_DOWN_INT, m->d.spow[level], a, st);
}

// BP_CLS, depth L-1: U[j] = T_j (prod_c T_c[:, x_c]) with the leaf-row product formed inside the GEMM's A producer
int ghm_wide_cls_leaf_fused(const ghm_model* m, int64_t B, const void* leaves, int leaf_dtype, float* U, cudaStream_t st) {
    if (m->d.s > 3 || m->d.L < 2) return GHM_EUNSUP;
    FusedArgs a{};
    a.level = m->d.L - 1; a.leaves = leaves; a.leaf_dtype = leaf_dtype; a.out = U;
    return launch_fused(m, B, LF_CLS, m->d.spow[m->d.L - 1], a, st);
}
