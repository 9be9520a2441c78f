// ubench.cu -- pipe-rate microbenchmarks behind the k_tree2 / k_dns2 design choices (development aid).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/scratch/ubench tools/scratch/ubench.cu
// Prints warp-instructions per cycle per SM sub-partition (SMSP) for: FFMA, FFMA2 (packed f32x2), FFMA2 with a
// constant-bank operand, IMAD.WIDE, and mixes.  16 warps per SMSP-quad (512 threads, 1 CTA/SM x 4 -> full issue).
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
struct CT { float v[64]; };

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, const __grid_constant__ CT ct, int iters) {
    float a[8], b = threadIdx.x * 1e-9f + 1.0f, c = 0.5f;
    unsigned long long p[8];
    uint32_t x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = i + threadIdx.x; p[i] = (unsigned long long)(i + threadIdx.x) << 20 | 5; x[i] = threadIdx.x * 7 + i; }
    const unsigned long long bb = ((unsigned long long)__float_as_uint(b) << 32) | __float_as_uint(b);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) {                         // FFMA reg
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
            } else if (MODE == 1) {                  // FFMA2 reg
                asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(bb));
            } else if (MODE == 2) {                  // FFMA2 with constant-bank operand (uniform index)
                const unsigned long long t = *reinterpret_cast<const unsigned long long*>(&ct.v[2 * ((it + i) & 31)]);
                asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[i]) : "l"(t), "l"(bb));
            } else if (MODE == 3) {                  // IMAD.WIDE.U32
                unsigned long long r;
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(x[i]), "r"(0xD2511F53u));
                x[i] = (uint32_t)(r >> 32) ^ (uint32_t)r;
            } else if (MODE == 4) {                  // FFMA2 + IMAD.WIDE interleaved 1:1
                asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(bb));
                unsigned long long r;
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(x[i]), "r"(0xD2511F53u));
                x[i] = (uint32_t)(r >> 32) ^ (uint32_t)r;
            } else if (MODE == 5) {                  // FFMA + FFMA2 interleaved 2:1 (same flops each)
                asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(bb));
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[(i + 4) & 7]) : "f"(b), "f"(c));
            } else if (MODE == 6) {                  // FMUL2
                asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(bb));
            } else if (MODE == 7) {                  // LOP3 / IADD (alu pipe)
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(x[(i + 1) & 7]), "r"(0x9E3779B9u));
            } else if (MODE == 8) {                  // FFMA2 + LOP3 1:1 (do fma and alu pipes dual-issue?)
                asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(bb));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(x[(i + 1) & 7]), "r"(0x9E3779B9u));
            } else if (MODE == 9) {                  // FFMA + LOP3 1:1
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(x[(i + 1) & 7]), "r"(0x9E3779B9u));
            } else if (MODE == 10) {                 // IMAD.WIDE + LOP3 1:1
                unsigned long long r;
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(x[i]), "r"(0xD2511F53u));
                asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(x[i]) : "r"((uint32_t)r), "r"((uint32_t)(r >> 32)), "r"(0x9E3779B9u));
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float((uint32_t)p[i]) + __uint_as_float((uint32_t)(p[i] >> 32)) + x[i];
    if (s == 12345.678f) out[0] = s;
}

template <int MODE>
void run(const char* name, int per_iter, float* d) {
    CT ct;
    for (int i = 0; i < 64; ++i) ct.v[i] = 1.0f + i * 1e-7f;
    cudaDeviceProp pr;
    cudaGetDeviceProperties(&pr, 0);
    const int ctas = pr.multiProcessorCount * 2, nt = 256;      // 16 warps per SM = 4 per SMSP
    k<MODE><<<ctas, nt>>>(d, ct, 64);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<ctas, nt>>>(d, ct, ITERS);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    int khz;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double cycles = ms * 1e-3 * khz * 1e3;
    const double winst_per_smsp = (double)ITERS * 8 * per_iter * 4;   // 4 warps per SMSP
    printf("%-34s %8.3f ms  %.3f warp-inst/clk/SMSP  (%.2f clk per inst)\n", name, ms, winst_per_smsp / cycles, cycles / winst_per_smsp);
}

int main() {
    float* d;
    cudaMalloc(&d, 4);
    run<0>("FFMA reg", 1, d);
    run<1>("FFMA2 reg", 1, d);
    run<2>("FFMA2 const-bank operand", 1, d);
    run<3>("IMAD.WIDE (+LOP)", 2, d);
    run<4>("FFMA2 + IMAD.WIDE (+LOP)", 3, d);
    run<5>("FFMA2 + 2 FFMA", 3, d);
    run<6>("FMUL2 reg", 1, d);
    run<7>("LOP3", 1, d);
    run<8>("FFMA2 + LOP3", 2, d);
    run<9>("FFMA + LOP3", 2, d);
    run<10>("IMAD.WIDE + LOP3", 2, d);
    cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
