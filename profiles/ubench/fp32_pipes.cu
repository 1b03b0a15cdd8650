// Micro-benchmark: issue rate of the fp32 instructions the VQ search is built from (sm_100a).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32_pipes fp32_pipes.cu && ./fp32_pipes
#include <cuda_runtime.h>
#include <cstdio>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float min3(float a, float b, float c) { float r; asm volatile("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float ffma(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float fmin2(float a, float b) { float r; asm volatile("min.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }

constexpr int CH = 8;  // independent chains per thread
template <int MODE>
__global__ void k(float *out, int iters, float seed)
{
    float x[CH]; u64 p[CH];
    for (int i = 0; i < CH; ++i) { x[i] = seed + i + threadIdx.x; p[i] = pack2(x[i], x[i] + 1.f); }
    const u64 bb = pack2(seed, seed);           // broadcast form (same value in both halves)
    const u64 pk = pack2(seed, seed * 1.5f);    // true packed operand
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            if (MODE == 0) p[i] = fma2(p[i], pk, pk);              // FFMA2 packed
            if (MODE == 1) p[i] = fma2(bb, p[i], pk);              // FFMA2 with a broadcast multiplicand
            if (MODE == 2) x[i] = ffma(x[i], seed, seed);          // scalar FFMA
            if (MODE == 3) x[i] = min3(x[i], seed, x[(i + 1) % CH]);  // FMNMX3
            if (MODE == 4) x[i] = fmin2(x[i], seed);               // FMNMX
            if (MODE == 5) { p[i] = fma2(p[i], pk, pk); x[i] = fmin2(x[i], seed); }   // FFMA2 + FMNMX (dual pipe?)
            if (MODE == 6) { p[i] = fma2(p[i], pk, pk); x[i] = ffma(x[i], seed, seed); } // FFMA2 + FFMA
            if (MODE == 7) { p[i] = fma2(p[i], pk, pk); x[i] = min3(x[i], seed, x[(i + 1) % CH]); } // FFMA2 + FMNMX3
        }
    }
    float s = 0; for (int i = 0; i < CH; ++i) { s += x[i]; s += (float)(p[i] & 0xffff); }
    if (s == 12345.678f) out[0] = s;
}
template <int MODE> void run(const char *name, int per_iter)
{
    float *d; cudaMalloc(&d, 4);
    const int iters = 4096, blocks = 148 * 4, threads = 256;
    k<MODE><<<blocks, threads>>>(d, 16, 1.0f);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a); k<MODE><<<blocks, threads>>>(d, iters, 1.0f); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double warp_instr = (double)blocks * threads / 32 * iters * CH * per_iter;
    // per SM per ns
    printf("%-28s %8.3f ms  %7.2f warp-instr/ns/SM  (at 1.965 GHz: %5.2f warp-instr/clk/SM = %5.2f per SMSP)\n", name, ms,
           warp_instr / (ms * 1e6) / 148, warp_instr / (ms * 1e6) / 148 / 1.965, warp_instr / (ms * 1e6) / 148 / 1.965 / 4);
    cudaFree(d);
}
int main()
{
    run<0>("FFMA2 packed", 1); run<1>("FFMA2 broadcast", 1); run<2>("FFMA scalar", 1); run<3>("FMNMX3", 1); run<4>("FMNMX", 1);
    run<5>("FFMA2 + FMNMX", 2); run<6>("FFMA2 + FFMA", 2); run<7>("FFMA2 + FMNMX3", 2);
    return 0;
}
