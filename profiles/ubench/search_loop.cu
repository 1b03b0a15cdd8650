// Micro-benchmark of the VQ search inner loop: where should the code table live (shared memory
// LDS.128 broadcast vs constant bank) and how many tokens per lane (T)?  Same math as
// vq_assign.cu; 80 640 tokens x 1024 codes = the leaders of bench config 2.
#include <cuda_runtime.h>
#include <cstdio>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(u64 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

#ifndef SCALE
#define SCALE 1
#endif
constexpr int NCH = 128;  // chunks of 8 codes
__constant__ ulonglong2 c_tab[NCH * 10];

template <int T, bool CONST, bool MIN3>
__global__ void __launch_bounds__(256, 2) search(const float4 *z, int n_tok, const ulonglong2 *g_tab, int *out)
{
    extern __shared__ ulonglong2 s_tab[];
    for (int i = threadIdx.x; i < NCH * 10; i += 256) s_tab[i] = g_tab[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int gw = blockIdx.x * 8 + (threadIdx.x >> 5), nw = gridDim.x * 8;
    const int groups = n_tok / (32 * T);
    const u64 minus2 = pack2(-2.f, -2.f);
    for (int g = gw; g < groups; g += nw) {
        u64 zd[T][4], zs[T];
        float bestd[T]; int bestc[T];
#pragma unroll
        for (int t = 0; t < T; ++t) {
            const float4 v = z[g * 32 * T + t * 32 + lane];
            zd[t][0] = pack2(v.x, v.x); zd[t][1] = pack2(v.y, v.y); zd[t][2] = pack2(v.z, v.z); zd[t][3] = pack2(v.w, v.w);
            const float s = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
            zs[t] = pack2(s, s); bestd[t] = 3e38f; bestc[t] = 0;
        }
        for (int c = 0; c < NCH; ++c) {
            float cm[T];
#pragma unroll
            for (int t = 0; t < T; ++t) cm[t] = 3e38f;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                ulonglong2 e0, e1, e2, e3, es;
                if (CONST) { e0 = c_tab[c * 10 + half]; e1 = c_tab[c * 10 + 2 + half]; e2 = c_tab[c * 10 + 4 + half]; e3 = c_tab[c * 10 + 6 + half]; es = c_tab[c * 10 + 8 + half]; }
                else { e0 = s_tab[c * 10 + half]; e1 = s_tab[c * 10 + 2 + half]; e2 = s_tab[c * 10 + 4 + half]; e3 = s_tab[c * 10 + 6 + half]; es = s_tab[c * 10 + 8 + half]; }
#pragma unroll
                for (int t = 0; t < T; ++t) {
                    u64 a = mul2(zd[t][0], e0.x), b = mul2(zd[t][0], e0.y);
                    a = fma2(zd[t][1], e1.x, a); b = fma2(zd[t][1], e1.y, b);
                    a = fma2(zd[t][2], e2.x, a); b = fma2(zd[t][2], e2.y, b);
                    a = fma2(zd[t][3], e3.x, a); b = fma2(zd[t][3], e3.y, b);
                    a = fma2(a, minus2, add2(zs[t], es.x)); b = fma2(b, minus2, add2(zs[t], es.y));
                    float a0, a1, b0, b1; unpack2(a, a0, a1); unpack2(b, b0, b1);
                    if (MIN3) {
                        asm("min.f32 %0, %1, %2, %3;" : "=f"(cm[t]) : "f"(cm[t]), "f"(a0), "f"(a1));
                        asm("min.f32 %0, %1, %2, %3;" : "=f"(cm[t]) : "f"(cm[t]), "f"(b0), "f"(b1));
                    } else {
                        cm[t] = fminf(fminf(cm[t], a0), fminf(a1, fminf(b0, b1)));
                    }
                }
            }
#pragma unroll
            for (int t = 0; t < T; ++t) if (cm[t] < bestd[t]) { bestd[t] = cm[t]; bestc[t] = c; }
        }
#pragma unroll
        for (int t = 0; t < T; ++t) out[g * 32 * T + t * 32 + lane] = bestc[t];
    }
}
template <int T, bool CONST, bool MIN3> void run(const char *name, const float4 *z, int n, const ulonglong2 *tab, int *out)
{
    cudaFuncSetAttribute(search<T, CONST, MIN3>, cudaFuncAttributeMaxDynamicSharedMemorySize, NCH * 160);
    search<T, CONST, MIN3><<<296, 256, NCH * 160>>>(z, n, tab, out);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    for (int i = 0; i < 10; ++i) search<T, CONST, MIN3><<<296, 256, NCH * 160>>>(z, n, tab, out);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("%-34s %7.2f us per pass   (%s)\n", name, ms * 100, cudaGetErrorString(cudaGetLastError()));
}
int main()
{
    const int n = 80640 * SCALE;
    float4 *z; ulonglong2 *tab; int *out;
    cudaMalloc(&z, n * 16); cudaMalloc(&tab, NCH * 160); cudaMalloc(&out, n * 4);
    float4 *hz = new float4[n]; for (int i = 0; i < n; ++i) hz[i] = make_float4(i * 1e-6f, 1e-3f, -i * 2e-6f, 5e-4f);
    float *ht = new float[NCH * 40]; for (int i = 0; i < NCH * 40; ++i) ht[i] = (i % 977) * 1e-6f;
    cudaMemcpy(z, hz, n * 16, cudaMemcpyHostToDevice); cudaMemcpy(tab, ht, NCH * 160, cudaMemcpyHostToDevice);
    cudaMemcpyToSymbol(c_tab, ht, NCH * 160);
    run<2, false, true>("smem  T=2 min3", z, n, tab, out);
    run<2, false, false>("smem  T=2 fmin", z, n, tab, out);
    run<4, false, true>("smem  T=4 min3", z, n, tab, out);
    run<4, false, false>("smem  T=4 fmin", z, n, tab, out);
    run<6, false, false>("smem  T=6 fmin", z, n, tab, out);
    run<2, true, false>("const T=2 fmin", z, n, tab, out);
    run<4, true, false>("const T=4 fmin", z, n, tab, out);
    run<1, true, false>("const T=1 fmin", z, n, tab, out);
    return 0;
}
