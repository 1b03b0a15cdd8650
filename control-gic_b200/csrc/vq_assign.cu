// a1  VectorQuantize2.forward  (CGIC/modules/vqvae/quantize.py:69-98) as three sm_100a kernels.
//
//   vq_classify : one thread per token.  A token whose 4 latent channels are bit-identical to the
//                 top-left token of its 4x4 (else 2x2) block is a FOLLOWER of that token; all other
//                 tokens are LEADERS and are appended to a compact work list.  The encoder's
//                 mask-mix (vqvae_blocks.py:364-366) makes coarse / medium regions constant over
//                 4x4 / 2x2 blocks, so only n_c + n_m + n_f of the h*w tokens need a search; the test
//                 is on the data itself, so the result is exact for ANY input.
//   vq_search   : persistent grid, exhaustive search of the leaders over the K codes with the
//                 reference's rounding sequence (see below), packed two codes per instruction
//                 (FMUL2/FFMA2/FADD2) with the codebook staged in shared memory by one TMA bulk
//                 copy.  Work = (32*T-token block) x (8-code chunk) units dealt evenly to warps.
//   vq_finalize : one thread per token: index of its leader -> idx (int64), z_q = fl(z + fl(e - z))
//                 in NCHW, and a deterministic two-stage reduction of sum((e-z)^2).
//
// Rounding contract (bit-exact against torch CPU, see oracle/cgic_oracle.c and SURVEY.md 7.1):
//     z2  = ((z0*z0 + z1*z1) + z2*z2) + z3*z3      every product and sum rounded to fp32
//     e2  likewise
//     dot = fma(z3,e3, fma(z2,e2, fma(z1,e1, fl(z0*e0))))
//     d   = fl(fl(z2 + e2) - 2*dot) = fma(-2, dot, fl(z2 + e2))
//     argmin with the LOWEST index among equal minima (torch.argmin).
#include "common.cuh"

namespace cgic {
namespace {

constexpr int VQ_T = 4;        // tokens per lane in the search
constexpr int VQ_CHUNK = 8;    // codes per chunk = 4 code pairs
constexpr int VQ_THREADS = 256;
constexpr int VQ_ROWS = 5;     // e0 e1 e2 e3 e^2
constexpr int VQ_MAX_K = 4096;

typedef unsigned long long u64;

__device__ __forceinline__ u64 pack2(float lo, float hi)
{
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(u64 v, float &lo, float &hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b)
{
    u64 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c)
{
    u64 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b)
{
    u64 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ float min3(float a, float b, float c)
{
    float r;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ float sumsq4(float a, float b, float c, float d)
{
    float s = __fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b));
    s = __fadd_rn(s, __fmul_rn(c, c));
    return __fadd_rn(s, __fmul_rn(d, d));
}

// monotone map float -> uint32 (a < b  <=>  key(a) < key(b)); d is never -0 nor NaN here
__device__ __forceinline__ uint32_t order_key(float d)
{
    const uint32_t u = __float_as_uint(d);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(VQ_THREADS)
vq_classify_kernel(const float *__restrict__ z, int64_t n_tokens, int h, int w, int32_t *__restrict__ leader_of,
                   int32_t *__restrict__ list, u64 *__restrict__ best, int32_t *__restrict__ counters)
{
    __shared__ int s_warp[VQ_THREADS / 32];
    __shared__ int s_base;
    const int64_t t = (int64_t)blockIdx.x * VQ_THREADS + threadIdx.x;
    const int plane = h * w;
    bool leader = false;
    if (t < n_tokens) {
        const int b = (int)(t / plane);
        const int p = (int)(t - (int64_t)b * plane);
        const int y = p / w, x = p - y * w;
        const float *zb = z + (int64_t)b * 4 * plane;
        const uint32_t v0 = __float_as_uint(zb[p]), v1 = __float_as_uint(zb[plane + p]),
                       v2 = __float_as_uint(zb[2 * plane + p]), v3 = __float_as_uint(zb[3 * plane + p]);
        int lead = p;
        const int p4 = (y & ~3) * w + (x & ~3);
        if (p4 != p && __float_as_uint(zb[p4]) == v0 && __float_as_uint(zb[plane + p4]) == v1 &&
            __float_as_uint(zb[2 * plane + p4]) == v2 && __float_as_uint(zb[3 * plane + p4]) == v3)
            lead = p4;
        if (lead == p) {
            const int p2 = (y & ~1) * w + (x & ~1);
            if (p2 != p && __float_as_uint(zb[p2]) == v0 && __float_as_uint(zb[plane + p2]) == v1 &&
                __float_as_uint(zb[2 * plane + p2]) == v2 && __float_as_uint(zb[3 * plane + p2]) == v3)
                lead = p2;
        }
        leader = (lead == p);
        leader_of[t] = (int32_t)((int64_t)b * plane + lead);
        if (leader) best[t] = ~0ull;
    }
    // block-aggregated append to the work list (order is irrelevant to the result)
    const unsigned m = __ballot_sync(0xffffffffu, leader);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) s_warp[wid] = __popc(m);
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
        for (int i = 0; i < VQ_THREADS / 32; ++i) {
            const int c = s_warp[i];
            s_warp[i] = tot;
            tot += c;
        }
        s_base = tot ? atomicAdd(&counters[0], tot) : 0;
    }
    __syncthreads();
    if (leader) list[s_base + s_warp[wid] + __popc(m & ((1u << lane) - 1u))] = (int32_t)t;
}

// ---------------------------------------------------------------------------------------------
// shared memory: [raw codebook Kpad*4 floats][chunk table (Kpad/8) * 5 rows * 4 pairs * float2]
__global__ void __launch_bounds__(VQ_THREADS, 2)
vq_search_kernel(const float *__restrict__ z, int h, int w, const float *__restrict__ codebook, int K, int Kpad,
                 const int32_t *__restrict__ list, const int32_t *__restrict__ counters, u64 *__restrict__ best)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) u64 mbar;
    float *raw = reinterpret_cast<float *>(smem);
    float2 *tab = reinterpret_cast<float2 *>(smem + (size_t)Kpad * 16);

    // --- stage the codebook with one TMA bulk copy (cp.async.bulk -> UBLKCP), mbarrier completion
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t bytes = (uint32_t)K * 16u;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(raw)),
                     "l"(codebook), "r"(bytes), "r"(smem_u32(&mbar))
                     : "memory");
    }
    {
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done)
                         : "r"(smem_u32(&mbar)), "r"(0u)
                         : "memory");
    }
    // --- chunk table: row r of chunk c holds (v[8c+0],v[8c+1]) (v[8c+2],v[8c+3]) ... for v = e_r or e^2
    const int npairs = Kpad / 2;
    for (int pr = threadIdx.x; pr < npairs; pr += VQ_THREADS) {
        const int k0 = 2 * pr, k1 = k0 + 1;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
        float sa = __int_as_float(0x7f800000), sb = sa;  // padding codes: d = +inf, never selected
        if (k0 < K) {
            a = reinterpret_cast<const float4 *>(raw)[k0];
            sa = sumsq4(a.x, a.y, a.z, a.w);
        }
        if (k1 < K) {
            b = reinterpret_cast<const float4 *>(raw)[k1];
            sb = sumsq4(b.x, b.y, b.z, b.w);
        }
        float2 *dst = tab + (size_t)(pr >> 2) * (VQ_ROWS * 4) + (pr & 3);
        dst[0] = make_float2(a.x, b.x);
        dst[4] = make_float2(a.y, b.y);
        dst[8] = make_float2(a.z, b.z);
        dst[12] = make_float2(a.w, b.w);
        dst[16] = make_float2(sa, sb);
    }
    __syncthreads();

    const int L = counters[0];
    const int plane = h * w;
    const int nchunks = Kpad / VQ_CHUNK;
    const int lane = threadIdx.x & 31;
    const int64_t n_blocks = ((int64_t)L + 32 * VQ_T - 1) / (32 * VQ_T);
    const int64_t total = n_blocks * nchunks;
    const int64_t n_warps = (int64_t)gridDim.x * (VQ_THREADS / 32);
    const int64_t per_warp = (total + n_warps - 1) / n_warps;
    const int64_t gw = (int64_t)blockIdx.x * (VQ_THREADS / 32) + (threadIdx.x >> 5);
    int64_t u = gw * per_warp;
    const int64_t u_end = min(total, u + per_warp);
    const u64 minus2 = pack2(-2.f, -2.f);
    const ulonglong2 *ctab = reinterpret_cast<const ulonglong2 *>(tab);
    const float *ftab = reinterpret_cast<const float *>(tab);

    while (u < u_end) {
        const int64_t blk = u / nchunks;
        const int c_begin = (int)(u - blk * nchunks);
        const int c_end = (int)min((int64_t)nchunks, c_begin + (u_end - u));
        // --- load this lane's T tokens
        int tok[VQ_T];
        float zf[VQ_T][4], z2[VQ_T];
        u64 zd[VQ_T][4], zs[VQ_T];
        float bestd[VQ_T];
        int bestc[VQ_T];
#pragma unroll
        for (int t = 0; t < VQ_T; ++t) {
            const int64_t li = blk * (32 * VQ_T) + t * 32 + lane;
            tok[t] = li < L ? list[li] : -1;
            if (tok[t] >= 0) {
                const int b = tok[t] / plane;
                const int p = tok[t] - b * plane;
                const float *zb = z + (int64_t)b * 4 * plane + p;
                zf[t][0] = zb[0];
                zf[t][1] = zb[plane];
                zf[t][2] = zb[2 * (int64_t)plane];
                zf[t][3] = zb[3 * (int64_t)plane];
            } else {
                zf[t][0] = zf[t][1] = zf[t][2] = zf[t][3] = 0.f;
            }
            z2[t] = sumsq4(zf[t][0], zf[t][1], zf[t][2], zf[t][3]);
#pragma unroll
            for (int c = 0; c < 4; ++c) zd[t][c] = pack2(zf[t][c], zf[t][c]);
            zs[t] = pack2(z2[t], z2[t]);
            bestd[t] = __int_as_float(0x7f800000);
            bestc[t] = c_begin;
        }
        // --- scan chunks [c_begin, c_end): per token the minimum distance of each chunk
        for (int c = c_begin; c < c_end; ++c) {
            float cm[VQ_T];
#pragma unroll
            for (int t = 0; t < VQ_T; ++t) cm[t] = __int_as_float(0x7f800000);
            const ulonglong2 *row = ctab + (size_t)c * (VQ_ROWS * 2);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const ulonglong2 e0 = row[0 + half], e1 = row[2 + half], e2 = row[4 + half], e3 = row[6 + half],
                                 es = row[8 + half];
#pragma unroll
                for (int t = 0; t < VQ_T; ++t) {
                    u64 a = mul2(zd[t][0], e0.x);
                    u64 b = mul2(zd[t][0], e0.y);
                    a = fma2(zd[t][1], e1.x, a);
                    b = fma2(zd[t][1], e1.y, b);
                    a = fma2(zd[t][2], e2.x, a);
                    b = fma2(zd[t][2], e2.y, b);
                    a = fma2(zd[t][3], e3.x, a);
                    b = fma2(zd[t][3], e3.y, b);
                    const u64 sa = add2(zs[t], es.x);
                    const u64 sb = add2(zs[t], es.y);
                    a = fma2(a, minus2, sa);
                    b = fma2(b, minus2, sb);
                    float a0, a1, b0, b1;
                    unpack2(a, a0, a1);
                    unpack2(b, b0, b1);
                    cm[t] = min3(cm[t], a0, a1);
                    cm[t] = min3(cm[t], b0, b1);
                }
            }
#pragma unroll
            for (int t = 0; t < VQ_T; ++t)
                if (cm[t] < bestd[t]) {
                    bestd[t] = cm[t];
                    bestc[t] = c;
                }
        }
        // --- resolve the first code of the winning chunk that attains the minimum, publish
#pragma unroll
        for (int t = 0; t < VQ_T; ++t) {
            if (tok[t] < 0 || !(bestd[t] < __int_as_float(0x7f800000))) continue;
            const float *rowf = ftab + (size_t)bestc[t] * (VQ_ROWS * 8);
            int k = 0;
#pragma unroll
            for (int j = VQ_CHUNK - 1; j >= 0; --j) {
                float dot = __fmul_rn(zf[t][0], rowf[j]);
                dot = __fmaf_rn(zf[t][1], rowf[8 + j], dot);
                dot = __fmaf_rn(zf[t][2], rowf[16 + j], dot);
                dot = __fmaf_rn(zf[t][3], rowf[24 + j], dot);
                const float d = __fmaf_rn(dot, -2.f, __fadd_rn(z2[t], rowf[32 + j]));
                if (d == bestd[t]) k = j;
            }
            const u64 key = ((u64)order_key(bestd[t]) << 32) | (uint32_t)(bestc[t] * VQ_CHUNK + k);
            atomicMin(&best[tok[t]], key);
        }
        u += (c_end - c_begin);
    }
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(VQ_THREADS)
vq_finalize_kernel(const float *__restrict__ z, int64_t n_tokens, int h, int w, const float *__restrict__ codebook,
                   const int32_t *__restrict__ leader_of, const u64 *__restrict__ best, int64_t *__restrict__ idx_out,
                   float *__restrict__ zq_out, double *__restrict__ partials, int32_t *__restrict__ counters,
                   double *__restrict__ sqerr_out)
{
    __shared__ double s_red[VQ_THREADS / 32];
    __shared__ bool s_last;
    const int64_t t = (int64_t)blockIdx.x * VQ_THREADS + threadIdx.x;
    const int plane = h * w;
    double sq = 0.0;
    if (t < n_tokens) {
        const u64 key = best[leader_of[t]];
        const int k = key == ~0ull ? 0 : (int)(uint32_t)key;
        idx_out[t] = k;
        if (zq_out || sqerr_out) {
            const float4 e = reinterpret_cast<const float4 *>(codebook)[k];
            const int b = (int)(t / plane);
            const int p = (int)(t - (int64_t)b * plane);
            const int64_t o = (int64_t)b * 4 * plane + p;
            const float ev[4] = {e.x, e.y, e.z, e.w};
            float acc = 0.f;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float zc = z[o + (int64_t)c * plane];
                const float diff = __fsub_rn(ev[c], zc);
                if (zq_out) zq_out[o + (int64_t)c * plane] = __fadd_rn(zc, diff);
                acc = __fmaf_rn(diff, diff, acc);
            }
            sq = (double)acc;
        }
    }
    if (!sqerr_out) return;
    // deterministic reduction: warp shuffle -> block -> per-block partial -> last block sums in order
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = sq;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int i = 0; i < VQ_THREADS / 32; ++i) tot += s_red[i];
        partials[blockIdx.x] = tot;
        __threadfence();
        s_last = (atomicAdd(&counters[1], 1) == (int)gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        double tot = 0.0;
        for (int i = threadIdx.x; i < (int)gridDim.x; i += VQ_THREADS) tot += __ldcg(&partials[i]);
        // fixed tree: thread-strided partial sums, then the same shuffle/serial order as above
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = tot;
        __syncthreads();
        if (threadIdx.x == 0) {
            double all = 0.0;
            for (int i = 0; i < VQ_THREADS / 32; ++i) all += s_red[i];
            *sqerr_out = all;
        }
    }
}

__global__ void vq_count_kernel(const int64_t *__restrict__ idx, int64_t n, float *__restrict__ counters, int K)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t k = idx[i];
    // one exact +1 per token, like the reference loop (fp32 counters stop growing at 2^24 there too)
    if (k >= 0 && k < K) atomicAdd(&counters[k], 1.0f);
}

struct VqCarve {
    int32_t *counters;  // [0] leader count, [1] finalize ticket   (64 bytes reserved)
    int32_t *leader_of;
    int32_t *list;
    u64 *best;
    double *partials;
    size_t bytes;
};

VqCarve carve(void *ws, int64_t n)
{
    auto up = [](size_t v) { return (v + 255) / 256 * 256; };
    VqCarve c{};
    unsigned char *p = static_cast<unsigned char *>(ws);
    size_t o = 0;
    c.counters = reinterpret_cast<int32_t *>(p + o);
    o += 256;
    c.leader_of = reinterpret_cast<int32_t *>(p + o);
    o += up((size_t)n * 4);
    c.list = reinterpret_cast<int32_t *>(p + o);
    o += up((size_t)n * 4);
    c.best = reinterpret_cast<u64 *>(p + o);
    o += up((size_t)n * 8);
    c.partials = reinterpret_cast<double *>(p + o);
    o += up((size_t)((n + VQ_THREADS - 1) / VQ_THREADS) * 8);
    c.bytes = o;
    return c;
}

}  // namespace
}  // namespace cgic

using namespace cgic;

extern "C" size_t cgic_vq_workspace_bytes(int64_t n_tokens)
{
    if (n_tokens <= 0) return 256;
    return carve(nullptr, n_tokens).bytes;
}

extern "C" int cgic_vq_assign(const float *z, int B, int h, int w, const float *codebook, int K, int64_t *idx_out,
                              float *zq_out, double *sqerr_out, void *workspace, size_t workspace_bytes,
                              cgic_stream_t stream_)
{
    CGIC_REQUIRE(z && codebook && idx_out && workspace, CGIC_EINVAL, "cgic_vq_assign: null argument");
    CGIC_REQUIRE(B >= 0 && h > 0 && w > 0, CGIC_EINVAL, "cgic_vq_assign: bad shape B=%d h=%d w=%d", B, h, w);
    CGIC_REQUIRE(K >= 1 && K <= VQ_MAX_K, CGIC_EINVAL, "cgic_vq_assign: K=%d outside [1, %d]", K, VQ_MAX_K);
    CGIC_REQUIRE((reinterpret_cast<uintptr_t>(codebook) & 15) == 0, CGIC_EINVAL, "cgic_vq_assign: codebook must be 16-byte aligned");
    const int64_t n = (int64_t)B * h * w;
    CGIC_REQUIRE(n < (int64_t)1 << 31, CGIC_EINVAL, "cgic_vq_assign: %lld tokens exceed 2^31", (long long)n);
    if (n == 0) {
        if (sqerr_out) CGIC_CUDA_CHECK(cudaMemsetAsync(sqerr_out, 0, sizeof(double), as_stream(stream_)));
        return CGIC_OK;
    }
    const VqCarve c = carve(workspace, n);
    CGIC_REQUIRE(workspace_bytes >= c.bytes, CGIC_ESPACE, "cgic_vq_assign: workspace %zu < %zu bytes", workspace_bytes, c.bytes);
    cudaStream_t stream = as_stream(stream_);

    static int n_sm = 0;
    if (!n_sm) {
        int dev = 0;
        CGIC_CUDA_CHECK(cudaGetDevice(&dev));
        CGIC_CUDA_CHECK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
        CGIC_CUDA_CHECK(cudaFuncSetAttribute(vq_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, VQ_MAX_K * 36));
    }
    const int Kpad = (K + VQ_CHUNK - 1) / VQ_CHUNK * VQ_CHUNK;
    const size_t smem = (size_t)Kpad * 16 + (size_t)(Kpad / VQ_CHUNK) * VQ_ROWS * 4 * sizeof(float2);
    const int blocks = (int)((n + VQ_THREADS - 1) / VQ_THREADS);

    CGIC_CUDA_CHECK(cudaMemsetAsync(c.counters, 0, 64, stream));
    {
        CGIC_PROF("vq_classify_kernel", stream);
        vq_classify_kernel<<<blocks, VQ_THREADS, 0, stream>>>(z, n, h, w, c.leader_of, c.list, c.best, c.counters);
    }
    CGIC_LAUNCH_CHECK();
    // persistent grid: 2 CTAs per SM, never more CTAs than there is work for in the worst case
    const int64_t max_units = ((n + 32 * VQ_T - 1) / (32 * VQ_T)) * (Kpad / VQ_CHUNK);
    const int64_t want = (max_units + VQ_THREADS / 32 - 1) / (VQ_THREADS / 32);
    const int grid = (int)(want < 2 * (int64_t)n_sm ? (want < 1 ? 1 : want) : 2 * (int64_t)n_sm);
    {
        CGIC_PROF("vq_search_kernel", stream);
        vq_search_kernel<<<grid, VQ_THREADS, smem, stream>>>(z, h, w, codebook, K, Kpad, c.list, c.counters, c.best);
    }
    CGIC_LAUNCH_CHECK();
    {
        CGIC_PROF("vq_finalize_kernel", stream);
        vq_finalize_kernel<<<blocks, VQ_THREADS, 0, stream>>>(z, n, h, w, codebook, c.leader_of, c.best, idx_out, zq_out,
                                                              c.partials, c.counters, sqerr_out);
    }
    CGIC_LAUNCH_CHECK();
    return CGIC_OK;
}

extern "C" int cgic_vq_count(const int64_t *idx, int64_t n, float *counters, int K, cgic_stream_t stream_)
{
    CGIC_REQUIRE(idx && counters && n >= 0 && K > 0, CGIC_EINVAL, "cgic_vq_count: bad argument");
    if (n == 0) return CGIC_OK;
    vq_count_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream_)>>>(idx, n, counters, K);
    CGIC_LAUNCH_CHECK();
    return CGIC_OK;
}
