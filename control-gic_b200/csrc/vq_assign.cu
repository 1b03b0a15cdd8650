// a1  VectorQuantize2.forward  (CGIC/modules/vqvae/quantize.py:69-98) as ONE persistent sm_100a kernel.
//
// Each CTA stages the codebook once (TMA bulk copy -> shared memory, re-laid out as packed code
// pairs), then loops over tiles of the token grid (R rows x C columns of one image, multiples
// of 4, at most 1024 tokens).  Per tile:
//   load     z of the tile, NCHW global -> one float4 per token in shared memory (coalesced);
//   classify a token whose 4 latent channels are bit-identical to the top-left token of its 4x4
//            (else 2x2) block is a FOLLOWER of that token; the others are LEADERS, compacted into
//            a list.  The encoder's mask-mix (vqvae_blocks.py:364-366) makes coarse / medium
//            regions constant over 4x4 / 2x2 blocks, so only n_c + n_m + n_f of the tokens need a
//            search; the test is on the data itself, hence exact for ANY input;
//   search   exhaustive over the K codes with the reference's rounding sequence (below), two codes
//            per instruction (FMUL2 / FFMA2 / FADD2).  Work units = (64 leaders) x (a range of
//            8-code chunks), dealt evenly to the warps; partial results meet in a 64-bit
//            shared-memory atomicMin on (ordered distance bits, code index);
//   finalize every token takes its leader's index: idx (int64), z_q = fl(z + fl(e - z)) in NCHW,
//            and sum((e-z)^2) accumulated per CTA; the last CTA adds the per-CTA partials in a
//            fixed order (deterministic).
// z is read once and idx / z_q written once: no intermediate ever goes to global memory.
// This kernel serves cgic_vq_assign (raw codebook).  cgic_vq_assign_indexed (prepared codebook, codebook.cu)
// runs vq_warp_kernel further down: same results from a handful of candidate codes per token.
//
// Rounding contract (bit-exact against torch CPU, see oracle/cgic_oracle.c and SURVEY.md 7.1):
//     z2  = ((z0*z0 + z1*z1) + z2*z2) + z3*z3      every product and sum rounded to fp32
//     e2  likewise
//     dot = fma(z3,e3, fma(z2,e2, fma(z1,e1, fl(z0*e0))))
//     d   = fl(fl(z2 + e2) - 2*dot) = fma(-2, dot, fl(z2 + e2))
//     argmin with the LOWEST index among equal minima (torch.argmin).
#include "vq_tile.cuh"

namespace cgic {

const unsigned char *codebook_blob(const cgic_codebook *cb);  // codebook.cu
int codebook_size(const cgic_codebook *cb);

namespace {

#ifndef CGIC_VQ_T
#define CGIC_VQ_T 2
#endif
#ifndef CGIC_VQ_TILE
#define CGIC_VQ_TILE 1024
#endif
#ifndef CGIC_VQ_THREADS
#define CGIC_VQ_THREADS 256
#endif
#ifndef CGIC_VQ_CTAS
#define CGIC_VQ_CTAS 2
#endif
constexpr int VQ_T = CGIC_VQ_T;  // tokens per lane in the search
constexpr int VQ_CHUNK = 8;    // codes per chunk = 4 code pairs
constexpr int VQ_THREADS = CGIC_VQ_THREADS;
constexpr int VQ_WARPS = VQ_THREADS / 32;
constexpr int VQ_ROWS = 5;     // e0 e1 e2 e3 e^2
constexpr int VQ_MAX_K = 4096;
constexpr int VQ_TILE = CGIC_VQ_TILE;  // tokens per batch of strips (shared-memory capacity)
constexpr int VQ_MAX_STRIPS = VQ_TILE / 16;  // a strip has at least 4 x 4 tokens
constexpr int VQ_CTAS = CGIC_VQ_CTAS;        // resident CTAs per SM (persistent grid = VQ_CTAS x #SM)

// dynamic shared memory of vq_fused_kernel: [raw codebook Kpad * 16][pair table Kpad * 20]
//                                           [z float4 x TILE][best u64 x TILE][lead u16 x TILE][list u16 x TILE]
__host__ __device__ inline size_t vq_smem_bytes(int Kpad) { return (size_t)Kpad * 36 + (size_t)VQ_TILE * (16 + 8 + 2 + 2); }
// staged head of a prepared codebook (header | bin -> cell table | codebook | e^2), vq_warp_kernel
__host__ __device__ inline size_t vq_stage_bytes(int K) { return (cb_layout(K).stage + 127) & ~(size_t)127; }

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

// monotone map float -> uint32 (a < b  <=>  key(a) < key(b)); d is never -0 nor NaN here
__device__ __forceinline__ uint32_t order_key(float d)
{
    const uint32_t u = __float_as_uint(d);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

#ifdef CGIC_VQ_TRACE
#define VQ_STAMP(k)                                                                                  \
    do {                                                                                             \
        if (threadIdx.x == 0 && blockIdx.x < 300) {                                                  \
            unsigned long long t__;                                                                  \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t__));                                  \
            reinterpret_cast<unsigned long long *>(partials)[512 + blockIdx.x * 8 + (k)] = t__;      \
            if ((k) == 0) {                                                                          \
                unsigned sm__;                                                                       \
                asm volatile("mov.u32 %0, %%smid;" : "=r"(sm__));                                    \
                reinterpret_cast<unsigned long long *>(partials)[512 + blockIdx.x * 8 + 7] = sm__;   \
            }                                                                                        \
        }                                                                                            \
    } while (0)
#else
#define VQ_STAMP(k)
#endif

// Work decomposition: a STRIP is 4 rows x C columns of one image (C a multiple of 4, <= 256), so
// every 4x4 / 2x2 block lies inside one strip.  The strips of the whole batch are numbered
// image-major and dealt to the CTAs as contiguous, equally long ranges (one CTA per SM: the
// per-SM load differs by at most one strip); a CTA processes its range in batches of NB strips.
struct VqTiles {
    int C;                 // columns per strip
    int SC;                // tokens per strip = 4 * C
    int tiles_x, tiles_y;  // strips per image row / strip rows per image
    int NB;                // strips per batch (NB * SC <= VQ_TILE)
    unsigned mSC, mC;      // ceil(2^32 / SC), ceil(2^32 / C): t / SC == __umulhi(t, mSC) for the small t used here
    int64_t n_strips;      // B * tiles_y * tiles_x
};

__host__ __device__ inline VqTiles make_tiles(int B, int h, int w)
{
    VqTiles t;
    const int w4 = (w + 3) & ~3;
    t.C = w4 < 256 ? w4 : 256;
    t.SC = 4 * t.C;
    t.tiles_x = (w + t.C - 1) / t.C;
    t.tiles_y = (h + 3) / 4;
    t.NB = VQ_TILE / t.SC;
    t.mSC = (unsigned)((((unsigned long long)1 << 32) + t.SC - 1) / t.SC);
    t.mC = (unsigned)((((unsigned long long)1 << 32) + t.C - 1) / t.C);
    t.n_strips = (int64_t)B * t.tiles_x * t.tiles_y;
    return t;
}

__global__ void __launch_bounds__(VQ_THREADS, VQ_CTAS)
vq_fused_kernel(const float *__restrict__ z, int B, int h, int w, const VqTiles tl, const float *__restrict__ codebook, int K, int Kpad,
                int64_t *__restrict__ idx_out, float *__restrict__ zq_out, double *__restrict__ partials,
                int32_t *__restrict__ counters, double *__restrict__ sqerr_out)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long mbar;
    __shared__ int s_count;
    __shared__ double s_red[VQ_WARPS];
    __shared__ bool s_last;
    __shared__ int s_sb[VQ_MAX_STRIPS], s_sy[VQ_MAX_STRIPS], s_sx[VQ_MAX_STRIPS];  // image, first row, first column of a strip
    float *raw = reinterpret_cast<float *>(smem);
    float2 *tab = reinterpret_cast<float2 *>(smem + (size_t)Kpad * 16);
    float4 *zs = reinterpret_cast<float4 *>(smem + (size_t)Kpad * 36);
    u64 *best = reinterpret_cast<u64 *>(zs + VQ_TILE);
    uint16_t *lead = reinterpret_cast<uint16_t *>(best + VQ_TILE);
    uint16_t *list = lead + VQ_TILE;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    VQ_STAMP(0);
    pdl_trigger_step<1>();
    pdl_wait();  // the codebook and z may come straight from a preceding kernel
    // --- stage the codebook with one TMA bulk copy (cp.async.bulk -> UBLKCP), mbarrier completion
    if (tid == 0) mbar_init(&mbar);
    __syncthreads();
    if (tid == 0) tma_load_1d(raw, codebook, (uint32_t)K * 16u, &mbar);
    bool staged = false;       // waited for the bulk copy (after the first batch's loads are in flight)
    bool table_ready = false;  // the pair table of the exhaustive search is built when a batch first needs it

    const int64_t plane = (int64_t)h * w;
    const int nchunks = Kpad / VQ_CHUNK;
    const u64 minus2 = pack2(-2.f, -2.f);
    const ulonglong2 *ctab = reinterpret_cast<const ulonglong2 *>(tab);
    const float *ftab = reinterpret_cast<const float *>(tab);
    const int C = tl.C, SC = tl.SC;
    double sq = 0.0;
    const int64_t s_begin = tl.n_strips * blockIdx.x / gridDim.x, s_end = tl.n_strips * (blockIdx.x + 1) / gridDim.x;

    for (int64_t s0 = s_begin; s0 < s_end; s0 += tl.NB) {
        const int ns = (int)min((int64_t)tl.NB, s_end - s0);
        const int ntok = ns * SC;
        __syncthreads();  // previous batch fully consumed (and the pair table complete on the first pass)
        VQ_STAMP(1);
        if (tid == 0) s_count = 0;
        if (tid < ns) {
            const int64_t id = s0 + tid;
            const int spi = tl.tiles_x * tl.tiles_y;
            const int b = (int)(id / spi), r = (int)(id - (int64_t)b * spi);
            s_sb[tid] = b;
            s_sy[tid] = (r / tl.tiles_x) * 4;
            s_sx[tid] = (r % tl.tiles_x) * C;
        }
        __syncthreads();
        // ---- load (4 tokens per thread and pass: 16 independent loads in flight)
        for (int t0 = tid; t0 < ntok; t0 += 4 * VQ_THREADS) {
            float4 v[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int t = t0 + i * VQ_THREADS;
                v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (t < ntok) {
                    const int sl = (int)__umulhi((unsigned)t, tl.mSC), r = t - sl * SC;
                    const int ly = (int)__umulhi((unsigned)r, tl.mC), lx = r - ly * C;
                    const int gy = s_sy[sl] + ly, gx = s_sx[sl] + lx;
                    if (gy < h && gx < w) {
                        const float *zb = z + (int64_t)s_sb[sl] * 4 * plane;
                        const int64_t p = (int64_t)gy * w + gx;
                        v[i] = make_float4(__ldg(zb + p), __ldg(zb + plane + p), __ldg(zb + 2 * plane + p), __ldg(zb + 3 * plane + p));
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (t0 + i * VQ_THREADS < ntok) zs[t0 + i * VQ_THREADS] = v[i];
        }
        if (!staged) {
            staged = true;
            mbar_wait(&mbar, 0);
        }
        __syncthreads();
        VQ_STAMP(2);
        // ---- classify + compact
        for (int t0 = 0; t0 < ntok; t0 += VQ_THREADS) {
            const int t = t0 + tid;
            bool leader = false;
            if (t < ntok) {
                const int sl = (int)__umulhi((unsigned)t, tl.mSC), r = t - sl * SC;
                const int ly = (int)__umulhi((unsigned)r, tl.mC), lx = r - ly * C;
                if (s_sy[sl] + ly < h && s_sx[sl] + lx < w) {
                    const uint4 v = reinterpret_cast<const uint4 *>(zs)[t];
                    int ld = t;
                    const int t4 = sl * SC + (lx & ~3);  // a strip is one row of 4x4 blocks
                    if (t4 != t) {
                        const uint4 u = reinterpret_cast<const uint4 *>(zs)[t4];
                        if (u.x == v.x && u.y == v.y && u.z == v.z && u.w == v.w) ld = t4;
                    }
                    if (ld == t) {
                        const int t2 = sl * SC + (ly & ~1) * C + (lx & ~1);
                        if (t2 != t) {
                            const uint4 u = reinterpret_cast<const uint4 *>(zs)[t2];
                            if (u.x == v.x && u.y == v.y && u.z == v.z && u.w == v.w) ld = t2;
                        }
                    }
                    leader = ld == t;
                    lead[t] = (uint16_t)ld;
                    if (leader) best[t] = ~0ull;
                }
            }
            const unsigned m = __ballot_sync(0xffffffffu, leader);
            int base = 0;
            if (lane == 0 && m) base = atomicAdd(&s_count, __popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (leader) list[base + __popc(m & ((1u << lane) - 1u))] = (uint16_t)t;  // order is irrelevant to the result
        }
        __syncthreads();
        VQ_STAMP(3);
        // ---- search: exhaustive, units = (64 leaders) x (chunk range) dealt evenly to the warps
        const int L = s_count;
        if (L > 0 && !table_ready) {
            table_ready = true;
        // --- pair table: row r of chunk c holds (v[8c+0],v[8c+1]) (v[8c+2],v[8c+3]) ... for v = e_r or e^2
        for (int pr = tid; pr < Kpad / 2; pr += VQ_THREADS) {
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
        }
        const int groups = (L + 32 * VQ_T - 1) / (32 * VQ_T);
        // (group, chunk) items are dealt to the warps as contiguous, equally long ranges: a warp
        // works on at most two groups and every warp gets the same number of chunks
        const int total_items = groups * nchunks;
        const int per_warp = (total_items + VQ_WARPS - 1) / VQ_WARPS;
        int u = warp * per_warp;
        const int u_end = min(total_items, u + per_warp);
        while (u < u_end) {
            const int g = u / nchunks;
            const int c_begin = u - g * nchunks;
            const int c_end = min(nchunks, c_begin + (u_end - u));
            u += c_end - c_begin;
            int tok[VQ_T];
            float zf[VQ_T][4], z2[VQ_T];
            u64 zd[VQ_T][4], zsum[VQ_T];
            float bestd[VQ_T];
            int bestc[VQ_T];
#pragma unroll
            for (int t = 0; t < VQ_T; ++t) {
                const int li = g * (32 * VQ_T) + t * 32 + lane;
                tok[t] = li < L ? (int)list[li] : -1;
                const float4 v = tok[t] >= 0 ? zs[tok[t]] : make_float4(0.f, 0.f, 0.f, 0.f);
                zf[t][0] = v.x;
                zf[t][1] = v.y;
                zf[t][2] = v.z;
                zf[t][3] = v.w;
                z2[t] = sumsq4(v.x, v.y, v.z, v.w);
#pragma unroll
                for (int c = 0; c < 4; ++c) zd[t][c] = pack2(zf[t][c], zf[t][c]);
                zsum[t] = pack2(z2[t], z2[t]);
                bestd[t] = __int_as_float(0x7f800000);
                bestc[t] = c_begin;
            }
            // per token the minimum distance of each chunk
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
                        u64 b2 = mul2(zd[t][0], e0.y);
                        a = fma2(zd[t][1], e1.x, a);
                        b2 = fma2(zd[t][1], e1.y, b2);
                        a = fma2(zd[t][2], e2.x, a);
                        b2 = fma2(zd[t][2], e2.y, b2);
                        a = fma2(zd[t][3], e3.x, a);
                        b2 = fma2(zd[t][3], e3.y, b2);
                        const u64 sa = add2(zsum[t], es.x);
                        const u64 sb = add2(zsum[t], es.y);
                        a = fma2(a, minus2, sa);
                        b2 = fma2(b2, minus2, sb);
                        float a0, a1, b0, b1;
                        unpack2(a, a0, a1);
                        unpack2(b2, b0, b1);
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
            // the first code of the winning chunk that attains the minimum; publish
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
        }
        __syncthreads();
        VQ_STAMP(4);
        // ---- finalize
        for (int t = tid; t < ntok; t += VQ_THREADS) {
            const int sl = (int)__umulhi((unsigned)t, tl.mSC), r = t - sl * SC;
            const int ly = (int)__umulhi((unsigned)r, tl.mC), lx = r - ly * C;
            const int gy = s_sy[sl] + ly, gx = s_sx[sl] + lx, b = s_sb[sl];
            if (gy >= h || gx >= w) continue;
            const u64 key = best[lead[t]];
            const int k = key == ~0ull ? 0 : (int)(uint32_t)key;  // every distance NaN / inf: torch.argmin gives 0 only
                                                                  // if all are equal; see tests (inputs are finite)
            const int64_t p = (int64_t)gy * w + gx;
            idx_out[(int64_t)b * plane + p] = k;
            if (zq_out || sqerr_out) {
                const float4 e = reinterpret_cast<const float4 *>(raw)[k];
                const float4 zv = zs[t];
                const float d0 = __fsub_rn(e.x, zv.x), d1 = __fsub_rn(e.y, zv.y), d2 = __fsub_rn(e.z, zv.z), d3 = __fsub_rn(e.w, zv.w);
                if (zq_out) {
                    float *q = zq_out + (int64_t)b * 4 * plane + p;
                    q[0] = __fadd_rn(zv.x, d0);
                    q[plane] = __fadd_rn(zv.y, d1);
                    q[2 * plane] = __fadd_rn(zv.z, d2);
                    q[3 * plane] = __fadd_rn(zv.w, d3);
                }
                float acc = __fmul_rn(d0, d0);
                acc = __fmaf_rn(d1, d1, acc);
                acc = __fmaf_rn(d2, d2, acc);
                acc = __fmaf_rn(d3, d3, acc);
                sq += (double)acc;
            }
        }
    }
    VQ_STAMP(5);
    if (!staged) mbar_wait(&mbar, 0);  // never leave with the bulk copy in flight
    if (!sqerr_out) return;
    // deterministic reduction: warp shuffle -> CTA -> per-CTA partial -> the last CTA sums them in order
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    if (lane == 0) s_red[warp] = sq;
    __syncthreads();
    if (tid == 0) {
        double tot = 0.0;
        for (int i = 0; i < VQ_WARPS; ++i) tot += s_red[i];
        partials[blockIdx.x] = tot;
        __threadfence();
        s_last = (atomicAdd(&counters[0], 1) == (int)gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        double tot = 0.0;
        for (int i = tid; i < (int)gridDim.x; i += VQ_THREADS) tot += __ldcg(&partials[i]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
        if (lane == 0) s_red[warp] = tot;
        __syncthreads();
        if (tid == 0) {
            double all = 0.0;
            for (int i = 0; i < VQ_WARPS; ++i) all += s_red[i];
            *sqerr_out = all;
            counters[0] = 0;  // leave the ticket zeroed for the next launch (workspace contract)
        }
    }
    VQ_STAMP(6);
}

// ---- indexed search, warp-autonomous ---------------------------------------------------------
// The prepared-codebook path needs so little arithmetic per token that CTA-wide phases (and their
// barriers) dominate; here every WARP owns a tile of 4 rows x 32 columns of one image and runs
// load -> classify -> search -> finalize on its own, synchronising with __syncwarp only, so the
// warps of an SM sit in different phases and hide each other's memory latency.
//   load      lane = column: 4 rows x 4 channels per lane, each warp load is one 128-byte row segment;
//             the tile is also written to the warp's shared-memory slice as one float4 per token;
//   classify  as in vq_fused_kernel (bit-identical to the top-left token of the 4x4 / 2x2 block =>
//             follower); leaders are compacted with ballots, no atomics;
//   search    one lane per leader: grid cell -> candidate record (4 x 16-byte loads) -> the reference's
//             rounding sequence on the candidates; leaders the index cannot serve are searched
//             exhaustively by their lane (rare; same result);
//   finalize  idx / z_q row segments, sum((e - z)^2) per lane, reduced per CTA at the end.
constexpr int VQW_THREADS = 256;
constexpr int VQW_WARPS = VQW_THREADS / 32;
__host__ __device__ inline size_t vqw_smem_bytes(int K) { return vq_stage_bytes(K) + (size_t)VQW_WARPS * (VQW_TILE_BYTES + VQW_REC_BYTES); }

#ifndef CGIC_VQW_CTAS
#define CGIC_VQW_CTAS 2
#endif
__global__ void __launch_bounds__(VQW_THREADS, CGIC_VQW_CTAS)
vq_warp_kernel(const float *__restrict__ z, int h, int w, int tiles_x, int tiles_y, int64_t n_tiles, const unsigned char *__restrict__ blob,
               int K, int64_t *__restrict__ idx_out, float *__restrict__ zq_out, double *__restrict__ partials,
               int32_t *__restrict__ counters, double *__restrict__ sqerr_out, int defer_reduce)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long mbar;
    __shared__ double s_red[VQW_WARPS];
    __shared__ bool s_last;
    const CbLayout CL = cb_layout(K);
    const CbHeader *hdr = reinterpret_cast<const CbHeader *>(smem);
    const unsigned char *lut = smem + CL.lut;
    const float4 *cbs = reinterpret_cast<const float4 *>(smem + CL.cb);
    const float *e2s = reinterpret_cast<const float *>(smem + CL.e2);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned char *wbase = smem + vq_stage_bytes(K) + (size_t)warp * (VQW_TILE_BYTES + VQW_REC_BYTES);
    float4 *zs = reinterpret_cast<float4 *>(wbase);                      // [128] the tile, row-major
    uint16_t *res = reinterpret_cast<uint16_t *>(zs + VQW_TILE);         // [128] code of a leader token
    uint8_t *list = reinterpret_cast<uint8_t *>(res + VQW_TILE);         // [128] leader tokens, compacted
    uint8_t *lead = list + VQW_TILE;                                     // [128] leader of every token
    const uint4 *recs = reinterpret_cast<const uint4 *>(blob + CL.rec);

    VQ_STAMP(0);
    pdl_trigger_step<1>();
    pdl_wait();  // the prepared blob and z may come straight from a preceding kernel
    VQ_STAMP(1);
    if (tid == 0) mbar_init(&mbar);
    __syncthreads();
    if (tid == 0) tma_load_1d(smem, blob, (uint32_t)CL.stage, &mbar);
    bool staged = false;

    const int64_t plane = (int64_t)h * w;
    const int tpi = tiles_x * tiles_y;
    double sq = 0.0;
    // tiles are dealt round-robin over the CTAs (tile i -> CTA i % grid), so every SM gets the same number of busy warps
    const int64_t wid = (int64_t)warp * gridDim.x + blockIdx.x, nwarps = (int64_t)gridDim.x * VQW_WARPS;
    for (int64_t tile = wid; tile < n_tiles; tile += nwarps) {
        const int b = (int)(tile / tpi), rt = (int)(tile - (int64_t)b * tpi);
        const int ty = rt / tiles_x, tx = rt - ty * tiles_x;
        VqTileCtx ctx{hdr, lut, cbs, e2s, recs, K, zs, res, list, lead, reinterpret_cast<uint4 *>(wbase + VQW_TILE_BYTES)};
        vq_process_tile(ctx, z + (int64_t)b * 4 * plane, h, w, ty * 4, tx * 32 + lane, lane, idx_out + (int64_t)b * plane,
                        zq_out ? zq_out + (int64_t)b * 4 * plane : nullptr, sqerr_out != nullptr, sq, (uint16_t *)nullptr, counters + 2,
                        [&]() {
                            if (!staged) {
                                staged = true;
                                mbar_wait(&mbar, 0);
                            }
                        },
                        [&](int k) {
                            (void)k;
                            if (tile == wid) VQ_STAMP(k);  // the warp's first tile (thread 0 stamps)
                        });
    }
    VQ_STAMP(5);
    if (!staged) mbar_wait(&mbar, 0);  // never leave with the bulk copy in flight
    if (!sqerr_out) return;
    // deterministic reduction: warp shuffle -> CTA -> per-CTA partial -> the last CTA sums them in order
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    if (lane == 0) s_red[warp] = sq;
    __syncthreads();
    if (tid == 0) {
        double tot = 0.0;
        for (int i = 0; i < VQW_WARPS; ++i) tot += s_red[i];
        partials[blockIdx.x] = tot;
        // defer_reduce: the per-CTA partials are summed (in CTA order) by the kernel that follows in the stream -- the packer of
        // cgic_encode -- instead of by the last CTA to finish here, whose ~2 us of ticket + reload + reduction would otherwise sit
        // on the step's critical path
        s_last = false;
        if (!defer_reduce) {
            __threadfence();
            s_last = (atomicAdd(&counters[0], 1) == (int)gridDim.x - 1);
        }
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        double tot = 0.0;
        for (int i = tid; i < (int)gridDim.x; i += VQW_THREADS) tot += __ldcg(&partials[i]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
        if (lane == 0) s_red[warp] = tot;
        __syncthreads();
        if (tid == 0) {
            double all = 0.0;
            for (int i = 0; i < VQW_WARPS; ++i) all += s_red[i];
            *sqerr_out = all;
            counters[0] = 0;  // leave the ticket zeroed for the next launch (workspace contract)
        }
    }
    VQ_STAMP(6);
}

__global__ void vq_count_kernel(const int64_t *__restrict__ idx, int64_t n, float *__restrict__ counters, int K)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t k = idx[i];
    // one exact +1 per token, like the reference loop (fp32 counters stop growing at 2^24 there too)
    if (k >= 0 && k < K) atomicAdd(&counters[k], 1.0f);
}

}  // namespace
}  // namespace cgic

using namespace cgic;

// workspace: [64 bytes of counters][one double per CTA]; sized for any grid this library launches.
// Contract: the first 64 bytes must be ZERO before the first use; the kernel leaves them zero.
extern "C" size_t cgic_vq_workspace_bytes(int64_t) { return 256 + 8 * 4096; }

static int vq_launch(const char *who, const float *z, int B, int h, int w, const float *codebook, int K, int64_t *idx_out, float *zq_out,
                     double *sqerr_out, void *workspace, size_t workspace_bytes, cgic_stream_t stream_)
{
    CGIC_REQUIRE(z && idx_out && workspace, CGIC_EINVAL, "%s: null argument", who);
    CGIC_REQUIRE(B >= 0 && h > 0 && w > 0, CGIC_EINVAL, "%s: bad shape B=%d h=%d w=%d", who, B, h, w);
    CGIC_REQUIRE(K >= 1 && K <= VQ_MAX_K, CGIC_EINVAL, "%s: K=%d outside [1, %d]", who, K, VQ_MAX_K);
    const int64_t n = (int64_t)B * h * w;
    CGIC_REQUIRE(n < (int64_t)1 << 31, CGIC_EINVAL, "%s: %lld tokens exceed 2^31", who, (long long)n);
    cudaStream_t stream = as_stream(stream_);
    if (n == 0) {
        if (sqerr_out) CGIC_CUDA_CHECK(cudaMemsetAsync(sqerr_out, 0, sizeof(double), stream));
        return CGIC_OK;
    }
    CGIC_REQUIRE(workspace_bytes >= cgic_vq_workspace_bytes(n), CGIC_ESPACE, "%s: workspace %zu < %zu bytes", who, workspace_bytes,
                 cgic_vq_workspace_bytes(n));
    int32_t *counters = static_cast<int32_t *>(workspace);
    double *partials = reinterpret_cast<double *>(static_cast<unsigned char *>(workspace) + 256);

    int n_sm = 0;
    {
        int rc = device_sm_count(&n_sm);
        if (!rc) rc = ensure_smem((const void *)vq_fused_kernel, vq_smem_bytes(VQ_MAX_K));
        if (rc) return rc;
    }
    const int Kpad = (K + VQ_CHUNK - 1) / VQ_CHUNK * VQ_CHUNK;
    const size_t smem = vq_smem_bytes(Kpad);
    const VqTiles tl = make_tiles(B, h, w);
    // persistent grid: VQ_CTAS CTAs per SM (when the codebook leaves room), never more CTAs than strips
    const int per_sm = smem * VQ_CTAS <= 220 * 1024 ? VQ_CTAS : 1;
    const int64_t cap = (int64_t)n_sm * per_sm;
    const int grid = (int)(tl.n_strips < cap ? tl.n_strips : cap);
    {
        CGIC_PROF("vq_fused_kernel", stream);
        CGIC_CUDA_CHECK(launch_pdl(vq_fused_kernel, dim3(grid), dim3(VQ_THREADS), smem, stream, z, B, h, w, tl, codebook, K, Kpad, idx_out, zq_out,
                                   partials, counters, sqerr_out));
    }
    CGIC_LAUNCH_CHECK();
    return CGIC_OK;
}

extern "C" int cgic_vq_assign(const float *z, int B, int h, int w, const float *codebook, int K, int64_t *idx_out,
                              float *zq_out, double *sqerr_out, void *workspace, size_t workspace_bytes,
                              cgic_stream_t stream)
{
    CGIC_REQUIRE(codebook, CGIC_EINVAL, "cgic_vq_assign: null argument");
    CGIC_REQUIRE((reinterpret_cast<uintptr_t>(codebook) & 15) == 0, CGIC_EINVAL, "cgic_vq_assign: codebook must be 16-byte aligned");
    return vq_launch("cgic_vq_assign", z, B, h, w, codebook, K, idx_out, zq_out, sqerr_out, workspace, workspace_bytes, stream);
}

namespace cgic {
int vq_assign_indexed_launch(const float *z, int B, int h, int w, const cgic_codebook *cb, int64_t *idx_out, float *zq_out, double *sqerr_out,
                             void *workspace, size_t workspace_bytes, cgic_stream_t stream_, bool defer_reduce, int *grid_out);
}

extern "C" int cgic_vq_assign_indexed(const float *z, int B, int h, int w, const cgic_codebook *cb, int64_t *idx_out, float *zq_out,
                                      double *sqerr_out, void *workspace, size_t workspace_bytes, cgic_stream_t stream_)
{
    return vq_assign_indexed_launch(z, B, h, w, cb, idx_out, zq_out, sqerr_out, workspace, workspace_bytes, stream_, false, nullptr);
}

// defer_reduce (needs sqerr_out): the kernel leaves one partial per CTA at workspace + 256 ([*grid_out] doubles, CTA order) and
// does NOT write *sqerr_out; the caller's next kernel sums them.
int cgic::vq_assign_indexed_launch(const float *z, int B, int h, int w, const cgic_codebook *cb, int64_t *idx_out, float *zq_out, double *sqerr_out,
                                   void *workspace, size_t workspace_bytes, cgic_stream_t stream_, bool defer_reduce, int *grid_out)
{
    if (grid_out) *grid_out = 0;
    const unsigned char *blob = codebook_blob(cb);
    CGIC_REQUIRE(blob, CGIC_EINVAL, "cgic_vq_assign_indexed: no prepared codebook (cgic_codebook_update has not been called)");
    CGIC_REQUIRE(z && idx_out && workspace, CGIC_EINVAL, "cgic_vq_assign_indexed: null argument");
    CGIC_REQUIRE(B >= 0 && h > 0 && w > 0, CGIC_EINVAL, "cgic_vq_assign_indexed: bad shape B=%d h=%d w=%d", B, h, w);
    const int K = codebook_size(cb);
    const int64_t n = (int64_t)B * h * w;
    CGIC_REQUIRE(n < (int64_t)1 << 31, CGIC_EINVAL, "cgic_vq_assign_indexed: %lld tokens exceed 2^31", (long long)n);
    cudaStream_t stream = as_stream(stream_);
    if (n == 0) {
        if (sqerr_out) CGIC_CUDA_CHECK(cudaMemsetAsync(sqerr_out, 0, sizeof(double), stream));
        return CGIC_OK;
    }
    CGIC_REQUIRE(workspace_bytes >= cgic_vq_workspace_bytes(n), CGIC_ESPACE, "cgic_vq_assign_indexed: workspace %zu < %zu bytes",
                 workspace_bytes, cgic_vq_workspace_bytes(n));
    int32_t *counters = static_cast<int32_t *>(workspace);
    double *partials = reinterpret_cast<double *>(static_cast<unsigned char *>(workspace) + 256);
    int n_sm = 0;
    {
        int rc = device_sm_count(&n_sm);
        if (!rc) rc = ensure_smem((const void *)vq_warp_kernel, vqw_smem_bytes(VQ_MAX_K));
        if (rc) return rc;
    }
    const int tiles_x = (w + 31) / 32, tiles_y = (h + 3) / 4;
    const int64_t n_tiles = (int64_t)B * tiles_x * tiles_y;
    // one tile per warp while the grid fits the machine (2 CTAs of 8 warps per SM), persistent beyond that; the grid is a
    // multiple of the SM count so that the round-robin deal leaves every SM with the same load
    const int64_t want = (n_tiles + VQW_WARPS - 1) / VQW_WARPS;
    int64_t per_sm = (want + n_sm - 1) / n_sm;
    if (per_sm > CGIC_VQW_CTAS) per_sm = CGIC_VQW_CTAS;
    const int64_t full = per_sm * n_sm;
    const int grid = (int)(n_tiles < full ? n_tiles : full);
    {
        CGIC_PROF("vq_warp_kernel", stream);
        CGIC_CUDA_CHECK(launch_pdl(vq_warp_kernel, dim3(grid), dim3(VQW_THREADS), vqw_smem_bytes(K), stream, z, h, w, tiles_x, tiles_y, n_tiles,
                                   blob, K, idx_out, zq_out, partials, counters, sqerr_out, (int)(defer_reduce && sqerr_out != nullptr)));
    }
    CGIC_LAUNCH_CHECK();
    if (grid_out) *grid_out = grid;
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
