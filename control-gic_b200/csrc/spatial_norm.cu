// f4 (second half): SpatialNorm, the conditioning every decoder block applies to its feature map
// (CGIC/modules/vqvae/decoder.py:34-53):
//     zq_up = nearest(zq -> f's H x W);  new_f = GroupNorm(f) * conv_y(zq_up) + conv_b(zq_up)      conv_y / conv_b are 1x1
// The reference materialises zq_up [B,Cz,H,W], conv_y(zq_up) and conv_b(zq_up) [B,C,H,W], GroupNorm(f), the product and
// the sum: six passes over a [B,C,H,W] tensor.  Here: one pass that reads f for the group statistics, one pass that
// reads f (from L2 when the feature map fits its 126 MB) and the tiny zq at its own resolution, evaluates both 1x1
// convolutions on the fly (2 * Cz FMAs per output) and writes new_f.  HBM-bound: 8 B (f resident in L2) or 12 B per element.
#include <algorithm>
#include <cmath>

#include "common.cuh"

namespace cgic {
namespace {

constexpr int SN_THREADS = 256;
constexpr int SN_MAX_CZ = 8;     // zq channels (Control-GIC: 4)
constexpr int SN_MAX_PARTS = 32; // CTAs one (image, group) slab may be split over for the statistics
constexpr int SN_ITEMS = 4;      // vectors per thread in the apply kernel

struct SnPartial {
    double sum, sumsq;  // of (x - pivot), pivot = first element of the slab
};

struct SnArgs {
    const float *f, *zq, *gamma, *beta, *wy, *by, *wb, *bb;
    float *out;
    const SnPartial *ws;
    int B, C, H, W, Cz, hz, wz, G, S;
    float eps, sh, sw;  // sh / sw: torch's nearest scales, (float)in / out
};

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// grid (S, G * B): CTA (s, g) sums part s of slab g.  Sums are taken of (x - pivot): the pivot is a sample of the slab,
// so the accumulated values are of the size of the spread, not of the mean -- E[d^2] - E[d]^2 does not cancel.
__global__ void __launch_bounds__(SN_THREADS) sn_stats_kernel(const float *__restrict__ f, int64_t slab, int64_t part, SnPartial *__restrict__ ws)
{
    pdl_launch_dependents();
    pdl_wait();
    const float *base = f + (int64_t)blockIdx.y * slab;
    const float pivot = __ldg(base);
    const int64_t lo = (int64_t)blockIdx.x * part, hi = min(slab, lo + part);
    float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
    if ((slab & 3) == 0) {  // every slab starts 16-byte aligned (part is a multiple of 4)
        const float4 *v = reinterpret_cast<const float4 *>(base + lo);
        const int64_t n4 = (hi - lo) >> 2;
        int64_t i = threadIdx.x;
        for (; i + 3 * SN_THREADS < n4; i += 4 * SN_THREADS) {  // four 16-byte loads in flight
            float4 t[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) t[u] = __ldg(v + i + u * SN_THREADS);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float d[4] = {t[u].x - pivot, t[u].y - pivot, t[u].z - pivot, t[u].w - pivot};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    s[k] += d[k];
                    q[k] = fmaf(d[k], d[k], q[k]);
                }
            }
        }
        for (; i < n4; i += SN_THREADS) {
            const float4 a = __ldg(v + i);
            const float d[4] = {a.x - pivot, a.y - pivot, a.z - pivot, a.w - pivot};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                s[k] += d[k];
                q[k] = fmaf(d[k], d[k], q[k]);
            }
        }
    } else {
        for (int64_t i = lo + threadIdx.x; i < hi; i += SN_THREADS) {
            const float d = __ldg(base + i) - pivot;
            s[0] += d;
            q[0] = fmaf(d, d, q[0]);
        }
    }
    double ds = ((double)s[0] + (double)s[1]) + ((double)s[2] + (double)s[3]);
    double dq = ((double)q[0] + (double)q[1]) + ((double)q[2] + (double)q[3]);
    ds = warp_sum(ds);
    dq = warp_sum(dq);
    __shared__ double sh_s[SN_THREADS / 32], sh_q[SN_THREADS / 32];
    if ((threadIdx.x & 31) == 0) {
        sh_s[threadIdx.x >> 5] = ds;
        sh_q[threadIdx.x >> 5] = dq;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
#pragma unroll
        for (int k = 0; k < SN_THREADS / 32; ++k) {  // fixed order: deterministic
            a += sh_s[k];
            b += sh_q[k];
        }
        ws[(int64_t)blockIdx.y * gridDim.x + blockIdx.x] = SnPartial{a, b};
    }
}

// torch's nearest source index (UpSample.h nearest_neighbor_compute_source_index): min(floor(dst * scale), in - 1)
__device__ __forceinline__ int nearest_src(int dst, float scale, int in) { return min((int)floorf((float)dst * scale), in - 1); }

// grid (ceil(H*W / VEC / (SN_THREADS * SN_ITEMS)), C, B): one channel plane (or a piece of it) per CTA, so that the
// statistics of its group and the channel's 2 * (Cz + 1) + 2 coefficients are CTA-uniform.
template <int VEC, int CZ>  // CZ: zq channels at compile time (4 = Control-GIC), 0 = a.Cz <= SN_MAX_CZ at run time
__global__ void __launch_bounds__(SN_THREADS) sn_apply_kernel(const SnArgs a)
{
    pdl_launch_dependents();
    pdl_wait();
    const int c = blockIdx.y, b = blockIdx.z;
    const int cpg = a.C / a.G, g = c / cpg;
    const int64_t hw = (int64_t)a.H * a.W, slab = (int64_t)cpg * hw;
    __shared__ float s_stat[2];
    if (threadIdx.x == 0) {
        const SnPartial *p = a.ws + ((int64_t)b * a.G + g) * a.S;
        double s = 0.0, q = 0.0;
        for (int i = 0; i < a.S; ++i) {
            s += p[i].sum;
            q += p[i].sumsq;
        }
        const double m = s / (double)slab;
        const double var = fmax(q / (double)slab - m * m, 0.0);  // biased, like GroupNorm
        s_stat[0] = (float)((double)__ldg(a.f + ((int64_t)b * a.G + g) * slab) + m);
        s_stat[1] = (float)(1.0 / sqrt(var + (double)a.eps));
    }
    constexpr int NZ = CZ ? CZ : SN_MAX_CZ;
    const int cz = CZ ? CZ : a.Cz;
    float wy[NZ], wb[NZ];
#pragma unroll
    for (int k = 0; k < NZ; ++k) {
        wy[k] = k < cz ? __ldg(a.wy + (int64_t)c * cz + k) : 0.f;
        wb[k] = k < cz ? __ldg(a.wb + (int64_t)c * cz + k) : 0.f;
    }
    const float by = a.by ? __ldg(a.by + c) : 0.f, bb = a.bb ? __ldg(a.bb + c) : 0.f;
    const float gam = a.gamma ? __ldg(a.gamma + c) : 1.f, bet = a.beta ? __ldg(a.beta + c) : 0.f;
    __syncthreads();
    const float mean = s_stat[0], rstd = s_stat[1];
    const float *fp = a.f + ((int64_t)b * a.C + c) * hw;
    float *op = a.out + ((int64_t)b * a.C + c) * hw;
    const float *zb = a.zq + (int64_t)b * cz * a.hz * a.wz;
    const int64_t zplane = (int64_t)a.hz * a.wz;
    const int64_t nvec = hw / VEC;
    // all of the thread's loads first (independent of the statistics: issued before the barrier above would be even
    // earlier, but the registers are better spent on occupancy), then the arithmetic
    float v[SN_ITEMS][VEC];
#pragma unroll
    for (int it = 0; it < SN_ITEMS; ++it) {
        const int64_t i = ((int64_t)blockIdx.x * SN_ITEMS + it) * SN_THREADS + threadIdx.x;
        if (i < nvec) {
            if constexpr (VEC == 4) {
                const float4 t = __ldg(reinterpret_cast<const float4 *>(fp) + i);
                v[it][0] = t.x, v[it][1] = t.y, v[it][2] = t.z, v[it][3] = t.w;
            } else {
                v[it][0] = __ldg(fp + i);
            }
        }
    }
#pragma unroll
    for (int it = 0; it < SN_ITEMS; ++it) {
        const int64_t i = ((int64_t)blockIdx.x * SN_ITEMS + it) * SN_THREADS + threadIdx.x;
        if (i >= nvec) break;
        const int64_t e0 = i * VEC;
        const int y = (int)(e0 / a.W), x0 = (int)(e0 - (int64_t)y * a.W);  // VEC == 4 only when W % 4 == 0: one row
        const float *zrow = zb + (int64_t)nearest_src(y, a.sh, a.hz) * a.wz;
        float sy = 0.f, sb = 0.f;
        int xs_prev = -1;
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            const int xs = nearest_src(x0 + j, a.sw, a.wz);
            if (xs != xs_prev) {  // the 1x1 convolutions at this source pixel
                sy = by, sb = bb;
#pragma unroll
                for (int k = 0; k < NZ; ++k)
                    if (k < cz) {
                        const float z = __ldg(zrow + k * zplane + xs);
                        sy = fmaf(wy[k], z, sy);
                        sb = fmaf(wb[k], z, sb);
                    }
                xs_prev = xs;
            }
            const float nrm = (v[it][j] - mean) * rstd * gam + bet;
            v[it][j] = nrm * sy + sb;
        }
        if constexpr (VEC == 4) reinterpret_cast<float4 *>(op)[i] = make_float4(v[it][0], v[it][1], v[it][2], v[it][3]);
        else op[i] = v[it][0];
    }
}

}  // namespace
}  // namespace cgic

using namespace cgic;

extern "C" size_t cgic_spatial_norm_workspace_bytes(int B, int groups)
{
    return (size_t)(B > 0 ? B : 0) * (size_t)(groups > 0 ? groups : 0) * SN_MAX_PARTS * sizeof(SnPartial) + 16;
}

extern "C" int cgic_spatial_norm(const float *f, const float *zq, const float *gn_weight, const float *gn_bias, const float *wy, const float *by,
                                 const float *wb, const float *bb, int B, int C, int H, int W, int Cz, int hz, int wz, int groups, float eps,
                                 float *out, void *workspace, size_t workspace_bytes, cgic_stream_t stream_)
{
    CGIC_REQUIRE(f && zq && wy && wb && out && workspace, CGIC_EINVAL, "cgic_spatial_norm: null argument");
    CGIC_REQUIRE(B >= 0 && C > 0 && H > 0 && W > 0 && hz > 0 && wz > 0 && groups > 0 && C % groups == 0, CGIC_EINVAL,
                 "cgic_spatial_norm: bad shape B=%d C=%d %dx%d zq %dx%d groups=%d", B, C, H, W, hz, wz, groups);
    CGIC_REQUIRE(Cz >= 1 && Cz <= SN_MAX_CZ, CGIC_EINVAL, "cgic_spatial_norm: zq channels %d outside 1..%d", Cz, SN_MAX_CZ);
    CGIC_REQUIRE(B <= 65535 && C <= 65535, CGIC_EINVAL, "cgic_spatial_norm: B and C must be <= 65535");
    CGIC_REQUIRE(workspace_bytes >= cgic_spatial_norm_workspace_bytes(B, groups), CGIC_EINVAL, "cgic_spatial_norm: workspace too small");
    for (const void *ptr : {(const void *)f, (const void *)out, (const void *)workspace})
        CGIC_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, CGIC_EINVAL, "cgic_spatial_norm: f, out and workspace must be 16-byte aligned");
    if (B == 0) return CGIC_OK;
    cudaStream_t stream = as_stream(stream_);
    const int64_t hw = (int64_t)H * W, slab = (int64_t)(C / groups) * hw;
    // statistics: enough CTAs for two per SM, at least 4096 elements each
    int64_t S = (2 * 148 + (int64_t)B * groups - 1) / ((int64_t)B * groups);
    S = std::min<int64_t>(std::min<int64_t>(S, SN_MAX_PARTS), std::max<int64_t>(1, slab / 4096));
    const int64_t part = ((slab + S - 1) / S + 3) & ~(int64_t)3;
    S = (slab + part - 1) / part;
    SnPartial *ws = static_cast<SnPartial *>(workspace);
    {
        CGIC_PROF("sn_stats_kernel", stream);
        CGIC_CUDA_CHECK(launch_pdl(sn_stats_kernel, dim3((unsigned)S, (unsigned)(B * groups)), dim3(SN_THREADS), 0, stream, f, slab, part, ws));
    }
    SnArgs a{f, zq, gn_weight, gn_bias, wy, by, wb, bb, out, ws, B, C, H, W, Cz, hz, wz, groups, (int)S, eps, (float)hz / (float)H, (float)wz / (float)W};
    {
        CGIC_PROF("sn_apply_kernel", stream);
        const int vec = (W & 3) == 0 ? 4 : 1;
        const int64_t nvec = hw / vec;
        const dim3 grid((unsigned)((nvec + SN_THREADS * SN_ITEMS - 1) / (SN_THREADS * SN_ITEMS)), (unsigned)C, (unsigned)B);
        void (*kernel)(SnArgs) = vec == 4 ? (Cz == 4 ? sn_apply_kernel<4, 4> : sn_apply_kernel<4, 0>) : (Cz == 4 ? sn_apply_kernel<1, 4> : sn_apply_kernel<1, 0>);
        CGIC_CUDA_CHECK(launch_pdl(kernel, grid, dim3(SN_THREADS), 0, stream, a));
    }
    CGIC_LAUNCH_CHECK();
    return CGIC_OK;
}
