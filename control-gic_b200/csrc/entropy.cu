// a4  Entropy.forward / Entropy.entropy  (CGIC/models/model.py:440-483), both patch sizes fused.
//
// Reference: gray image -> unfold into p x p patches -> for each patch a 32-bin soft histogram
//   pdf_j = mean_i exp(-0.5 * ((v_i - bin_j) / 0.01)^2),  pdf = pdf / (sum + 1e-40) + 1e-40,
//   H = -sum_j pdf_j * log(pdf_j);  it materialises [patches, p*p, 32] temporaries per scale.
// Here one CTA owns one 16x16 pixel block (= one p=16 patch = four p=8 patches) and reads the
// image exactly once:
//   phase 1  one thread per pixel: gray value, then the kernel value for the 7 bins around the
//            nearest bin.  sigma = 0.01 against a bin spacing of 2/31 makes every other term
//            exp(-130) or smaller, which IS 0.0f in fp32 -- skipping them is exact, not an
//            approximation.
//   phase 2  one warp per 8x8 patch, lane = bin: a fixed-order (row-major) sum of the 64 pixel
//            contributions -> pdf8; the p=16 histogram is the sum of the four (deterministic).
//   phase 3  normalisation and -sum p*log(p) by warp shuffles.
// fp32 throughout, denormals kept (eps = 1e-40 is a denormal; flush-to-zero would turn every
// entropy into NaN).  Float-tolerance parity (the reference's own summation order is torch's).
#include "common.cuh"

namespace cgic {
namespace {

struct Bins {
    float v[32];
};

constexpr int EN_WIN = 7;

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float entropy_of(float pdf_lane)
{
    const float eps = 1e-40f;
    const float norm = warp_sum(pdf_lane) + eps;
    const float p = pdf_lane / norm + eps;
    return -warp_sum(p * logf(p));
}

__global__ void __launch_bounds__(256)
entropy_kernel(const float *__restrict__ x, int H, int W, const Bins bins, float *__restrict__ e8, float *__restrict__ e16)
{
    __shared__ float s_val[256][EN_WIN];
    __shared__ signed char s_lo[256];
    __shared__ float s_bins[32];
    __shared__ float s_sum[4][32];
    const int tid = threadIdx.x;
    const int b = blockIdx.z;
    const int64_t plane = (int64_t)H * W;
    if (tid < 32) s_bins[tid] = bins.v[tid];
    __syncthreads();
    {
        const int py = tid >> 4, px = tid & 15;
        const int64_t o = (int64_t)(blockIdx.y * 16 + py) * W + blockIdx.x * 16 + px;
        const float *xb = x + (int64_t)b * 3 * plane;
        // 0.2989*R + 0.5870*G + 0.1140*B, each product and sum rounded (model.py:471)
        float g = __fadd_rn(__fmul_rn(0.2989f, xb[o]), __fmul_rn(0.5870f, xb[plane + o]));
        g = __fadd_rn(g, __fmul_rn(0.1140f, xb[2 * plane + o]));
        int jc = __float2int_rn((g + 1.0f) * 15.5f);
        jc = max(0, min(31, jc));
        const int lo = jc - EN_WIN / 2;
        s_lo[tid] = (signed char)lo;
#pragma unroll
        for (int r = 0; r < EN_WIN; ++r) {
            const int j = lo + r;
            float v = 0.f;
            if (j >= 0 && j < 32) {
                const float q = __fdiv_rn(__fsub_rn(g, s_bins[j]), 0.01f);
                v = expf(__fmul_rn(-0.5f, __fmul_rn(q, q)));
            }
            s_val[tid][r] = v;
        }
    }
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31;
    if (warp < 4) {
        const int sy = warp >> 1, sx = warp & 1;
        float acc = 0.f;
        for (int i = 0; i < 64; ++i) {
            const int t = ((sy * 8 + (i >> 3)) << 4) + sx * 8 + (i & 7);
            const int r = lane - (int)s_lo[t];
            if (r >= 0 && r < EN_WIN) acc += s_val[t][r];
        }
        s_sum[warp][lane] = acc;
        if (e8) {
            const float ent = entropy_of(acc / 64.0f);
            if (lane == 0)
                e8[((int64_t)b * (H / 8) + blockIdx.y * 2 + sy) * (W / 8) + blockIdx.x * 2 + sx] = ent;
        }
    }
    __syncthreads();
    if (warp == 0 && e16) {
        const float tot = (s_sum[0][lane] + s_sum[1][lane]) + (s_sum[2][lane] + s_sum[3][lane]);
        const float ent = entropy_of(tot / 256.0f);
        if (lane == 0) e16[((int64_t)b * (H / 16) + blockIdx.y) * (W / 16) + blockIdx.x] = ent;
    }
}

}  // namespace
}  // namespace cgic

using namespace cgic;

extern "C" int cgic_entropy_maps(const float *x, int B, int H, int W, const float *bins32_host, float *e8_out, float *e16_out,
                                 cgic_stream_t stream)
{
    CGIC_REQUIRE(x && bins32_host && (e8_out || e16_out), CGIC_EINVAL, "cgic_entropy_maps: null argument");
    CGIC_REQUIRE(B >= 0 && H > 0 && W > 0 && H % 16 == 0 && W % 16 == 0, CGIC_EINVAL,
                 "cgic_entropy_maps: image %dx%d must be multiples of 16", H, W);
    CGIC_REQUIRE(B <= 65535 && H / 16 <= 65535, CGIC_EINVAL, "cgic_entropy_maps: grid too large");
    if (B == 0) return CGIC_OK;
    Bins bins;
    for (int i = 0; i < 32; ++i) bins.v[i] = bins32_host[i];
    {
        CGIC_PROF("entropy_kernel", as_stream(stream));
        entropy_kernel<<<dim3(W / 16, H / 16, B), 256, 0, as_stream(stream)>>>(x, H, W, bins, e8_out, e16_out);
    }
    CGIC_LAUNCH_CHECK();
    return CGIC_OK;
}
