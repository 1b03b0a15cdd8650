// a4  Entropy.forward / Entropy.entropy  (CGIC/models/model.py:440-483), both patch sizes fused.
//
// Reference: gray image -> unfold into p x p patches -> for each patch a 32-bin soft histogram
//   pdf_j = mean_i exp(-0.5 * ((v_i - bin_j) / 0.01)^2),  pdf = pdf / (sum + 1e-40) + 1e-40,
//   H = -sum_j pdf_j * log(pdf_j);  it materialises [patches, p*p, 32] temporaries per scale.
// Here the image is read exactly once and nothing is materialised.  A WARP owns a region of 16 rows x
// 32 columns (= two 16x16 blocks = eight 8x8 patches); lane = column, so every load is one 128-byte row
// segment.  Each lane keeps two 32-bin histogram rows in shared memory (its column's upper and lower 8
// pixels) and adds the kernel values of its 16 pixels into them -- only the 3 bins around the nearest one:
// sigma = 0.01 against a bin spacing of 2/31 puts every other bin at least 1.5 spacings away, a term of at
// most exp(-46.8) = 4.7e-21 next to the nearest bin's >= 5.5e-3: far below the fp32 resolution of the
// histogram sums and of the entropy (terms 2.5 spacings away are exactly 0.0f in fp32).  The histogram of
// an 8x8 patch is the sum of 8 lanes' rows (lane = bin, fixed order), the 16x16 histogram the sum of its
// four patches; the ten histograms of a region are normalised and turned into -sum p*log(p) together
// (lane = bin for the element-wise part, one lane per histogram for the two sums).  __syncwarp only.
// fp32 throughout, denormals kept (eps = 1e-40 is a denormal; flush-to-zero would turn every entropy into
// NaN).  Float-tolerance parity (rtol 2e-5 in the tests): the summation order is ours, and
// exp(-0.5 ((v - bin) / 0.01)^2) is evaluated as ex2.approx(-((v - bin) * c)^2), c = 100 sqrt(log2(e) / 2).
#include "common.cuh"
#ifdef CGIC_TRACE
// phases of the routing tail, one row per image: 0 tail starts (the image's last entropy CTA), 1 routing starts, 2 coarse
// threshold, 3 coarse mask written, 4 medium threshold, 5 done   (profiles/trace_phases.py --route)
static __device__ unsigned long long g_trace_route[1024 * 8];
extern "C" __attribute__((visibility("default"))) int cgic_trace_route(unsigned long long *host)
{
    return (int)cudaMemcpyFromSymbol(host, g_trace_route, sizeof(g_trace_route));
}
#define RS_STAMP(k)                                                       \
    do {                                                                  \
        if (threadIdx.x == 0 && blockIdx.y < 1024) {                      \
            unsigned long long t__;                                       \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t__));       \
            g_trace_route[blockIdx.y * 8 + (k)] = t__;                    \
        }                                                                 \
    } while (0)
#endif
#include "router_select.cuh"

namespace cgic {
namespace {

struct Bins {
    float v[32];
};

constexpr int EN_WIN = 3;
#ifndef CGIC_EN_WARPS
#define CGIC_EN_WARPS 8  // 8: the routing tail of the fused kernel runs on 256 threads (4 warps: 62 us, 8: 54 us, 16: 54 us for entropy + routing of 64 images; the entropy maps alone take 37 us with 4 or 8, 41 us with 16)
#endif
constexpr int EN_WARPS = CGIC_EN_WARPS;  // warps (= regions in flight) per CTA
constexpr int EN_STRIDE = 33;            // words per histogram row: lanes hitting the same bin fall into different banks

__device__ __forceinline__ float ex2_approx(float x)  // no flush-to-zero: denormal results are kept
{
    float y;
    asm("ex2.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Routing request of the fused kernel (f1, first half): when m_c is set, the LAST CTA of an image to finish its entropy maps
// also runs TripleGrainFixedEntropyRouter.forward for that image (per-image thresholds; route_image, router_select.cuh).
struct RouteReq {
    int32_t *m_c, *m_m;   // [B, n16], [B, 4 n16]; m_c == nullptr: entropy maps only
    int32_t *near;        // [B, 2] entropies within tolerance of the coarse / medium threshold (nullable)
    int32_t *tickets;     // [B] zero before the launch, left zero
    int mode;
    int64_t k_c, k_m;
    float rtol, atol;
};

// grid (CTAs per image, B): the regions of an image are dealt to its own CTAs, so that an image's last CTA is well defined
__global__ void __launch_bounds__(EN_WARPS * 32, 32 / EN_WARPS)
entropy_kernel(const float *__restrict__ x, int H, int W, int regions_x, int regions_img, const Bins bins, float *__restrict__ e8,
               float *__restrict__ e16, const RouteReq rq)
{
    extern __shared__ __align__(16) float s_rows_dyn[];                   // per warp: 32 histogram rows (slice a multiple of 16 bytes)
    float(*s_rows)[32 * EN_STRIDE + 4] = reinterpret_cast<float(*)[32 * EN_STRIDE + 4]>(s_rows_dyn);
    constexpr int S_ROWS_WORDS = EN_WARPS * (32 * EN_STRIDE + 4);
    __shared__ float s_bins[32];
    __shared__ uint32_t s_state[RS_STATE];
    __shared__ int s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    pdl_launch_dependents();
    if (threadIdx.x < 32) s_bins[threadIdx.x] = bins.v[threadIdx.x];
    __syncthreads();
    pdl_wait();  // the image may come straight from a preceding kernel
    const int b = blockIdx.y;
    const int rr = (int)blockIdx.x * EN_WARPS + warp;
    if (rr < regions_img) {
    const int ry = rr / regions_x, rx = rr - ry * regions_x;
    const int gx = rx * 32 + lane;
    const bool col_ok = gx < W;
    const int64_t plane = (int64_t)H * W;
    const float *xb = x + (int64_t)b * 3 * plane + (int64_t)(ry * 16) * W + gx;
    const float *xc[3] = {xb, xb + plane, xb + 2 * plane};  // one base pointer per channel: the 48 loads below then need 32-bit offsets only
    float *rows = s_rows[warp];
    float *row = rows + lane * EN_STRIDE;
    // the two halves (upper / lower 8 rows) go through the same 32 histogram rows one after the other; the
    // lower half's pixels are already in flight while the upper half is accumulated
    float raw[2][8][3];
#pragma unroll
    for (int half = 0; half < 2; ++half)
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int o = (half * 8 + r) * W;
#pragma unroll
            for (int c = 0; c < 3; ++c) raw[half][r][c] = col_ok ? __ldg(xc[c] + o) : 0.f;
        }
    float acc[2][4];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        // zero the 32 rows (16-byte stores over the whole slice)
        for (int i = lane; i < (32 * EN_STRIDE + 4) / 4; i += 32) reinterpret_cast<float4 *>(rows)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncwarp();
        if (col_ok) {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                // 0.2989*R + 0.5870*G + 0.1140*B, each product and sum rounded (model.py:471)
                const float g = __fadd_rn(__fadd_rn(__fmul_rn(0.2989f, raw[half][r][0]), __fmul_rn(0.5870f, raw[half][r][1])),
                                          __fmul_rn(0.1140f, raw[half][r][2]));
                int jc = __float2int_rn((g + 1.0f) * 15.5f);
                jc = max(0, min(31, jc));
                if (jc >= EN_WIN / 2 && jc < 32 - EN_WIN / 2) {  // the whole window is inside the histogram (always, for images in [0, 1])
#pragma unroll
                    for (int k = 0; k < EN_WIN; ++k) {
                        const int j = jc - EN_WIN / 2 + k;
                        const float q = __fmul_rn(__fsub_rn(g, s_bins[j]), 84.93218002880191f);  // 100 * sqrt(log2(e) / 2)
                        row[j] += ex2_approx(-__fmul_rn(q, q));  // same thread owns the row: plain read-modify-write
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < EN_WIN; ++k) {
                        const int j = jc - EN_WIN / 2 + k;
                        if (j >= 0 && j < 32) {
                            const float q = __fmul_rn(__fsub_rn(g, s_bins[j]), 84.93218002880191f);
                            row[j] += ex2_approx(-__fmul_rn(q, q));
                        }
                    }
                }
            }
        }
        __syncwarp();
        // the half's four 8x8 patches (pq = column group of 8): lane = bin, fixed order over the 8 columns
#pragma unroll
        for (int pq = 0; pq < 4; ++pq) {
            float a = 0.f;
#pragma unroll
            for (int c = 0; c < 8; ++c) a += rows[(pq * 8 + c) * EN_STRIDE + lane];
            acc[half][pq] = a;
        }
        __syncwarp();
    }
    const int valid_pq = min(4, (W - rx * 32) / 8);  // a region may hang over the right edge by one 16-pixel block
    // ten histograms per region: 0..7 = the 8x8 patches (py * 4 + pq), 8..9 = the 16x16 blocks; pdf = sum / patch size
    float pdf[10];
#pragma unroll
    for (int py = 0; py < 2; ++py)
#pragma unroll
        for (int pq = 0; pq < 4; ++pq) pdf[py * 4 + pq] = acc[py][pq] / 64.0f;
#pragma unroll
    for (int kx = 0; kx < 2; ++kx)
        pdf[8 + kx] = ((acc[0][2 * kx] + acc[0][2 * kx + 1]) + (acc[1][2 * kx] + acc[1][2 * kx + 1])) / 256.0f;
    // the histogram rows are dead: reuse the slice as T[10][33] (+ norms at [10 * 33 ..])
    float *T = rows;
#pragma unroll
    for (int p = 0; p < 10; ++p) T[p * EN_STRIDE + lane] = pdf[p];
    __syncwarp();
    const float eps = 1e-40f;
    if (lane < 10) {
        float sum = 0.f;
        for (int j = 0; j < 32; ++j) sum += T[lane * EN_STRIDE + j];
        T[10 * EN_STRIDE + lane] = sum + eps;
    }
    __syncwarp();
#pragma unroll
    for (int p = 0; p < 10; ++p) {
        const float q = pdf[p] / T[10 * EN_STRIDE + p] + eps;
        T[p * EN_STRIDE + lane] = q * logf(q);
    }
    __syncwarp();
    if (lane < 10) {
        float sum = 0.f;
        for (int j = 0; j < 32; ++j) sum += T[lane * EN_STRIDE + j];
        const float ent = -sum;
        if (lane < 8) {
            const int py = lane >> 2, pq = lane & 3;
            if (e8 && pq < valid_pq) e8[((int64_t)b * (H / 8) + ry * 2 + py) * (W / 8) + rx * 4 + pq] = ent;
        } else {
            const int kx = lane - 8;
            if (e16 && 2 * kx < valid_pq) e16[((int64_t)b * (H / 16) + ry) * (W / 16) + rx * 2 + kx] = ent;
        }
    }
    }  // rr < regions_img
    if (!rq.m_c) return;
    // ---- routing by the last CTA of the image (all its entropies are then in global memory)
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = atomicAdd(&rq.tickets[b], 1) == (int)gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    RS_STAMP(0);
    __threadfence();
    const int h16 = H / 16, w16 = W / 16, n16 = h16 * w16;
    // the histogram rows are dead: the select's histogram (8 KB) and, when it fits behind it, the key cache take their place
    uint32_t *s_hist = reinterpret_cast<uint32_t *>(&s_rows[0][0]);
    uint32_t *s_keys = RS_BINS + 4 * n16 <= S_ROWS_WORDS ? s_hist + RS_BINS : nullptr;
    route_image(e16 + (int64_t)b * n16, e8 + (int64_t)b * 4 * n16, h16, w16, rq.mode, rq.k_c, rq.k_m, rq.m_c + (int64_t)b * n16,
                rq.m_m + (int64_t)b * 4 * n16, rq.near ? rq.near + 2 * b : nullptr, rq.rtol, rq.atol, s_hist, s_state, s_keys);
    RS_STAMP(5);
    if (threadIdx.x == 0) rq.tickets[b] = 0;  // (workspace contract: zero between launches)
}

}  // namespace
}  // namespace cgic

using namespace cgic;

static int entropy_launch(const char *who, const float *x, int B, int H, int W, const float *bins32_host, float *e8_out, float *e16_out,
                          const RouteReq &rq, cudaStream_t stream)
{
    CGIC_REQUIRE(B >= 0 && H > 0 && W > 0 && H % 16 == 0 && W % 16 == 0, CGIC_EINVAL, "%s: image %dx%d must be multiples of 16", who, H, W);
    if (B == 0) return CGIC_OK;
    const int regions_x = (W + 31) / 32;
    const int64_t regions_img = (int64_t)(H / 16) * regions_x;
    CGIC_REQUIRE(regions_img < ((int64_t)1 << 30) && B <= 65535, CGIC_EINVAL, "%s: grid too large (B <= 65535)", who);
    Bins bins;
    for (int i = 0; i < 32; ++i) bins.v[i] = bins32_host[i];
    {
        CGIC_PROF("entropy_kernel", stream);
        const size_t smem = (size_t)EN_WARPS * (32 * EN_STRIDE + 4) * sizeof(float);
        const int rc = ensure_smem((const void *)entropy_kernel, smem);
        if (rc) return rc;
        CGIC_CUDA_CHECK(launch_pdl(entropy_kernel, dim3((unsigned)((regions_img + EN_WARPS - 1) / EN_WARPS), (unsigned)B), dim3(EN_WARPS * 32), smem, stream, x,
                                   H, W, regions_x, (int)regions_img, bins, e8_out, e16_out, rq));
    }
    CGIC_LAUNCH_CHECK();
    return CGIC_OK;
}

extern "C" int cgic_entropy_maps(const float *x, int B, int H, int W, const float *bins32_host, float *e8_out, float *e16_out,
                                 cgic_stream_t stream)
{
    CGIC_REQUIRE(x && bins32_host && (e8_out || e16_out), CGIC_EINVAL, "cgic_entropy_maps: null argument");
    if (B > 65535) {  // (grid y limit) -- slices of the batch
        for (int b0 = 0; b0 < B; b0 += 65535) {
            const int nb = B - b0 < 65535 ? B - b0 : 65535;
            const int rc = cgic_entropy_maps(x + (int64_t)b0 * 3 * H * W, nb, H, W, bins32_host, e8_out ? e8_out + (int64_t)b0 * (H / 8) * (W / 8) : nullptr,
                                             e16_out ? e16_out + (int64_t)b0 * (H / 16) * (W / 16) : nullptr, stream);
            if (rc) return rc;
        }
        return CGIC_OK;
    }
    return entropy_launch("cgic_entropy_maps", x, B, H, W, bins32_host, e8_out, e16_out, RouteReq{}, as_stream(stream));
}

extern "C" size_t cgic_entropy_route_workspace_bytes(int B) { return ((size_t)(B > 0 ? B : 0) * 4 + 255) / 256 * 256 + 256; }

extern "C" int cgic_entropy_route(const float *x, int B, int H, int W, const float *bins32_host, float *e8_out, float *e16_out, int mode, int64_t k_c,
                                  int64_t k_m, float rtol, float atol, int32_t *m_c, int32_t *m_m, int32_t *near_out, void *workspace,
                                  size_t workspace_bytes, cgic_stream_t stream)
{
    CGIC_REQUIRE(x && bins32_host && e8_out && e16_out && m_c && m_m && workspace, CGIC_EINVAL, "cgic_entropy_route: null argument");
    CGIC_REQUIRE(mode >= 0 && mode <= 6 && k_c >= 0 && k_m >= 0, CGIC_EINVAL, "cgic_entropy_route: bad mode / ranks");
    CGIC_REQUIRE(workspace_bytes >= cgic_entropy_route_workspace_bytes(B), CGIC_ESPACE, "cgic_entropy_route: workspace %zu < %zu bytes", workspace_bytes,
                 cgic_entropy_route_workspace_bytes(B));
    if (near_out && B > 0) CGIC_CUDA_CHECK(cudaMemsetAsync(near_out, 0, (size_t)B * 2 * sizeof(int32_t), as_stream(stream)));
    RouteReq rq{m_c, m_m, near_out, static_cast<int32_t *>(workspace), mode, k_c, k_m, rtol, atol};
    return entropy_launch("cgic_entropy_route", x, B, H, W, bins32_host, e8_out, e16_out, rq, as_stream(stream));
}
