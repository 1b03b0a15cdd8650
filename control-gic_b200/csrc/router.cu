// a5  TripleGrainFixedEntropyRouter.forward  (CGIC/modules/vqvae/RouterTriple.py:15-96)
// a6  mask-mix tail of Encoder.forward        (CGIC/modules/vqvae/vqvae_blocks.py:361-366)
//
// The reference sorts the flattened entropy map and thresholds on `<` against the k-th smallest
// value.  Only that one order statistic is needed, so the kernel runs an exact 4-pass radix
// select (8 bits per pass, shared-memory histogram) over order-preserving integer keys: the
// threshold VALUE is identical to sorted[k-1], hence so are the masks.  The ranks k_c / k_m and
// the mode come from the host (Python round() on doubles, banker's rounding).
//   per_image = 1 : one CTA per image, thresholds per image (B independent B == 1 calls)
//   per_image = 0 : one CTA, thresholds over the whole batch (reference behaviour for B > 1)
#include "router_select.cuh"

namespace cgic {
namespace {

constexpr int RT_THREADS = 512;

// grid = B (per_image) or 1; writes m_c and m_m.
__global__ void __launch_bounds__(RT_THREADS)
router_select_kernel(const float *__restrict__ e16, const float *__restrict__ e8, int B, int h16, int w16, int mode,
                     int64_t k_c, int64_t k_m, int per_image, int32_t *__restrict__ m_c, int32_t *__restrict__ m_m,
                     int32_t *__restrict__ m_f, float *__restrict__ gate, int key_cache)
{
    extern __shared__ uint32_t s_keys_dyn[];  // key cache (n8 keys) when the launch provided it
    __shared__ uint32_t s_hist[RS_BINS];
    __shared__ uint32_t s_state[RS_STATE];
    const int h8 = 2 * h16, w8 = 2 * w16;
    const int64_t n16_img = (int64_t)h16 * w16, n8_img = (int64_t)h8 * w8;
    const int nb = per_image ? 1 : B;
    const int64_t img0 = per_image ? blockIdx.x : 0;
    const float *a16 = e16 + img0 * n16_img;
    const float *a8 = e8 + img0 * n8_img;
    int32_t *c = m_c + img0 * n16_img;
    int32_t *m = m_m + img0 * n8_img;
    const int64_t n16 = nb * n16_img, n8 = nb * n8_img;
    const bool use_c = mode == 0 || mode == 2 || mode == 3;
    uint32_t *s_keys = key_cache ? s_keys_dyn : nullptr;
    pdl_launch_dependents();
    pdl_wait();  // the entropy maps usually come straight from entropy_kernel

    if (use_c) {
        const float thr = select_rank([&](int64_t i) { return a16[i]; }, n16, k_c != 0 ? k_c - 1 : 0, s_hist, s_state, s_keys);
        for (int64_t i = threadIdx.x; i < n16; i += blockDim.x) c[i] = a16[i] < thr;
    } else {
        for (int64_t i = threadIdx.x; i < n16; i += blockDim.x) c[i] = (mode == 4);
    }
    __syncthreads();  // c[] (global) is re-read below by other threads of this CTA
    auto parent = [&](int64_t i) -> int64_t {  // coarse cell above medium cell i
        const int64_t img = i / n8_img, p = i - img * n8_img;
        const int y = (int)(p / w8), x = (int)(p - (int64_t)y * w8);
        return img * n16_img + (int64_t)(y >> 1) * w16 + (x >> 1);
    };
    if (mode == 0) {
        // entropy of cells under a coarse patch is zeroed before the sort (RouterTriple.py:27)
        const float thr = select_rank(
            [&](int64_t i) { return __fmul_rn(a8[i], __fsub_rn(1.0f, (float)c[parent(i)])); }, n8,
            k_m != 0 ? k_m - 1 : 0, s_hist, s_state, s_keys);
        for (int64_t i = threadIdx.x; i < n8; i += blockDim.x) m[i] = (a8[i] < thr) && !c[parent(i)];
    } else if (mode == 1) {
        const float thr = select_rank([&](int64_t i) { return a8[i]; }, n8, k_m != 0 ? k_m - 1 : 0, s_hist, s_state, s_keys);
        for (int64_t i = threadIdx.x; i < n8; i += blockDim.x) m[i] = a8[i] < thr;
    } else if (mode == 3) {
        for (int64_t i = threadIdx.x; i < n8; i += blockDim.x) m[i] = 1 - c[parent(i)];
    } else {
        for (int64_t i = threadIdx.x; i < n8; i += blockDim.x) m[i] = (mode == 5);
    }
    if (!m_f) return;
    // per-image launches also write the image's fine mask and gate (saves the second launch); 4 tokens per thread
    __syncthreads();  // c[] and m[] (global) are re-read below by other threads of this CTA
    const int h = 4 * h16, w = 4 * w16;
    const int64_t plane = (int64_t)h * w;
    int32_t *f = m_f + img0 * plane;
    for (int64_t q = threadIdx.x; q < plane / 4; q += blockDim.x) {
        const int y = (int)(q / w16), xq = (int)(q - (int64_t)y * w16);
        const int cc = c[(int64_t)(y >> 2) * w16 + xq];
        const int2 mm = *reinterpret_cast<const int2 *>(&m[(int64_t)(y >> 1) * w8 + 2 * xq]);
        int4 fv;
        if (mode <= 2) fv = make_int4((1 - cc - mm.x) != 0, (1 - cc - mm.x) != 0, (1 - cc - mm.y) != 0, (1 - cc - mm.y) != 0);
        else fv = make_int4(mode == 6, mode == 6, mode == 6, mode == 6);
        *reinterpret_cast<int4 *>(&f[(int64_t)y * w + 4 * xq]) = fv;
        if (gate) {
            float *row = gate + (img0 * h + y) * (int64_t)(3 * w);
            const float fc = (float)cc;
            *reinterpret_cast<float4 *>(&row[4 * xq]) = make_float4(fc, fc, fc, fc);
            *reinterpret_cast<float4 *>(&row[w + 4 * xq]) = make_float4((float)mm.x, (float)mm.x, (float)mm.y, (float)mm.y);
            *reinterpret_cast<float4 *>(&row[2 * w + 4 * xq]) =
                mode <= 2 ? make_float4((float)(1 - cc - mm.x), (float)(1 - cc - mm.x), (float)(1 - cc - mm.y), (float)(1 - cc - mm.y))
                          : make_float4((float)fv.x, (float)fv.y, (float)fv.z, (float)fv.w);
        }
    }
}

// one thread per fine token: m_f and the optional gate tensor [B,1,h,3w]
__global__ void router_fine_kernel(const int32_t *__restrict__ m_c, const int32_t *__restrict__ m_m, int64_t n_tokens, int h,
                                   int w, int mode, int32_t *__restrict__ m_f, float *__restrict__ gate)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    pdl_launch_dependents();
    pdl_wait();
    if (t >= n_tokens) return;
    const int64_t plane = (int64_t)h * w;
    const int64_t b = t / plane, p = t - b * plane;
    const int y = (int)(p / w), x = (int)(p - (int64_t)y * w);
    const int c = m_c[b * (plane / 16) + (int64_t)(y >> 2) * (w / 4) + (x >> 2)];
    const int m = m_m[b * (plane / 4) + (int64_t)(y >> 1) * (w / 2) + (x >> 1)];
    int f;
    if (mode <= 2) f = (1 - c - m) != 0;
    else f = (mode == 6);
    m_f[t] = f;
    if (gate) {
        float *row = gate + (b * h + y) * (int64_t)(3 * w);
        row[x] = (float)c;
        row[w + x] = (float)m;
        row[2 * w + x] = (mode <= 2) ? (float)(1 - c - m) : (float)f;
    }
}

// one thread per 4 consecutive fine tokens of a row (they share one coarse cell and two medium cells): 16-byte loads / stores
__global__ void __launch_bounds__(256)
mask_mix_kernel(const float *__restrict__ h_c, const float *__restrict__ h_m, const float *__restrict__ h_f, const int32_t *__restrict__ m_c,
                const int32_t *__restrict__ m_m, const int32_t *__restrict__ m_f, int64_t n_quads, int C, int h, int w,
                float *__restrict__ out)
{
    const int64_t qi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    pdl_launch_dependents();
    pdl_wait();  // masks (and usually the heads) come from preceding kernels
    if (qi >= n_quads) return;
    const unsigned wq = (unsigned)w / 4u;
    const int64_t rowi = qi / wq;              // (b*C + c) * h + y
    const int xq = (int)(qi - rowi * wq);      // = x / 4
    const int64_t bc = rowi / h;
    const int y = (int)(rowi - bc * h);
    const int64_t b = bc / C;
    const int h8 = h / 2, w8 = w / 2, h16 = h / 4, w16 = w / 4;
    const float hc = __ldg(&h_c[(bc * h16 + (y >> 2)) * w16 + xq]);
    const float mc = (float)__ldg(&m_c[(b * h16 + (y >> 2)) * w16 + xq]);
    const float2 hm = __ldg(reinterpret_cast<const float2 *>(&h_m[(bc * h8 + (y >> 1)) * w8 + 2 * xq]));
    const int2 mm = __ldg(reinterpret_cast<const int2 *>(&m_m[(b * h8 + (y >> 1)) * w8 + 2 * xq]));
    const float4 hf = __ldg(reinterpret_cast<const float4 *>(&h_f[rowi * w + 4 * xq]));
    const int4 mf = __ldg(reinterpret_cast<const int4 *>(&m_f[(b * h + y) * (int64_t)w + 4 * xq]));
    const float a = __fmul_rn(hc, mc);
    const float m0 = __fmul_rn(hm.x, (float)mm.x), m1 = __fmul_rn(hm.y, (float)mm.y);
    float4 o;
    o.x = __fadd_rn(__fadd_rn(a, m0), __fmul_rn(hf.x, (float)mf.x));
    o.y = __fadd_rn(__fadd_rn(a, m0), __fmul_rn(hf.y, (float)mf.y));
    o.z = __fadd_rn(__fadd_rn(a, m1), __fmul_rn(hf.z, (float)mf.z));
    o.w = __fadd_rn(__fadd_rn(a, m1), __fmul_rn(hf.w, (float)mf.w));
    *reinterpret_cast<float4 *>(&out[rowi * w + 4 * xq]) = o;
}

// f1, second half: fine mask (+ gate) + mask-mix in one launch.  One thread per 4 consecutive fine tokens of a row and
// channel; the fine mask follows from the coarse / medium cells (RouterTriple.py:34), so it is computed, not loaded; the
// threads of channel 0 also write m_f and the gate tensor.
__global__ void __launch_bounds__(256)
route_mix_kernel(const float *__restrict__ h_c, const float *__restrict__ h_m, const float *__restrict__ h_f, const int32_t *__restrict__ m_c,
                 const int32_t *__restrict__ m_m, int mode, int64_t n_quads, int C, int h, int w, int32_t *__restrict__ m_f_out,
                 float *__restrict__ gate, float *__restrict__ out)
{
    const int64_t qi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    pdl_launch_dependents();
    pdl_wait();
    if (qi >= n_quads) return;
    const unsigned wq = (unsigned)w / 4u;
    const int64_t rowi = qi / wq;              // (b*C + c) * h + y
    const int xq = (int)(qi - rowi * wq);      // = x / 4
    const int64_t bc = rowi / h;
    const int y = (int)(rowi - bc * h);
    const int64_t b = bc / C;
    const int ch = (int)(bc - b * C);
    const int h8 = h / 2, w8 = w / 2, h16 = h / 4, w16 = w / 4;
    const float hc = __ldg(&h_c[(bc * h16 + (y >> 2)) * w16 + xq]);
    const int cc = __ldg(&m_c[(b * h16 + (y >> 2)) * w16 + xq]);
    const float2 hm = __ldg(reinterpret_cast<const float2 *>(&h_m[(bc * h8 + (y >> 1)) * w8 + 2 * xq]));
    const int2 mm = __ldg(reinterpret_cast<const int2 *>(&m_m[(b * h8 + (y >> 1)) * w8 + 2 * xq]));
    const float4 hf = __ldg(reinterpret_cast<const float4 *>(&h_f[rowi * w + 4 * xq]));
    int4 mf;
    float4 gf;  // the gate's fine third: 1 - up4(c) - up2(m) as a FLOAT in modes 0..2 (RouterTriple.py:34), the mask itself otherwise
    if (mode <= 2) {
        const int f0 = 1 - cc - mm.x, f1 = 1 - cc - mm.y;
        mf = make_int4(f0 != 0, f0 != 0, f1 != 0, f1 != 0);
        gf = make_float4((float)f0, (float)f0, (float)f1, (float)f1);
    } else {
        const int f = mode == 6;
        mf = make_int4(f, f, f, f);
        gf = make_float4((float)f, (float)f, (float)f, (float)f);
    }
    const float a = __fmul_rn(hc, (float)cc);
    const float m0 = __fmul_rn(hm.x, (float)mm.x), m1 = __fmul_rn(hm.y, (float)mm.y);
    float4 o;
    o.x = __fadd_rn(__fadd_rn(a, m0), __fmul_rn(hf.x, (float)mf.x));
    o.y = __fadd_rn(__fadd_rn(a, m0), __fmul_rn(hf.y, (float)mf.y));
    o.z = __fadd_rn(__fadd_rn(a, m1), __fmul_rn(hf.z, (float)mf.z));
    o.w = __fadd_rn(__fadd_rn(a, m1), __fmul_rn(hf.w, (float)mf.w));
    *reinterpret_cast<float4 *>(&out[rowi * w + 4 * xq]) = o;
    if (ch == 0) {
        *reinterpret_cast<int4 *>(&m_f_out[(b * h + y) * (int64_t)w + 4 * xq]) = mf;
        if (gate) {
            float *row = gate + (b * h + y) * (int64_t)(3 * w);
            const float fc = (float)cc;
            *reinterpret_cast<float4 *>(&row[4 * xq]) = make_float4(fc, fc, fc, fc);
            *reinterpret_cast<float4 *>(&row[w + 4 * xq]) = make_float4((float)mm.x, (float)mm.x, (float)mm.y, (float)mm.y);
            *reinterpret_cast<float4 *>(&row[2 * w + 4 * xq]) = gf;
        }
    }
}

// f4  decoder entry (CGIC/modules/vqvae/decoder.py:373-382): mask-gated merge of the decoder's branches, two elements per thread.
//   LEVEL 2: out = h * up2(m_c) + other * m_m                       tensors [B,C,hh,ww], m_c [B,1,hh/2,ww/2], m_m [B,1,hh,ww]
//   LEVEL 3: out = (h * up4(m_c) + h * up2(m_m)) + other * m_f      tensors [B,C,hh,ww], m_c [B,1,hh/4,ww/4], m_m [B,1,hh/2,ww/2], m_f [B,1,hh,ww]
// Products and sums rounded one by one in torch's order; masks int32, int64 or float32 (converted like `.float()`).
template <typename M, int LEVEL>
__global__ void __launch_bounds__(256)
decoder_merge_kernel(const float *__restrict__ h, const float *__restrict__ other, const M *__restrict__ m_c, const M *__restrict__ m_m,
                     const M *__restrict__ m_f, int64_t n_pairs, int C, int hh, int ww, float *__restrict__ out)
{
    const int64_t pi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    pdl_launch_dependents();
    pdl_wait();
    if (pi >= n_pairs) return;
    const unsigned wp = (unsigned)ww / 2u;
    const int64_t rowi = pi / wp;              // (b*C + c) * hh + y
    const int xp = (int)(pi - rowi * wp);      // x / 2
    const int64_t bc = rowi / hh;
    const int y = (int)(rowi - bc * hh);
    const int64_t b = bc / C;
    const float2 hv = __ldg(reinterpret_cast<const float2 *>(&h[rowi * ww + 2 * xp]));
    const float2 ov = __ldg(reinterpret_cast<const float2 *>(&other[rowi * ww + 2 * xp]));
    float2 o;
    if (LEVEL == 2) {
        const float mc = (float)m_c[(b * (hh / 2) + (y >> 1)) * (ww / 2) + xp];
        const float m0 = (float)m_m[(b * hh + y) * (int64_t)ww + 2 * xp], m1 = (float)m_m[(b * hh + y) * (int64_t)ww + 2 * xp + 1];
        o.x = __fadd_rn(__fmul_rn(hv.x, mc), __fmul_rn(ov.x, m0));
        o.y = __fadd_rn(__fmul_rn(hv.y, mc), __fmul_rn(ov.y, m1));
    } else {
        const float mc = (float)m_c[(b * (hh / 4) + (y >> 2)) * (ww / 4) + (xp >> 1)];
        const float mm = (float)m_m[(b * (hh / 2) + (y >> 1)) * (ww / 2) + xp];
        const float f0 = (float)m_f[(b * hh + y) * (int64_t)ww + 2 * xp], f1 = (float)m_f[(b * hh + y) * (int64_t)ww + 2 * xp + 1];
        o.x = __fadd_rn(__fadd_rn(__fmul_rn(hv.x, mc), __fmul_rn(hv.x, mm)), __fmul_rn(ov.x, f0));
        o.y = __fadd_rn(__fadd_rn(__fmul_rn(hv.y, mc), __fmul_rn(hv.y, mm)), __fmul_rn(ov.y, f1));
    }
    *reinterpret_cast<float2 *>(&out[rowi * ww + 2 * xp]) = o;
}

}  // namespace
}  // namespace cgic

using namespace cgic;

template <typename M>
static int decoder_merge_launch(const float *h, const float *other, const void *m_c, const void *m_m, const void *m_f, int level, int64_t n_pairs,
                                int C, int hh, int ww, float *out, cudaStream_t stream)
{
    const M *c = static_cast<const M *>(m_c), *m = static_cast<const M *>(m_m), *f = static_cast<const M *>(m_f);
    const dim3 grid((unsigned)((n_pairs + 255) / 256)), block(256);
    CGIC_PROF("decoder_merge_kernel", stream);
    if (level == 2) CGIC_CUDA_CHECK(launch_pdl(decoder_merge_kernel<M, 2>, grid, block, 0, stream, h, other, c, m, f, n_pairs, C, hh, ww, out));
    else CGIC_CUDA_CHECK(launch_pdl(decoder_merge_kernel<M, 3>, grid, block, 0, stream, h, other, c, m, f, n_pairs, C, hh, ww, out));
    return CGIC_OK;
}

extern "C" int cgic_decoder_merge(const float *h, const float *other, const void *m_c, const void *m_m, const void *m_f, int mask_elem,
                                  int level, int B, int C, int hh, int ww, float *out, cgic_stream_t stream_)
{
    CGIC_REQUIRE(h && other && m_c && m_m && out && (level == 2 || (level == 3 && m_f)), CGIC_EINVAL, "cgic_decoder_merge: null argument or level %d", level);
    const int div = level == 2 ? 2 : 4;
    CGIC_REQUIRE(B >= 0 && C > 0 && hh > 0 && ww > 0 && hh % div == 0 && ww % div == 0, CGIC_EINVAL,
                 "cgic_decoder_merge: %dx%d must be multiples of %d at level %d", hh, ww, div, level);
    CGIC_REQUIRE(mask_elem == 4 || mask_elem == 8 || mask_elem == -4, CGIC_EINVAL, "cgic_decoder_merge: mask_elem must be 4, 8 or -4 (float)");
    for (const void *ptr : {(const void *)h, (const void *)other, (const void *)out})
        CGIC_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 7) == 0, CGIC_EINVAL, "cgic_decoder_merge: tensors must be 8-byte aligned");
    const int64_t n_pairs = (int64_t)B * C * hh * ww / 2;
    if (n_pairs == 0) return CGIC_OK;
    cudaStream_t stream = as_stream(stream_);
    int rc;
    if (mask_elem == 4) rc = decoder_merge_launch<int32_t>(h, other, m_c, m_m, m_f, level, n_pairs, C, hh, ww, out, stream);
    else if (mask_elem == 8) rc = decoder_merge_launch<int64_t>(h, other, m_c, m_m, m_f, level, n_pairs, C, hh, ww, out, stream);
    else rc = decoder_merge_launch<float>(h, other, m_c, m_m, m_f, level, n_pairs, C, hh, ww, out, stream);
    if (rc) return rc;
    CGIC_LAUNCH_CHECK();
    return CGIC_OK;
}

extern "C" size_t cgic_router_workspace_bytes(int, int, int) { return 256; }

extern "C" int cgic_router(const float *e16, const float *e8, int B, int h16, int w16, int mode, int64_t k_c, int64_t k_m,
                           int per_image, int32_t *m_c, int32_t *m_m, int32_t *m_f, float *gate_out, void *, size_t,
                           cgic_stream_t stream_)
{
    CGIC_REQUIRE(e16 && e8 && m_c && m_m && m_f, CGIC_EINVAL, "cgic_router: null argument");
    CGIC_REQUIRE(B >= 0 && h16 > 0 && w16 > 0 && mode >= 0 && mode <= 6 && k_c >= 0 && k_m >= 0, CGIC_EINVAL,
                 "cgic_router: bad argument B=%d h16=%d w16=%d mode=%d", B, h16, w16, mode);
    if (B == 0) return CGIC_OK;
    cudaStream_t stream = as_stream(stream_);
    // per-image thresholds: the select kernel (one CTA per image) also writes the fine mask / gate when the rows are 16-byte aligned
    const bool fuse_fine = per_image && (reinterpret_cast<uintptr_t>(m_m) & 7) == 0 && (reinterpret_cast<uintptr_t>(m_f) & 15) == 0 &&
                           (!gate_out || (reinterpret_cast<uintptr_t>(gate_out) & 15) == 0);
    {
        CGIC_PROF("router_select_kernel", stream);
        const int64_t n8_cta = (int64_t)(per_image ? 1 : B) * 4 * h16 * w16;
        // key cache: static (histogram, state: ~8.5 KB) + dynamic shared memory must stay within the 48 KB a kernel gets without
        // opting in, so 9 728 keys (38 KB) at most -- larger maps are re-fetched from global memory in every pass
        const int key_cache = n8_cta <= 9728;
        CGIC_CUDA_CHECK(launch_pdl(router_select_kernel, dim3(per_image ? B : 1), dim3(RT_THREADS), key_cache ? (size_t)n8_cta * 4 : 0, stream, e16, e8, B,
                                   h16, w16, mode, k_c, k_m, per_image, m_c, m_m, fuse_fine ? m_f : (int32_t *)nullptr, gate_out, key_cache));
    }
    CGIC_LAUNCH_CHECK();
    if (fuse_fine) return CGIC_OK;
    const int64_t n = (int64_t)B * 16 * h16 * w16;
    {
        CGIC_PROF("router_fine_kernel", stream);
        CGIC_CUDA_CHECK(launch_pdl(router_fine_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, stream, (const int32_t *)m_c, (const int32_t *)m_m, n,
                                   4 * h16, 4 * w16, mode, m_f, gate_out));
    }
    CGIC_LAUNCH_CHECK();
    return CGIC_OK;
}

extern "C" int cgic_mask_mix(const float *h_c, const float *h_m, const float *h_f, const int32_t *m_c, const int32_t *m_m,
                             const int32_t *m_f, int B, int C, int h, int w, float *out, cgic_stream_t stream)
{
    CGIC_REQUIRE(h_c && h_m && h_f && m_c && m_m && m_f && out, CGIC_EINVAL, "cgic_mask_mix: null argument");
    CGIC_REQUIRE(B >= 0 && C > 0 && h > 0 && w > 0 && h % 4 == 0 && w % 4 == 0, CGIC_EINVAL, "cgic_mask_mix: bad shape");
    const int64_t n = (int64_t)B * C * h * w / 4;
    if (n == 0) return CGIC_OK;
    for (const void *ptr : {(const void *)h_m, (const void *)h_f, (const void *)m_m, (const void *)m_f, (const void *)out})
        CGIC_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, CGIC_EINVAL, "cgic_mask_mix: buffers must be 16-byte aligned");
    {
        CGIC_PROF("mask_mix_kernel", as_stream(stream));
        CGIC_CUDA_CHECK(launch_pdl(mask_mix_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, as_stream(stream), h_c, h_m, h_f, m_c, m_m, m_f, n, C, h,
                                   w, out));
    }
    CGIC_LAUNCH_CHECK();
    return CGIC_OK;
}

extern "C" int cgic_route_mix(const float *h_c, const float *h_m, const float *h_f, const int32_t *m_c, const int32_t *m_m, int mode, int B, int C,
                              int h, int w, int32_t *m_f_out, float *gate_out, float *out, cgic_stream_t stream)
{
    CGIC_REQUIRE(h_c && h_m && h_f && m_c && m_m && m_f_out && out, CGIC_EINVAL, "cgic_route_mix: null argument");
    CGIC_REQUIRE(B >= 0 && C > 0 && h > 0 && w > 0 && h % 4 == 0 && w % 4 == 0 && mode >= 0 && mode <= 6, CGIC_EINVAL, "cgic_route_mix: bad shape / mode");
    const int64_t n = (int64_t)B * C * h * w / 4;
    if (n == 0) return CGIC_OK;
    for (const void *ptr : {(const void *)h_m, (const void *)h_f, (const void *)m_m, (const void *)m_f_out, (const void *)out, (const void *)gate_out})
        CGIC_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, CGIC_EINVAL, "cgic_route_mix: buffers must be 16-byte aligned");
    {
        CGIC_PROF("route_mix_kernel", as_stream(stream));
        CGIC_CUDA_CHECK(launch_pdl(route_mix_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, as_stream(stream), h_c, h_m, h_f, m_c, m_m, mode, n, C,
                                   h, w, m_f_out, gate_out, out));
    }
    CGIC_LAUNCH_CHECK();
    return CGIC_OK;
}
