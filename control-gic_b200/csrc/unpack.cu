// a10 + a11 + a13 + a14: stream decode, mask / index re-assembly and codebook gather.
//
//   reference: CGIC/models/model.py:269-392;  HuffmanCoding.decompress_string
//              CGIC/tools/indices_coding.py:153-168 (131-138, 140-151);  BinaryCoding
//              CGIC/tools/mask_coding.py:81-96.
//
// Stage 1 (unpack_decode_kernel, one CTA per (stream, image)):
//   streams 0..2  prefix decode of the Huffman payload into a symbol list (table-driven: the
//                 first lut_bits bits index a LUT that either names the symbol or the tree node
//                 to continue from, so codes of any length -- the untrained model has 224-bit
//                 codes -- decode correctly);
//   stream 3 / 4  granularity masks -> bitmaps (one bit per grid cell) plus an exclusive
//                 popcount prefix per 32-bit word, for the coarse (CTA 3) and the medium and
//                 derived fine level (CTA 4): fine = 1 - up2(medium) - up4(coarse).
// Stage 2 (unpack_assemble_kernel, one thread per fine token): rank of the token's cell among
//   the set cells of its level = prefix[word] + popc(bits below) -> the decoded symbol that the
//   reference's masked assignment puts there (row-major order); ind = fine + up2(medium) +
//   up4(coarse); quant = codebook[ind] written NCHW; masks written as int64 like the reference.
#include "common.cuh"

namespace cgic {
namespace {

constexpr int UP_THREADS = 128;

struct UnpackWs {
    uint16_t *sym;     // [B][n16 + n8 + n4]
    int32_t *count;    // [B][3]  decoded symbols per index stream (-1: empty stream)
    int32_t *pop;      // [B][3]  population of each mask level
    uint32_t *bits;    // [B][nw16 + nw8 + nw4] bitmaps
    uint32_t *prefix;  // same shape: exclusive popcount prefix
    size_t bytes;
};

struct Geo {
    int h, w, h8, w8, h16, w16;
    int64_t n4, n8, n16;
    int nw16, nw8, nw4;
};

__host__ __device__ inline Geo make_geo(int h, int w)
{
    Geo g;
    g.h = h;
    g.w = w;
    g.h8 = h / 2;
    g.w8 = w / 2;
    g.h16 = h / 4;
    g.w16 = w / 4;
    g.n4 = (int64_t)h * w;
    g.n8 = (int64_t)g.h8 * g.w8;
    g.n16 = (int64_t)g.h16 * g.w16;
    g.nw16 = (int)((g.n16 + 31) / 32);
    g.nw8 = (int)((g.n8 + 31) / 32);
    g.nw4 = (int)((g.n4 + 31) / 32);
    return g;
}

UnpackWs carve_unpack(void *ws, int B, const Geo &g)
{
    auto up = [](size_t v) { return (v + 255) / 256 * 256; };
    UnpackWs c{};
    unsigned char *p = static_cast<unsigned char *>(ws);
    size_t o = 0;
    c.count = reinterpret_cast<int32_t *>(p + o);
    o += up((size_t)B * 3 * 4);
    c.pop = reinterpret_cast<int32_t *>(p + o);
    o += up((size_t)B * 3 * 4);
    c.sym = reinterpret_cast<uint16_t *>(p + o);
    o += up((size_t)B * (g.n16 + g.n8 + g.n4) * 2);
    c.bits = reinterpret_cast<uint32_t *>(p + o);
    o += up((size_t)B * (g.nw16 + g.nw8 + g.nw4) * 4);
    c.prefix = reinterpret_cast<uint32_t *>(p + o);
    o += up((size_t)B * (g.nw16 + g.nw8 + g.nw4) * 4);
    c.bytes = o;
    return c;
}

struct UnpackArgs {
    const uint8_t *bytes;
    const int32_t *sizes;
    int64_t image_stride;
    int64_t slot_off[5];
    int mode;
    Geo g;
    DevTable T;
    UnpackWs ws;
    const float *codebook;
    int64_t *mc_out, *mm_out, *mf_out, *ind_out;
    float *quant_out;
    int32_t *status;
};

// ---- serial prefix decoder (one thread).  Returns the number of symbols, -1 for an empty
// stream (the reference returns None), -2 on overflow of `cap`.
template <typename Out>
__device__ int decode_stream(const uint8_t *in, int64_t nbytes, const DevTable &T, Out *out, int64_t cap)
{
    if (nbytes <= 0) return -1;
    const int pad = in[0];
    int64_t nbits = (nbytes - 1) * 8 - pad;
    if (pad == 0 || nbits < 0) nbits = 0;  // text[:-0] is empty in the reference
    const uint8_t *p = in + 1;
    const int64_t npay = nbytes - 1;
    unsigned long long buf = 0;  // MSB-aligned window
    int avail = 0;
    int64_t next = 0;  // next payload byte to load
    int64_t pos = 0;   // consumed bits
    int cnt = 0;
    const int L = T.lut_bits;
    while (pos < nbits) {
        while (avail <= 56) {
            const unsigned long long byte = next < npay ? p[next] : 0;
            ++next;
            buf |= byte << (56 - avail);
            avail += 8;
        }
        const uint32_t e = __ldg(&T.lut[(uint32_t)(buf >> (64 - L))]);
        int len = e & 0xFF;
        int sym;
        if (len != 0xFF) {
            sym = e >> 8;
        } else {
            // long code: continue bit by bit from the tree node reached after L bits
            if (pos + L >= nbits) break;
            int node = e >> 8;
            buf <<= L;
            avail -= L;
            pos += L;
            len = 0;
            for (;;) {
                if (pos + len >= nbits) return cnt;  // trailing incomplete code is dropped
                if (avail == 0) {
                    const unsigned long long byte = next < npay ? p[next] : 0;
                    ++next;
                    buf = byte << 56;
                    avail = 8;
                }
                const int bit = (int)(buf >> 63);
                buf <<= 1;
                --avail;
                ++len;
                node = __ldg(&T.child[2 * node + bit]);
                if (node < T.K) break;
            }
            if (cnt >= cap) return -2;
            out[cnt++] = (Out)node;
            pos += len;
            continue;
        }
        if (pos + len > nbits) break;  // code completed only thanks to pad bits: not a symbol
        if (cnt >= cap) return -2;
        out[cnt++] = (Out)sym;
        buf <<= len;
        avail -= len;
        pos += len;
    }
    return cnt;
}

// mask cell value per mode (model.py:278-280, 301-303, 319-321, 338-340, 360-387)
__device__ __forceinline__ int stream_bit(const uint8_t *in, int64_t p) { return (in[1 + (p >> 3)] >> (7 - (int)(p & 7))) & 1; }

__device__ __forceinline__ int coarse_cell(const UnpackArgs &a, const uint8_t *mc, int y16, int x16)
{
    if (a.mode == 0 || a.mode == 2 || a.mode == 3) return stream_bit(mc, (int64_t)y16 * a.g.w16 + x16);
    return a.mode == 4;
}
__device__ __forceinline__ int medium_cell(const UnpackArgs &a, const uint8_t *mc, const uint8_t *mm, int y8, int x8)
{
    if (a.mode == 0 || a.mode == 1) return stream_bit(mm, (int64_t)y8 * a.g.w8 + x8);
    if (a.mode == 3) return 1 - coarse_cell(a, mc, y8 >> 1, x8 >> 1);
    return a.mode == 5;
}
__device__ __forceinline__ int fine_cell(const UnpackArgs &a, const uint8_t *mc, const uint8_t *mm, int y, int x)
{
    if (a.mode <= 2) return (1 - medium_cell(a, mc, mm, y >> 1, x >> 1) - coarse_cell(a, mc, y >> 2, x >> 2)) == 1;
    return a.mode == 6;
}

// bitmap + exclusive popcount prefix of one level, by the whole CTA
template <typename F>
__device__ void build_level(int gw, int64_t n, int nw, uint32_t *bits, uint32_t *prefix, int32_t *pop_out, F cell)
{
    __shared__ int s_warp[UP_THREADS / 32 + 1];
    __shared__ int s_run;
    if (threadIdx.x == 0) s_run = 0;
    __syncthreads();
    for (int base = 0; base < nw; base += UP_THREADS) {
        const int wi = base + threadIdx.x;
        uint32_t word = 0;
        if (wi < nw) {
            const int64_t p0 = (int64_t)wi * 32;
            int y = (int)(p0 / gw), x = (int)(p0 - (int64_t)y * gw);
            for (int i = 0; i < 32 && p0 + i < n; ++i) {
                word |= (uint32_t)cell(y, x) << i;
                if (++x == gw) {
                    x = 0;
                    ++y;
                }
            }
            bits[wi] = word;
        }
        const int v = __popc(word);
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) s_warp[wid] = inc;
        __syncthreads();
        int woff = 0;
        for (int i = 0; i < wid; ++i) woff += s_warp[i];
        int tot = 0;
        for (int i = 0; i < UP_THREADS / 32; ++i) tot += s_warp[i];
        const int run = s_run;
        if (wi < nw) prefix[wi] = (uint32_t)(run + woff + inc - v);
        __syncthreads();
        if (threadIdx.x == 0) s_run = run + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) *pop_out = s_run;
}

__global__ void __launch_bounds__(UP_THREADS) unpack_decode_kernel(const UnpackArgs a)
{
    const int s = blockIdx.x, b = blockIdx.y;
    const Geo &g = a.g;
    const uint8_t *img = a.bytes + (int64_t)b * a.image_stride;
    const int32_t *sz = a.sizes + b * 5;
    const int nwt = g.nw16 + g.nw8 + g.nw4;
    if (s < 3) {
        if (threadIdx.x != 0) return;
        const int64_t soff = s == 0 ? 0 : (s == 1 ? g.n16 : g.n16 + g.n8);
        const int64_t cap = s == 0 ? g.n16 : (s == 1 ? g.n8 : g.n4);
        int cnt = -1;
        if (stream_present(a.mode, s))
            cnt = decode_stream<uint16_t>(img + a.slot_off[s], sz[s], a.T, a.ws.sym + (int64_t)b * (g.n16 + g.n8 + g.n4) + soff, cap);
        a.ws.count[b * 3 + s] = cnt;
        if (cnt == -2) atomicExch(&a.status[b], CGIC_EFORMAT);
        return;
    }
    const uint8_t *mc = img + a.slot_off[3];
    const uint8_t *mm = img + a.slot_off[4];
    // framing check of the mask streams this mode reads
    if (threadIdx.x == 0) {
        const bool need_c = a.mode == 0 || a.mode == 2 || a.mode == 3;
        const bool need_m = a.mode == 0 || a.mode == 1;
        if (s == 3 && need_c && (sz[3] != g.n16 / 8 + 2 || mc[0] != 8 - (int)(g.n16 & 7))) atomicExch(&a.status[b], CGIC_EFORMAT);
        if (s == 4 && need_m && (sz[4] != g.n8 / 8 + 2 || mm[0] != 8 - (int)(g.n8 & 7))) atomicExch(&a.status[b], CGIC_EFORMAT);
    }
    uint32_t *bits = a.ws.bits + (int64_t)b * nwt;
    uint32_t *prefix = a.ws.prefix + (int64_t)b * nwt;
    int32_t *pop = a.ws.pop + b * 3;
    if (s == 3) {
        build_level(g.w16, g.n16, g.nw16, bits, prefix, pop + 0, [&](int y, int x) { return coarse_cell(a, mc, y, x); });
    } else {
        build_level(g.w8, g.n8, g.nw8, bits + g.nw16, prefix + g.nw16, pop + 1,
                    [&](int y, int x) { return medium_cell(a, mc, mm, y, x); });
        build_level(g.w, g.n4, g.nw4, bits + g.nw16 + g.nw8, prefix + g.nw16 + g.nw8, pop + 2,
                    [&](int y, int x) { return fine_cell(a, mc, mm, y, x); });
    }
}

__global__ void __launch_bounds__(256) unpack_assemble_kernel(const UnpackArgs a)
{
    const Geo &g = a.g;
    const int b = blockIdx.y;
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int nwt = g.nw16 + g.nw8 + g.nw4;
    const uint32_t *bits = a.ws.bits + (int64_t)b * nwt;
    const uint32_t *prefix = a.ws.prefix + (int64_t)b * nwt;
    const int32_t *cnt = a.ws.count + b * 3;
    const int32_t *pop = a.ws.pop + b * 3;
    if (p == 0) {
        // the reference's masked assignment raises unless #symbols == #set cells (an empty
        // coarse / medium stream stands for zeros, model.py:284-290)
        for (int s = 0; s < 3; ++s) {
            if (!stream_present(a.mode, s)) continue;
            const bool empty_ok = cnt[s] == -1 && (s < 2 || pop[s] == 0);
            if (!empty_ok && cnt[s] != pop[s]) atomicExch(&a.status[b], CGIC_EFORMAT);
        }
    }
    if (p >= g.n4) return;
    const int y = (int)(p / g.w), x = (int)(p - (int64_t)y * g.w);
    const uint16_t *sym = a.ws.sym + (int64_t)b * (g.n16 + g.n8 + g.n4);
    const int64_t p16 = (int64_t)(y >> 2) * g.w16 + (x >> 2);
    const int64_t p8 = (int64_t)(y >> 1) * g.w8 + (x >> 1);
    const uint32_t wc = bits[p16 >> 5], wm = bits[g.nw16 + (p8 >> 5)], wf = bits[g.nw16 + g.nw8 + (p >> 5)];
    const int cbit = (wc >> (p16 & 31)) & 1, mbit = (wm >> (p8 & 31)) & 1, fbit = (wf >> (p & 31)) & 1;
    int64_t ind = 0;
    if (cbit && cnt[0] > 0) {
        const int r = prefix[p16 >> 5] + __popc(wc & ((1u << (p16 & 31)) - 1u));
        if (r < cnt[0]) ind += sym[r];
    }
    if (mbit && cnt[1] > 0) {
        const int r = prefix[g.nw16 + (p8 >> 5)] + __popc(wm & ((1u << (p8 & 31)) - 1u));
        if (r < cnt[1]) ind += sym[g.n16 + r];
    }
    if (fbit && cnt[2] > 0) {
        const int r = prefix[g.nw16 + g.nw8 + (p >> 5)] + __popc(wf & ((1u << (p & 31)) - 1u));
        if (r < cnt[2]) ind += sym[g.n16 + g.n8 + r];
    }
    a.ind_out[(int64_t)b * g.n4 + p] = ind;
    a.mf_out[(int64_t)b * g.n4 + p] = fbit;
    if (((y | x) & 1) == 0) a.mm_out[(int64_t)b * g.n8 + p8] = mbit;
    if (((y | x) & 3) == 0) a.mc_out[(int64_t)b * g.n16 + p16] = cbit;
    if (a.quant_out) {
        const int64_t k = ind < a.T.K ? ind : 0;  // sums of overlapping levels cannot occur with valid masks
        const float4 e = __ldg(reinterpret_cast<const float4 *>(a.codebook) + k);
        float *q = a.quant_out + (int64_t)b * 4 * g.n4 + p;
        q[0] = e.x;
        q[g.n4] = e.y;
        q[2 * g.n4] = e.z;
        q[3 * g.n4] = e.w;
        if (ind >= a.T.K) atomicExch(&a.status[b], CGIC_EFORMAT);
    }
}

__global__ void huff_decode_single_kernel(const uint8_t *bytes, int64_t nbytes, DevTable T, int32_t *out, int64_t cap,
                                          int32_t *count_out)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) *count_out = decode_stream<int32_t>(bytes, nbytes, T, out, cap);
}

__global__ void bits_decode_single_kernel(const uint8_t *bytes, int64_t nbytes, int32_t *out, int64_t cap, int32_t *count_out)
{
    if (nbytes <= 0) {
        if (threadIdx.x == 0 && blockIdx.x == 0) *count_out = -1;
        return;
    }
    const int pad = bytes[0];
    int64_t nbits = (nbytes - 1) * 8 - pad;
    if (pad == 0 || nbits < 0) nbits = 0;
    if (nbits > cap) {
        if (threadIdx.x == 0 && blockIdx.x == 0) *count_out = -2;
        return;
    }
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nbits; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = stream_bit(bytes, i);
    if (threadIdx.x == 0 && blockIdx.x == 0) *count_out = (int32_t)nbits;
}

}  // namespace
}  // namespace cgic

using namespace cgic;

extern "C" size_t cgic_unpack_workspace_bytes(int B, int h, int w)
{
    if (B <= 0 || h <= 0 || w <= 0) return 256;
    return carve_unpack(nullptr, B, make_geo(h, w)).bytes;
}

extern "C" int cgic_unpack(const uint8_t *bytes, const int32_t *sizes, int B, int h, int w, int mode, const cgic_table *t,
                           const float *codebook, int64_t *mc_out, int64_t *mm_out, int64_t *mf_out, int64_t *ind_out,
                           float *quant_out, int32_t *status_out, void *workspace, size_t workspace_bytes,
                           cgic_stream_t stream_)
{
    CGIC_REQUIRE(bytes && sizes && mc_out && mm_out && mf_out && ind_out && status_out && workspace, CGIC_EINVAL,
                 "cgic_unpack: null argument");
    CGIC_REQUIRE(!quant_out || codebook, CGIC_EINVAL, "cgic_unpack: quant_out needs the codebook");
    CGIC_REQUIRE(B >= 0 && h > 0 && w > 0 && h % 4 == 0 && w % 4 == 0, CGIC_EINVAL, "cgic_unpack: token grid %dx%d must be multiples of 4", h, w);
    CGIC_REQUIRE(mode >= 0 && mode <= 6, CGIC_EINVAL, "cgic_unpack: mode %d", mode);
    CGIC_REQUIRE(!codebook || (reinterpret_cast<uintptr_t>(codebook) & 15) == 0, CGIC_EINVAL, "cgic_unpack: codebook must be 16-byte aligned");
    if (B == 0) return CGIC_OK;
    UnpackArgs a{};
    int rc = table_device_view(t, &a.T);
    if (rc) return rc;
    CGIC_REQUIRE(a.T.K <= 65536, CGIC_EINVAL, "cgic_unpack: K too large");
    a.g = make_geo(h, w);
    a.ws = carve_unpack(workspace, B, a.g);
    CGIC_REQUIRE(workspace_bytes >= a.ws.bytes, CGIC_ESPACE, "cgic_unpack: workspace %zu < %zu bytes", workspace_bytes, a.ws.bytes);
    const PackLayout L = make_pack_layout(a.T.max_len, h, w);
    a.bytes = bytes;
    a.sizes = sizes;
    a.image_stride = L.stride;
    for (int s = 0; s < 5; ++s) a.slot_off[s] = L.off[s];
    a.mode = mode;
    a.codebook = codebook;
    a.mc_out = mc_out;
    a.mm_out = mm_out;
    a.mf_out = mf_out;
    a.ind_out = ind_out;
    a.quant_out = quant_out;
    a.status = status_out;
    cudaStream_t stream = as_stream(stream_);
    CGIC_CUDA_CHECK(cudaMemsetAsync(status_out, 0, (size_t)B * 4, stream));
    {
        CGIC_PROF("unpack_decode_kernel", stream);
        unpack_decode_kernel<<<dim3(5, B), UP_THREADS, 0, stream>>>(a);
    }
    CGIC_LAUNCH_CHECK();
    {
        CGIC_PROF("unpack_assemble_kernel", stream);
        unpack_assemble_kernel<<<dim3((unsigned)((a.g.n4 + 255) / 256), B), 256, 0, stream>>>(a);
    }
    CGIC_LAUNCH_CHECK();
    return CGIC_OK;
}

extern "C" int cgic_huff_decode(const uint8_t *bytes, int64_t nbytes, const cgic_table *t, int32_t *symbols_out, int64_t cap,
                                int32_t *count_out, cgic_stream_t stream)
{
    CGIC_REQUIRE(count_out && nbytes >= 0 && (bytes || nbytes == 0) && (symbols_out || cap == 0), CGIC_EINVAL,
                 "cgic_huff_decode: bad argument");
    DevTable T;
    int rc = table_device_view(t, &T);
    if (rc) return rc;
    huff_decode_single_kernel<<<1, 32, 0, as_stream(stream)>>>(bytes, nbytes, T, symbols_out, cap, count_out);
    CGIC_LAUNCH_CHECK();
    return CGIC_OK;
}

extern "C" int cgic_bits_decode(const uint8_t *bytes, int64_t nbytes, int32_t *values_out, int64_t cap, int32_t *count_out,
                                cgic_stream_t stream)
{
    CGIC_REQUIRE(count_out && nbytes >= 0 && (bytes || nbytes == 0) && (values_out || cap == 0), CGIC_EINVAL,
                 "cgic_bits_decode: bad argument");
    const int64_t nb = nbytes * 8;
    const int grid = (int)(nb / 256 + 1 > 1024 ? 1024 : nb / 256 + 1);
    bits_decode_single_kernel<<<grid, 256, 0, as_stream(stream)>>>(bytes, nbytes, values_out, cap, count_out);
    CGIC_LAUNCH_CHECK();
    return CGIC_OK;
}
