// Prepared codebook for VectorQuantize2.forward (CGIC/modules/vqvae/quantize.py:69-98) and the
// indexed nearest-code search on top of it.
//
// The exhaustive search of vq_assign.cu is fp32-issue bound (K distance evaluations per token).
// The reference's argmin, however, is decided among a handful of codes: a code can only win at z
// if its TRUE squared distance is within 2E of the smallest true distance, where E bounds the
// rounding error of the reference's fp32 evaluation  d = fl(fl(|z|^2 + |e|^2) - 2 fl(z.e)):
//     |d_k - |z - e_k|^2|  <=  6u (|z| + |e_k|)^2,  u = 2^-24            (we use 10u)
// (|z|^2, |e|^2: 4 roundings on positive terms; z.e: 4 roundings per term; one rounding for each
// of the two final operations).  So the index is a 4-D grid over the codebook's bounding box:
// for every cell R it stores the ascending list of codes that are NOT dominated over the whole
// cell, where j dominates k when
//     min_{z in R} (|z - e_k|^2 - |z - e_j|^2)  >  2 E_max(R)
// -- the left side is linear in z, so its minimum over a box is closed-form.  The search then
// evaluates the reference's exact rounding sequence on the cell's list only (lowest index wins
// ties, as torch.argmin): every code that could be the reference's answer is in the list, hence
// the result is bit-identical to the exhaustive search for ANY input.  Tokens outside the grid,
// in a cell whose list overflows its record, or with a codebook the index cannot describe
// (non-finite / huge rows) take the exhaustive path inside the same kernel.
//
// Grid: per dimension 256 uniform bins over [min - pad, max + pad] mapped by a lookup table to
// G cells whose edges follow the (clipped) histogram of the code coordinates, so cells adapt to
// where the codes are.  bin(z) is computed in fp32 exactly as the search kernel does it; the cell
// boxes used for the lists are widened by 1e-3 bin (the bin computation is off by < 1e-4 bin).
#include <new>

#include "codebook.cuh"

namespace cgic {
namespace {

__device__ __forceinline__ unsigned f2ord(float f)
{
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned o) { return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o); }

// ---- build, step 1: copy the codebook, e^2, choose the grid ----------------------------------
__global__ void __launch_bounds__(1024, 1) cb_grid_kernel(const float *__restrict__ codebook, int K, unsigned char *__restrict__ blob)
{
    __shared__ unsigned s_min[4], s_max[4];
    __shared__ int s_bad;
    __shared__ unsigned s_hist[4][CB_NB];
    __shared__ float s_lo[4], s_inv[4];
    const CbLayout L = cb_layout(K);
    CbHeader *hdr = reinterpret_cast<CbHeader *>(blob);
    unsigned char *lut = blob + L.lut;
    float4 *cb = reinterpret_cast<float4 *>(blob + L.cb);
    float *e2 = reinterpret_cast<float *>(blob + L.e2);
    const int tid = threadIdx.x;
    if (tid < 4) {
        s_min[tid] = 0xffffffffu;
        s_max[tid] = 0u;
    }
    if (tid == 0) s_bad = 0;
    for (int i = tid; i < 4 * CB_NB; i += blockDim.x) (&s_hist[0][0])[i] = 0u;
    __syncthreads();
    const int K4 = (K + 3) & ~3;
    for (int k = tid; k < K4; k += blockDim.x) {
        float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < K) {
            e = reinterpret_cast<const float4 *>(codebook)[k];
            const float v[4] = {e.x, e.y, e.z, e.w};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (!(fabsf(v[c]) < 1e18f)) atomicOr(&s_bad, 1);  // NaN, inf or a square that would overflow
                atomicMin(&s_min[c], f2ord(v[c]));
                atomicMax(&s_max[c], f2ord(v[c]));
            }
        }
        cb[k] = e;
        e2[k] = k < K ? sumsq4f(e.x, e.y, e.z, e.w) : __int_as_float(0x7f800000);
    }
    __syncthreads();
    if (tid < 4) {
        const double mn = (double)ord2f(s_min[tid]), mx = (double)ord2f(s_max[tid]);
        double range = mx - mn;
        if (!(range > 0.0)) range = fmax(fabs(mn), 1e-20) * 1e-3;  // every code equal along this dimension
        const double binw = range / (double)(CB_NB - 2 * CB_PAD);
        const float lo = (float)(mn - CB_PAD * binw), inv = (float)(1.0 / binw);
        s_lo[tid] = lo;
        s_inv[tid] = inv;
        if (!(fabsf(lo) < 1e30f) || !(inv > 1e-30f && inv < 1e30f) || !(range > 1e-30)) atomicOr(&s_bad, 1);
    }
    __syncthreads();
    const bool bad = s_bad != 0;
    if (!bad) {
        for (int k = tid; k < K; k += blockDim.x) {
            const float4 e = cb[k];
            const float v[4] = {e.x, e.y, e.z, e.w};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int b = cb_bin(v[c], s_lo[c], s_inv[c]);
                if (b < 0) atomicOr(&s_bad, 2);  // cannot happen with the pads above; checked anyway
                else atomicAdd(&s_hist[c][b], 1u);
            }
        }
    }
    __syncthreads();
    if (tid < 4) {
        // cell edges along dimension `tid`: equal shares of the clipped histogram (a clump of codes cannot claim
        // more than half a cell's share; empty bins weigh a little so that empty space is shared as well)
        const int c = tid;
        int *edges = hdr->edges[c];
        double tot = 0.0;
        const double clip = 0.5 * (double)K / CB_G;
        for (int b = 0; b < CB_NB; ++b) tot += fmin((double)s_hist[c][b], clip) + 1e-3;
        double cum = 0.0;
        int cell = 0;
        edges[0] = 0;
        for (int b = 0; b < CB_NB; ++b) {
            lut[c * CB_NB + b] = (unsigned char)cell;
            cum += fmin((double)s_hist[c][b], clip) + 1e-3;
            // close the cell after bin b when its share is reached, keeping at least one bin for every later cell
            const bool must = (CB_NB - 1 - b) == (CB_G - 1 - cell);
            if (cell < CB_G - 1 && (must || cum >= (cell + 1) * tot / CB_G)) {
                ++cell;
                edges[cell] = b + 1;
            }
        }
        edges[CB_G] = CB_NB;
        hdr->lo[c] = s_lo[c];
        hdr->inv[c] = s_inv[c];
    }
    if (tid == 0) {
        hdr->G = CB_G;
        hdr->NB = CB_NB;
        hdr->RW = CB_RW;
        hdr->K = K;
        hdr->max_count = 0;
        hdr->overflow_cells = 0;
        hdr->pad0 = 0;
    }
    __syncthreads();
    if (tid == 0) {
        hdr->valid = s_bad == 0 ? 1 : 0;
        double em = 0.0;
        for (int k = 0; k < K; ++k) {
            const float4 e = cb[k];
            const double n = (double)e.x * e.x + (double)e.y * e.y + (double)e.z * e.z + (double)e.w * e.w;
            em = fmax(em, n);
        }
        hdr->emax = sqrt(em);
    }
}

// ---- build, step 2: the candidate list of one cell per CTA ---------------------------------
__global__ void __launch_bounds__(256) cb_cells_kernel(unsigned char *__restrict__ blob, int K)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *es = reinterpret_cast<float4 *>(smem_raw);                       // [K]
    unsigned short *list = reinterpret_cast<unsigned short *>(es + K);       // [2][K]: survivors of stage 1 / stage 2
    __shared__ int s_dom[CB_NDOM];
    __shared__ int s_wcnt[8];
    __shared__ int s_base;
    const CbLayout L = cb_layout(K);
    CbHeader *hdr = reinterpret_cast<CbHeader *>(blob);
    unsigned short *rec = reinterpret_cast<unsigned short *>(blob + L.rec) + (size_t)blockIdx.x * CB_RW;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (!hdr->valid) {
        if (tid == 0) rec[0] = 0xffffu;
        return;
    }
    for (int k = tid; k < K; k += blockDim.x) es[k] = reinterpret_cast<const float4 *>(blob + L.cb)[k];
    // the cell's box, widened by 1e-3 bin on every side
    double lo[4], hi[4];
    {
        int c3 = blockIdx.x;
        int cc[4];
        for (int c = 3; c >= 0; --c) {
            cc[c] = c3 % CB_G;
            c3 /= CB_G;
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const double l0 = (double)hdr->lo[c], iv = (double)hdr->inv[c];
            lo[c] = l0 + ((double)hdr->edges[c][cc[c]] - 1e-3) / iv;
            hi[c] = l0 + ((double)hdr->edges[c][cc[c] + 1] + 1e-3) / iv;
        }
    }
    double zmax2 = 0.0;
#pragma unroll
    for (int c = 0; c < 4; ++c) zmax2 += fmax(lo[c] * lo[c], hi[c] * hi[c]);
    const double reach = sqrt(zmax2) + hdr->emax;
    const double u = 5.9604644775390625e-08;  // 2^-24
    const double margin = 2.0 * 10.0 * u * reach * reach * 1.001 + 1e-40;
    __syncthreads();

    // domination test: is code k worse than dominator j by more than the rounding slack EVERYWHERE in the cell?
    auto dominated = [&](int k, const int *dom, int ndom) {
        const float4 ek = es[k];
        const double e[4] = {ek.x, ek.y, ek.z, ek.w};
        const double n_k = e[0] * e[0] + e[1] * e[1] + e[2] * e[2] + e[3] * e[3];
        for (int i = 0; i < ndom; ++i) {
            const float4 ejf = es[dom[i]];
            const double ej[4] = {ejf.x, ejf.y, ejf.z, ejf.w};
            double mn = n_k - (ej[0] * ej[0] + ej[1] * ej[1] + ej[2] * ej[2] + ej[3] * ej[3]);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const double dl = e[c] - ej[c];
                mn -= fmax(2.0 * lo[c] * dl, 2.0 * hi[c] * dl);  // min over the box of a linear function: per-coordinate extremes
            }
            if (mn > margin) return true;
        }
        return false;
    };
    // ordered compaction of the codes src[0..n) (ascending) that survive `dom` into dst; whole CTA; returns the count
    auto survivors = [&](const unsigned short *src, int n, const int *dom, int ndom, unsigned short *dst) {
        if (tid == 0) s_base = 0;
        __syncthreads();
        for (int i0 = 0; i0 < n; i0 += blockDim.x) {
            const int i = i0 + tid;
            const int k = i < n ? (src ? (int)src[i] : i) : -1;
            const bool alive = k >= 0 && !dominated(k, dom, ndom);
            const unsigned m = __ballot_sync(0xffffffffu, alive);
            if (lane == 0) s_wcnt[warp] = __popc(m);
            __syncthreads();
            int before = s_base;
            for (int w = 0; w < warp; ++w) before += s_wcnt[w];
            if (alive) dst[before + __popc(m & ((1u << lane) - 1u))] = (unsigned short)k;
            __syncthreads();
            if (tid == 0) {
                int t = s_base;
                for (int w = 0; w < 8; ++w) t += s_wcnt[w];
                s_base = t;
            }
            __syncthreads();
        }
        return s_base;
    };
    // nearest code to p among src[0..n) (src == nullptr: all K codes), one warp; ties -> lowest index
    auto nearest = [&](const double p[4], const unsigned short *src, int n) {
        double bd = 1e300;
        int bk = 0x7fffffff;
        for (int i = lane; i < n; i += 32) {
            const int k = src ? (int)src[i] : i;
            const float4 e = es[k];
            const double d0 = e.x - p[0], d1 = e.y - p[1], d2 = e.z - p[2], d3 = e.w - p[3];
            const double d = d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
            if (d < bd || (d == bd && k < bk)) {
                bd = d;
                bk = k;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double od = __shfl_xor_sync(0xffffffffu, bd, o);
            const int ok = __shfl_xor_sync(0xffffffffu, bk, o);
            if (od < bd || (od == bd && ok < bk)) {
                bd = od;
                bk = ok;
            }
        }
        return bk;
    };

    // ---- stage 1: eight dominators = each warp's nearest code to the cell centre within its slice of the codebook
    //      (the global nearest is among them); they already remove all but a few dozen codes
    {
        double p[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) p[c] = 0.5 * (lo[c] + hi[c]);
        const int per = (K + 7) / 8, k0 = warp * per, k1 = min(K, k0 + per);
        double bd = 1e300;
        int bk = 0x7fffffff;
        for (int k = k0 + lane; k < k1; k += 32) {
            const float4 e = es[k];
            const double d0 = e.x - p[0], d1 = e.y - p[1], d2 = e.z - p[2], d3 = e.w - p[3];
            const double d = d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
            if (d < bd) {
                bd = d;
                bk = k;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double od = __shfl_xor_sync(0xffffffffu, bd, o);
            const int ok = __shfl_xor_sync(0xffffffffu, bk, o);
            if (od < bd || (od == bd && ok < bk)) {
                bd = od;
                bk = ok;
            }
        }
        if (lane == 0) s_dom[warp] = bk == 0x7fffffff ? 0 : bk;  // an empty slice (K < 8) repeats code 0: harmless
    }
    __syncthreads();
    unsigned short *list2 = list + K;
    int count = survivors(nullptr, K, s_dom, 8, list);
    // ---- stage 2: the nearest survivor to each of the 16 corners (the nearest code to a point of the cell is never
    //      dominated, so it is among the survivors); a second pass with these dominators
    for (int q = warp; q < 16; q += 8) {
        double p[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) p[c] = ((q >> c) & 1) ? hi[c] : lo[c];
        const int k = nearest(p, list, count);
        if (lane == 0) s_dom[q] = k == 0x7fffffff ? 0 : k;  // (no survivor cannot happen; any code is a valid dominator)
    }
    __syncthreads();
    count = survivors(list, count, s_dom, 16, list2);
    list = list2;
    if (count >= 1 && count <= CB_RW - 1) {
        for (int i = tid; i < CB_RW; i += blockDim.x) rec[i] = i == 0 ? (unsigned short)count : list[min(i, count) - 1];
    } else {
        for (int i = tid; i < CB_RW; i += blockDim.x) rec[i] = i == 0 ? 0xffffu : 0;
        if (tid == 0) atomicAdd(&hdr->overflow_cells, 1);
    }
    if (tid == 0) atomicMax(&hdr->max_count, count);
}

// ---- staleness guard: is the live weight tensor still the one the index was built from? -------
// Writes through `weight.data` (the reference's own LitEma.copy_to / restore, CGIC/models/ema.py:51,76) do not move
// torch's version counter, so the host cannot know.  One CTA compares the live rows with the blob's copy bit for bit;
// on a mismatch it refreshes the copy and e^2, marks the index unusable (every latent then takes the exhaustive path
// of vq_warp_kernel -- same results, slower) and raises a flag in mapped host memory that the next call polls.
__global__ void __launch_bounds__(1024, 1)
cb_check_kernel(const float *__restrict__ live, int K, unsigned char *__restrict__ blob, int generation, volatile int *stale_flag)
{
    __shared__ int s_diff;
    const CbLayout L = cb_layout(K);
    CbHeader *hdr = reinterpret_cast<CbHeader *>(blob);
    uint4 *cb = reinterpret_cast<uint4 *>(blob + L.cb);
    float *e2 = reinterpret_cast<float *>(blob + L.e2);
    if (threadIdx.x == 0) s_diff = 0;
    __syncthreads();
    bool diff = false;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        const uint4 a = reinterpret_cast<const uint4 *>(live)[k], b = cb[k];
        diff |= a.x != b.x || a.y != b.y || a.z != b.z || a.w != b.w;
    }
    if (diff) s_diff = 1;
    __syncthreads();
    if (!s_diff) return;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        const float4 e = reinterpret_cast<const float4 *>(live)[k];
        reinterpret_cast<float4 *>(cb)[k] = e;
        e2[k] = sumsq4f(e.x, e.y, e.z, e.w);
    }
    if (threadIdx.x == 0) {
        hdr->valid = 0;
        *stale_flag = generation;
    }
}

}  // namespace
}  // namespace cgic

using namespace cgic;

struct cgic_codebook {
    int K = 0;
    int device = -1;
    unsigned char *blob = nullptr;
    size_t bytes = 0;
    bool built = false;
    int generation = 0;          // bumped by every update; a check kernel reports staleness by writing its generation
    int *stale_host = nullptr;   // mapped, pinned: written by cb_check_kernel, polled by cgic_codebook_is_stale
    int *stale_dev = nullptr;
};

extern "C" int cgic_codebook_create(int K, cgic_codebook **out)
{
    CGIC_REQUIRE(out, CGIC_EINVAL, "cgic_codebook_create: null argument");
    CGIC_REQUIRE(K >= 1 && K <= CB_MAX_K, CGIC_EINVAL, "cgic_codebook_create: K=%d outside [1, %d]", K, CB_MAX_K);
    auto *cb = new (std::nothrow) cgic_codebook();
    CGIC_REQUIRE(cb, CGIC_ENOMEM, "cgic_codebook_create: out of memory");
    cb->K = K;
    cb->bytes = cb_layout(K).total;
    cudaError_t e = cudaGetDevice(&cb->device);
    if (e == cudaSuccess) e = cudaMalloc(&cb->blob, cb->bytes);
    if (e == cudaSuccess) e = cudaHostAlloc(reinterpret_cast<void **>(&cb->stale_host), sizeof(int), cudaHostAllocMapped);
    if (e == cudaSuccess) {
        *cb->stale_host = -1;
        e = cudaHostGetDevicePointer(reinterpret_cast<void **>(&cb->stale_dev), cb->stale_host, 0);
    }
    if (e != cudaSuccess) {
        set_error("cgic_codebook_create: %s", cudaGetErrorString(e));
        if (cb->blob) cudaFree(cb->blob);
        if (cb->stale_host) cudaFreeHost(cb->stale_host);
        delete cb;
        return e == cudaErrorMemoryAllocation ? CGIC_ENOMEM : CGIC_ECUDA;
    }
    *out = cb;
    return CGIC_OK;
}

extern "C" void cgic_codebook_free(cgic_codebook *cb)
{
    if (!cb) return;
    if (cb->blob) cudaFree(cb->blob);
    if (cb->stale_host) cudaFreeHost(cb->stale_host);
    delete cb;
}

extern "C" int cgic_codebook_update(cgic_codebook *cb, const float *codebook, cgic_stream_t stream_)
{
    CGIC_REQUIRE(cb && codebook, CGIC_EINVAL, "cgic_codebook_update: null argument");
    CGIC_REQUIRE((reinterpret_cast<uintptr_t>(codebook) & 15) == 0, CGIC_EINVAL, "cgic_codebook_update: codebook must be 16-byte aligned");
    cudaStream_t stream = as_stream(stream_);
    const int K = cb->K;
    const size_t smem = (size_t)K * (16 + 4);
    {
        const int rc = ensure_smem((const void *)cb_cells_kernel, CB_MAX_K * (16 + 4));
        if (rc) return rc;
    }
    {
        CGIC_PROF("cb_grid_kernel", stream);
        cb_grid_kernel<<<1, 1024, 0, stream>>>(codebook, K, cb->blob);
    }
    CGIC_LAUNCH_CHECK();
    {
        CGIC_PROF("cb_cells_kernel", stream);
        cb_cells_kernel<<<CB_G * CB_G * CB_G * CB_G, 256, smem, stream>>>(cb->blob, K);
    }
    CGIC_LAUNCH_CHECK();
    cb->built = true;
    ++cb->generation;  // reports of check kernels launched before this rebuild no longer count
    return CGIC_OK;
}

extern "C" int cgic_codebook_check(cgic_codebook *cb, const float *codebook, cgic_stream_t stream_)
{
    CGIC_REQUIRE(cb && codebook && cb->built, CGIC_EINVAL, "cgic_codebook_check: no built codebook");
    CGIC_REQUIRE((reinterpret_cast<uintptr_t>(codebook) & 15) == 0, CGIC_EINVAL, "cgic_codebook_check: codebook must be 16-byte aligned");
    cudaStream_t stream = as_stream(stream_);
    {
        CGIC_PROF("cb_check_kernel", stream);
        cb_check_kernel<<<1, 1024, 0, stream>>>(codebook, cb->K, cb->blob, cb->generation, cb->stale_dev);
    }
    CGIC_LAUNCH_CHECK();
    return CGIC_OK;
}

extern "C" int cgic_codebook_is_stale(const cgic_codebook *cb)
{
    if (!cb || !cb->built) return CGIC_EINVAL;
    return *reinterpret_cast<volatile int *>(cb->stale_host) == cb->generation ? 1 : 0;
}

extern "C" int cgic_codebook_stats_host(const cgic_codebook *cb, int32_t out[4])
{
    CGIC_REQUIRE(cb && out && cb->built, CGIC_EINVAL, "cgic_codebook_stats_host: no built codebook");
    CbHeader h;
    CGIC_CUDA_CHECK(cudaMemcpy(&h, cb->blob, sizeof(h), cudaMemcpyDeviceToHost));
    out[0] = h.valid;
    out[1] = h.G * h.G * h.G * h.G;
    out[2] = h.max_count;
    out[3] = h.overflow_cells;
    return CGIC_OK;
}

namespace cgic {
const unsigned char *codebook_blob(const cgic_codebook *cb) { return cb && cb->built ? cb->blob : nullptr; }
int codebook_size(const cgic_codebook *cb) { return cb ? cb->K : 0; }
}  // namespace cgic

