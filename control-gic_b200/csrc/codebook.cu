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

#include "common.cuh"

namespace cgic {
namespace {

constexpr int CB_NB = 256;        // fine bins per dimension
constexpr int CB_PAD = 48;        // bins beyond the codebook's min / max on each side
constexpr int CB_G = 8;           // cells per dimension
constexpr int CB_RW = 64;         // u16 words per cell record: count + up to 63 candidates
constexpr int CB_MAX_K = 4096;
constexpr int CB_NDOM = 20;       // dominators tried per cell: nearest code to the 16 corners + 4 nearest to the centre
constexpr int CB_HDR = 256;       // header bytes
constexpr int CB_LUT = 4 * CB_NB; // lookup-table bytes

struct CbHeader {
    float lo[4];
    float inv[4];
    int G, NB, RW, valid;
    int K, max_count, overflow_cells, pad0;
    double emax;
    int edges[4][CB_G + 1];
};
static_assert(sizeof(CbHeader) <= CB_HDR, "header does not fit");

struct CbLayout {
    size_t lut, cb, e2, rec, total, stage;  // byte offsets; `stage` = bytes the search kernel copies to shared memory
};

__host__ __device__ inline CbLayout cb_layout(int K)
{
    CbLayout L;
    const size_t K4 = ((size_t)K + 3) & ~(size_t)3;
    L.lut = CB_HDR;
    L.cb = L.lut + CB_LUT;
    L.e2 = L.cb + K4 * 16;
    L.stage = L.e2 + K4 * 4;
    L.rec = (L.stage + 127) & ~(size_t)127;
    L.total = L.rec + (size_t)CB_G * CB_G * CB_G * CB_G * CB_RW * 2;
    return L;
}

__device__ __forceinline__ float sumsq4f(float a, float b, float c, float d)
{
    float s = __fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b));
    s = __fadd_rn(s, __fmul_rn(c, c));
    return __fadd_rn(s, __fmul_rn(d, d));
}

// the bin of coordinate v along one dimension, or -1 when outside the grid (also for NaN)
__device__ __forceinline__ int cb_bin(float v, float lo, float inv)
{
    const float t = __fmul_rn(__fsub_rn(v, lo), inv);
    return (t >= 0.f && t < (float)CB_NB) ? (int)t : -1;
}

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
        double w[1];
        double tot = 0.0;
        const double clip = 0.5 * (double)K / CB_G;
        for (int b = 0; b < CB_NB; ++b) tot += fmin((double)s_hist[c][b], clip) + 1e-3;
        (void)w;
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
    unsigned short *list = reinterpret_cast<unsigned short *>(es + K);       // [K]
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

    // ---- dominators: the nearest code to each corner (2 per warp), then the 4 codes nearest to the centre (warp 0)
    auto nearest = [&](const double p[4], const int *excl, int nexcl) {
        double bd = 1e300;
        int bk = 0x7fffffff;
        for (int k = lane; k < K; k += 32) {
            bool skip = false;
            for (int i = 0; i < nexcl; ++i) skip |= excl[i] == k;
            if (skip) continue;
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
        return bk;
    };
    for (int q = warp; q < 16; q += 8) {
        double p[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) p[c] = ((q >> c) & 1) ? hi[c] : lo[c];
        const int k = nearest(p, nullptr, 0);
        if (lane == 0) s_dom[q] = k;
    }
    if (warp == 0) {
        double p[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) p[c] = 0.5 * (lo[c] + hi[c]);
        int ex[4];
        for (int i = 0; i < 4; ++i) {
            int k = nearest(p, ex, i);
            if (k == 0x7fffffff) k = ex[0];  // fewer than 4 codes
            ex[i] = k;
            if (lane == 0) s_dom[16 + i] = k;
        }
    }
    if (tid == 0) s_base = 0;
    __syncthreads();

    // ---- domination tests + ordered compaction (ascending code index)
    for (int k0 = 0; k0 < K; k0 += blockDim.x) {
        const int k = k0 + tid;
        bool alive = k < K;
        if (alive) {
            const float4 ek = es[k];
            const double e[4] = {ek.x, ek.y, ek.z, ek.w};
            const double n_k = e[0] * e[0] + e[1] * e[1] + e[2] * e[2] + e[3] * e[3];
            for (int i = 0; i < CB_NDOM && alive; ++i) {
                const int j = s_dom[i];
                const float4 ejf = es[j];
                const double ej[4] = {ejf.x, ejf.y, ejf.z, ejf.w};
                double mn = n_k - (ej[0] * ej[0] + ej[1] * ej[1] + ej[2] * ej[2] + ej[3] * ej[3]);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const double dl = e[c] - ej[c];
                    mn -= fmax(2.0 * lo[c] * dl, 2.0 * hi[c] * dl);
                }
                if (mn > margin) alive = false;  // e_j is closer than e_k by more than the rounding slack everywhere in the cell
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, alive);
        if (lane == 0) s_wcnt[warp] = __popc(m);
        __syncthreads();
        int before = s_base;
        for (int i = 0; i < warp; ++i) before += s_wcnt[i];
        if (alive) list[before + __popc(m & ((1u << lane) - 1u))] = (unsigned short)k;
        __syncthreads();
        if (tid == 0) {
            int t = s_base;
            for (int i = 0; i < 8; ++i) t += s_wcnt[i];
            s_base = t;
        }
        __syncthreads();
    }
    const int count = s_base;
    if (count >= 1 && count <= CB_RW - 1) {
        for (int i = tid; i < CB_RW; i += blockDim.x) rec[i] = i == 0 ? (unsigned short)count : list[min(i, count) - 1];
    } else {
        for (int i = tid; i < CB_RW; i += blockDim.x) rec[i] = i == 0 ? 0xffffu : 0;
        if (tid == 0) atomicAdd(&hdr->overflow_cells, 1);
    }
    if (tid == 0) atomicMax(&hdr->max_count, count);
}

// ---- the search --------------------------------------------------------------------------
constexpr int VQI_THREADS = 512;
constexpr int VQI_WARPS = VQI_THREADS / 32;

// exact reference distance of one code (rounding sequence of quantize.py:73-75, see vq_assign.cu)
__device__ __forceinline__ float ref_dist(const float z[4], float z2, const float4 e, float e2)
{
    float dot = __fmul_rn(z[0], e.x);
    dot = __fmaf_rn(z[1], e.y, dot);
    dot = __fmaf_rn(z[2], e.z, dot);
    dot = __fmaf_rn(z[3], e.w, dot);
    return __fmaf_rn(dot, -2.f, __fadd_rn(z2, e2));
}

__device__ __forceinline__ void eval_cand(unsigned k, const float z[4], float z2, const float4 *__restrict__ cbs,
                                          const float *__restrict__ e2s, float &best, int &bk)
{
    const float d = ref_dist(z, z2, cbs[k], e2s[k]);
    if (d < best) {  // lists ascend, so the lowest index wins ties (torch.argmin)
        best = d;
        bk = (int)k;
    }
}

__device__ __forceinline__ void eval_word(unsigned wd, const float z[4], float z2, const float4 *cbs, const float *e2s, float &best, int &bk)
{
    eval_cand(wd & 0xffffu, z, z2, cbs, e2s, best, bk);
    eval_cand(wd >> 16, z, z2, cbs, e2s, best, bk);
}

// One thread per ROW UNIT = 4 horizontally adjacent tokens (w % 4 == 0): four 16-byte loads (one per
// channel), up to four searches (a token bit-identical to its left neighbour reuses that result:
// coarse rows search once, medium rows twice), 16-byte stores of idx / z_q.
__global__ void __launch_bounds__(VQI_THREADS, 2)
vq_indexed_kernel(const float *__restrict__ z, int h, int w, int64_t n_units, const unsigned char *__restrict__ blob, int K,
                  int64_t *__restrict__ idx_out, float *__restrict__ zq_out, double *__restrict__ partials,
                  int32_t *__restrict__ counters, double *__restrict__ sqerr_out)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long mbar;
    __shared__ int s_qn;
    __shared__ double s_red[VQI_WARPS];
    __shared__ bool s_last;
    const CbLayout L = cb_layout(K);
    const CbHeader *hdr = reinterpret_cast<const CbHeader *>(smem);
    const unsigned char *lut = smem + L.lut;
    const float4 *cbs = reinterpret_cast<const float4 *>(smem + L.cb);
    const float *e2s = reinterpret_cast<const float *>(smem + L.e2);
    unsigned short *queue = reinterpret_cast<unsigned short *>(smem + ((L.stage + 15) & ~(size_t)15));  // [4 * VQI_THREADS] local token ids
    unsigned short *qres = queue + 4 * VQI_THREADS;                                                     // [4 * VQI_THREADS] their codes
    const uint4 *recs = reinterpret_cast<const uint4 *>(blob + L.rec);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    pdl_launch_dependents();
    pdl_wait();  // the blob and z may come straight from a preceding kernel
    if (tid == 0) {
        mbar_init(&mbar);
        s_qn = 0;
    }
    __syncthreads();
    if (tid == 0) tma_load_1d(smem, blob, (uint32_t)L.stage, &mbar);

    const int64_t plane = (int64_t)h * w;
    const int64_t upi = plane >> 2;  // units per image
    double sq = 0.0;
    bool staged = false;
    for (int64_t u0 = (int64_t)blockIdx.x * VQI_THREADS; u0 < n_units; u0 += (int64_t)gridDim.x * VQI_THREADS) {
        const int64_t u = u0 + tid;
        const bool active = u < n_units;
        float zt[4][4];  // [token][channel]
        int64_t b = 0, r = 0;
        if (active) {
            b = u / upi;
            r = u - b * upi;
            const float *zb = z + b * 4 * plane + 4 * r;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float4 v = __ldg(reinterpret_cast<const float4 *>(zb + c * plane));
                zt[0][c] = v.x;
                zt[1][c] = v.y;
                zt[2][c] = v.z;
                zt[3][c] = v.w;
            }
        }
        if (!staged) {
            staged = true;
            mbar_wait(&mbar, 0);
        }
        int code[4] = {0, 0, 0, 0};
        int qpos[4] = {-1, -1, -1, -1};
        if (active) {
            bool need[4];
            need[0] = true;
#pragma unroll
            for (int t = 1; t < 4; ++t) {
                bool same = true;
#pragma unroll
                for (int c = 0; c < 4; ++c) same &= __float_as_uint(zt[t][c]) == __float_as_uint(zt[t - 1][c]);
                need[t] = !same;
            }
            // cells and the first 32 bytes of their records (count + 15 candidates), all loads in flight together
            int cell[4];
            uint4 r0[4], r1[4];
            const bool valid = hdr->valid != 0;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                cell[t] = -1;
                if (need[t] && valid) {
                    int cidx = 0;
                    bool in = true;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int bn = cb_bin(zt[t][c], hdr->lo[c], hdr->inv[c]);
                        in &= bn >= 0;
                        cidx = cidx * CB_G + (int)lut[c * CB_NB + max(bn, 0)];
                    }
                    if (in) cell[t] = cidx;
                }
                if (cell[t] >= 0) {
                    const uint4 *rp = recs + (size_t)cell[t] * (CB_RW / 8);
                    r0[t] = __ldg(rp);
                    r1[t] = __ldg(rp + 1);
                }
            }
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                if (!need[t]) {
                    code[t] = code[t - (t > 0)];
                    qpos[t] = qpos[t - (t > 0)];
                    continue;
                }
                const unsigned count = cell[t] >= 0 ? (r0[t].x & 0xffffu) : 0xffffu;
                if (count == 0xffffu) {  // outside the grid / overflowing cell / no index: exhaustive search below
                    qpos[t] = atomicAdd(&s_qn, 1);
                    queue[qpos[t]] = (unsigned short)(tid * 4 + t);
                    continue;
                }
                const float z2 = sumsq4f(zt[t][0], zt[t][1], zt[t][2], zt[t][3]);
                float best = __int_as_float(0x7f800000);
                int bk = 0;
                eval_cand(r0[t].x >> 16, zt[t], z2, cbs, e2s, best, bk);
                eval_word(r0[t].y, zt[t], z2, cbs, e2s, best, bk);
                eval_word(r0[t].z, zt[t], z2, cbs, e2s, best, bk);
                eval_word(r0[t].w, zt[t], z2, cbs, e2s, best, bk);
                if (count > 7) {
                    eval_word(r1[t].x, zt[t], z2, cbs, e2s, best, bk);
                    eval_word(r1[t].y, zt[t], z2, cbs, e2s, best, bk);
                    eval_word(r1[t].z, zt[t], z2, cbs, e2s, best, bk);
                    eval_word(r1[t].w, zt[t], z2, cbs, e2s, best, bk);
                    const uint4 *rp = recs + (size_t)cell[t] * (CB_RW / 8);
                    for (unsigned pc = 2; pc * 8 < count + 1; ++pc) {
                        const uint4 q = __ldg(rp + pc);
                        eval_word(q.x, zt[t], z2, cbs, e2s, best, bk);
                        eval_word(q.y, zt[t], z2, cbs, e2s, best, bk);
                        eval_word(q.z, zt[t], z2, cbs, e2s, best, bk);
                        eval_word(q.w, zt[t], z2, cbs, e2s, best, bk);
                    }
                }
                code[t] = bk;
            }
        }
        // ---- exhaustive search of the queued tokens: one warp per token, lanes stride over the codes
        __syncthreads();
        const int qn = s_qn;
        if (qn > 0) {
            for (int q = warp; q < qn; q += VQI_WARPS) {
                const int id = queue[q];
                const int64_t uu = u0 + (id >> 2);
                const int64_t bb = uu / upi, rr = uu - bb * upi;
                const float *zp = z + bb * 4 * plane + 4 * rr + (id & 3);
                const float zz[4] = {__ldg(zp), __ldg(zp + plane), __ldg(zp + 2 * plane), __ldg(zp + 3 * plane)};
                const float z2 = sumsq4f(zz[0], zz[1], zz[2], zz[3]);
                float best = __int_as_float(0x7f800000);
                int bk = 0x7fffffff;
                for (int k = lane; k < K; k += 32) {
                    const float d = ref_dist(zz, z2, cbs[k], e2s[k]);
                    if (d < best) {
                        best = d;
                        bk = k;
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float od = __shfl_xor_sync(0xffffffffu, best, o);
                    const int ok = __shfl_xor_sync(0xffffffffu, bk, o);
                    if (od < best || (od == best && ok < bk)) {
                        best = od;
                        bk = ok;
                    }
                }
                if (lane == 0) qres[q] = (unsigned short)(bk == 0x7fffffff ? 0 : bk);  // no finite distance: index 0, as vq_assign.cu
            }
            __syncthreads();
#pragma unroll
            for (int t = 0; t < 4; ++t)
                if (qpos[t] >= 0) code[t] = qres[qpos[t]];
            __syncthreads();
            if (tid == 0) s_qn = 0;
        }
        // ---- finalize: idx, z_q = fl(z + fl(e - z)), sum (e - z)^2
        if (active) {
            int64_t *ip = idx_out + b * plane + 4 * r;
            reinterpret_cast<longlong2 *>(ip)[0] = make_longlong2(code[0], code[1]);
            reinterpret_cast<longlong2 *>(ip)[1] = make_longlong2(code[2], code[3]);
            if (zq_out || sqerr_out) {
                float q[4][4];
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const float4 e = cbs[code[t]];
                    const float d0 = __fsub_rn(e.x, zt[t][0]), d1 = __fsub_rn(e.y, zt[t][1]), d2 = __fsub_rn(e.z, zt[t][2]),
                                d3 = __fsub_rn(e.w, zt[t][3]);
                    q[t][0] = __fadd_rn(zt[t][0], d0);
                    q[t][1] = __fadd_rn(zt[t][1], d1);
                    q[t][2] = __fadd_rn(zt[t][2], d2);
                    q[t][3] = __fadd_rn(zt[t][3], d3);
                    float acc = __fmul_rn(d0, d0);
                    acc = __fmaf_rn(d1, d1, acc);
                    acc = __fmaf_rn(d2, d2, acc);
                    acc = __fmaf_rn(d3, d3, acc);
                    sq += (double)acc;
                }
                if (zq_out) {
                    float *qb = zq_out + b * 4 * plane + 4 * r;
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        *reinterpret_cast<float4 *>(qb + c * plane) = make_float4(q[0][c], q[1][c], q[2][c], q[3][c]);
                }
            }
        }
    }
    if (!staged) mbar_wait(&mbar, 0);  // never leave with the bulk copy in flight
    if (!sqerr_out) return;
    // deterministic reduction: warp shuffle -> CTA -> per-CTA partial -> the last CTA sums them in order
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    if (lane == 0) s_red[warp] = sq;
    __syncthreads();
    if (tid == 0) {
        double tot = 0.0;
        for (int i = 0; i < VQI_WARPS; ++i) tot += s_red[i];
        partials[blockIdx.x] = tot;
        __threadfence();
        s_last = (atomicAdd(&counters[0], 1) == (int)gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        double tot = 0.0;
        for (int i = tid; i < (int)gridDim.x; i += VQI_THREADS) tot += __ldcg(&partials[i]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
        if (lane == 0) s_red[warp] = tot;
        __syncthreads();
        if (tid == 0) {
            double all = 0.0;
            for (int i = 0; i < VQI_WARPS; ++i) all += s_red[i];
            *sqerr_out = all;
            counters[0] = 0;  // leave the ticket zeroed for the next launch (workspace contract)
        }
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
    if (e != cudaSuccess) {
        set_error("cgic_codebook_create: %s", cudaGetErrorString(e));
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
    delete cb;
}

extern "C" int cgic_codebook_update(cgic_codebook *cb, const float *codebook, cgic_stream_t stream_)
{
    CGIC_REQUIRE(cb && codebook, CGIC_EINVAL, "cgic_codebook_update: null argument");
    CGIC_REQUIRE((reinterpret_cast<uintptr_t>(codebook) & 15) == 0, CGIC_EINVAL, "cgic_codebook_update: codebook must be 16-byte aligned");
    cudaStream_t stream = as_stream(stream_);
    const int K = cb->K;
    const size_t smem = (size_t)K * (16 + 2);
    static bool attr_done = false;
    if (!attr_done) {
        CGIC_CUDA_CHECK(cudaFuncSetAttribute(cb_cells_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CB_MAX_K * (16 + 2)));
        attr_done = true;
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
    return CGIC_OK;
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
const float *codebook_rows(const cgic_codebook *cb) { return reinterpret_cast<const float *>(cb->blob + cb_layout(cb->K).cb); }
int codebook_size(const cgic_codebook *cb) { return cb->K; }
}  // namespace cgic

extern "C" int cgic_vq_assign_indexed(const float *z, int B, int h, int w, const cgic_codebook *cb, int64_t *idx_out, float *zq_out,
                                      double *sqerr_out, void *workspace, size_t workspace_bytes, cgic_stream_t stream_)
{
    CGIC_REQUIRE(z && cb && idx_out && workspace, CGIC_EINVAL, "cgic_vq_assign_indexed: null argument");
    CGIC_REQUIRE(cb->built, CGIC_EINVAL, "cgic_vq_assign_indexed: cgic_codebook_update has not been called");
    CGIC_REQUIRE(B >= 0 && h > 0 && w > 0, CGIC_EINVAL, "cgic_vq_assign_indexed: bad shape B=%d h=%d w=%d", B, h, w);
    const bool vec_ok = w % 4 == 0 && (reinterpret_cast<uintptr_t>(z) & 15) == 0 && (reinterpret_cast<uintptr_t>(idx_out) & 15) == 0 &&
                        (!zq_out || (reinterpret_cast<uintptr_t>(zq_out) & 15) == 0);
    if (!vec_ok)  // ragged rows: the generic exhaustive kernel (same results)
        return cgic_vq_assign(z, B, h, w, codebook_rows(cb), cb->K, idx_out, zq_out, sqerr_out, workspace, workspace_bytes, stream_);
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
    static int n_sm = 0;
    if (!n_sm) {
        int dev = 0, sm = 0;
        CGIC_CUDA_CHECK(cudaGetDevice(&dev));
        CGIC_CUDA_CHECK(cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev));
        CGIC_CUDA_CHECK(cudaFuncSetAttribute(vq_indexed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)(cb_layout(CB_MAX_K).stage + 16 + 4 * VQI_THREADS * 4)));
        n_sm = sm;
    }
    const CbLayout L = cb_layout(cb->K);
    const size_t smem = ((L.stage + 15) & ~(size_t)15) + (size_t)4 * VQI_THREADS * 4;
    const int64_t n_units = n / 4;
    const int64_t want = (n_units + VQI_THREADS - 1) / VQI_THREADS;
    const int grid = (int)(want < (int64_t)2 * n_sm ? want : (int64_t)2 * n_sm);
    {
        CGIC_PROF("vq_indexed_kernel", stream);
        CGIC_CUDA_CHECK(launch_pdl(vq_indexed_kernel, dim3(grid), dim3(VQI_THREADS), smem, stream, z, h, w, n_units,
                                   (const unsigned char *)cb->blob, cb->K, idx_out, zq_out, partials, counters, sqerr_out));
    }
    CGIC_LAUNCH_CHECK();
    return CGIC_OK;
}
