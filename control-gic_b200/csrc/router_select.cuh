// Exact order-statistic select shared by the router kernels (router.cu) and the fused entropy + routing kernel (entropy.cu).
#pragma once
#include "common.cuh"

#ifndef RS_STAMP
#define RS_STAMP(k)  // (tracing builds of entropy.cu stamp the phases of the routing tail)
#endif

namespace cgic {
namespace {

__device__ __forceinline__ uint32_t float_key(float v)
{
    if (v != v) return 0xFFFFFFFFu;  // torch.sort places NaN last
    const uint32_t u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_float(uint32_t k)
{
    if (k == 0xFFFFFFFFu) return __int_as_float(0x7fc00000);
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}

// Exact select: the value of rank `rank` (0-based, ascending) among fetch(0..n-1); whole CTA must call (any block size that
// is a multiple of 32, at most 1024 threads).  Order-preserving integer keys; s_keys (nullable): room for n keys in shared
// memory -- the values are then fetched and converted once.
//   1. block min / max of the keys: only the bits in which they differ matter (entropies share sign and most of the exponent);
//   2. a histogram over the top RS_BITS of the remaining range [lo, lo + 2^bits): one shared-memory atomicAdd per key;
//   3. the bucket that holds the rank (block scan of the histogram);
//   4. with n keys over 2048 buckets that bucket usually holds a handful of keys: up to 32 are collected and one warp ranks
//      them directly; a fuller bucket (ties, clusters) goes round again with the bucket as the new range.
// s_hist: RS_BINS words, s_state: RS_STATE words.  Five block barriers in the common case (one round): the histogram is zeroed
// before the keys are fetched, every thread derives the range from the per-warp minima / maxima itself, and every warp ranks
// the final handful of keys for itself.
constexpr int RS_BITS = 11;
constexpr int RS_BINS = 1 << RS_BITS;
constexpr int RS_STATE = 176;  // [0] lo, [1] bits, [2] rank in range, [3] keys in bucket, [4] small-list counter, [8..39] small list, [48..79] per-warp min, [80..111] per-warp max, [112..143] per-warp min over keys > key(+0), [144..175] per-warp count of keys <= key(+0)

template <typename Fetch>
__device__ float select_rank(Fetch fetch, int64_t n, int64_t rank, uint32_t *s_hist, uint32_t *s_state, uint32_t *s_keys)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthreads = blockDim.x, nwarps = nthreads >> 5;
    if (rank > n - 1) rank = n - 1;
    if (rank < 0) rank = 0;
    auto key_at = [&](int64_t i) { return s_keys ? s_keys[i] : float_key(fetch(i)); };
    // (a previous select's readers of s_hist are behind its last barriers; its answer, read after them, lives in a register)
    for (int i = tid; i < RS_BINS; i += nthreads) s_hist[i] = 0u;
    // ---- keys, block min / max; the values <= +0 are counted apart: order-preserving keys are dense in the exponent, so a
    //      range that reaches from 0 up to the entropies (the router zeroes the cells under a coarse patch) would spend almost
    //      all histogram bins on the empty stretch in between and need two or three rounds
    constexpr uint32_t KZ = 0x80000000u;  // float_key(+0.0f)
    uint32_t kmin = 0xFFFFFFFFu, kmax = 0u, kminp = 0xFFFFFFFFu, nle = 0u;
#pragma unroll 4  // (independent global loads: several in flight)
    for (int64_t i = tid; i < n; i += nthreads) {
        const uint32_t k = float_key(fetch(i));
        if (s_keys) s_keys[i] = k;
        kmin = min(kmin, k);
        kmax = max(kmax, k);
        if (k > KZ) kminp = min(kminp, k);
        else ++nle;
    }
    kmin = __reduce_min_sync(0xffffffffu, kmin);
    kmax = __reduce_max_sync(0xffffffffu, kmax);
    kminp = __reduce_min_sync(0xffffffffu, kminp);
    nle = __reduce_add_sync(0xffffffffu, nle);
    if (lane == 0) {
        s_state[48 + warp] = kmin;
        s_state[80 + warp] = kmax;
        s_state[112 + warp] = kminp;
        s_state[144 + warp] = nle;
    }
    if (tid == 0) s_state[4] = 0u;
    __syncthreads();
    // the answer lies in [lo, lo + 2^bits) at rank r of that range: derived by every warp for itself
    uint32_t lo, bits, r = (uint32_t)rank;
    {
        uint32_t a = lane < nwarps ? s_state[48 + lane] : 0xFFFFFFFFu, b = lane < nwarps ? s_state[80 + lane] : 0u;
        uint32_t ap = lane < nwarps ? s_state[112 + lane] : 0xFFFFFFFFu, c = lane < nwarps ? s_state[144 + lane] : 0u;
        a = __reduce_min_sync(0xffffffffu, a);
        b = __reduce_max_sync(0xffffffffu, b);
        ap = __reduce_min_sync(0xffffffffu, ap);
        c = __reduce_add_sync(0xffffffffu, c);
        if (c != 0u && (int64_t)c != n) {
            if (r >= c) {  // among the positive values: the c smaller keys lie below the range
                a = ap;
                r -= c;
            } else {       // among the values <= +0: the range starts at the global minimum, so ranks in it are global ranks
                b = KZ;
            }
        }
        lo = a;
        bits = a == b ? 0u : 32u - (uint32_t)__clz((int)(b - a));
    }
    for (bool first = true;; first = false) {
        if (bits == 0) return key_float(lo);
        const uint32_t shift = bits > (uint32_t)RS_BITS ? bits - RS_BITS : 0u;
        // a key belongs to the current range iff (k - lo) >> bits == 0 (and k >= lo); bits == 32 only in the first round: all keys
        auto in_range = [&](uint32_t k) { return k >= lo && (bits >= 32u || ((k - lo) >> bits) == 0u); };
        if (!first) {  // (the first round's histogram was zeroed before the keys were fetched)
            for (int i = tid; i < RS_BINS; i += nthreads) s_hist[i] = 0u;
            if (tid == 0) s_state[4] = 0u;
            __syncthreads();
        }
        for (int64_t i = tid; i < n; i += nthreads) {
            const uint32_t k = key_at(i);
            if (in_range(k)) atomicAdd(&s_hist[(k - lo) >> shift], 1u);
        }
        __syncthreads();
        // ---- the bucket holding rank r: every thread sums its slice of bins, block scan of the slice sums, the owner walks its slice
        const int per = (RS_BINS + nthreads - 1) / nthreads, b0 = tid * per, b1 = min(RS_BINS, b0 + per);
        uint32_t mine = 0;
        for (int i = b0; i < b1; ++i) mine += s_hist[i];
        uint32_t inc = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += v;
        }
        __shared__ uint32_t s_wtot[32];
        if (lane == 31) s_wtot[warp] = inc;
        __syncthreads();
        uint32_t woff = 0;
        for (int w = 0; w < warp; ++w) woff += s_wtot[w];
        const uint32_t exc = woff + inc - mine;
        if (r >= exc && r < exc + mine) {  // exactly one thread (r < number of keys in range)
            uint32_t rem = r - exc;
            int i = b0;
            for (; i < b1 - 1; ++i) {
                if (rem < s_hist[i]) break;
                rem -= s_hist[i];
            }
            s_state[0] = lo + ((uint32_t)i << shift);
            s_state[1] = shift;
            s_state[2] = rem;
            s_state[3] = s_hist[i];
        }
        __syncthreads();
        lo = s_state[0];
        bits = s_state[1];
        r = s_state[2];
        const uint32_t cnt = s_state[3];
        if (bits == 0) return key_float(lo);
        if (cnt > 32u) {  // a full bucket: another round over [lo, lo + 2^bits)
            __syncthreads();  // (everybody has read the state before the next round rewrites it)
            continue;
        }
        // ---- a handful of keys left: collect them, every warp ranks them for itself
        for (int64_t i = tid; i < n; i += nthreads) {
            const uint32_t k = key_at(i);
            if (k >= lo && ((k - lo) >> bits) == 0u) s_state[8 + atomicAdd(&s_state[4], 1u)] = k;
        }
        __syncthreads();
        const uint32_t k = lane < cnt ? s_state[8 + lane] : 0xFFFFFFFFu;
        uint32_t less = 0, leq = 0;
        for (uint32_t j = 0; j < cnt; ++j) {
            const uint32_t o = s_state[8 + j];
            less += o < k;
            leq += o <= k;
        }
        const unsigned hit = __ballot_sync(0xffffffffu, lane < cnt && less <= r && r < leq);  // (equal keys: any of them)
        return key_float(__shfl_sync(0xffffffffu, k, __ffs(hit) - 1));
    }
}

// TripleGrainFixedEntropyRouter.forward for ONE image by the whole CTA (RouterTriple.py:15-96, per-image thresholds =
// the reference's B == 1 call): thresholds by exact select, then m_c [n16] and m_m [4 n16] (global, int32).  e16 / e8 are
// read with ld.global.cg (they may just have been written by other CTAs).  near[0..1] += entropies within
// rtol * |thr| + atol of the coarse / medium threshold (the cells a float-tolerance difference in Entropy could flip).
// s_keys (nullable): room for 4 n16 keys.
__device__ __forceinline__ void route_image(const float *e16, const float *e8, int h16, int w16, int mode, int64_t k_c, int64_t k_m, int32_t *c,
                                            int32_t *m, int32_t *near, float rtol, float atol, uint32_t *s_hist, uint32_t *s_state, uint32_t *s_keys)
{
    const int n16 = h16 * w16, n8 = 4 * n16, w8 = 2 * w16;
    const bool use_c = mode == 0 || mode == 2 || mode == 3;
    auto parent = [&](int i) { return ((i / w8) >> 1) * w16 + ((i % w8) >> 1); };  // coarse cell above medium cell i
    RS_STAMP(1);
    if (use_c) {
        const float thr = select_rank([&](int64_t i) { return __ldcg(e16 + i); }, n16, k_c != 0 ? k_c - 1 : 0, s_hist, s_state, s_keys);
        RS_STAMP(2);
        const float tol = rtol * fabsf(thr) + atol;
        int cnt = 0;
#pragma unroll 4
        for (int i = threadIdx.x; i < n16; i += blockDim.x) {
            const float v = __ldcg(e16 + i);
            c[i] = v < thr;
            cnt += fabsf(v - thr) <= tol;
        }
        if (near && cnt) atomicAdd(&near[0], cnt);
    } else {
        for (int i = threadIdx.x; i < n16; i += blockDim.x) c[i] = (mode == 4);
    }
    __syncthreads();  // c[] (global) is re-read below by other threads of this CTA
    RS_STAMP(3);
    if (mode == 0 || mode == 1) {
        // mode 0: the entropy of cells under a coarse patch is zeroed before the sort (RouterTriple.py:27)
        const float thr = mode == 0 ? select_rank([&](int64_t i) { return __fmul_rn(__ldcg(e8 + i), __fsub_rn(1.0f, (float)c[parent((int)i)])); }, n8,
                                                  k_m != 0 ? k_m - 1 : 0, s_hist, s_state, s_keys)
                                    : select_rank([&](int64_t i) { return __ldcg(e8 + i); }, n8, k_m != 0 ? k_m - 1 : 0, s_hist, s_state, s_keys);
        RS_STAMP(4);
        const float tol = rtol * fabsf(thr) + atol;
        int cnt = 0;
#pragma unroll 4
        for (int i = threadIdx.x; i < n8; i += blockDim.x) {
            const float v = __ldcg(e8 + i);
            const bool under = mode == 0 && c[parent(i)];
            m[i] = (v < thr) && !under;
            cnt += !under && fabsf(v - thr) <= tol;
        }
        if (near && cnt) atomicAdd(&near[1], cnt);
    } else if (mode == 3) {
        for (int i = threadIdx.x; i < n8; i += blockDim.x) m[i] = 1 - c[parent(i)];
    } else {
        for (int i = threadIdx.x; i < n8; i += blockDim.x) m[i] = (mode == 5);
    }
}

}  // namespace
}  // namespace cgic
