// Exact order-statistic select shared by the router kernels (router.cu) and the fused entropy + routing kernel (entropy.cu).
#pragma once
#include "common.cuh"

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

// value of rank `rank` (0-based) among fetch(0..n-1); whole CTA must call.  s_keys (nullable): room for n keys in
// shared memory -- the values are fetched and converted once instead of once per pass.
template <typename Fetch>
__device__ float select_rank(Fetch fetch, int64_t n, int64_t rank, uint32_t *s_hist, uint32_t *s_state, uint32_t *s_keys)
{
    uint32_t prefix = 0, mask = 0;
    if (rank > n - 1) rank = n - 1;
    if (rank < 0) rank = 0;
    if (threadIdx.x == 0) {
        s_state[0] = 0;
        s_state[1] = (uint32_t)rank;
        s_state[2] = (uint32_t)((uint64_t)rank >> 32);
    }
    if (s_keys)
        for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s_keys[i] = float_key(fetch(i));
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        for (int i = threadIdx.x; i < 256; i += blockDim.x) s_hist[i] = 0;
        __syncthreads();
        for (int64_t i0 = 0; i0 < n; i0 += blockDim.x) {  // uniform trip count: the warp votes below need every lane
            const int64_t i = i0 + threadIdx.x;
            uint32_t digit = 0xFFFFFFFFu;
            if (i < n) {
                const uint32_t k = s_keys ? s_keys[i] : float_key(fetch(i));
                if ((k & mask) == prefix) digit = (k >> shift) & 255u;
            }
            // one atomic per distinct digit and warp (entropies share their exponent: the first pass would otherwise
            // serialise hundreds of increments on one or two counters)
            const unsigned same = __match_any_sync(0xffffffffu, digit);
            if (digit != 0xFFFFFFFFu && (threadIdx.x & 31) == __ffs(same) - 1) atomicAdd(&s_hist[digit], (uint32_t)__popc(same));
        }
        __syncthreads();
        if (threadIdx.x < 32) {
            // the bucket holding the wanted rank: lane l owns buckets 8l .. 8l+7; inclusive scan of the lane totals by
            // shuffles, then the owning lane walks its 8 buckets (a serial walk over 256 buckets cost 4 us per pass)
            const int lane = threadIdx.x;
            uint64_t rk = ((uint64_t)s_state[2] << 32) | s_state[1];
            uint32_t cnt[8], tot = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                cnt[i] = s_hist[lane * 8 + i];
                tot += cnt[i];
            }
            uint32_t inc = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += v;
            }
            const uint32_t exc = inc - tot;
            // rank < n always, so exactly one lane has exc <= rk < inc (the last lane takes everything beyond: d <= 255)
            const bool mine = (rk >= exc && rk < inc) || (lane == 31 && rk >= inc);
            __syncwarp();  // every lane has read the rank before the owning lane replaces it
            if (mine) {
                uint64_t rem = rk - exc;
                uint32_t d = 0;
                for (; d < 7; ++d) {
                    if (rem < cnt[d]) break;
                    rem -= cnt[d];
                }
                s_state[0] = prefix | ((uint32_t)(lane * 8 + d) << shift);
                s_state[1] = (uint32_t)rem;
                s_state[2] = (uint32_t)(rem >> 32);
            }
        }
        __syncthreads();
        prefix = s_state[0];
        mask |= 0xFFu << shift;
        __syncthreads();
    }
    return key_float(prefix);
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
    if (use_c) {
        const float thr = select_rank([&](int64_t i) { return __ldcg(e16 + i); }, n16, k_c != 0 ? k_c - 1 : 0, s_hist, s_state, s_keys);
        const float tol = rtol * fabsf(thr) + atol;
        int cnt = 0;
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
    if (mode == 0 || mode == 1) {
        // mode 0: the entropy of cells under a coarse patch is zeroed before the sort (RouterTriple.py:27)
        const float thr = mode == 0 ? select_rank([&](int64_t i) { return __fmul_rn(__ldcg(e8 + i), __fsub_rn(1.0f, (float)c[parent((int)i)])); }, n8,
                                                  k_m != 0 ? k_m - 1 : 0, s_hist, s_state, s_keys)
                                    : select_rank([&](int64_t i) { return __ldcg(e8 + i); }, n8, k_m != 0 ? k_m - 1 : 0, s_hist, s_state, s_keys);
        const float tol = rtol * fabsf(thr) + atol;
        int cnt = 0;
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
