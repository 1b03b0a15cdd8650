// The warp-autonomous tile search shared by vq_warp_kernel (vq_assign.cu) and the fused small-grid encoder
// (encode_small_kernel, pack.cu): a WARP owns a tile of 4 rows x 32 columns of one image and runs
// load -> classify -> search -> finalize on its own, synchronising with __syncwarp only.
//   load      lane = column: 4 rows x 4 channels per lane, each warp load is one 128-byte row segment; the tile is kept in
//             the warp's shared-memory slice as one float4 per token (registers are not held across the search);
//   classify  a token whose 4 latent channels are bit-identical to the top-left token of its 4x4 (else 2x2) block is a
//             FOLLOWER of that token (the mask-mix of vqvae_blocks.py:364-366 makes coarse / medium regions constant
//             over blocks; the test is on the data itself, hence exact for ANY input); leaders are compacted with
//             ballots, no atomics;
//   search    one lane per leader: grid cell -> candidate record (4 x 16-byte loads) -> the reference's rounding
//             sequence on the candidates; leaders the index cannot serve are searched exhaustively by the warp;
//   finalize  idx / z_q row segments, sum((e - z)^2) per lane.
#pragma once
#include "codebook.cuh"

namespace cgic {
namespace {

constexpr int VQW_TILE = 128;  // tokens per warp tile
// per-warp shared memory: the tile (float4 per token), code of a leader token, compacted leaders, leader of every token
constexpr int VQW_TILE_BYTES = VQW_TILE * (16 + 2 + 1 + 1);
constexpr int VQW_REC_BYTES = VQW_TILE * 64;  // optional: the first 64 bytes of every leader's cell record, staged with cp.async

__device__ __forceinline__ float sumsq4(float a, float b, float c, float d) { return sumsq4f(a, b, c, d); }

// exact reference distance of one code (rounding sequence of quantize.py:73-75, see the header)
__device__ __forceinline__ void eval_cand(unsigned k, const float4 &zv, float z2, const float4 *__restrict__ cbs,
                                          const float *__restrict__ e2s, float &best, int &bk)
{
    const float4 e = cbs[k];
    float dot = __fmul_rn(zv.x, e.x);
    dot = __fmaf_rn(zv.y, e.y, dot);
    dot = __fmaf_rn(zv.z, e.z, dot);
    dot = __fmaf_rn(zv.w, e.w, dot);
    const float d = __fmaf_rn(dot, -2.f, __fadd_rn(z2, e2s[k]));
    if (d < best) {  // candidate lists ascend, so the lowest index wins ties (torch.argmin)
        best = d;
        bk = (int)k;
    }
}
__device__ __forceinline__ void eval_word(unsigned wd, const float4 &zv, float z2, const float4 *cbs, const float *e2s, float &best, int &bk)
{
    eval_cand(wd & 0xffffu, zv, z2, cbs, e2s, best, bk);
    eval_cand(wd >> 16, zv, z2, cbs, e2s, best, bk);
}
__device__ __forceinline__ void eval_piece(const uint4 &q, const float4 &zv, float z2, const float4 *cbs, const float *e2s, float &best, int &bk)
{
    eval_word(q.x, zv, z2, cbs, e2s, best, bk);
    eval_word(q.y, zv, z2, cbs, e2s, best, bk);
    eval_word(q.z, zv, z2, cbs, e2s, best, bk);
    eval_word(q.w, zv, z2, cbs, e2s, best, bk);
}

// Leaders the index could not serve (bits of `todo` = positions in list[]): the WARP searches all K codes, up to four
// leaders per sweep (every code row is loaded once for all of them; lane k, k + 32, ...; per-lane ascending order + an
// index tie-break in the reduction = the lowest index among equal minima; no finite distance at all leaves index 0,
// like vq_fused_kernel).  Out of line: its registers must not weigh on the indexed path.
__device__ __noinline__ void vq_exhaustive_sweep(unsigned todo, const uint8_t *list, const float4 *zs, uint16_t *res, const float4 *cbs,
                                          const float *e2s, int K, int lane)
{
    while (todo) {
        // up to four leaders per sweep over the codebook: every code row is loaded once for all of them
        int tq[4];
        float4 vq[4];
        float z2q[4], bdq[4];
        int bkq[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            tq[i] = -1;
            if (todo) {
                tq[i] = list[__ffs(todo) - 1];
                todo &= todo - 1;
            }
            vq[i] = zs[tq[i] >= 0 ? tq[i] : 0];
            z2q[i] = sumsq4(vq[i].x, vq[i].y, vq[i].z, vq[i].w);
            bdq[i] = __int_as_float(0x7f800000);
            bkq[i] = 0x7fffffff;
        }
        for (int k = lane; k < K; k += 32) {
            const float4 e = cbs[k];
            const float e2 = e2s[k];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float dot = __fmul_rn(vq[i].x, e.x);
                dot = __fmaf_rn(vq[i].y, e.y, dot);
                dot = __fmaf_rn(vq[i].z, e.z, dot);
                dot = __fmaf_rn(vq[i].w, e.w, dot);
                const float d = __fmaf_rn(dot, -2.f, __fadd_rn(z2q[i], e2));
                if (d < bdq[i]) {
                    bdq[i] = d;
                    bkq[i] = k;
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float bd = bdq[i];
            int bk = bkq[i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float od = __shfl_xor_sync(0xffffffffu, bd, o);
                const int ok = __shfl_xor_sync(0xffffffffu, bk, o);
                if (od < bd || (od == bd && ok < bk)) {
                    bd = od;
                    bk = ok;
                }
            }
            if (lane == 0 && tq[i] >= 0) res[tq[i]] = (uint16_t)(bk == 0x7fffffff ? 0 : bk);
        }
    }
}

struct VqTileCtx {
    const CbHeader *hdr;         // staged head of the prepared codebook (shared memory)
    const unsigned char *lut;    // bin -> cell
    const float4 *cbs;           // codebook rows
    const float *e2s;            // |e|^2
    const uint4 *recs;           // cell records (global)
    int K;
    float4 *zs;                  // [128] the warp's tile, row-major
    uint16_t *res;               // [128] code of a leader token
    uint8_t *list;               // [128] leader tokens, compacted
    uint8_t *lead;               // [128] leader of every token
    uint4 *rec_s;                // [128][4] staged head of every leader's record, or null (records are then read from global)
};

// One tile: rows gy0 .. gy0+3, column gx = tile column 0 + lane of an h x w token grid whose channel planes start at zb
// (NCHW, plane = h*w).  idx_img / zq_img: the image's outputs; idx_s (nullable): the image's indices as u16 in shared
// memory (fused encoder).  exhaustive_count (nullable): leaders that took the exhaustive path are added there.
// `staged()` is called once the loads are in flight: it waits for the codebook's bulk copy on first use.
struct VqNoStamp {
    __device__ __forceinline__ void operator()(int) const {}
};
// `stamp(k)` (tracing builds): k = 2 tile loaded and classified, 3 cells found, 4 leaders searched.
template <typename Staged, typename Stamp = VqNoStamp>
__device__ __forceinline__ void vq_process_tile(const VqTileCtx &c, const float *__restrict__ zb, int h, int w, int gy0, int gx, int lane,
                                                int64_t *__restrict__ idx_img, float *__restrict__ zq_img, bool want_sq, double &sq,
                                                uint16_t *idx_s, int32_t *exhaustive_count, Staged staged, Stamp stamp = Stamp())
{
    const int64_t plane = (int64_t)h * w;
    const bool col_ok = gx < w;
    float4 *zs = c.zs;
    uint16_t *res = c.res;
    uint8_t *list = c.list, *lead = c.lead;
    // ---- load
    {
        float zt[4][4];  // [row][channel]
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const bool ok = col_ok && gy0 + r < h;
            const int64_t p = (int64_t)(gy0 + r) * w + gx;
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) zt[r][ch] = ok ? __ldg(zb + ch * plane + p) : 0.f;
        }
        __syncwarp();  // the previous tile's slice is no longer read
#pragma unroll
        for (int r = 0; r < 4; ++r) zs[r * 32 + lane] = make_float4(zt[r][0], zt[r][1], zt[r][2], zt[r][3]);
    }
    stamp(7);
    staged();
    stamp(1);
    __syncwarp();
    // ---- classify + compact (branch-free: the seven vectors a lane compares are loaded up front)
    int nlead = 0;
    unsigned lmask[4];
    {
        const uint4 *zs4 = reinterpret_cast<const uint4 *>(zs);
        const int t4 = lane & ~3;  // row 0 of the tile is the top row of every 4x4 block
        const uint4 top4 = zs4[t4], top2a = zs4[lane & ~1], top2b = zs4[64 + (lane & ~1)];
        uint4 own[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) own[r] = zs4[r * 32 + lane];
        auto same = [](const uint4 &u, const uint4 &v) { return u.x == v.x && u.y == v.y && u.z == v.z && u.w == v.w; };
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int t = r * 32 + lane;
            const bool ok = col_ok && gy0 + r < h;
            const int t2 = (r & ~1) * 32 + (lane & ~1);
            const bool eq4 = t4 != t && same(top4, own[r]);
            const bool eq2 = t2 != t && same(r < 2 ? top2a : top2b, own[r]);
            const int ld = !ok ? t : (eq4 ? t4 : (eq2 ? t2 : t));
            lead[t] = (uint8_t)ld;
            lmask[r] = __ballot_sync(0xffffffffu, ok && ld == t);
        }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        if ((lmask[r] >> lane) & 1u) list[nlead + __popc(lmask[r] & ((1u << lane) - 1u))] = (uint8_t)(r * 32 + lane);
        nlead += __popc(lmask[r]);
    }
    __syncwarp();
    stamp(2);
    // ---- search: one lane per leader.  Pass 1 finds every leader's grid cell and prefetches its record, so that
    //      the memory latency of all rounds overlaps; pass 2 evaluates.
    const bool usable = c.hdr->valid != 0;
    for (int j = lane; j < nlead; j += 32) {
        const int t = list[j];
        const float4 v = zs[t];
        int cell = 0xffff;
        if (usable) {
            const int b0 = cb_bin(v.x, c.hdr->lo[0], c.hdr->inv[0]), b1 = cb_bin(v.y, c.hdr->lo[1], c.hdr->inv[1]),
                      b2 = cb_bin(v.z, c.hdr->lo[2], c.hdr->inv[2]), b3 = cb_bin(v.w, c.hdr->lo[3], c.hdr->inv[3]);
            if ((b0 | b1 | b2 | b3) >= 0) {
                cell = (((int)c.lut[b0] * CB_G + (int)c.lut[CB_NB + b1]) * CB_G + (int)c.lut[2 * CB_NB + b2]) * CB_G + (int)c.lut[3 * CB_NB + b3];
                const uint4 *rp = c.recs + (size_t)cell * (CB_RW / 8);
                if (c.rec_s) {
                    // no registers held, one wait for all rounds: pass 2 reads the records from shared memory
                    // piece i of leader j sits in slot i ^ ((j >> 1) & 3) of its 64-byte row: the eight lanes of a quarter warp,
                    // which read piece i of eight consecutive leaders together, then touch eight different 16-byte bank groups
                    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(c.rec_s + j * 4);
                    const int sw = (j >> 1) & 3;
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16 * (i ^ sw)), "l"(rp + i) : "memory");
                } else {
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(rp));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + 2));
                }
            }
        }
        res[t] = (uint16_t)cell;  // parked here until pass 2 overwrites it with the code
    }
    if (c.rec_s) asm volatile("cp.async.wait_all;" ::: "memory");
    __syncwarp();
    stamp(3);
    bool unserved = false;  // some leader of this lane must be searched exhaustively
    for (int j = lane; j < nlead; j += 32) {
        const int t = list[j];
        const float4 v = zs[t];
        const int cell = res[t] == 0xffffu ? -1 : (int)res[t];
        unsigned count = 0xffffu;
        uint4 q0, q1, q2, q3;
        const uint4 *rp = c.recs + (size_t)max(cell, 0) * (CB_RW / 8);
        if (cell >= 0) {
            if (c.rec_s) {
                const uint4 *rs = c.rec_s + j * 4;
                const int sw = (j >> 1) & 3;
                q0 = rs[sw];
                q1 = rs[1 ^ sw];
                q2 = rs[2 ^ sw];
                q3 = rs[3 ^ sw];
            } else {
                q0 = __ldg(rp);
                q1 = __ldg(rp + 1);
                q2 = __ldg(rp + 2);
                q3 = __ldg(rp + 3);
            }
            count = q0.x & 0xffffu;
        }
        const float z2 = sumsq4(v.x, v.y, v.z, v.w);
        float bd = __int_as_float(0x7f800000);
        int bk = 0;
        if (count != 0xffffu) {
            // entries past `count` repeat the last candidate, so whole 16-byte pieces are evaluated
            eval_cand(q0.x >> 16, v, z2, c.cbs, c.e2s, bd, bk);
            eval_word(q0.y, v, z2, c.cbs, c.e2s, bd, bk);
            eval_word(q0.z, v, z2, c.cbs, c.e2s, bd, bk);
            eval_word(q0.w, v, z2, c.cbs, c.e2s, bd, bk);
            if (count > 7) eval_piece(q1, v, z2, c.cbs, c.e2s, bd, bk);
            if (count > 15) eval_piece(q2, v, z2, c.cbs, c.e2s, bd, bk);
            if (count > 23) eval_piece(q3, v, z2, c.cbs, c.e2s, bd, bk);
            for (unsigned pc = 4; pc * 8 < count + 1; ++pc) eval_piece(__ldg(rp + pc), v, z2, c.cbs, c.e2s, bd, bk);
        } else {
            bk = 0xffff;  // outside the grid / overflowing cell / no usable index: searched exhaustively by the whole warp below
            unserved = true;
        }
        res[t] = (uint16_t)bk;
    }
    __syncwarp();
    stamp(4);
    // ---- leaders the index could not serve: the WARP searches all K codes (lane k, k + 32, ...; per-lane ascending order +
    //      an index tie-break in the reduction = the lowest index among equal minima; no finite distance at all leaves
    //      index 0, like vq_fused_kernel)
    for (int j0 = 0; __any_sync(0xffffffffu, unserved) && j0 < nlead; j0 += 32) {
        const int j = j0 + lane;
        unsigned todo = __ballot_sync(0xffffffffu, j < nlead && res[list[j < nlead ? j : 0]] == 0xffffu);
        if (todo) {
            if (exhaustive_count && lane == 0) atomicAdd(exhaustive_count, __popc(todo));
            vq_exhaustive_sweep(todo, list + j0, zs, res, c.cbs, c.e2s, c.K, lane);  // rare: kept out of line
        }
    }
    __syncwarp();
    // ---- finalize
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        if (!(col_ok && gy0 + r < h)) continue;
        const int k = res[lead[r * 32 + lane]];
        const int64_t p = (int64_t)(gy0 + r) * w + gx;
        idx_img[p] = k;
        if (idx_s) idx_s[p] = (uint16_t)k;
        if (zq_img || want_sq) {
            const float4 e = c.cbs[k];
            const float4 zv = zs[r * 32 + lane];
            const float d0 = __fsub_rn(e.x, zv.x), d1 = __fsub_rn(e.y, zv.y), d2 = __fsub_rn(e.z, zv.z), d3 = __fsub_rn(e.w, zv.w);
            if (zq_img) {
                float *q = zq_img + p;
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

}  // namespace
}  // namespace cgic
