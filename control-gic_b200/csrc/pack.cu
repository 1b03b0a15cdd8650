// a7 + a9 + a11 + a12: granularity index selection and the five-stream bit packer.
//
//   reference: CGIC/models/model.py:217-260 (selection, which streams exist per mode, bpp),
//              HuffmanCoding.compress  CGIC/tools/indices_coding.py:113-126 (78-82, 91-98, 101-110),
//              BinaryCoding.compress   CGIC/tools/mask_coding.py:40-55.
//
// One CTA per (index stream, image); the coarse stream's CTA also packs the two mask streams.  An index stream walks its level's grid in row-major order in
// tiles of 1024 positions: mask test + symbol fetch (the block's top-left token), a block-wide
// exclusive scan of the code lengths gives every symbol its bit offset, the codes are OR-ed into
// a shared-memory staging window (atomicOr on 32-bit words), and complete words leave as
// coalesced big-endian 32-bit stores.  The partial last word is carried into the next tile, so
// global memory sees every output word exactly once and needs no pre-zeroing.
// Framing (must be byte exact): [pad count byte][payload, MSB first][pad zero bits],
// pad = 8 - nbits % 8 in 1..8, empty symbol list -> 0 bytes.
#include <algorithm>

#include "vq_tile.cuh"

namespace cgic {
namespace {

CGIC_TRACE_DECL(pack)


constexpr int PK_THREADS = 512;

struct PackArgs {
    const int64_t *idx;  // [B,h,w]  (or the symbol list for the single-stream entry point)
    const int32_t *mask[3];
    int h, w, mode;
    int64_t n_direct;  // >= 0: single stream of n_direct symbols taken from idx, no mask
    DevTable T;
    uint8_t *out;
    int64_t image_stride;
    int64_t slot_off[5];
    int64_t slot_cap[5];
    int32_t *sizes;
    // large token grids: the tiles of a stream are dealt over `nslots` CTAs and chained through two self-clearing
    // 64-bit records per tile in the workspace (bit position / carry word), see pack_index_stream_chained
    unsigned long long *chain;  // [B][3] x (max_tiles x max_tiles + max_tiles) records
    int max_tiles, nslots;
    int gather;                 // 1: every tile publishes its bit count to all later tiles (long chains), 0: tile-to-tile chain
    // cgic_encode: the per-CTA partial sums of (e - z)^2 the search kernel left behind (its deferred reduction); the mask CTA of
    // image 0 adds them in CTA order and writes *sq_out
    const double *sq_partials;
    int sq_n;
    double *sq_out;
};

// deterministic sum of n doubles by the whole CTA (thread t takes t, t + T, ...; lanes by xor shuffles; warps in order)
__device__ void reduce_partials(const double *partials, int n, double *out)
{
    __shared__ double s_part[PK_THREADS / 32];
    double tot = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) tot += __ldcg(&partials[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = tot;
    __syncthreads();
    if (threadIdx.x == 0) {
        double all = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) all += s_part[i];
        *out = all;
    }
}

__device__ __forceinline__ uint32_t to_big_endian(uint32_t v) { return __byte_perm(v, 0, 0x0123); }

// block-wide exclusive scan of one int per thread; returns the exclusive prefix, total in *total
__device__ __forceinline__ int block_exscan(int v, int *s_warp, int *total)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
    }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    int woff = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < PK_THREADS / 32; ++i) {
        const int t = s_warp[i];
        if (i < wid) woff += t;
        tot += t;
    }
    *total = tot;
    __syncthreads();
    return woff + inc - v;
}

// Codes of one thread's 8 consecutive positions of a level grid (lean path: 32-bit arithmetic, shifts for the level's step,
// vector loads).  Requires (grid width of the level) % 8 == 0 -- a thread's positions then lie in one grid row, 16-byte aligned --
// and the (code, length) table in shared memory.  pos0 >= n_pos: nothing.  Returns the bits; *bad |= symbol outside the table.
struct LevelGrid {
    int sh, gw, n_pos;             // log2(step), level grid width, level grid cells
    const int32_t *mask;           // [n_pos] of this image
    const int64_t *src;            // [h * w] of this image
    int w;                         // fine grid width
};
__device__ __forceinline__ LevelGrid level_grid(const PackArgs &a, int s, int b)
{
    LevelGrid g;
    g.sh = 2 - s;
    g.gw = a.w >> g.sh;
    g.n_pos = (a.h >> g.sh) * g.gw;
    g.mask = a.mask[s] + (size_t)b * g.n_pos;
    g.src = a.idx + (size_t)b * a.h * a.w;
    g.w = a.w;
    return g;
}
__device__ __forceinline__ bool level_grid_fast(const PackArgs &a, int s) { return a.n_direct < 0 && a.T.enc != nullptr && ((a.w >> (2 - s)) & 7) == 0; }

__device__ __forceinline__ void load_tile_inputs(const LevelGrid &g, int pos0, int4 &m0, int4 &m1, long long (&v)[8])
{
    m0 = make_int4(0, 0, 0, 0);
    m1 = m0;
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 0;
    if (pos0 >= g.n_pos) return;
    m0 = __ldg(reinterpret_cast<const int4 *>(g.mask + pos0));
    m1 = __ldg(reinterpret_cast<const int4 *>(g.mask + pos0) + 1);
    const int y = pos0 / g.gw, x = pos0 - y * g.gw;
    const long long *p = reinterpret_cast<const long long *>(g.src) + (size_t)(y << g.sh) * g.w + (x << g.sh);
    if (g.sh == 0) {  // fine level: the symbols are the 8 consecutive tokens themselves
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const longlong2 t = __ldg(reinterpret_cast<const longlong2 *>(p) + i);
            v[2 * i] = t.x;
            v[2 * i + 1] = t.y;
        }
    } else {          // the block's top-left token (model.py:219-221)
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = __ldg(p + (i << g.sh));
    }
}
__device__ __forceinline__ int tile_codes(const int4 &m0, const int4 &m1, const long long (&v)[8], const uint2 *s_enc, int K, uint32_t (&code)[8],
                                          uint32_t (&len)[8], int *bad)
{
    const int on[8] = {m0.x == 1, m0.y == 1, m0.z == 1, m0.w == 1, m1.x == 1, m1.y == 1, m1.z == 1, m1.w == 1};
    int tsum = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        code[i] = 0;
        len[i] = 0;
        if (on[i]) {
            if ((unsigned long long)v[i] >= (unsigned long long)K) {
                *bad = 1;
            } else {
                const uint2 e = s_enc[(int)v[i]];
                code[i] = e.x;
                len[i] = e.y;
            }
        }
        tsum += (int)len[i];
    }
    return tsum;
}

// One index stream by the whole CTA.  ITEMS consecutive grid positions per thread and tile; the
// code table is read from shared memory (s_enc: (code bits left aligned, length) per symbol) when
// every code fits 32 bits, else from the global code pool.  `stage` holds one tile of output bits.
template <int ITEMS>
__device__ void pack_index_stream(const PackArgs &a, int s, int b, uint32_t *stage, const uint2 *s_enc, unsigned long long *mbar)
{
    constexpr int TILE = PK_THREADS * ITEMS;
    __shared__ int s_warp[PK_THREADS / 32];
    __shared__ uint32_t s_carry;
    __shared__ int s_bad;
    const int tid = threadIdx.x;
    int gw, step;
    int64_t n_pos;
    const int32_t *mask = nullptr;
    const int64_t *src;
    uint8_t *out;
    int32_t *size_out;
    int64_t cap;
    if (a.n_direct >= 0) {
        gw = 1;
        step = 1;
        n_pos = a.n_direct;
        src = a.idx;
        out = a.out;
        size_out = a.sizes;
        cap = a.slot_cap[0];
    } else {
        step = 4 >> s;
        gw = a.w / step;
        n_pos = (int64_t)(a.h / step) * gw;
        mask = a.mask[s] + (int64_t)b * n_pos;
        src = a.idx + (int64_t)b * a.h * a.w;
        out = a.out + (int64_t)b * a.image_stride + a.slot_off[s];
        size_out = a.sizes + b * 5 + s;
        cap = a.slot_cap[s];
    }
    uint32_t *out32 = reinterpret_cast<uint32_t *>(out);
    if (tid == 0) {
        s_carry = 0;
        s_bad = 0;
    }
    __syncthreads();
    int64_t P = 8;  // stream bit position (header byte first)
    const bool fast = ITEMS == 8 && level_grid_fast(a, s);
    const LevelGrid lg = fast ? level_grid(a, s, b) : LevelGrid{};
    for (int64_t tile = 0; tile < n_pos; tile += TILE) {
        int sym[ITEMS];
        uint32_t len[ITEMS], code[ITEMS];
        const int64_t pos0 = tile + (int64_t)tid * ITEMS;
        int tsum = 0;
        if (ITEMS == 8 && fast) {
            int4 m0, m1;
            long long v[8];
            load_tile_inputs(lg, (int)pos0, m0, m1, v);
            if (mbar) {  // the code table's bulk copy was started before the loads above; first use is below
                mbar_wait(mbar, 0);
                mbar = nullptr;
            }
            int bad = 0;
            tsum = tile_codes(m0, m1, v, s_enc, a.T.K, reinterpret_cast<uint32_t (&)[8]>(code), reinterpret_cast<uint32_t (&)[8]>(len), &bad);
            if (bad) s_bad = 1;
        } else {
        // ---- which positions emit a symbol (vector loads of the mask where possible)
        bool on[ITEMS];
        if (!mask) {
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) on[i] = pos0 + i < n_pos;
        } else if (ITEMS % 4 == 0 && pos0 + ITEMS <= n_pos && (reinterpret_cast<uintptr_t>(mask + pos0) & 15) == 0) {
#pragma unroll
            for (int v = 0; v < ITEMS / 4; ++v) {
                const int4 m = __ldg(reinterpret_cast<const int4 *>(mask + pos0) + v);
                on[4 * v] = m.x == 1;
                on[4 * v + 1] = m.y == 1;
                on[4 * v + 2] = m.z == 1;
                on[4 * v + 3] = m.w == 1;
            }
        } else {
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) on[i] = pos0 + i < n_pos && mask[pos0 + i] == 1;
        }
        // ---- symbols: the block's top-left token (model.py:219-221).  The loads do not wait for the mask
        //      values (issued for every in-range position), so mask and index latencies overlap.
        int y = 0, x = 0;
        if (mask) {
            y = (int)(pos0 / gw);
            x = (int)(pos0 - (int64_t)y * gw);
        }
        int64_t val[ITEMS];
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            val[i] = 0;
            if (pos0 + i < n_pos) {
                const int64_t at = mask ? (int64_t)(y * step) * a.w + x * step : pos0 + i;
                val[i] = src[at];
            }
            if (++x == gw) {
                x = 0;
                ++y;
            }
        }
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            sym[i] = -1;
            if (on[i]) {
                if (val[i] < 0 || val[i] >= a.T.K) s_bad = 1;
                else sym[i] = (int)val[i];
            }
        }
        if (mbar) {  // the code table's bulk copy was started before the loads above; first use is below
            mbar_wait(mbar, 0);
            mbar = nullptr;
        }
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            len[i] = 0;
            code[i] = 0;
            if (sym[i] >= 0) {
                if (s_enc) {
                    const uint2 e = s_enc[sym[i]];
                    code[i] = e.x;
                    len[i] = e.y;
                } else {
                    len[i] = a.T.len[sym[i]];
                }
            }
            tsum += (int)len[i];
        }
        }  // generic path
        int tot;
        int o = block_exscan(tsum, s_warp, &tot);
        const int r0 = (int)(P & 31);
        const int nwords = (r0 + tot + 31) >> 5;
        for (int j = tid; j < nwords; j += PK_THREADS) stage[j] = (j == 0 && r0) ? s_carry : 0u;
        __syncthreads();
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            if (len[i]) {
                int q = r0 + o;
                if (s_enc) {
                    const int sh = q & 31, wi = q >> 5;
                    atomicOr(&stage[wi], code[i] >> sh);
                    if (sh + (int)len[i] > 32) atomicOr(&stage[wi + 1], code[i] << (32 - sh));
                } else {
                    const uint32_t *cw_p = a.T.pool + a.T.off[sym[i]];
                    for (int rem = (int)len[i]; rem > 0; rem -= 32, q += 32, ++cw_p) {
                        const uint32_t cw = *cw_p;
                        const int sh = q & 31, wi = q >> 5;
                        atomicOr(&stage[wi], cw >> sh);
                        if (sh && (sh + min(rem, 32) > 32)) atomicOr(&stage[wi + 1], cw << (32 - sh));
                    }
                }
                o += (int)len[i];
            }
        }
        __syncthreads();
        const int full = (r0 + tot) >> 5;
        const int64_t w0 = P >> 5;
        if ((w0 + full) * 4 <= cap)
            for (int j = tid; j < full; j += PK_THREADS) out32[w0 + j] = to_big_endian(stage[j]);
        if (tid == 0) s_carry = ((r0 + tot) & 31) ? stage[full] : 0u;
        P += tot;
        __syncthreads();
    }
    if (mbar) mbar_wait(mbar, 0);  // empty stream: never leave with the bulk copy in flight
    // tail: bytes not yet written, pad, header
    const int64_t nbits = P - 8;
    if (nbits == 0 || s_bad) {
        if (tid == 0) *size_out = s_bad ? -1 : 0;
        return;
    }
    const int64_t total = nbits / 8 + 2;
    const int64_t written = (P >> 5) * 4;
    const uint32_t carry = (P & 31) ? s_carry : 0u;
    if (total > cap) {
        if (tid == 0) *size_out = -2;
        return;
    }
    if (tid < 8) {
        const int64_t i = written + tid;
        if (i < total) out[i] = tid < 4 ? (uint8_t)(carry >> (24 - 8 * tid)) : (uint8_t)0;
    }
    __syncthreads();
    if (tid == 0) {
        out[0] = (uint8_t)(8 - (int)(nbits & 7));
        *size_out = (int32_t)total;
    }
}

// ---- the same for one stream spread over several CTAs (large token grids) -----------------------------------------
// Slot j of a stream packs tiles j, j + nslots, ...  A tile needs two things from its predecessor: the bit position
// where it starts (known as soon as the predecessor has scanned its code lengths) and the bits of the 32-bit word the
// two tiles share (known once the predecessor has staged its codes).  Both travel through self-clearing records:
//   recA[t]: bit 63 valid, bit 62 "a symbol was outside the table so far", bits 0..47 bit position after tile t
//   recB[t]: bit 63 valid, bits 0..31 the partial last word of tile t (0 when it ended on a word boundary)
// Long chains (a.gather, >= 8 tiles in the fine stream): the bit position does not travel from tile to tile; every tile
// publishes its own bit count to ALL later tiles at once (cnt[p][t], t > p: one copy per reader, which clears it) as soon as
// it has scanned its code lengths, and gathers the counts of all earlier tiles -- one publish + one gather whatever the
// number of tiles (2032 x 1344 as 768-pixel tiles, 9-tile chains: -2.4 us; on 6-tile chains the plain chain is faster).
// The boundary word is written by the LATER tile.  The CTA that packs the last tile finishes the stream (pad, header,
// size).  Tiles are handed out in block-id order, so a CTA only ever waits for CTAs dispatched before it.
__device__ __forceinline__ unsigned long long pk_wait_clear(unsigned long long *p)
{
    unsigned long long v;
    do {
        asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
        if (!(v >> 63)) __nanosleep(40);
    } while (!(v >> 63));
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(0ull) : "memory");
    return v;
}
__device__ __forceinline__ void pk_publish(unsigned long long *p, unsigned long long payload)
{
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"((1ull << 63) | payload) : "memory");
}

template <int ITEMS>
__device__ void pack_index_stream_chained(const PackArgs &a, int s, int b, int slot, uint32_t *stage, const uint2 *s_enc, unsigned long long *mbar)
{
    constexpr int TILE = PK_THREADS * ITEMS;
    __shared__ int s_warp[PK_THREADS / 32];
    __shared__ uint32_t s_carry;
    __shared__ int s_bad, s_badall;
    __shared__ unsigned long long s_P;
    const int tid = threadIdx.x;
    const int step = 4 >> s;
    const int gw = a.w / step;
    const int64_t n_pos = (int64_t)(a.h / step) * gw;
    const int32_t *mask = a.mask[s] + (int64_t)b * n_pos;
    const int64_t *src = a.idx + (int64_t)b * a.h * a.w;
    uint8_t *out = a.out + (int64_t)b * a.image_stride + a.slot_off[s];
    int32_t *size_out = a.sizes + b * 5 + s;
    const int64_t cap = a.slot_cap[s];
    uint32_t *out32 = reinterpret_cast<uint32_t *>(out);
    // per stream: recA[max_tiles] (chain) aliasing the first row of cnt[max_tiles][max_tiles] (gather), then recB[max_tiles]
    unsigned long long *cnt = a.chain + ((int64_t)b * 3 + s) * ((int64_t)a.max_tiles * a.max_tiles + a.max_tiles);
    unsigned long long *recA = cnt, *recB = cnt + (int64_t)a.max_tiles * a.max_tiles;
    const int n_tiles = (int)((n_pos + TILE - 1) / TILE);
    const bool mine_last = (n_tiles - 1) % a.nslots == slot;
    if (tid == 0) {
        s_carry = 0;
        s_bad = 0;
        s_badall = 0;
    }
    __syncthreads();
    int64_t P_end = 8;
    const bool fast = ITEMS == 8 && level_grid_fast(a, s);
    const LevelGrid lg = fast ? level_grid(a, s, b) : LevelGrid{};
    for (int t = slot; t < n_tiles; t += a.nslots) {
        const int64_t tile = (int64_t)t * TILE;
        uint32_t len[ITEMS], code[ITEMS];
        const int64_t pos0 = tile + (int64_t)tid * ITEMS;
        int tsum = 0;
        if (ITEMS == 8 && fast) {
            int4 m0, m1;
            long long v[8];
            load_tile_inputs(lg, (int)pos0, m0, m1, v);
            if (mbar) {
                mbar_wait(mbar, 0);
                mbar = nullptr;
            }
            int bad = 0;
            tsum = tile_codes(m0, m1, v, s_enc, a.T.K, reinterpret_cast<uint32_t (&)[8]>(code), reinterpret_cast<uint32_t (&)[8]>(len), &bad);
            if (bad) s_bad = 1;
        } else {
        int sym[ITEMS];
        bool on[ITEMS];
        if (ITEMS % 4 == 0 && pos0 + ITEMS <= n_pos && (reinterpret_cast<uintptr_t>(mask + pos0) & 15) == 0) {
#pragma unroll
            for (int v = 0; v < ITEMS / 4; ++v) {
                const int4 m = __ldg(reinterpret_cast<const int4 *>(mask + pos0) + v);
                on[4 * v] = m.x == 1;
                on[4 * v + 1] = m.y == 1;
                on[4 * v + 2] = m.z == 1;
                on[4 * v + 3] = m.w == 1;
            }
        } else {
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) on[i] = pos0 + i < n_pos && mask[pos0 + i] == 1;
        }
        int y = (int)(pos0 / gw), x = (int)(pos0 - (int64_t)y * gw);
        int64_t val[ITEMS];
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            val[i] = 0;
            if (pos0 + i < n_pos) val[i] = src[(int64_t)(y * step) * a.w + x * step];
            if (++x == gw) {
                x = 0;
                ++y;
            }
        }
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            sym[i] = -1;
            if (on[i]) {
                if (val[i] < 0 || val[i] >= a.T.K) s_bad = 1;
                else sym[i] = (int)val[i];
            }
        }
        if (mbar) {
            mbar_wait(mbar, 0);
            mbar = nullptr;
        }
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            len[i] = 0;
            code[i] = 0;
            if (sym[i] >= 0) {
                const uint2 e = s_enc[sym[i]];
                code[i] = e.x;
                len[i] = e.y;
            }
            tsum += (int)len[i];
        }
        }  // generic path
        CGIC_STAMP(pack, 3);  // codes ready
        int tot;
        int o = block_exscan(tsum, s_warp, &tot);  // (its barriers also order the s_bad writes above before the read below)
        CGIC_STAMP(pack, 4);  // scanned
        if (a.gather) {
            if (tid < 32) {
                // this tile's count to every later tile first (nobody waits for it longer than necessary) ...
                const unsigned long long mine = ((unsigned long long)(s_bad != 0) << 62) | (unsigned long long)tot;
                for (int r = t + 1 + tid; r < n_tiles; r += 32) pk_publish(cnt + (int64_t)t * a.max_tiles + r, mine);
                // ... then the counts of all earlier tiles (each record is this tile's own copy: cleared after reading)
                unsigned long long sum = 0;
                unsigned bad = 0;
                for (int p = tid; p < t; p += 32) {
                    const unsigned long long r = pk_wait_clear(cnt + (int64_t)p * a.max_tiles + t);
                    sum += r & 0xFFFFFFFFFFFFull;
                    bad |= (unsigned)(r >> 62) & 1u;
                }
#pragma unroll
                for (int o2 = 16; o2 > 0; o2 >>= 1) {
                    sum += __shfl_xor_sync(0xffffffffu, sum, o2);
                    bad |= __shfl_xor_sync(0xffffffffu, bad, o2);
                }
                if (tid == 0) {
                    s_P = 8ull + sum;
                    if (bad || s_bad) s_badall = 1;
                }
            }
        } else if (tid == 0) {
            unsigned long long P = 8, bad = (unsigned long long)(s_bad != 0);
            if (t > 0) {
                const unsigned long long r = pk_wait_clear(recA + (t - 1));
                P = r & 0xFFFFFFFFFFFFull;
                bad |= (r >> 62) & 1ull;
            }
            if (t + 1 < n_tiles) pk_publish(recA + t, (bad << 62) | (P + (unsigned long long)tot));
            s_P = P;
            s_badall = (int)bad;
        }
        __syncthreads();
        CGIC_STAMP(pack, 5);  // bit position known
        const int64_t P = (int64_t)s_P;
        const int r0 = (int)(P & 31);
        const int nwords = (r0 + tot + 31) >> 5;
        for (int j = tid; j < nwords; j += PK_THREADS) stage[j] = 0u;
        __syncthreads();
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            if (len[i]) {
                const int q = r0 + o;
                const int sh = q & 31, wi = q >> 5;
                atomicOr(&stage[wi], code[i] >> sh);
                if (sh + (int)len[i] > 32) atomicOr(&stage[wi + 1], code[i] << (32 - sh));
                o += (int)len[i];
            }
        }
        __syncthreads();
        CGIC_STAMP(pack, 6);  // staged
        const int full = (r0 + tot) >> 5;
        const int64_t w0 = P >> 5;
        if (tid == 0) {
            const bool partial = ((r0 + tot) & 31) != 0;
            // a tile that completes at least one word knows its own last partial word without its predecessor
            if (full > 0 && t + 1 < n_tiles) pk_publish(recB + t, partial ? stage[full] : 0u);
            uint32_t carry_in = 0;
            if (t > 0) carry_in = (uint32_t)pk_wait_clear(recB + (t - 1));  // (0 when this tile starts on a word boundary)
            stage[0] |= carry_in;
            if (full == 0 && t + 1 < n_tiles) pk_publish(recB + t, partial ? stage[0] : 0u);  // everything so far sits in one word
            s_carry = partial ? stage[full] : 0u;
        }
        __syncthreads();
        if ((w0 + full) * 4 <= cap)
            for (int j = tid; j < full; j += PK_THREADS) out32[w0 + j] = to_big_endian(stage[j]);
        P_end = P + tot;
        __syncthreads();
        CGIC_STAMP(pack, 7);  // written
    }
    if (mbar) mbar_wait(mbar, 0);
    if (!mine_last) return;
    // tail by the CTA of the last tile: bytes not yet written, pad, header
    const int64_t P = P_end;
    const int64_t nbits = P - 8;
    const bool bad = s_badall != 0 || s_bad != 0;
    if (nbits == 0 || bad) {
        if (tid == 0) *size_out = bad ? -1 : 0;
        return;
    }
    const int64_t total = nbits / 8 + 2;
    const int64_t written = (P >> 5) * 4;
    const uint32_t carry = (P & 31) ? s_carry : 0u;
    if (total > cap) {
        if (tid == 0) *size_out = -2;
        return;
    }
    if (tid < 8) {
        const int64_t i = written + tid;
        if (i < total) out[i] = tid < 4 ? (uint8_t)(carry >> (24 - 8 * tid)) : (uint8_t)0;
    }
    __syncthreads();
    if (tid == 0) {
        out[0] = (uint8_t)(8 - (int)(nbits & 7));
        *size_out = (int32_t)total;
    }
}

// mask / raw bit stream: one output byte per thread
__device__ void pack_bit_stream(const int32_t *v, int64_t n, uint8_t *out, int64_t cap, int32_t *size_out)
{
    if (n == 0) {
        if (threadIdx.x == 0) *size_out = 0;
        return;
    }
    const int64_t total = n / 8 + 2;
    if (total > cap) {
        if (threadIdx.x == 0) *size_out = -2;
        return;
    }
    const int64_t nbytes = total - 1;  // payload bytes incl. the (possibly all-pad) last one
    for (int64_t j = threadIdx.x; j < nbytes; j += blockDim.x) {
        uint32_t byte = 0;
        const int64_t base = j * 8;
        if (base + 8 <= n && (reinterpret_cast<uintptr_t>(v) & 15) == 0) {
            const int4 lo = *reinterpret_cast<const int4 *>(v + base);
            const int4 hi = *reinterpret_cast<const int4 *>(v + base + 4);
            byte = ((lo.x != 0) << 7) | ((lo.y != 0) << 6) | ((lo.z != 0) << 5) | ((lo.w != 0) << 4) | ((hi.x != 0) << 3) |
                   ((hi.y != 0) << 2) | ((hi.z != 0) << 1) | (hi.w != 0);
        } else {
            for (int i = 0; i < 8 && base + i < n; ++i) byte |= (uint32_t)(v[base + i] != 0) << (7 - i);
        }
        out[1 + j] = (uint8_t)byte;
    }
    if (threadIdx.x == 0) {
        out[0] = (uint8_t)(8 - (int)(n & 7));
        *size_out = (int32_t)total;
    }
}

// dynamic shared memory: [code table K x 8 bytes when max_len <= 32][one tile of output bits]
template <int ITEMS>
__device__ __forceinline__ void pack_stream_entry(const PackArgs &a, int s, int b, unsigned char *dyn, unsigned long long *mbar)
{
    const uint2 *s_enc = nullptr;
    uint32_t *stage = reinterpret_cast<uint32_t *>(dyn);
    bool staged = false;
    if (a.T.enc) {  // stage the code table with one TMA bulk copy; waited for inside, after the first loads are in flight
        const uint32_t bytes = (uint32_t)((a.T.K + 1) / 2 * 2) * 8u;
        if (threadIdx.x == 0) mbar_init(mbar);
        __syncthreads();
        if (threadIdx.x == 0) tma_load_1d(dyn, a.T.enc, bytes, mbar);
        s_enc = reinterpret_cast<const uint2 *>(dyn);
        stage = reinterpret_cast<uint32_t *>(dyn + bytes);
        staged = true;
    }
    pdl_wait();  // the code table above is immutable; indices and masks come from the predecessor
    CGIC_STAMP(pack, 2);
    pack_index_stream<ITEMS>(a, s, b, stage, s_enc, staged ? mbar : nullptr);
}

// grid (4, B): per image one CTA for each index stream and one for the two mask streams.  Every index-stream CTA costs
// about the same whatever its stream's length (one tile: loads, scan, staging, barriers), so the mask streams riding
// with the coarse CTA made that CTA the kernel's critical path by ~1.7 us (in-graph timeline, profiles/trace_graph.py);
// a 64-image batch is 256 CTAs = still one wave at two CTAs per SM.
template <int ITEMS>
__global__ void __launch_bounds__(PK_THREADS) pack_kernel(const PackArgs a)
{
    extern __shared__ __align__(128) unsigned char dyn[];
    __shared__ __align__(8) unsigned long long mbar;
    // linear block order = hand-out order: the fine streams (4096 grid positions each) first, then medium, coarse, and
    // the light mask CTAs last, so that the second CTA of an SM is a short one
    const int lin = (int)(blockIdx.y * gridDim.x + blockIdx.x), nb = (int)gridDim.y;
    const int k = lin / nb, s = 2 - k, b = lin - k * nb;
    CGIC_STAMP(pack, 0);
    pdl_trigger_step<2>();
    if (k < 3) {
        if (stream_present(a.mode, s)) {
            pack_stream_entry<ITEMS>(a, s, b, dyn, &mbar);   // waits for the predecessor grid inside
        } else {
            pdl_wait();
            if (threadIdx.x == 0) a.sizes[b * 5 + s] = 0;
        }
        CGIC_STAMP(pack, 1);
        return;
    }
    pdl_wait();
    CGIC_STAMP(pack, 2);
    if (a.sq_out && b == 0) reduce_partials(a.sq_partials, a.sq_n, a.sq_out);
    for (int ms = 3; ms < 5; ++ms) {
        if (!stream_present(a.mode, ms)) {
            if (threadIdx.x == 0) a.sizes[b * 5 + ms] = 0;
            continue;
        }
        const int lvl = ms - 3;  // 0 coarse, 1 medium
        const int div = lvl == 0 ? 4 : 2;
        const int64_t n = (int64_t)(a.h / div) * (a.w / div);
        pack_bit_stream(a.mask[lvl] + (int64_t)b * n, n, a.out + (int64_t)b * a.image_stride + a.slot_off[ms], a.slot_cap[ms],
                        a.sizes + b * 5 + ms);
    }
    CGIC_STAMP(pack, 1);
}

// grid (3 * nslots, B): nslots CTAs per index stream (consecutive block ids); large token grids with <= 32-bit codes only
template <int ITEMS>
__global__ void __launch_bounds__(PK_THREADS) pack_chained_kernel(const PackArgs a)
{
    extern __shared__ __align__(128) unsigned char dyn[];
    __shared__ __align__(8) unsigned long long mbar;
    const int s = blockIdx.x / a.nslots, slot = blockIdx.x - s * a.nslots, b = blockIdx.y;
    CGIC_STAMP(pack, 0);
    pdl_trigger_step<2>();
    if (stream_present(a.mode, s)) {
        const uint32_t bytes = (uint32_t)((a.T.K + 1) / 2 * 2) * 8u;
        if (threadIdx.x == 0) mbar_init(&mbar);
        __syncthreads();
        if (threadIdx.x == 0) tma_load_1d(dyn, a.T.enc, bytes, &mbar);
        pdl_wait();
        CGIC_STAMP(pack, 2);
        pack_index_stream_chained<ITEMS>(a, s, b, slot, reinterpret_cast<uint32_t *>(dyn + bytes), reinterpret_cast<const uint2 *>(dyn), &mbar);
    } else {
        pdl_wait();
        if (threadIdx.x == 0 && slot == 0) a.sizes[b * 5 + s] = 0;
    }
    CGIC_STAMP(pack, 1);
    if (blockIdx.x != 0) return;
    if (a.sq_out && b == 0) reduce_partials(a.sq_partials, a.sq_n, a.sq_out);
    for (int ms = 3; ms < 5; ++ms) {
        if (!stream_present(a.mode, ms)) {
            if (threadIdx.x == 0) a.sizes[b * 5 + ms] = 0;
            continue;
        }
        const int lvl = ms - 3;
        const int div = lvl == 0 ? 4 : 2;
        const int64_t n = (int64_t)(a.h / div) * (a.w / div);
        pack_bit_stream(a.mask[lvl] + (int64_t)b * n, n, a.out + (int64_t)b * a.image_stride + a.slot_off[ms], a.slot_cap[ms],
                        a.sizes + b * 5 + ms);
    }
    CGIC_STAMP(pack, 8);
}

template <int ITEMS>
__global__ void __launch_bounds__(PK_THREADS) pack_single_kernel(const PackArgs a)
{
    extern __shared__ __align__(128) unsigned char dyn[];
    __shared__ __align__(8) unsigned long long mbar;
    pack_stream_entry<ITEMS>(a, 0, 0, dyn, &mbar);
}

__global__ void __launch_bounds__(PK_THREADS)
bits_single_kernel(const int32_t *v, int64_t n, uint8_t *out, int64_t cap, int32_t *size_out)
{
    pack_bit_stream(v, n, out, cap, size_out);
}

// ======================================================================================================================
// Small token grids (<= ES_MAX_N4 fine tokens, codes of at most 32 bits): VectorQuantize2.forward (a1) + index selection
// (a7) + the five-stream pack (a9, a11, a12) in ONE launch, one CTA per image.  The 32 warps first run the warp-tile
// search of vq_tile.cuh on the image's 4 x 32 tiles (idx / z_q / sum((e-z)^2) go to global memory as before, the indices
// also to shared memory as u16), then the CTA packs the three index streams in one pass over the concatenated level
// grids [coarse | medium | fine]: one exclusive scan of the code lengths serves all three (a stream's bit offsets are
// relative to the scan value at its first position), codes are OR-ed into per-stream staging windows in shared memory
// and leave as coalesced big-endian words; header byte and padding are part of the staged words.  The indices never
// make the round trip through global memory and the step loses one launch and one hand-over.
#ifndef CGIC_ES_THREADS
#define CGIC_ES_THREADS 1024
#endif
constexpr int ES_THREADS = CGIC_ES_THREADS;
constexpr int ES_WARPS = ES_THREADS / 32;
constexpr int ES_MAX_N4 = 4096;
constexpr int ES_ITEMS = (ES_MAX_N4 + ES_MAX_N4 / 4 + ES_MAX_N4 / 16 + ES_THREADS - 1) / ES_THREADS;  // (n16 + n8 + n4) <= 5376 positions per image

struct EncodeArgs {
    const float *z;
    const unsigned char *blob;  // prepared codebook
    int K, h, w, mode;
    const int32_t *mask[3];
    DevTable T;
    int64_t *idx_out;
    float *zq_out;
    double *partials, *sqerr_out;
    int32_t *counters;
    uint8_t *out;
    int64_t image_stride;
    int64_t slot_off[5];
    int64_t slot_cap[5];
    int32_t *sizes;
};

struct EsLayout {
    size_t enc, tiles, idx, stage[3], total;
};
__host__ __device__ inline EsLayout es_layout(int K, int h, int w, const int64_t cap[5])
{
    auto up = [](size_t v) { return (v + 127) / 128 * 128; };
    EsLayout L;
    size_t o = up(cb_layout(K).stage);
    L.enc = o;
    o += up((size_t)((K + 1) / 2 * 2) * 8);
    L.idx = o;
    o += up((size_t)h * w * 2);
    L.tiles = o;
    // the staging windows of the pack phase reuse the warps' tile slices
    size_t so = o;
    for (int s = 0; s < 3; ++s) {
        L.stage[s] = so;
        so += up((size_t)cap[s] + 8);
    }
    o += (size_t)ES_WARPS * VQW_TILE_BYTES;
    L.total = o > so ? o : so;
    return L;
}

__global__ void __launch_bounds__(ES_THREADS, 1) encode_small_kernel(const EncodeArgs a)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long mbar;
    __shared__ double s_red[ES_WARPS];
    __shared__ int s_wsum[ES_WARPS];
    __shared__ int s_base[4];
    __shared__ bool s_last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, b = blockIdx.x;
    const int h = a.h, w = a.w, K = a.K;
    const CbLayout CL = cb_layout(K);
    const EsLayout SL = es_layout(K, h, w, a.slot_cap);
    const CbHeader *hdr = reinterpret_cast<const CbHeader *>(smem);
    const unsigned char *lut = smem + CL.lut;
    const float4 *cbs = reinterpret_cast<const float4 *>(smem + CL.cb);
    const float *e2s = reinterpret_cast<const float *>(smem + CL.e2);
    const uint2 *s_enc = reinterpret_cast<const uint2 *>(smem + SL.enc);
    uint16_t *idx_s = reinterpret_cast<uint16_t *>(smem + SL.idx);
    unsigned char *wbase = smem + SL.tiles + (size_t)warp * VQW_TILE_BYTES;
    float4 *zs = reinterpret_cast<float4 *>(wbase);
    uint16_t *res = reinterpret_cast<uint16_t *>(zs + VQW_TILE);
    uint8_t *list = reinterpret_cast<uint8_t *>(res + VQW_TILE);
    uint8_t *lead = list + VQW_TILE;
    const uint4 *recs = reinterpret_cast<const uint4 *>(a.blob + CL.rec);

    CGIC_STAMP(pack, 0);
    pdl_trigger_step<1>();
    // immutable tables first: the code table of the pack phase may be staged while the predecessor still runs
    if (tid == 0) mbar_init(&mbar);
    __syncthreads();
    const uint32_t enc_bytes = (uint32_t)((a.T.K + 1) / 2 * 2) * 8u;
    pdl_wait();  // the prepared blob and z may come straight from a preceding kernel
    if (tid == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&mbar)), "r"((uint32_t)CL.stage + enc_bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(smem)),
                     "l"(a.blob), "r"((uint32_t)CL.stage), "r"(smem_addr(&mbar))
                     : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(smem + SL.enc)),
                     "l"(a.T.enc), "r"(enc_bytes), "r"(smem_addr(&mbar))
                     : "memory");
    }
    bool staged = false;
    CGIC_STAMP(pack, 2);

    // ---- a1: warp tiles of this image
    const int64_t plane = (int64_t)h * w;
    const int tiles_x = (w + 31) / 32, tiles_y = (h + 3) / 4, n_tiles = tiles_x * tiles_y;
    double sq = 0.0;
    const VqTileCtx ctx{hdr, lut, cbs, e2s, recs, K, zs, res, list, lead, nullptr};
    for (int tile = warp; tile < n_tiles; tile += ES_WARPS) {
        const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
        vq_process_tile(ctx, a.z + (int64_t)b * 4 * plane, h, w, ty * 4, tx * 32 + lane, lane, a.idx_out + (int64_t)b * plane,
                        a.zq_out ? a.zq_out + (int64_t)b * 4 * plane : nullptr, a.sqerr_out != nullptr, sq, idx_s, a.counters + 2, [&]() {
                            if (!staged) {
                                staged = true;
                                mbar_wait(&mbar, 0);
                            }
                        });
    }
    if (!staged) mbar_wait(&mbar, 0);  // (warps without a tile still need the code table below)
    CGIC_STAMP(pack, 7);
    // sum((e - z)^2): warp -> CTA partial; the last CTA to arrive adds the partials in image order (deterministic)
    if (a.sqerr_out) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        if (lane == 0) s_red[warp] = sq;
    }
    __syncthreads();  // idx_s complete; the tile slices are free: the staging windows take their place
    CGIC_STAMP(pack, 3);

    // ---- a7 + a9: the three index streams in one pass
    const int n4 = h * w, n8 = n4 / 4, n16 = n4 / 16;
    const int n_all = n16 + n8 + n4;
    uint32_t *const stage0 = reinterpret_cast<uint32_t *>(smem + SL.stage[0]), *const stage1 = reinterpret_cast<uint32_t *>(smem + SL.stage[1]),
                    *const stage2 = reinterpret_cast<uint32_t *>(smem + SL.stage[2]);
    auto stage_of = [&](int s) { return s == 0 ? stage0 : (s == 1 ? stage1 : stage2); };
    {
        // zero the staging windows (16 bytes per thread and pass)
        const int nz = (int)((SL.stage[2] + ((size_t)a.slot_cap[2] + 8 + 127) / 128 * 128 - SL.stage[0]) / 16);
        uint4 *zp = reinterpret_cast<uint4 *>(smem + SL.stage[0]);
        for (int i = tid; i < nz; i += ES_THREADS) zp[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    uint32_t code[ES_ITEMS], len[ES_ITEMS];
    int lvl[ES_ITEMS];
    int tsum = 0;
    const int u0 = tid * ES_ITEMS;
#pragma unroll
    for (int i = 0; i < ES_ITEMS; ++i) {
        const int u = u0 + i;
        code[i] = 0;
        len[i] = 0;
        lvl[i] = u < n16 ? 0 : (u < n16 + n8 ? 1 : 2);
        if (u < n_all) {
            const int s = lvl[i];
            const int p = u - (s == 0 ? 0 : (s == 1 ? n16 : n16 + n8));
            const int step = 4 >> s, gw = w / step;
            const int np = s == 0 ? n16 : (s == 1 ? n8 : n4);
            if (stream_present(a.mode, s) && __ldg(a.mask[s] + (int64_t)b * np + p) == 1) {
                const int y = p / gw, x = p - y * gw;
                const uint2 e = s_enc[idx_s[(y * step) * w + x * step]];  // the block's top-left token (model.py:219-221)
                code[i] = e.x;
                len[i] = e.y;
            }
        }
        tsum += (int)len[i];
    }
    // block-wide exclusive scan of the per-thread bit counts
    int inc = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
    }
    if (lane == 31) s_wsum[warp] = inc;
    __syncthreads();
    int woff = 0, total = 0;
#pragma unroll
    for (int i = 0; i < ES_WARPS; ++i) {
        const int t = s_wsum[i];
        if (i < warp) woff += t;
        total += t;
    }
    int o = woff + inc - tsum;
    // the scan value at a stream's first position = the bits of the streams before it
    {
        int run = o;
#pragma unroll
        for (int i = 0; i < ES_ITEMS; ++i) {
            const int u = u0 + i;
            if (u == 0) s_base[0] = run;
            if (u == n16) s_base[1] = run;
            if (u == n16 + n8) s_base[2] = run;
            run += (int)len[i];
        }
        if (tid == 0) s_base[3] = total;
    }
    __syncthreads();
    CGIC_STAMP(pack, 4);
    const int base0 = s_base[0], base1 = s_base[1], base2 = s_base[2];
#pragma unroll
    for (int i = 0; i < ES_ITEMS; ++i) {
        if (len[i]) {
            const int s = lvl[i];
            const int q = 8 + o - (s == 0 ? base0 : (s == 1 ? base1 : base2));  // the header byte is bits 0..7
            const int sh = q & 31, wi = q >> 5;
            uint32_t *st = stage_of(s);
            atomicOr(&st[wi], code[i] >> sh);
            if (sh + (int)len[i] > 32) atomicOr(&st[wi + 1], code[i] << (32 - sh));
            o += (int)len[i];
        }
    }
    // header byte (pad count, 1..8) of the non-empty streams
    if (tid < 3) {
        const int nbits = s_base[tid + 1] - s_base[tid];
        if (nbits > 0) atomicOr(stage_of(tid), (uint32_t)(8 - (nbits & 7)) << 24);
    }
    __syncthreads();
    CGIC_STAMP(pack, 5);
    uint8_t *img = a.out + (int64_t)b * a.image_stride;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
        const int nbits = s_base[s + 1] - s_base[s];
        const int total_bytes = nbits > 0 ? nbits / 8 + 2 : 0;  // empty symbol list -> 0 bytes (the reference's empty file)
        uint32_t *out32 = reinterpret_cast<uint32_t *>(img + a.slot_off[s]);
        const uint32_t *st = stage_of(s);
        for (int j = tid; j < (total_bytes + 3) / 4; j += ES_THREADS) out32[j] = to_big_endian(st[j]);
        if (tid == 0) a.sizes[b * 5 + s] = total_bytes;
    }
    CGIC_STAMP(pack, 6);
    // ---- a11: the two mask streams
    for (int ms = 3; ms < 5; ++ms) {
        if (!stream_present(a.mode, ms)) {
            if (tid == 0) a.sizes[b * 5 + ms] = 0;
            continue;
        }
        const int lv = ms - 3;
        const int64_t n = lv == 0 ? n16 : n8;
        pack_bit_stream(a.mask[lv] + (int64_t)b * n, n, img + a.slot_off[ms], a.slot_cap[ms], a.sizes + b * 5 + ms);
    }
    CGIC_STAMP(pack, 1);
    if (!a.sqerr_out) return;
    if (tid == 0) {
        double tot = 0.0;
        for (int i = 0; i < ES_WARPS; ++i) tot += s_red[i];
        a.partials[b] = tot;
        __threadfence();
        s_last = (atomicAdd(&a.counters[0], 1) == (int)gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        double tot = 0.0;
        for (int i = tid; i < (int)gridDim.x; i += ES_THREADS) tot += __ldcg(&a.partials[i]);
#pragma unroll
        for (int o2 = 16; o2 > 0; o2 >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o2);
        __syncthreads();
        if (lane == 0) s_red[warp] = tot;
        __syncthreads();
        if (tid == 0) {
            double all = 0.0;
            for (int i = 0; i < ES_WARPS; ++i) all += s_red[i];
            *a.sqerr_out = all;
            a.counters[0] = 0;  // leave the ticket zeroed for the next launch (workspace contract)
        }
    }
}

// ======================================================================================================================
// One CTA per IMAGE packs all five streams (token grids of at most PI_MAX_GROUPS * 8 level cells, codes of at most 32 bits,
// level grid widths multiples of 8).  pack_kernel gives every stream a CTA of its own -- four CTAs per image that each pay
// the whole fixed cost (table staging, scan, barriers) whatever their stream's length: 41 k warp instructions per 256x256
// image, most of them in the nearly empty coarse and mask CTAs.  Here the work items are groups of 8 consecutive cells of the
// concatenated level grids [coarse | medium | fine] (= stream order), two consecutive groups per thread; ONE exclusive scan
// of the code lengths serves the three streams (a stream's bit offsets are relative to the scan value at its first group),
// codes are OR-ed into per-stream staging windows and leave as coalesced big-endian words with header byte and padding.
constexpr int PI_THREADS = 512;
constexpr int PI_GPT = 2;                                  // groups per thread
constexpr int PI_MAX_GROUPS = PI_THREADS * PI_GPT;         // 1024 groups = 8192 level cells (a 256x256 image has 5376)

struct PiLayout {
    size_t stage[3], total;
};
__host__ __device__ inline PiLayout pi_layout(int K, const int64_t cap[5])
{
    auto up = [](size_t v) { return (v + 127) / 128 * 128; };
    PiLayout L;
    size_t o = up((size_t)((K + 1) / 2 * 2) * 8);
    for (int s = 0; s < 3; ++s) {
        L.stage[s] = o;
        o += up((size_t)cap[s] + 8);
    }
    L.total = o;
    return L;
}

__global__ void __launch_bounds__(PI_THREADS, 2) pack_image_kernel(const PackArgs a)
{
    extern __shared__ __align__(128) unsigned char dyn[];
    __shared__ __align__(8) unsigned long long mbar;
    __shared__ int s_wsum[PI_THREADS / 32];
    __shared__ int s_base[4];
    __shared__ int s_badv[3];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, b = blockIdx.x;
    const PiLayout SL = pi_layout(a.T.K, a.slot_cap);
    const uint2 *s_enc = reinterpret_cast<const uint2 *>(dyn);
    uint32_t *const stage0 = reinterpret_cast<uint32_t *>(dyn + SL.stage[0]), *const stage1 = reinterpret_cast<uint32_t *>(dyn + SL.stage[1]),
                    *const stage2 = reinterpret_cast<uint32_t *>(dyn + SL.stage[2]);
    auto stage_of = [&](int s) { return s == 0 ? stage0 : (s == 1 ? stage1 : stage2); };
    CGIC_STAMP(pack, 0);
    pdl_trigger_step<2>();
    if (tid == 0) mbar_init(&mbar);
    if (tid < 3) s_badv[tid] = 0;
    __syncthreads();
    if (tid == 0) tma_load_1d(dyn, a.T.enc, (uint32_t)((a.T.K + 1) / 2 * 2) * 8u, &mbar);  // the code table is immutable: staged before the wait
    // zero the staging windows while the table travels
    {
        const int nz = (int)((SL.total - SL.stage[0]) / 16);
        uint4 *zp = reinterpret_cast<uint4 *>(dyn + SL.stage[0]);
        for (int i = tid; i < nz; i += PI_THREADS) zp[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    pdl_wait();  // indices and masks come from the predecessor
    CGIC_STAMP(pack, 2);
    const int n4 = a.h * a.w, n8 = n4 >> 2, n16 = n4 >> 4;
    const int g16 = n16 >> 3, g8 = n8 >> 3, g4 = n4 >> 3, n_groups = g16 + g8 + g4;
    // ---- loads of both groups first, then the table
    int4 m0[PI_GPT], m1[PI_GPT];
    long long v[PI_GPT][8];
    int lvl[PI_GPT];
#pragma unroll
    for (int g = 0; g < PI_GPT; ++g) {
        const int item = PI_GPT * tid + g;
        lvl[g] = item < g16 ? 0 : (item < g16 + g8 ? 1 : 2);
        const int pos0 = (item - (lvl[g] == 0 ? 0 : (lvl[g] == 1 ? g16 : g16 + g8))) * 8;
        if (item < n_groups && stream_present(a.mode, lvl[g])) {
            const LevelGrid lg = level_grid(a, lvl[g], b);
            load_tile_inputs(lg, pos0, m0[g], m1[g], v[g]);
        } else {
            m0[g] = make_int4(0, 0, 0, 0);
            m1[g] = m0[g];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[g][i] = 0;
        }
    }
    if (a.sq_out && b == 0) reduce_partials(a.sq_partials, a.sq_n, a.sq_out);  // (cgic_encode: the search kernel's deferred reduction)
    mbar_wait(&mbar, 0);
    uint32_t code[PI_GPT][8], len[PI_GPT][8];
    int gsum[PI_GPT], tsum = 0;
#pragma unroll
    for (int g = 0; g < PI_GPT; ++g) {
        int bad = 0;
        gsum[g] = tile_codes(m0[g], m1[g], v[g], s_enc, a.T.K, code[g], len[g], &bad);
        if (bad) s_badv[lvl[g]] = 1;
        tsum += gsum[g];
    }
    // ---- one exclusive scan over all groups
    int inc = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
    }
    if (lane == 31) s_wsum[warp] = inc;
    __syncthreads();
    int woff = 0, total = 0;
#pragma unroll
    for (int i = 0; i < PI_THREADS / 32; ++i) {
        const int t = s_wsum[i];
        if (i < warp) woff += t;
        total += t;
    }
    const int o0 = woff + inc - tsum;
    // the scan value at a level's first group = the bits of the streams before it
#pragma unroll
    for (int g = 0; g < PI_GPT; ++g) {
        const int item = PI_GPT * tid + g, og = g == 0 ? o0 : o0 + gsum[0];
        if (item == 0) s_base[0] = og;
        if (item == g16) s_base[1] = og;
        if (item == g16 + g8) s_base[2] = og;
    }
    if (tid == 0) s_base[3] = total;
    __syncthreads();
    CGIC_STAMP(pack, 4);
#pragma unroll
    for (int g = 0; g < PI_GPT; ++g) {
        int o = (g == 0 ? o0 : o0 + gsum[0]) - s_base[lvl[g]] + 8;  // the header byte is bits 0..7
        uint32_t *st = stage_of(lvl[g]);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (len[g][i]) {
                const int sh = o & 31, wi = o >> 5;
                atomicOr(&st[wi], code[g][i] >> sh);
                if (sh + (int)len[g][i] > 32) atomicOr(&st[wi + 1], code[g][i] << (32 - sh));
                o += (int)len[g][i];
            }
        }
    }
    // header byte (pad count, 1..8) of the non-empty streams
    if (tid < 3) {
        const int nbits = s_base[tid + 1] - s_base[tid];
        if (nbits > 0) atomicOr(stage_of(tid), (uint32_t)(8 - (nbits & 7)) << 24);
    }
    __syncthreads();
    CGIC_STAMP(pack, 5);
    uint8_t *img = a.out + (int64_t)b * a.image_stride;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
        const int nbits = s_base[s + 1] - s_base[s];
        const int total_bytes = nbits > 0 ? nbits / 8 + 2 : 0;  // empty symbol list -> 0 bytes (the reference's empty file)
        const bool bad = s_badv[s] != 0;
        uint32_t *out32 = reinterpret_cast<uint32_t *>(img + a.slot_off[s]);
        const uint32_t *st = stage_of(s);
        if (!bad)
            for (int j = tid; j < (total_bytes + 3) / 4; j += PI_THREADS) out32[j] = to_big_endian(st[j]);
        if (tid == 0) a.sizes[b * 5 + s] = bad ? -1 : total_bytes;
    }
    CGIC_STAMP(pack, 6);
    for (int ms = 3; ms < 5; ++ms) {
        if (!stream_present(a.mode, ms)) {
            if (tid == 0) a.sizes[b * 5 + ms] = 0;
            continue;
        }
        const int64_t n = ms == 3 ? n16 : n8;
        pack_bit_stream(a.mask[ms - 3] + (int64_t)b * n, n, img + a.slot_off[ms], a.slot_cap[ms], a.sizes + b * 5 + ms);
    }
    CGIC_STAMP(pack, 1);
}

// tile of PK_THREADS * items positions: output bits + 2 words, plus the staged code table
size_t pack_smem_bytes(const DevTable &T, int items)
{
    size_t b = ((size_t)PK_THREADS * items * T.max_len / 32 + 4) * 4;
    if (T.enc) b += (size_t)((T.K + 1) / 2 * 2) * 8;
    return b;
}

}  // namespace
}  // namespace cgic

using namespace cgic;

static int pack_max_tiles(int h, int w) { return (int)(((int64_t)h * w + PK_THREADS * 8 - 1) / (PK_THREADS * 8)) + 1; }

extern "C" size_t cgic_pack_workspace_bytes(int B, int h, int w)
{
    if (B <= 0 || h <= 0 || w <= 0) return 256;
    const size_t mt = (size_t)pack_max_tiles(h, w);
    return ((size_t)B * 3 * (mt * mt + mt) * 8 + 255) / 256 * 256;  // count records (one per ordered pair of tiles) + carry records
}

extern "C" int cgic_pack(const int64_t *idx, const int32_t *m_c, const int32_t *m_m, const int32_t *m_f, int B, int h, int w,
                         int mode, const cgic_table *t, uint8_t *bytes_out, int32_t *sizes_out, cgic_stream_t stream)
{
    return cgic_pack_ws(idx, m_c, m_m, m_f, B, h, w, mode, t, bytes_out, sizes_out, nullptr, 0, stream);
}

static int pack_launch(const int64_t *idx, const int32_t *m_c, const int32_t *m_m, const int32_t *m_f, int B, int h, int w, int mode,
                       const cgic_table *t, uint8_t *bytes_out, int32_t *sizes_out, void *workspace, size_t workspace_bytes, cgic_stream_t stream,
                       const double *sq_partials, int sq_n, double *sq_out);

extern "C" int cgic_pack_ws(const int64_t *idx, const int32_t *m_c, const int32_t *m_m, const int32_t *m_f, int B, int h, int w,
                            int mode, const cgic_table *t, uint8_t *bytes_out, int32_t *sizes_out, void *workspace,
                            size_t workspace_bytes, cgic_stream_t stream)
{
    return pack_launch(idx, m_c, m_m, m_f, B, h, w, mode, t, bytes_out, sizes_out, workspace, workspace_bytes, stream, nullptr, 0, nullptr);
}

static int pack_launch(const int64_t *idx, const int32_t *m_c, const int32_t *m_m, const int32_t *m_f, int B, int h, int w, int mode,
                       const cgic_table *t, uint8_t *bytes_out, int32_t *sizes_out, void *workspace, size_t workspace_bytes, cgic_stream_t stream,
                       const double *sq_partials, int sq_n, double *sq_out)
{
    CGIC_REQUIRE(idx && m_c && m_m && m_f && bytes_out && sizes_out, CGIC_EINVAL, "cgic_pack: null argument");
    CGIC_REQUIRE(B >= 0 && h > 0 && w > 0 && h % 4 == 0 && w % 4 == 0, CGIC_EINVAL, "cgic_pack: token grid %dx%d must be multiples of 4", h, w);
    CGIC_REQUIRE(mode >= 0 && mode <= 6, CGIC_EINVAL, "cgic_pack: mode %d", mode);
    CGIC_REQUIRE((reinterpret_cast<uintptr_t>(bytes_out) & 15) == 0 && (reinterpret_cast<uintptr_t>(m_c) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(m_m) & 15) == 0 && (reinterpret_cast<uintptr_t>(m_f) & 15) == 0,
                 CGIC_EINVAL, "cgic_pack: bytes_out and masks must be 16-byte aligned");
    if (B == 0) return CGIC_OK;
    PackArgs a{};
    int rc = table_device_view(t, &a.T);
    if (rc) return rc;
    const PackLayout L = make_pack_layout(a.T.max_len, h, w);
    a.idx = idx;
    a.mask[0] = m_c;
    a.mask[1] = m_m;
    a.mask[2] = m_f;
    a.h = h;
    a.w = w;
    a.mode = mode;
    a.n_direct = -1;
    a.out = bytes_out;
    a.image_stride = L.stride;
    for (int s = 0; s < 5; ++s) {
        a.slot_off[s] = L.off[s];
        a.slot_cap[s] = L.cap[s];
    }
    a.sizes = sizes_out;
    a.sq_partials = sq_partials;
    a.sq_n = sq_n;
    a.sq_out = sq_out;
    const int items = a.T.enc ? 8 : 1;
    const size_t smem = pack_smem_bytes(a.T, items);
    CGIC_REQUIRE(smem <= 200 * 1024, CGIC_EINVAL, "cgic_pack: code length %d needs %zu bytes of staging", a.T.max_len, smem);
    rc = ensure_smem(items == 8 ? (const void *)pack_kernel<8> : (const void *)pack_kernel<1>, smem);
    if (rc) return rc;
    // small token grids: one CTA per image packs all five streams (pack_image_kernel) -- every level's grid width a multiple of 8
    // (so is then its cell count), all groups within one pass
    {
        const int64_t n4 = (int64_t)h * w;
        // Measured on B200 (256x256 images): 2048 images 70 us against 123 us for one CTA per stream, 512 images 26 against 35 us;
        // 64 images 12.7 against 11.2 us -- a batch that leaves SMs idle is better served by four CTAs per image.
        int n_sm = 0;
        rc = device_sm_count(&n_sm);
        if (rc) return rc;
        const int mode_pi = tune_pack_image();
        const bool eligible = items == 8 && (w & 31) == 0 && (n4 + n4 / 4 + n4 / 16) / 8 <= PI_MAX_GROUPS && (mode_pi == 1 || (mode_pi == 0 && B > n_sm));
        if (eligible) {
            const size_t smem_pi = pi_layout(a.T.K, a.slot_cap).total;
            if (smem_pi <= 100 * 1024) {
                rc = ensure_smem((const void *)pack_image_kernel, smem_pi);
                if (rc) return rc;
                {
                    CGIC_PROF("pack_image_kernel", as_stream(stream));
                    CGIC_CUDA_CHECK(launch_pdl(pack_image_kernel, dim3(B), dim3(PI_THREADS), smem_pi, as_stream(stream), a));
                }
                CGIC_LAUNCH_CHECK();
                return CGIC_OK;
            }
        }
    }
    // large token grids (more than one tile per fine stream) with a workspace: tiles chained over several CTAs per stream
    const int fine_tiles = (int)(((int64_t)h * w + PK_THREADS * 8 - 1) / (PK_THREADS * 8));
    if (items == 8 && fine_tiles > 1 && workspace && workspace_bytes >= cgic_pack_workspace_bytes(B, h, w)) {
        a.chain = static_cast<unsigned long long *>(workspace);
        a.max_tiles = pack_max_tiles(h, w);
        // one CTA per tile of the fine stream up to 16 (a 768 x 768 tile has 9: with 8 CTAs one of them packed two tiles in
        // a row and the stream finished 4 us later)
        static const int slots_env = getenv("CGIC_PACK_SLOTS") ? atoi(getenv("CGIC_PACK_SLOTS")) : 0;  // A-B runs only
        const int slots_max = slots_env >= 1 && slots_env <= 32 ? slots_env : 16;
        a.nslots = fine_tiles < slots_max ? fine_tiles : slots_max;
        static const int gather_env = getenv("CGIC_PACK_GATHER") ? atoi(getenv("CGIC_PACK_GATHER")) : -1;        // A-B runs only
        a.gather = gather_env >= 0 ? (gather_env != 0) : fine_tiles >= 8;
        rc = ensure_smem((const void *)pack_chained_kernel<8>, smem);
        if (rc) return rc;
        {
            CGIC_PROF("pack_chained_kernel", as_stream(stream));
            CGIC_CUDA_CHECK(launch_pdl(pack_chained_kernel<8>, dim3(3 * a.nslots, B), dim3(PK_THREADS), smem, as_stream(stream), a));
        }
        CGIC_LAUNCH_CHECK();
        return CGIC_OK;
    }
    {
        CGIC_PROF("pack_kernel", as_stream(stream));
        if (items == 8) CGIC_CUDA_CHECK(launch_pdl(pack_kernel<8>, dim3(4, B), dim3(PK_THREADS), smem, as_stream(stream), a));
        else CGIC_CUDA_CHECK(launch_pdl(pack_kernel<1>, dim3(4, B), dim3(PK_THREADS), smem, as_stream(stream), a));
    }
    CGIC_LAUNCH_CHECK();
    return CGIC_OK;
}

namespace cgic {
const unsigned char *codebook_blob(const cgic_codebook *cb);  // codebook.cu
int codebook_size(const cgic_codebook *cb);
int vq_assign_indexed_launch(const float *z, int B, int h, int w, const cgic_codebook *cb, int64_t *idx_out, float *zq_out, double *sqerr_out,
                             void *workspace, size_t workspace_bytes, cgic_stream_t stream_, bool defer_reduce, int *grid_out);  // vq_assign.cu
}  // namespace cgic

extern "C" size_t cgic_encode_workspace_bytes(int B, int h, int w)
{
    const size_t a = cgic_vq_workspace_bytes((int64_t)B * h * w), b = 256 + 8 * (size_t)(B > 0 ? B : 0);
    return ((a > b ? a : b) + 255) / 256 * 256 + cgic_pack_workspace_bytes(B, h, w);
}

extern "C" int cgic_encode(const float *z, const int32_t *m_c, const int32_t *m_m, const int32_t *m_f, int B, int h, int w, int mode,
                           const cgic_codebook *cb, const cgic_table *t, int64_t *idx_out, float *zq_out, double *sqerr_out,
                           uint8_t *bytes_out, int32_t *sizes_out, void *workspace, size_t workspace_bytes, cgic_stream_t stream_)
{
    CGIC_REQUIRE(z && m_c && m_m && m_f && idx_out && bytes_out && sizes_out && workspace, CGIC_EINVAL, "cgic_encode: null argument");
    CGIC_REQUIRE(B >= 0 && h > 0 && w > 0 && h % 4 == 0 && w % 4 == 0, CGIC_EINVAL, "cgic_encode: token grid %dx%d must be multiples of 4", h, w);
    CGIC_REQUIRE(mode >= 0 && mode <= 6, CGIC_EINVAL, "cgic_encode: mode %d", mode);
    CGIC_REQUIRE(workspace_bytes >= cgic_encode_workspace_bytes(B, h, w), CGIC_ESPACE, "cgic_encode: workspace %zu < %zu bytes", workspace_bytes,
                 cgic_encode_workspace_bytes(B, h, w));
    const unsigned char *blob = codebook_blob(cb);
    CGIC_REQUIRE(blob, CGIC_EINVAL, "cgic_encode: no prepared codebook (cgic_codebook_update has not been called)");
    if (B == 0) return CGIC_OK;
    EncodeArgs a{};
    int rc = table_device_view(t, &a.T);
    if (rc) return rc;
    const int K = codebook_size(cb);
    const size_t vq_ws = ((std::max(cgic_vq_workspace_bytes((int64_t)B * h * w), (size_t)256 + 8 * (size_t)B)) + 255) / 256 * 256;
    cudaStream_t stream = as_stream(stream_);
    // The one-CTA-per-image encoder is opt-in (cgic_tune("fused_encode", 1) or CGIC_FUSED_ENCODE=1): measured on B200 it loses to the two launches at every
    // batch size tried (64 images: 26.6 us against 17.9 us; 2048 images: 382 us against 324 us) -- the warp-tile search is
    // instruction-issue bound, and one SM per image serialises what the two-launch path spreads over the whole machine.
    const bool fused = tune_fused_encode() == 1;
    const PackLayout L = make_pack_layout(a.T.max_len, h, w);
    bool small = fused && (int64_t)h * w <= ES_MAX_N4 && a.T.enc != nullptr && a.T.K == K;
    size_t smem = 0;
    if (small) {
        smem = es_layout(K, h, w, L.cap).total;
        small = smem <= 200 * 1024;
    }
    if (!small) {
        // large token grids / long codes: the two launches (warp-tile search, then the chained packer)
        // the sum of (e - z)^2 over the batch is finished by the packer (its mask CTA of image 0) from the search kernel's per-CTA
        // partials: the search kernel's own last-CTA reduction would add ~2 us to the step's critical path
        int vq_grid = 0;
        rc = vq_assign_indexed_launch(z, B, h, w, cb, idx_out, zq_out, sqerr_out, workspace, vq_ws, stream_, true, &vq_grid);
        if (rc) return rc;
        return pack_launch(idx_out, m_c, m_m, m_f, B, h, w, mode, t, bytes_out, sizes_out, static_cast<unsigned char *>(workspace) + vq_ws,
                           workspace_bytes - vq_ws, stream_, reinterpret_cast<const double *>(static_cast<unsigned char *>(workspace) + 256),
                           sqerr_out ? vq_grid : 0, sqerr_out);
    }
    for (const void *ptr : {(const void *)bytes_out, (const void *)m_c, (const void *)m_m, (const void *)m_f})
        CGIC_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, CGIC_EINVAL, "cgic_encode: bytes_out and masks must be 16-byte aligned");
    a.z = z;
    a.blob = blob;
    a.K = K;
    a.h = h;
    a.w = w;
    a.mode = mode;
    a.mask[0] = m_c;
    a.mask[1] = m_m;
    a.mask[2] = m_f;
    a.idx_out = idx_out;
    a.zq_out = zq_out;
    a.sqerr_out = sqerr_out;
    a.counters = static_cast<int32_t *>(workspace);
    a.partials = reinterpret_cast<double *>(static_cast<unsigned char *>(workspace) + 256);
    a.out = bytes_out;
    a.image_stride = L.stride;
    for (int s = 0; s < 5; ++s) {
        a.slot_off[s] = L.off[s];
        a.slot_cap[s] = L.cap[s];
    }
    a.sizes = sizes_out;
    rc = ensure_smem((const void *)encode_small_kernel, smem);
    if (rc) return rc;
    {
        CGIC_PROF("encode_small_kernel", stream);
        CGIC_CUDA_CHECK(launch_pdl(encode_small_kernel, dim3(B), dim3(ES_THREADS), smem, stream, a));
    }
    CGIC_LAUNCH_CHECK();
    return CGIC_OK;
}

extern "C" int cgic_huff_encode(const int64_t *symbols, int64_t n, const cgic_table *t, uint8_t *out, int64_t cap,
                                int32_t *size_out, cgic_stream_t stream)
{
    CGIC_REQUIRE(out && size_out && n >= 0 && (symbols || n == 0), CGIC_EINVAL, "cgic_huff_encode: bad argument");
    CGIC_REQUIRE((reinterpret_cast<uintptr_t>(out) & 3) == 0, CGIC_EINVAL, "cgic_huff_encode: out must be 4-byte aligned");
    PackArgs a{};
    int rc = table_device_view(t, &a.T);
    if (rc) return rc;
    a.idx = symbols;
    a.n_direct = n;
    a.out = out;
    a.slot_cap[0] = cap;
    a.sizes = size_out;
    const int items = a.T.enc ? 8 : 1;
    const size_t smem = pack_smem_bytes(a.T, items);
    CGIC_REQUIRE(smem <= 200 * 1024, CGIC_EINVAL, "cgic_huff_encode: code length %d too long", a.T.max_len);
    rc = ensure_smem(items == 8 ? (const void *)pack_single_kernel<8> : (const void *)pack_single_kernel<1>, smem);
    if (rc) return rc;
    if (items == 8) pack_single_kernel<8><<<1, PK_THREADS, smem, as_stream(stream)>>>(a);
    else pack_single_kernel<1><<<1, PK_THREADS, smem, as_stream(stream)>>>(a);
    CGIC_LAUNCH_CHECK();
    return CGIC_OK;
}

extern "C" int cgic_bits_encode(const int32_t *values, int64_t n, uint8_t *out, int64_t cap, int32_t *size_out,
                                cgic_stream_t stream)
{
    CGIC_REQUIRE(out && size_out && n >= 0 && (values || n == 0), CGIC_EINVAL, "cgic_bits_encode: bad argument");
    CGIC_REQUIRE((reinterpret_cast<uintptr_t>(values) & 15) == 0, CGIC_EINVAL, "cgic_bits_encode: values must be 16-byte aligned");
    bits_single_kernel<<<1, PK_THREADS, 0, as_stream(stream)>>>(values, n, out, cap, size_out);
    CGIC_LAUNCH_CHECK();
    return CGIC_OK;
}
