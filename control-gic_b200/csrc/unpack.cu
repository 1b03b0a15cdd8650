// a10 + a11 + a13 + a14: stream decode, mask / index re-assembly and codebook gather.
//
//   reference: CGIC/models/model.py:269-392;  HuffmanCoding.decompress_string
//              CGIC/tools/indices_coding.py:153-168 (131-138, 140-151);  BinaryCoding
//              CGIC/tools/mask_coding.py:81-96.
//
// Stage 1 (unpack_decode_kernel, four CTAs per image):
//   streams 0..2  prefix decode of the Huffman payload into a symbol list (table-driven: the
//                 first lut_bits bits index a LUT that either names the symbol or the tree node
//                 to continue from, so codes of any length -- the untrained model has 224-bit
//                 codes -- decode correctly);
//   stream 3 / 4  granularity masks -> bitmaps (one bit per grid cell) plus an exclusive
//                 popcount prefix per 32-bit word, for the coarse, the medium and the derived
//                 fine level (all by CTA 3): fine = 1 - up2(medium) - up4(coarse).
// Stage 2 (unpack_assemble_kernel, one thread per fine token): rank of the token's cell among
//   the set cells of its level = prefix[word] + popc(bits below) -> the decoded symbol that the
//   reference's masked assignment puts there (row-major order); ind = fine + up2(medium) +
//   up4(coarse); quant = codebook[ind] written NCHW; masks written as int64 like the reference.
#include <algorithm>

#include <cooperative_groups.h>

#include "common.cuh"

CGIC_TRACE_DECL(unpack)
CGIC_TRACE_DECL(assemble)

namespace cgic {
namespace {

constexpr int UP_THREADS = 512;  // == DEC_THREADS

struct UnpackWs {
    uint16_t *sym;     // [B][n16 + n8 + n4]
    int32_t *count;    // [B][3]  decoded symbols per index stream (-1: empty stream)
    int32_t *pop;      // [B][3]  population of each mask level
    int32_t *flag;     // [B][5]  per decode CTA: 0 or CGIC_EFORMAT (no zeroing needed: every CTA writes its slot)
    unsigned long long *chain;  // [B][3][max_chunks][chain_row] chunk hand-over tables (zero between launches)
    int32_t *ticket;            // [B][3] CTAs of a chained stream that are done (zero between launches)
    int max_chunks, chain_row;
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

// chunk records per stream: enough for a stream of `n4` symbols of 64 bits each in chunks of 128 subsequences
__host__ __device__ inline int unpack_max_chunks(const Geo &g) { return (int)(g.n4 * 64 / (128 * 128)) + 2; }
constexpr int DEC_SINGLE_N4 = 4096;  // token grids up to this size: one CTA per stream (no hand-over tables)
constexpr int DEC_CHAIN_D = 32;      // streams are spread over several CTAs for codes of at most this many bits
constexpr int DEC_CHAIN_SLOTS = 16;  // at most this many CTAs per stream

UnpackWs carve_unpack(void *ws, int B, const Geo &g)
{
    auto up = [](size_t v) { return (v + 255) / 256 * 256; };
    UnpackWs c{};
    unsigned char *p = static_cast<unsigned char *>(ws);
    size_t o = 0;
    c.max_chunks = unpack_max_chunks(g);
    c.chain_row = g.n4 > DEC_SINGLE_N4 ? DEC_CHAIN_D : 1;       // only large token grids spread a stream over several CTAs
    c.chain = reinterpret_cast<unsigned long long *>(p + o);  // first: the part of the workspace that must start zeroed
    o += up((size_t)B * 3 * c.max_chunks * c.chain_row * 8);
    c.ticket = reinterpret_cast<int32_t *>(p + o);
    o += up((size_t)B * 3 * 4);
    c.count = reinterpret_cast<int32_t *>(p + o);
    o += up((size_t)B * 3 * 4);
    c.pop = reinterpret_cast<int32_t *>(p + o);
    o += up((size_t)B * 3 * 4);
    c.flag = reinterpret_cast<int32_t *>(p + o);
    o += up((size_t)B * 5 * 4);
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
    int64_t slot_cap[5];
    int mode;
    Geo g;
    DevTable T;
    UnpackWs ws;
    const float *codebook;
    int64_t *mc_out, *mm_out, *mf_out, *ind_out;
    float *quant_out;
    int32_t *status;
    int mask_stage;  // 1: mask CTAs copy the mask streams to shared memory first
    int ch;          // subsequences per chunk of the candidate decoder (shared-memory budget)
    int nslots;      // CTAs per index stream (chunks dealt round-robin; hand-over tables in ws.chain)
};

// ---- parallel prefix decoder (whole CTA, one stream) ------------------------------------------
// A Huffman stream carries no synchronisation points, but prefix codes self-synchronise: a decoder
// started at a wrong bit falls into step with the true codeword sequence after a few codewords.
// The stream is cut into 128-bit subsequences, one per thread.  Every thread decodes the
// codewords that START inside its subsequence, first from a guessed start (the subsequence
// boundary), then again from the end position its left neighbour reports, until no start
// changes any more; thread 0 always starts from a known codeword boundary, so after k rounds
// the first k subsequences are exact whatever the data -- typically everything settles in 2-4
// rounds.  Counts are then prefix-summed and a last pass writes the symbols in stream order.
// Streams longer than one CTA-load of subsequences are processed chunk by chunk, the exact end
// of a chunk seeding the next.  Result = the reference's greedy decode (indices_coding.py:140-151),
// including its dropping of a trailing incomplete code.
constexpr int DEC_THREADS = 512;
constexpr int DEC_SUB_BITS = 128;
constexpr int64_t DEC_STOP = (int64_t)1 << 60;  // "decoding ended": beyond every subsequence

// bulk copy of the decode tables in flight (stage_decode_tables): wait() once, before the first look-up
struct TableStage {
    unsigned long long *mbar;
    bool pending;
    __device__ __forceinline__ void wait()
    {
        if (pending) {
            pending = false;
            mbar_wait(mbar, 0);
        }
    }
};

struct BitWindow {
    const uint8_t *in;  // 4-byte aligned stream start (the header byte is bit 0..7)
    int64_t nbytes;
    int64_t wi;         // index of the cached word pair
    uint32_t w0, w1;
    __device__ __forceinline__ uint32_t word(int64_t i) const
    {
        const int64_t b = i * 4;
        if (b + 4 <= nbytes) return __byte_perm(*reinterpret_cast<const uint32_t *>(in + b), 0, 0x0123);
        uint32_t v = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (b + k < nbytes) v |= (uint32_t)in[b + k] << (24 - 8 * k);
        return v;
    }
    // the 32 stream bits starting at bit q, MSB first
    __device__ __forceinline__ uint32_t peek(int64_t q)
    {
        const int64_t i = q >> 5;
        if (i != wi) {
            w0 = (i == wi + 1) ? w1 : word(i);
            w1 = word(i + 1);
            wi = i;
        }
        return __funnelshift_l(w1, w0, (int)(q & 31));
    }
};

// Decodes the codewords starting in [q, q_sub_end) (and before q_end = end of the payload).
// Returns their number; *q_next = start of the next codeword (DEC_STOP when decoding is over).
// s_lut: first-level table in shared memory; lut2: second-level tables (shared or global).
template <bool WRITE, typename Out>
__device__ __forceinline__ int decode_range(BitWindow &bw, int64_t q, int64_t q_sub_end, int64_t q_end, const uint32_t *s_lut,
                                            const uint32_t *lut2, const DevTable &T, Out *out, int64_t *q_next)
{
    int cnt = 0;
    const int L = T.lut_bits;
    while (q < q_sub_end) {
        if (q >= q_end) {
            q = DEC_STOP;
            break;
        }
        const uint32_t win = bw.peek(q);
        const uint32_t e = s_lut[win >> (32 - L)];
        int len = (int)(e & 0xFFu);
        int sym = (int)(e >> 8);
        if (len & 0x80) {
            if (len != 0xFF) {  // second-level table on the next `hgt` bits (L + hgt <= 20 window bits)
                const int hgt = len & 0x7F;
                const uint32_t e2 = lut2[sym + ((win << L) >> (32 - hgt))];
                sym = (int)(e2 >> 8);
                len = L + (int)(e2 & 0xFFu);
            } else {            // long code: walk the tree bit by bit
                int node = sym;
                int64_t qq = q + L;
                while (node >= T.K && qq < q_end) {
                    const uint32_t bit = bw.peek(qq) >> 31;
                    node = __ldg(&T.child[2 * node + (int)bit]);
                    ++qq;
                }
                if (node >= T.K) {  // trailing incomplete code: dropped, decoding ends
                    q = DEC_STOP;
                    break;
                }
                sym = node;
                len = (int)(qq - q);
            }
        }
        if (q + len > q_end) {  // completed only by bits past the payload: not a symbol
            q = DEC_STOP;
            break;
        }
        if (WRITE) out[cnt] = (Out)sym;
        ++cnt;
        q += len;
    }
    *q_next = q;
    return cnt;
}

// Whole CTA (DEC_THREADS threads).  Returns (same value in every thread) the number of symbols,
// -1 for an empty stream (the reference returns None), -2 when `cap` would overflow.
// `in` must be 4-byte aligned.
template <typename Out>
__device__ int decode_stream_cta(const uint8_t *in, int64_t nbytes, const DevTable &T, const uint32_t *s_lut, const uint32_t *lut2,
                                 Out *out, int64_t cap, TableStage &ts)
{
    ts.wait();
    __shared__ int64_t s_start[DEC_THREADS + 1];
    __shared__ int s_wsum[DEC_THREADS / 32];
    if (nbytes <= 0) return -1;
    const int tid = threadIdx.x;
    const int pad = in[0];
    int64_t nbits = (nbytes - 1) * 8 - pad;
    if (pad == 0 || nbits < 0) nbits = 0;  // text[:-0] is empty in the reference
    const int64_t q_end = 8 + nbits;
    BitWindow bw{in, nbytes, -2, 0u, 0u};
    int64_t chunk_q0 = 8;     // first bit of this chunk's first subsequence
    int64_t chunk_start = 8;  // exact start of the first codeword at or after chunk_q0
    int64_t total = 0;
    while (chunk_q0 < q_end && chunk_start < DEC_STOP) {
        const int64_t sub0 = chunk_q0 + (int64_t)tid * DEC_SUB_BITS;
        const int64_t sub1 = sub0 + DEC_SUB_BITS;
        const bool live = sub0 < q_end;
        int64_t my_start = tid == 0 ? chunk_start : sub0;
        int64_t my_end = my_start;
        int cnt = 0;
        if (live) cnt = decode_range<false, Out>(bw, my_start, sub1, q_end, s_lut, lut2, T, nullptr, &my_end);
        for (;;) {
            s_start[tid + 1] = my_end;
            __syncthreads();
            const int64_t s = tid == 0 ? chunk_start : s_start[tid];
            const bool changed = live && s != my_start;
            if (!__syncthreads_or(changed)) break;
            if (changed) {
                my_start = s;
                my_end = s;  // a start beyond this subsequence passes through unchanged
                cnt = decode_range<false, Out>(bw, my_start, sub1, q_end, s_lut, lut2, T, nullptr, &my_end);
            }
        }
        // exclusive scan of the counts
        const int lane = tid & 31, wid = tid >> 5;
        int inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += n;
        }
        if (lane == 31) s_wsum[wid] = inc;
        __syncthreads();
        int woff = 0, ctot = 0;
#pragma unroll
        for (int i = 0; i < DEC_THREADS / 32; ++i) {
            const int v = s_wsum[i];
            if (i < wid) woff += v;
            ctot += v;
        }
        if (total + ctot > cap) return -2;
        if (live && cnt) {
            int64_t dummy;
            decode_range<true, Out>(bw, my_start, sub1, q_end, s_lut, lut2, T, out + total + woff + inc - cnt, &dummy);
        }
        total += ctot;
        // dead threads hold my_end == their sub0 (no codeword can start there), so the last
        // slot is exact only when the last thread was live; otherwise decoding is finished
        const int64_t next_start = s_start[DEC_THREADS];
        __syncthreads();
        chunk_q0 += (int64_t)DEC_THREADS * DEC_SUB_BITS;
        chunk_start = next_start;
    }
    return (int)total;
}

// ---- candidate-start decoder (codes of at most DEC_MAX_D bits) --------------------------------
// Near-uniform code lengths (a trained codebook used evenly: lengths 9..12) barely
// self-synchronise, and the rounds above then degenerate into a serial walk.  The first codeword
// of a subsequence can only start at one of D = max_len offsets (the codeword straddling the
// boundary is at most max_len bits long), so instead of guessing:
//   A  every (subsequence, candidate offset) pair is decoded independently -> (end offset in the
//      next subsequence, number of codewords): a table f[sub][offset], D x redundant work but
//      perfectly parallel and data independent;
//   B  the true start offsets follow by composing the f[sub] along the stream -- blocks of
//      subsequences are composed per candidate in parallel, one thread chains the block results,
//      then every block replays its subsequences from its now known start;
//   C  counts are prefix-summed and each subsequence is decoded once more from its true start,
//      writing its symbols.
// Chunks of `ch` subsequences (shared-memory capacity) are processed in sequence.
constexpr int DEC_MAX_D = 128;       // candidate path handles max_len <= 128 (== DEC_SUB_BITS)
constexpr int DEC_MAX_CH = 1024;     // subsequences per chunk
constexpr int DEC_LOOKAHEAD_WORDS = 16;  // staged words past the chunk: straddling codeword + a 64-bit window
constexpr uint32_t DEC_OFF_STOP = 0xFF;

// Decoding from a chunk staged in shared memory as native-endian words (bit 31 of word 0 is stream
// bit 0 of the chunk; words past the stream are zero), 32-bit positions local to the chunk.
// One codeword at bit q: returns (sym << 8) | len.  Branch-free for one- and two-level codes (the
// second lookup is always issued, on entry 0 when unused); only codes deeper than the second
// level walk the tree.  LUT2S: the second-level tables are in shared memory too.
template <bool LUT2S>
__device__ __forceinline__ uint32_t decode_one(const uint32_t *s_words, uint32_t q, const uint32_t *s_lut, const uint32_t *lut2,
                                               const DevTable &T, int L)
{
    const uint32_t i = q >> 5;
    const uint32_t win = __funnelshift_l(s_words[i + 1], s_words[i], q & 31);
    const uint32_t e = s_lut[win >> (32 - L)];
    const uint32_t f = e & 0xFFu;
    if (f == 0xFFu) {  // long code (<= DEC_MAX_D bits): walk the tree; bits past the payload read as 0
        int node = (int)(e >> 8);
        uint32_t qq = q + L;
        while (node >= T.K) {
            const uint32_t bit = (s_words[qq >> 5] >> (31 - (qq & 31))) & 1u;
            node = __ldg(&T.child[2 * node + (int)bit]);
            ++qq;
        }
        return ((uint32_t)node << 8) | (qq - q);
    }
    const bool two = (f & 0x80u) != 0;
    const uint32_t hgt = two ? (f & 0x7Fu) : 1u;
    const uint32_t i2 = two ? (e >> 8) + ((win << L) >> (32 - hgt)) : 0u;
    const uint32_t e2 = LUT2S ? lut2[i2] : __ldg(&lut2[i2]);
    return two ? e2 + (uint32_t)L : e;  // second-level entries hold the bits used beyond L
}

// Codewords starting in [q, q_sub_end) of a chunk whose length table len8[] is ready (candidate decoder, phase C);
// returns their number.  A trailing incomplete code, or one completed only by bits past the payload, is dropped like in
// the reference: len8[] is 0 there.  The walk from codeword to codeword goes through len8[] alone -- one dependent
// shared-memory byte load per codeword instead of the whole window -> first-level -> second-level lookup chain --
// and the symbol lookups of four codewords at a time are independent of each other and of the next group's walk, so
// their latencies overlap.
template <bool LUT2S, typename Out>
__device__ __forceinline__ int decode_write_len(const uint32_t *s_words, const uint8_t *s_len, uint32_t q, uint32_t q_sub_end,
                                                const uint32_t *s_lut, const uint32_t *lut2, const DevTable &T, Out *out)
{
    constexpr int G = 4;
    constexpr uint32_t DEAD = 0xFFFFFFFFu;
    const int L = T.lut_bits;
    int cnt = 0;
    uint32_t qn[G];
    auto walk = [&]() {
#pragma unroll
        for (int k = 0; k < G; ++k) {
            const uint32_t len = q < q_sub_end ? (uint32_t)s_len[q] : 0u;
            qn[k] = len ? q : DEAD;
            q = len ? q + len : q_sub_end;  // frozen once the subsequence (or the payload) is over
        }
    };
    walk();
    while (qn[0] != DEAD) {
        uint32_t qc[G], r[G];
#pragma unroll
        for (int k = 0; k < G; ++k) qc[k] = qn[k];
        walk();
        bool deep = false;
#pragma unroll
        for (int k = 0; k < G; ++k) {
            const uint32_t pos = qc[k] != DEAD ? qc[k] : 8u;
            const uint32_t i = pos >> 5;
            const uint32_t win = __funnelshift_l(s_words[i + 1], s_words[i], pos & 31);
            const uint32_t e = s_lut[win >> (32 - L)];
            const uint32_t f = e & 0xFFu;
            const bool two = (f & 0x80u) != 0 && f != 0xFFu;
            const uint32_t hgt = two ? (f & 0x7Fu) : 1u;
            const uint32_t i2 = two ? (e >> 8) + ((win << L) >> (32 - hgt)) : 0u;
            const uint32_t e2 = LUT2S ? lut2[i2] : __ldg(&lut2[i2]);
            r[k] = two ? e2 : e;
            deep |= f == 0xFFu && qc[k] != DEAD;
        }
        if (deep) {  // rare: a code deeper than the second level
#pragma unroll
            for (int k = 0; k < G; ++k)
                if (qc[k] != DEAD) r[k] = decode_one<LUT2S>(s_words, qc[k], s_lut, lut2, T, L);
        }
#pragma unroll
        for (int k = 0; k < G; ++k)
            if (qc[k] != DEAD) out[cnt + k] = (Out)(r[k] >> 8);
#pragma unroll
        for (int k = 0; k < G; ++k) cnt += qc[k] != DEAD;
    }
    return cnt;
}

// Chunks of one stream may be spread over `nslots` CTAs (slot j takes chunks j, j + nslots, ...).  What a chunk
// needs from its predecessors -- the offset of its first codeword and the number of symbols decoded before it -- does
// NOT travel down a chain of CTAs (one global round trip per chunk): every chunk publishes what it does to EVERY possible
// start offset, a table of D 64-bit words in global memory (bit 63 = valid, bits 32..39 = start offset of the next chunk,
// bits 0..31 = symbols of this chunk), as soon as its own phase A is done; a chunk then gathers the tables of the
// predecessors it has not folded in yet (at most nslots - 1) and composes them locally.  All chunks of a stream thus run
// at the same time and the serial part is one publish + one gather, whatever the number of chunks.
// The last CTA of the stream to finish (ticket) zeroes the tables again (workspace contract: zero between launches).
struct DecChain {
    unsigned long long *rec;  // [max chunks][row] tables of this stream; null = single CTA per stream
    int32_t *ticket;          // CTAs of this stream that are done
    int slot, nslots, row;
};
constexpr int DEC_NOT_MINE = -2147483647 - 1;  // returned by the CTAs that did not decode a stream's last chunk

__device__ __forceinline__ unsigned long long chain_wait(const unsigned long long *p)
{
    unsigned long long v;
    do {
        asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
        if (!(v >> 63)) __nanosleep(20);
    } while (!(v >> 63));
    return v;
}
__device__ __forceinline__ void chain_publish(unsigned long long *p, uint32_t next_start, uint32_t symbols)
{
    const unsigned long long v = (1ull << 63) | ((unsigned long long)(next_start & 0xFFu) << 32) | symbols;
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Phase A of the candidate decoder: NCH chains per thread in lock step (each step is one dependent
// shared-memory load, no branches); pairs tid, tid + DEC_THREADS, ... ; pair = subsequence * D + candidate offset.
template <int NCH>
__device__ __forceinline__ void follow_chains(const uint8_t *s_len, uint16_t *s_fn, int npairs, int D, int tid)
{
    for (int pair0 = tid; pair0 < npairs; pair0 += NCH * DEC_THREADS) {
        uint32_t q[NCH], se[NCH], n[NCH];
#pragma unroll
        for (int k = 0; k < NCH; ++k) {
            const int pair = pair0 + k * DEC_THREADS;
            const bool ok = pair < npairs;
            const int i = ok ? pair / D : 0, c = ok ? pair - i * D : 0;
            se[k] = ok ? 8u + (uint32_t)(i + 1) * DEC_SUB_BITS : 0u;  // dead slot: q >= se from the start
            q[k] = 8u + (uint32_t)i * DEC_SUB_BITS + (uint32_t)c;
            n[k] = 0;
        }
        for (;;) {
            uint32_t len[NCH];
#pragma unroll
            for (int k = 0; k < NCH; ++k) len[k] = s_len[q[k]];
            uint32_t moved = 0;
#pragma unroll
            for (int k = 0; k < NCH; ++k) {
                const uint32_t adv = q[k] < se[k] ? len[k] : 0u;  // 0 also where decoding ends inside the subsequence
                q[k] += adv;
                n[k] += adv != 0u;
                moved |= adv;
            }
            if (!moved) break;
        }
#pragma unroll
        for (int k = 0; k < NCH; ++k) {
            const int pair = pair0 + k * DEC_THREADS;
            // a chain that stopped before the subsequence end met a position without a complete codeword: decoding is over
            if (pair < npairs) s_fn[pair] = (uint16_t)(((q[k] < se[k] ? DEC_OFF_STOP : q[k] - se[k]) << 8) | n[k]);
        }
    }
}

// dynamic shared memory behind the staged tables: the chunk's words, len8[] (code length at every
// bit position of the chunk, 0 = no complete codeword before the payload end), f[ch * D] (uint16)
template <bool LUT2S, typename Out>
__device__ int decode_stream_cta_cand(const uint8_t *in, int64_t nbytes, const DevTable &T, const uint32_t *s_lut, const uint32_t *lut2,
                                      uint32_t *s_words, uint8_t *s_len, uint16_t *s_fn, int ch, Out *out, int64_t cap, TableStage &ts)
{
    __shared__ uint8_t s_blockfn[32 * DEC_MAX_D];
    __shared__ uint8_t s_blkstart[32];
    __shared__ uint8_t s_substart[DEC_MAX_CH];
    __shared__ int s_wsum[DEC_THREADS / 32];
    __shared__ uint32_t s_next;
    if (nbytes <= 0) return -1;
    const int tid = threadIdx.x;
    const int pad = in[0];  // (consumed after the first chunk's loads have been issued: one global round trip, not two)
    const int64_t nsub_ub = ((nbytes - 1) * 8 + DEC_SUB_BITS - 1) / DEC_SUB_BITS;  // the pad byte takes at most one subsequence off
    const int D = T.max_len, L = T.lut_bits;
    uint32_t start_off = 0;
    int64_t total = 0;
    int64_t nbits = 0, q_end_abs = 8, nsub_total = nsub_ub;
    for (int64_t c0 = 0; c0 < nsub_total && start_off != DEC_OFF_STOP; c0 += ch) {
        CGIC_STAMP(unpack, 2);
        // ---- stage the chunk: bytes [c0 * 16, ...) of the stream as big-endian words, zero past the end
        {
            const int64_t byte0 = c0 * (DEC_SUB_BITS / 8);
            const int nw = (int)min((int64_t)ch, nsub_total - c0) * (DEC_SUB_BITS / 32) + DEC_LOOKAHEAD_WORDS;
            for (int wi = tid; wi < nw; wi += DEC_THREADS) {
                const int64_t bo = byte0 + (int64_t)wi * 4;
                uint32_t v = 0;
                if (bo + 4 <= nbytes) {
                    v = __byte_perm(__ldg(reinterpret_cast<const uint32_t *>(in + bo)), 0, 0x0123);
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (bo + k < nbytes) v |= (uint32_t)in[bo + k] << (24 - 8 * k);
                }
                s_words[wi] = v;
            }
        }
        if (c0 == 0) {  // now the header byte: payload bits, exact number of subsequences
            nbits = (nbytes - 1) * 8 - pad;
            if (pad == 0 || nbits < 0) nbits = 0;
            q_end_abs = 8 + nbits;  // stream bit coordinates (header byte = bits 0..7)
            nsub_total = (nbits + DEC_SUB_BITS - 1) / DEC_SUB_BITS;
            if (nsub_total == 0) break;
        }
        const int nsub = (int)min((int64_t)ch, nsub_total - c0);
        const int npos_words = nsub * (DEC_SUB_BITS / 32) + (DEC_MAX_D + 8 + 31) / 32;  // positions a chain can visit
        ts.wait();  // (first chunk: the tables' bulk copy ran beside the loads above)
        __syncthreads();
        const int64_t rel_end = q_end_abs - c0 * DEC_SUB_BITS;  // payload end, local to the chunk
        const uint32_t q_end = (uint32_t)min(rel_end, (int64_t)(nsub * DEC_SUB_BITS + 8 + DEC_MAX_D + 64));
        CGIC_STAMP(unpack, 3);
        // ---- A0: code length at EVERY bit position (independent lookups: 32 positions per thread and
        //      pass sharing one word pair); 0 where the codeword would run past the payload.  Fast path:
        //      the first-level entry's low byte IS the length; positions whose entry carries the
        //      second-level / tree-walk flag (bit 7), and the few words near the payload end, are
        //      resolved by a second sweep with the full decode.
        for (int it = tid; it < npos_words * 4; it += DEC_THREADS) {  // work item: 8 positions of one word
            const int wi = it >> 2, qtr = it & 3;
            const uint32_t w0 = s_words[wi], w1 = s_words[wi + 1];
            uint32_t *dst = reinterpret_cast<uint32_t *>(s_len) + wi * 8 + qtr * 2;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int p0 = qtr * 8 + j * 4;
                uint32_t f[4], e[4], win[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    win[k] = __funnelshift_l(w1, w0, p0 + k);
                    e[k] = s_lut[win[k] >> (32 - L)];
                    f[k] = e[k] & 0xFFu;
                }
                if (LUT2S) {
                    // second-level codes without a branch (a warp covers 128 positions per pass: "some lane needs the second
                    // table" is the normal case even at a 1 % share of two-level codes); only flag 0xFF walks the tree
                    uint32_t e2[4], deep = 0;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const bool two = f[k] >= 0x80u && f[k] != 0xFFu;
                        const uint32_t hgt = two ? (f[k] & 0x7Fu) : 1u;
                        e2[k] = lut2[two ? (e[k] >> 8) + ((win[k] << L) >> (32 - hgt)) : 0u];
                        deep |= f[k] == 0xFFu;
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (f[k] >= 0x80u && f[k] != 0xFFu) f[k] = (e2[k] & 0xFFu) + (uint32_t)L;
                    if (deep) {
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            if (f[k] == 0xFFu) {
                                f[k] = decode_one<LUT2S>(s_words, (uint32_t)wi * 32u + (uint32_t)(p0 + k), s_lut, lut2, T, L) & 0xFFu;
                                if (f[k] == 0xFFu) f[k] = 0;  // (DEC_MAX_D = 128 < 0xFF)
                            }
                    }
                } else if ((f[0] | f[1] | f[2] | f[3]) & 0x80u) {  // second-level table in global memory / tree walk
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (f[k] & 0x80u) {
                            if (f[k] != 0xFFu) {
                                const uint32_t hgt = f[k] & 0x7Fu;
                                const uint32_t i2 = (e[k] >> 8) + ((win[k] << L) >> (32 - hgt));  // win holds 32 bits from the position on
                                f[k] = (__ldg(&lut2[i2]) & 0xFFu) + (uint32_t)L;
                            } else {
                                f[k] = decode_one<LUT2S>(s_words, (uint32_t)wi * 32u + (uint32_t)(p0 + k), s_lut, lut2, T, L) & 0xFFu;
                                if (f[k] == 0xFFu) f[k] = 0;  // (DEC_MAX_D = 128 < 0xFF)
                            }
                        }
                }
                dst[j] = f[0] | (f[1] << 8) | (f[2] << 16) | (f[3] << 24);
            }
        }
        __syncthreads();
        // no codeword may run past the payload: only positions within D bits of its end can
        for (uint32_t q = (q_end > (uint32_t)D ? q_end - (uint32_t)D : 0u) + (uint32_t)tid; q < (uint32_t)npos_words * 32u; q += DEC_THREADS)
            if (q + s_len[q] > q_end) s_len[q] = 0;
        __syncthreads();
        CGIC_STAMP(unpack, 7);
        // ---- A: every (subsequence, candidate offset) pair follows its chain through len8[];
        //      four chains per thread in lock step (each step is one dependent shared-memory load, no branches)
        {
            const int npairs = nsub * D;
            const int per = (npairs + DEC_THREADS - 1) / DEC_THREADS;
            if (per <= 1) follow_chains<1>(s_len, s_fn, npairs, D, tid);
            else if (per == 2) follow_chains<2>(s_len, s_fn, npairs, D, tid);
            else if (per == 3) follow_chains<3>(s_len, s_fn, npairs, D, tid);
            else follow_chains<4>(s_len, s_fn, npairs, D, tid);
        }
        __syncthreads();
        CGIC_STAMP(unpack, 4);
        // ---- B: true start offset of every subsequence
        const int bs = nsub <= 128 ? 8 : 32;  // subsequences per block; at most 32 blocks either way
        const int nblk = (nsub + bs - 1) / bs;
        for (int item = tid; item < nblk * D; item += DEC_THREADS) {
            const int blk = item / D, c = item - blk * D;
            uint32_t x = (uint32_t)c;
            const int j1 = min(nsub, (blk + 1) * bs);
            for (int j = blk * bs; j < j1 && x != DEC_OFF_STOP; ++j) x = s_fn[j * D + x] >> 8;
            s_blockfn[item] = (uint8_t)x;
        }
        __syncthreads();
        if (tid == 0) {
            uint32_t x = start_off;
            for (int blk = 0; blk < nblk; ++blk) {
                s_blkstart[blk] = (uint8_t)x;
                if (x != DEC_OFF_STOP) x = s_blockfn[blk * D + x];
            }
            s_next = x;
        }
        __syncthreads();
        if (tid < nblk) {
            uint32_t x = s_blkstart[tid];
            const int j1 = min(nsub, (tid + 1) * bs);
            for (int j = tid * bs; j < j1; ++j) {
                s_substart[j] = (uint8_t)x;
                if (x != DEC_OFF_STOP) x = s_fn[j * D + x] >> 8;
            }
        }
        __syncthreads();
        CGIC_STAMP(unpack, 5);
        // ---- C: counts -> offsets -> symbols
        for (int base = 0; base < nsub; base += DEC_THREADS) {
            const int i = base + tid;
            uint32_t st = DEC_OFF_STOP;
            int cnt = 0;
            if (i < nsub) {
                st = s_substart[i];
                if (st != DEC_OFF_STOP) cnt = s_fn[i * D + st] & 0xFF;
            }
            const int lane = tid & 31, wid = tid >> 5;
            int inc = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int n = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += n;
            }
            if (lane == 31) s_wsum[wid] = inc;
            __syncthreads();
            int woff = 0, ctot = 0;
#pragma unroll
            for (int k = 0; k < DEC_THREADS / 32; ++k) {
                const int v = s_wsum[k];
                if (k < wid) woff += v;
                ctot += v;
            }
            if (total + ctot > cap) return -2;
            if (cnt) {
                const uint32_t sub0 = 8u + (uint32_t)i * DEC_SUB_BITS;
                decode_write_len<LUT2S, Out>(s_words, s_len, sub0 + st, sub0 + DEC_SUB_BITS, s_lut, lut2, T, out + total + woff + inc - cnt);
            }
            total += ctot;
            __syncthreads();
        }
        start_off = s_next;
    }
    return (int)total;
}

// The same decoder for streams whose chunks are spread over several CTAs (MULTI) -- kept as a separate function so
// that the single-CTA decoder above stays exactly the code that was tuned on the small-grid batch.
// Chunk size and number of chunks of a chained stream of `nsub_total` subsequences: one chunk per CTA whenever the stream
// allows it (all its chunks are then decoded at the same time and only the gather of the tables is serial), but not
// below 64 subsequences -- a chunk has fixed costs (staging, barriers) that smaller ones no longer repay.  Every CTA of
// the stream (and the kernel, for the CTAs that get no chunk) computes the same plan from the stream's length.
__device__ __forceinline__ int64_t chain_plan(int64_t nsub_total, int nslots, int &ch)
{
    const int64_t per = (nsub_total + nslots - 1) / nslots;
    const int64_t fit = per < 64 ? 64 : ((per + 31) & ~(int64_t)31);
    if (fit < ch) ch = (int)fit;
    return (nsub_total + ch - 1) / ch;
}
__device__ __forceinline__ int64_t stream_subsequences(const uint8_t *in, int64_t nbytes)
{
    const int pad = in[0];
    int64_t nbits = (nbytes - 1) * 8 - pad;
    if (pad == 0 || nbits < 0) nbits = 0;
    return (nbits + DEC_SUB_BITS - 1) / DEC_SUB_BITS;
}

// Shared memory of decode_stream_cta_chain, ONE instance per kernel (the two table-shape instantiations of the function share
// it).  Chained streams (MULTI) carry codes of at most DEC_CHAIN_D bits and chunks of at most DEC_CHAIN_CH subsequences, so
// their kernel's footprint stays small enough for four CTAs per SM.
constexpr int DEC_CHAIN_CH = 128;
template <bool MULTI>
struct ChainShared {
    static constexpr int MAXD = MULTI ? DEC_CHAIN_D : DEC_MAX_D;
    uint32_t aggn[MULTI ? MAXD : 1];  // chained streams: symbols of the chunk per candidate start
    uint32_t tabn[MULTI ? (DEC_CHAIN_SLOTS - 1) * DEC_CHAIN_D : 1];  // tables of the predecessors not folded in yet
    uint32_t next, start, base, res_st, res_base;  // res_*: state after the last chunk this CTA has resolved
    int wsum[DEC_THREADS / 32];
    uint16_t blockcnt[MULTI ? 32 * MAXD : 1];  // chained streams: symbols of a block per candidate start
    uint8_t blockfn[32 * MAXD];
    uint8_t aggx[MULTI ? MAXD : 1];   // chained streams: chunk exit offset per candidate start
    uint8_t tabx[MULTI ? (DEC_CHAIN_SLOTS - 1) * DEC_CHAIN_D : 1];
    uint8_t blkstart[32];
    uint8_t substart[MULTI ? DEC_CHAIN_CH : DEC_MAX_CH];
    bool sweep;
};

template <bool LUT2S, bool MULTI, typename Out>
__device__ int decode_stream_cta_chain(const uint8_t *in, int64_t nbytes, const DevTable &T, const uint32_t *s_lut, const uint32_t *lut2,
                                      uint32_t *s_words, uint8_t *s_len, uint16_t *s_fn, int ch, Out *out, int64_t cap, const DecChain chain,
                                      TableStage &ts, ChainShared<MULTI> &sh)
{
    uint8_t *s_blockfn = sh.blockfn, *s_aggx = sh.aggx, *s_blkstart = sh.blkstart, *s_substart = sh.substart, *s_tabx = sh.tabx;
    uint16_t *s_blockcnt = sh.blockcnt;
    uint32_t *s_aggn = sh.aggn, *s_tabn = sh.tabn;
    int *s_wsum = sh.wsum;
    uint32_t &s_next = sh.next, &s_start = sh.start, &s_base = sh.base, &s_res_st = sh.res_st, &s_res_base = sh.res_base;
    bool &s_sweep = sh.sweep;
    constexpr bool multi = MULTI;
    bool overflow = false;
    if (nbytes <= 0) return chain.slot == 0 ? -1 : DEC_NOT_MINE;
    const int tid = threadIdx.x;
    const int pad = in[0];
    int64_t nbits = (nbytes - 1) * 8 - pad;
    if (pad == 0 || nbits < 0) nbits = 0;
    const int64_t q_end_abs = 8 + nbits;  // stream bit coordinates (header byte = bits 0..7)
    const int D = T.max_len, L = T.lut_bits;
    const int64_t nsub_total = (nbits + DEC_SUB_BITS - 1) / DEC_SUB_BITS;
    uint32_t start_off = 0;
    int64_t total = 0;
    const int nslots = multi ? chain.nslots : 1, slot = multi ? chain.slot : 0;
    const bool tables = multi && nslots > 1;  // (the predicate is uniform over the CTAs of a stream)
    // (the shared-memory carve-up is the one of the largest chunk)
    const int64_t nchunks = tables ? chain_plan(nsub_total, nslots, ch) : (nsub_total + ch - 1) / ch;
    int64_t res_chunk = -1;                   // last chunk whose outgoing state s_res_* holds
    if (tables && threadIdx.x == 0) {
        s_res_st = 0u;
        s_res_base = 0u;
    }
    // the CTA that decodes the last chunk reports the stream's symbol count (slot 0 when there is no chunk at all)
    const bool mine_last = nchunks == 0 ? slot == 0 : (int)((nchunks - 1) % nslots) == slot;
    for (int64_t chunk = slot; chunk < nchunks && start_off != DEC_OFF_STOP; chunk += nslots) {
        const int64_t c0 = chunk * ch;
        const int nsub = (int)min((int64_t)ch, nsub_total - c0);
        CGIC_STAMP(unpack, 2);
        // ---- stage the chunk: bytes [c0 * 16, ...) of the stream as big-endian words, zero past the end
        const int npos_words = nsub * (DEC_SUB_BITS / 32) + (DEC_MAX_D + 8 + 31) / 32;  // positions a chain can visit
        {
            const int64_t byte0 = c0 * (DEC_SUB_BITS / 8);
            const int nw = nsub * (DEC_SUB_BITS / 32) + DEC_LOOKAHEAD_WORDS;
            for (int wi = tid; wi < nw; wi += DEC_THREADS) {
                const int64_t bo = byte0 + (int64_t)wi * 4;
                uint32_t v = 0;
                if (bo + 4 <= nbytes) {
                    v = __byte_perm(__ldg(reinterpret_cast<const uint32_t *>(in + bo)), 0, 0x0123);
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (bo + k < nbytes) v |= (uint32_t)in[bo + k] << (24 - 8 * k);
                }
                s_words[wi] = v;
            }
        }
        ts.wait();  // (first chunk: the tables' bulk copy ran beside the loads above)
        __syncthreads();
        const int64_t rel_end = q_end_abs - c0 * DEC_SUB_BITS;  // payload end, local to the chunk
        const uint32_t q_end = (uint32_t)min(rel_end, (int64_t)(nsub * DEC_SUB_BITS + 8 + DEC_MAX_D + 64));
        CGIC_STAMP(unpack, 3);
        // ---- A0: code length at EVERY bit position (independent lookups: 32 positions per thread and
        //      pass sharing one word pair); 0 where the codeword would run past the payload.  Fast path:
        //      the first-level entry's low byte IS the length; positions whose entry carries the
        //      second-level / tree-walk flag (bit 7), and the few words near the payload end, are
        //      resolved by a second sweep with the full decode.
        for (int it = tid; it < npos_words * 4; it += DEC_THREADS) {  // work item: 8 positions of one word
            const int wi = it >> 2, qtr = it & 3;
            const uint32_t w0 = s_words[wi], w1 = s_words[wi + 1];
            uint32_t *dst = reinterpret_cast<uint32_t *>(s_len) + wi * 8 + qtr * 2;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int p0 = qtr * 8 + j * 4;
                uint32_t f[4], e[4], win[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    win[k] = __funnelshift_l(w1, w0, p0 + k);
                    e[k] = s_lut[win[k] >> (32 - L)];
                    f[k] = e[k] & 0xFFu;
                }
                if (LUT2S) {
                    // second-level codes without a branch (a warp covers 128 positions per pass: "some lane needs the second
                    // table" is the normal case even at a 1 % share of two-level codes); only flag 0xFF walks the tree
                    uint32_t e2[4], deep = 0;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const bool two = f[k] >= 0x80u && f[k] != 0xFFu;
                        const uint32_t hgt = two ? (f[k] & 0x7Fu) : 1u;
                        e2[k] = lut2[two ? (e[k] >> 8) + ((win[k] << L) >> (32 - hgt)) : 0u];
                        deep |= f[k] == 0xFFu;
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (f[k] >= 0x80u && f[k] != 0xFFu) f[k] = (e2[k] & 0xFFu) + (uint32_t)L;
                    if (deep) {
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            if (f[k] == 0xFFu) {
                                f[k] = decode_one<LUT2S>(s_words, (uint32_t)wi * 32u + (uint32_t)(p0 + k), s_lut, lut2, T, L) & 0xFFu;
                                if (f[k] == 0xFFu) f[k] = 0;  // (DEC_MAX_D = 128 < 0xFF)
                            }
                    }
                } else if ((f[0] | f[1] | f[2] | f[3]) & 0x80u) {  // second-level table in global memory / tree walk
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (f[k] & 0x80u) {
                            if (f[k] != 0xFFu) {
                                const uint32_t hgt = f[k] & 0x7Fu;
                                const uint32_t i2 = (e[k] >> 8) + ((win[k] << L) >> (32 - hgt));  // win holds 32 bits from the position on
                                f[k] = (__ldg(&lut2[i2]) & 0xFFu) + (uint32_t)L;
                            } else {
                                f[k] = decode_one<LUT2S>(s_words, (uint32_t)wi * 32u + (uint32_t)(p0 + k), s_lut, lut2, T, L) & 0xFFu;
                                if (f[k] == 0xFFu) f[k] = 0;  // (DEC_MAX_D = 128 < 0xFF)
                            }
                        }
                }
                dst[j] = f[0] | (f[1] << 8) | (f[2] << 16) | (f[3] << 24);
            }
        }
        __syncthreads();
        // no codeword may run past the payload: only positions within D bits of its end can
        for (uint32_t q = (q_end > (uint32_t)D ? q_end - (uint32_t)D : 0u) + (uint32_t)tid; q < (uint32_t)npos_words * 32u; q += DEC_THREADS)
            if (q + s_len[q] > q_end) s_len[q] = 0;
        __syncthreads();
        CGIC_STAMP(unpack, 7);
        // ---- A: every (subsequence, candidate offset) pair follows its chain through len8[];
        //      four chains per thread in lock step (each step is one dependent shared-memory load, no branches)
        {
            const int npairs = nsub * D;
            const int per = (npairs + DEC_THREADS - 1) / DEC_THREADS;
            if (per <= 1) follow_chains<1>(s_len, s_fn, npairs, D, tid);
            else if (per == 2) follow_chains<2>(s_len, s_fn, npairs, D, tid);
            else if (per == 3) follow_chains<3>(s_len, s_fn, npairs, D, tid);
            else follow_chains<4>(s_len, s_fn, npairs, D, tid);
        }
        __syncthreads();
        CGIC_STAMP(unpack, 4);
        // ---- B: true start offset of every subsequence
        const int bs = nsub <= 128 ? 8 : 32;  // subsequences per block; at most 32 blocks either way
        const int nblk = (nsub + bs - 1) / bs;
        for (int item = tid; item < nblk * D; item += DEC_THREADS) {
            const int blk = item / D, c = item - blk * D;
            uint32_t x = (uint32_t)c;
            const int j1 = min(nsub, (blk + 1) * bs);
            uint32_t nsym = 0;
            for (int j = blk * bs; j < j1 && x != DEC_OFF_STOP; ++j) {
                const uint32_t e = s_fn[j * D + x];
                nsym += e & 0xFFu;
                x = e >> 8;
            }
            s_blockfn[item] = (uint8_t)x;
            if (multi) s_blockcnt[item] = (uint16_t)nsym;
        }
        __syncthreads();
        if (multi) {
            // what this chunk does to EVERY possible start offset, ready before the predecessor's answer arrives
            if (tid < D) {
                uint32_t x = (uint32_t)tid, nsym = 0;
                for (int blk = 0; blk < nblk && x != DEC_OFF_STOP; ++blk) {
                    nsym += s_blockcnt[blk * D + x];
                    x = s_blockfn[blk * D + x];
                }
                s_aggx[tid] = (uint8_t)x;
                s_aggn[tid] = nsym;
            }
            __syncthreads();
            if (tables) {
                // this chunk's table first (nobody waits for us longer than necessary) ...
                if (tid < D) chain_publish(chain.rec + chunk * chain.row + tid, s_aggx[tid], s_aggn[tid]);
                // ... then the tables of chunks res_chunk + 1 .. chunk - 1 (at most nslots - 1 of them)
                const int np = (int)(chunk - 1 - res_chunk);
                for (int item = tid; item < np * D; item += DEC_THREADS) {
                    const int p = item / D, d = item - p * D;
                    const unsigned long long r = chain_wait(chain.rec + (res_chunk + 1 + p) * chain.row + d);
                    s_tabx[item] = (uint8_t)(r >> 32);
                    s_tabn[item] = (uint32_t)r;
                }
                __syncthreads();
                if (tid == 0) {
                    uint32_t st = s_res_st, base = s_res_base;
                    for (int p = 0; p < np && st != DEC_OFF_STOP; ++p) {
                        base += s_tabn[p * D + st];
                        st = s_tabx[p * D + st];
                    }
                    s_start = st;
                    s_base = base;
                    s_res_st = st == DEC_OFF_STOP ? DEC_OFF_STOP : s_aggx[st];
                    s_res_base = base + (st == DEC_OFF_STOP ? 0u : s_aggn[st]);
                }
                res_chunk = chunk;
            } else if (tid == 0) {
                s_start = start_off;
                s_base = (uint32_t)total;
            }
            __syncthreads();
            start_off = s_start;
            total = s_base;
            if (start_off == DEC_OFF_STOP) {  // decoding ended in an earlier chunk: nothing here (and nothing after)
                start_off = 0;                // (keeps the loop going: this slot's later chunks must still publish their tables)
                continue;
            }
        }
        if (tid == 0) {
            uint32_t x = start_off;
            for (int blk = 0; blk < nblk; ++blk) {
                s_blkstart[blk] = (uint8_t)x;
                if (x != DEC_OFF_STOP) x = s_blockfn[blk * D + x];
            }
            s_next = x;
        }
        __syncthreads();
        if (tid < nblk) {
            uint32_t x = s_blkstart[tid];
            const int j1 = min(nsub, (tid + 1) * bs);
            for (int j = tid * bs; j < j1; ++j) {
                s_substart[j] = (uint8_t)x;
                if (x != DEC_OFF_STOP) x = s_fn[j * D + x] >> 8;
            }
        }
        __syncthreads();
        CGIC_STAMP(unpack, 5);
        // ---- C: counts -> offsets -> symbols
        for (int base = 0; base < nsub; base += DEC_THREADS) {
            const int i = base + tid;
            uint32_t st = DEC_OFF_STOP;
            int cnt = 0;
            if (i < nsub) {
                st = s_substart[i];
                if (st != DEC_OFF_STOP) cnt = s_fn[i * D + st] & 0xFF;
            }
            const int lane = tid & 31, wid = tid >> 5;
            int inc = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int n = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += n;
            }
            if (lane == 31) s_wsum[wid] = inc;
            __syncthreads();
            int woff = 0, ctot = 0;
#pragma unroll
            for (int k = 0; k < DEC_THREADS / 32; ++k) {
                const int v = s_wsum[k];
                if (k < wid) woff += v;
                ctot += v;
            }
            if (total + ctot > cap) {
                if (!multi) return -2;
                overflow = true;  // a chained CTA keeps going (later chunks must still publish); the stream's last CTA sees the overflow too
            }
            if (cnt && !overflow) {
                const uint32_t sub0 = 8u + (uint32_t)i * DEC_SUB_BITS;
                decode_write_len<LUT2S, Out>(s_words, s_len, sub0 + st, sub0 + DEC_SUB_BITS, s_lut, lut2, T, out + total + woff + inc - cnt);
            }
            total += ctot;
            __syncthreads();
        }
        start_off = tables ? 0u : s_next;
    }
    if (tables && nchunks > 0) {
        // the stream's last CTA to get here zeroes the hand-over tables for the next launch (every CTA of the stream is
        // past its last gather when it takes its ticket)
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            s_sweep = atomicAdd(chain.ticket, 1) == (int)min((int64_t)nslots, nchunks) - 1;  // (CTAs without a chunk never get here)
        }
        __syncthreads();
        if (s_sweep) {
            __threadfence();
            for (int64_t i = tid; i < nchunks * chain.row; i += DEC_THREADS) chain.rec[i] = 0ull;
            if (tid == 0) *chain.ticket = 0;
        }
    }
    if (overflow) return -2;
    return mine_last ? (int)total : DEC_NOT_MINE;
}

// subsequences per chunk for a table: words 16 B + len8[] 128 B + f[] 2 D bytes per subsequence, <= 48 KB
// Small token grids (<= 4096 fine tokens: the typical stream fits 128 subsequences) take half the budget so
// that three decode CTAs fit one SM and the whole (5, B) grid of a 64-image batch is resident at once.
__host__ __device__ inline int cand_chunk_subs(int max_len, int64_t n4 = (int64_t)1 << 40)
{
    int ch = (48 * 1024 / (16 + DEC_SUB_BITS + max_len * 2)) & ~31;
    ch = ch > DEC_MAX_CH ? DEC_MAX_CH : ch;
    if (n4 <= 4096 && ch > 128) ch = 128;
    return ch;
}
__host__ __device__ inline int cand_words(int ch) { return ch * (DEC_SUB_BITS / 32) + DEC_LOOKAHEAD_WORDS; }
__host__ __device__ inline int cand_len_bytes(int ch) { return (ch * (DEC_SUB_BITS / 32) + (DEC_MAX_D + 8 + 31) / 32 + 1) * 32; }

template <typename Out, bool MULTI = false>
__device__ __forceinline__ int decode_stream_any(const uint8_t *in, int64_t nbytes, const DevTable &T, const uint32_t *s_dec,
                                                 const uint32_t *lut2, Out *out, int64_t cap, int ch, TableStage &ts,
                                                 const DecChain chain = DecChain{nullptr, nullptr, 0, 1, 1})
{
    if (T.max_len <= DEC_MAX_D) {
        // f[] lives right behind the staged tables in dynamic shared memory
        uint32_t *s_words = const_cast<uint32_t *>(s_dec) + T.dec_stage_words;
        uint8_t *s_len = reinterpret_cast<uint8_t *>(s_words + cand_words(ch));
        uint16_t *s_fn = reinterpret_cast<uint16_t *>(s_len + cand_len_bytes(ch));
        if constexpr (MULTI) {
            __shared__ ChainShared<true> sh;
            if (T.dec_stage_words > T.lut_pad)
                return decode_stream_cta_chain<true, true, Out>(in, nbytes, T, s_dec, lut2, s_words, s_len, s_fn, ch, out, cap, chain, ts, sh);
            return decode_stream_cta_chain<false, true, Out>(in, nbytes, T, s_dec, lut2, s_words, s_len, s_fn, ch, out, cap, chain, ts, sh);
        }
        if (T.dec_stage_words > T.lut_pad)
            return decode_stream_cta_cand<true, Out>(in, nbytes, T, s_dec, lut2, s_words, s_len, s_fn, ch, out, cap, ts);
        return decode_stream_cta_cand<false, Out>(in, nbytes, T, s_dec, lut2, s_words, s_len, s_fn, ch, out, cap, ts);
    }
    if constexpr (MULTI) {
        return chain.slot == 0 ? -2 : DEC_NOT_MINE;  // unreachable: the chained kernel is launched for codes of <= DEC_CHAIN_D bits only
    } else {
        if (chain.slot != 0) return DEC_NOT_MINE;  // codes longer than a subsequence: one CTA per stream
        return decode_stream_cta<Out>(in, nbytes, T, s_dec, lut2, out, cap, ts);
    }
}

// Stages the decode tables with one TMA bulk copy; every thread of the CTA must call.  The copy is only STARTED here: the
// decoders wait for it (TableStage::wait) right before their first table look-up, after their own global loads (stream
// sizes, header byte, the chunk's words) have been issued, so that the two latencies overlap.
// Returns the second-level table pointer (shared when it was staged, else global).
__device__ __forceinline__ const uint32_t *stage_decode_tables(const DevTable &T, uint32_t *s_dec, unsigned long long *mbar)
{
    if (threadIdx.x == 0) mbar_init(mbar);
    __syncthreads();
    if (threadIdx.x == 0) tma_load_1d(s_dec, T.lut, T.dec_stage_words * 4u, mbar);
    return T.dec_stage_words > T.lut_pad ? s_dec + T.lut_pad : T.lut2;
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

// bitmap + exclusive popcount prefix of one level, by the whole CTA; wordfn(wi) = the 32 cells
// wi*32 .. wi*32+31 (bit i = cell wi*32+i; bits past n are masked here)
template <int NT = UP_THREADS, typename F>
__device__ void build_level(int64_t n, int nw, uint32_t *bits, uint32_t *prefix, int32_t *pop_out, F wordfn)
{
    __shared__ int s_warp[NT / 32];
    int run = 0;
    for (int base = 0; base < nw; base += NT) {
        const int wi = base + threadIdx.x;
        uint32_t word = 0;
        if (wi < nw) {
            word = wordfn(wi);
            const int64_t left = n - (int64_t)wi * 32;
            if (left < 32) word &= (1u << left) - 1u;
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
        int woff = 0, tot = 0;
#pragma unroll
        for (int i = 0; i < NT / 32; ++i) {
            const int t = s_warp[i];
            if (i < wid) woff += t;
            tot += t;
        }
        if (wi < nw) prefix[wi] = (uint32_t)(run + woff + inc - v);
        run += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) *pop_out = run;
}

// 32 cells of a level from its per-cell function (any geometry)
template <typename F>
__device__ __forceinline__ uint32_t word_from_cells(int wi, int gw, int64_t n, F cell)
{
    const int64_t p0 = (int64_t)wi * 32;
    int y = (int)(p0 / gw), x = (int)(p0 - (int64_t)y * gw);
    const int lim = (int)min((int64_t)32, n - p0);
    uint32_t word = 0;
    for (int i = 0; i < lim; ++i) {
        word |= (uint32_t)cell(y, x) << i;
        if (++x == gw) {
            x = 0;
            ++y;
        }
    }
    return word;
}

// 32 consecutive cells straight from a mask stream (cell j = payload bit j, MSB first): bit i = cell wi*32+i
__device__ __forceinline__ uint32_t word_from_stream(const uint8_t *in, int cap, int wi)
{
    uint32_t be = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int o = 1 + 4 * wi + k;
        if (o < cap) be |= (uint32_t)in[o] << (24 - 8 * k);
    }
    return __brev(be);
}

// every bit of the low 16 bits doubled into 2 adjacent bits
__device__ __forceinline__ uint32_t spread2(uint32_t x)
{
    x &= 0xFFFFu;
    x = (x | (x << 8)) & 0x00FF00FFu;
    x = (x | (x << 4)) & 0x0F0F0F0Fu;
    x = (x | (x << 2)) & 0x33333333u;
    x = (x | (x << 1)) & 0x55555555u;
    return x | (x << 1);
}

// copies a 16-byte aligned slot prefix into shared memory (whole CTA); returns the shared pointer
__device__ __forceinline__ const uint8_t *stage_bytes(const uint8_t *src, int nbytes, unsigned char *dst)
{
    const int n16 = (nbytes + 15) >> 4;
    for (int i = threadIdx.x; i < n16; i += blockDim.x) reinterpret_cast<uint4 *>(dst)[i] = __ldg(reinterpret_cast<const uint4 *>(src) + i);
    return dst;
}

// bitmaps + exclusive popcount prefixes + populations of the coarse, medium and derived fine level of one image, by the
// whole CTA (NT threads); mc / mm = the mask streams (header byte first), bits / prefix [nw16 + nw8 + nw4], pop [3]
template <int NT>
__device__ __forceinline__ void build_mask_levels(const UnpackArgs &a, const uint8_t *mc, const uint8_t *mm, int cap_c, int cap_m, bool need_c,
                                                  bool need_m, uint32_t *bits, uint32_t *prefix, int32_t *pop)
{
    const Geo &g = a.g;
    {
        if (need_c)
            build_level<NT>(g.n16, g.nw16, bits, prefix, pop + 0, [&](int wi) { return word_from_stream(mc, cap_c, wi); });
        else
            build_level<NT>(g.n16, g.nw16, bits, prefix, pop + 0, [&](int) { return a.mode == 4 ? 0xFFFFFFFFu : 0u; });
    }
    {
        if (need_m)
            build_level<NT>(g.n8, g.nw8, bits + g.nw16, prefix + g.nw16, pop + 1, [&](int wi) { return word_from_stream(mm, cap_m, wi); });
        else if (a.mode == 3)
            build_level<NT>(g.n8, g.nw8, bits + g.nw16, prefix + g.nw16, pop + 1, [&](int wi) {
                return word_from_cells(wi, g.w8, g.n8, [&](int y, int x) { return medium_cell(a, mc, mm, y, x); });
            });
        else
            build_level<NT>(g.n8, g.nw8, bits + g.nw16, prefix + g.nw16, pop + 1, [&](int) { return a.mode == 5 ? 0xFFFFFFFFu : 0u; });
        uint32_t *fb = bits + g.nw16 + g.nw8, *fp = prefix + g.nw16 + g.nw8;
        if (a.mode > 2) {
            build_level<NT>(g.n4, g.nw4, fb, fp, pop + 2, [&](int) { return a.mode == 6 ? 0xFFFFFFFFu : 0u; });
        } else if (g.w % 32 == 0) {
            // a bitmap word = 32 cells of one row = 16 medium cells (2 stream bytes) and 8 coarse cells (1 byte)
            build_level<NT>(g.n4, g.nw4, fb, fp, pop + 2, [&](int wi) {
                const int64_t p0 = (int64_t)wi * 32;
                const int y = (int)(p0 / g.w), x0 = (int)(p0 - (int64_t)y * g.w);
                uint32_t m16 = 0, c8 = 0;
                if (need_m) {
                    const int64_t p8 = (int64_t)(y >> 1) * g.w8 + (x0 >> 1);
                    const uint8_t *q = mm + 1 + (p8 >> 3);
                    m16 = __brev(((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16));
                }
                if (need_c) {
                    const int64_t p16 = (int64_t)(y >> 2) * g.w16 + (x0 >> 2);
                    c8 = __brev((uint32_t)mc[1 + (p16 >> 3)] << 24);
                }
                return ~(spread2(m16) | spread2(spread2(c8)));
            });
        } else {
            build_level<NT>(g.n4, g.nw4, fb, fp, pop + 2, [&](int wi) {
                return word_from_cells(wi, g.w, g.n4, [&](int y, int x) { return fine_cell(a, mc, mm, y, x); });
            });
        }
    }
}

// grid (4, B): CTAs 0..2 decode the index streams, CTA 3 builds the coarse, medium and fine mask levels
// (four CTAs per image keep a 64-image batch within one wave at two CTAs per SM).  Dynamic shared memory: decode tables (CTAs 0..2) / mask stream bytes (3, 4).
template <bool MULTI>
__device__ __forceinline__ void unpack_decode_cta(const UnpackArgs &a, unsigned char *dyn, unsigned long long &mbar_ref)
{
    unsigned long long *mbar_p = &mbar_ref;
    const int nslots = MULTI ? a.nslots : 1;
    int s, slot, b;
    if (MULTI) {
        // linear block order = hand-out order: slot 0 of every stream of every image first, then slot 1, ...
        // A short stream leaves its high slots without a chunk: those CTAs come late and leave at once, instead of
        // holding SM slots that the busy CTAs of later images need.  (A chunk only ever waits for chunks of lower slots or
        // earlier rounds, all handed out before it.)
        const int lin = (int)(blockIdx.y * gridDim.x + blockIdx.x), nb = (int)gridDim.y;
        if (lin < nb) {
            s = 3, slot = 0, b = lin;  // the mask CTAs (always busy) lead
        } else {
            slot = (lin - nb) / (3 * nb);
            const int r = lin - nb - slot * 3 * nb;
            b = r / 3;
            s = r - b * 3;
        }
    } else {
        // CTAs are handed to the SMs in linear block order, breadth first: the first 148 get an SM of their own.  Long streams
        // first (all medium, then all fine), the short coarse streams and the mask CTAs last, so that at two CTAs per SM a long
        // stream shares its SM with a short one instead of with another long one (block b*4+s pairs equal s: 148 % 4 == 0).
        const int lin = (int)(blockIdx.y * gridDim.x + blockIdx.x), nb = (int)gridDim.y;
        const int k = lin / nb;
        s = k == 0 ? 1 : (k == 1 ? 2 : (k == 2 ? 0 : 3)), slot = 0, b = lin - k * nb;
    }
    const Geo &g = a.g;
    CGIC_STAMP(unpack, 0);
    pdl_trigger_step<4>();
    const uint8_t *img = a.bytes + (int64_t)b * a.image_stride;
    const int32_t *sz = a.sizes + b * 5;
    const int nwt = g.nw16 + g.nw8 + g.nw4;
    if (s < 3) {
        const int64_t soff = s == 0 ? 0 : (s == 1 ? g.n16 : g.n16 + g.n8);
        const int64_t cap = s == 0 ? g.n16 : (s == 1 ? g.n8 : g.n4);
        int cnt = -1;
        // the (immutable) decode tables are staged while the predecessor kernel may still be running
        uint32_t *s_dec = reinterpret_cast<uint32_t *>(dyn);
        const uint32_t *lut2 = nullptr;
        TableStage ts{mbar_p, false};
        if (stream_present(a.mode, s)) {
            lut2 = stage_decode_tables(a.T, s_dec, mbar_p);
            ts.pending = true;
        }
        pdl_wait();
        int nbytes = stream_present(a.mode, s) ? sz[s] : 0;
        CGIC_STAMP(unpack, 1);
        if (nbytes < 0 || nbytes > a.slot_cap[s]) {  // a size no packer can have produced: never read past the slot
            nbytes = 0;
            cnt = slot == 0 ? -2 : DEC_NOT_MINE;
        } else if (nbytes > 0) {
            // a stream whose payload may exceed the chain's capacity (only a corrupt size can) is decoded by slot 0 alone
            const bool chained = MULTI && nslots > 1 && ((int64_t)nbytes * 8 + DEC_SUB_BITS - 1) / DEC_SUB_BITS <= (int64_t)a.ws.max_chunks * a.ch;
            const DecChain chain{chained ? a.ws.chain + ((int64_t)b * 3 + s) * a.ws.max_chunks * a.ws.chain_row : nullptr,
                                 a.ws.ticket + b * 3 + s, chained ? slot : 0, chained ? nslots : 1, a.ws.chain_row};
            bool idle = false;  // a chained stream with fewer chunks than slots: the high slots have nothing to do
            if (chained) {
                int ch_plan = a.ch;
                const int64_t nch = chain_plan(stream_subsequences(img + a.slot_off[s], nbytes), nslots, ch_plan);
                idle = slot >= (nch > 0 ? nch : 1);
            }
            if (idle)
                cnt = DEC_NOT_MINE;
            else if (chained || slot == 0)
                cnt = decode_stream_any<uint16_t, MULTI>(img + a.slot_off[s], nbytes, a.T, s_dec, lut2,
                                                         a.ws.sym + (int64_t)b * (g.n16 + g.n8 + g.n4) + soff, cap, a.ch, ts, chain);
            else
                cnt = DEC_NOT_MINE;
        } else if (slot != 0) {
            cnt = DEC_NOT_MINE;
        }
        ts.wait();  // never leave with the bulk copy in flight (empty / absent streams decode nothing)
        CGIC_STAMP(unpack, 6);
        if (threadIdx.x == 0) {
            if (cnt == -2) a.ws.flag[b * 5 + s] = CGIC_EFORMAT;  // symbol capacity exceeded (any of the stream's CTAs may see it)
            else if (cnt != DEC_NOT_MINE) a.ws.flag[b * 5 + s] = 0;
            if (cnt != DEC_NOT_MINE) a.ws.count[b * 3 + s] = cnt;
        }
        return;
    }
    pdl_wait();
    const bool need_c = a.mode == 0 || a.mode == 2 || a.mode == 3;
    const bool need_m = a.mode == 0 || a.mode == 1;
    const int cap_c = (int)(g.n16 / 8 + 2), cap_m = (int)(g.n8 / 8 + 2);
    // framing check of the mask streams this mode reads
    bool ok_c = true, ok_m = true;
    if (need_c) ok_c = sz[3] == cap_c && img[a.slot_off[3]] == 8 - (int)(g.n16 & 7);
    if (need_m) ok_m = sz[4] == cap_m && img[a.slot_off[4]] == 8 - (int)(g.n8 & 7);
    if (threadIdx.x == 0) {
        a.ws.flag[b * 5 + 3] = !ok_c ? CGIC_EFORMAT : 0;
        a.ws.flag[b * 5 + 4] = !ok_m ? CGIC_EFORMAT : 0;
        a.status[b] = 0;  // the assemble kernel (which runs after this grid) raises it atomically
    }
    // mask bytes: shared copies when they fit (a.mask_stage), else straight from global
    const uint8_t *mc = img + a.slot_off[3];
    const uint8_t *mm = img + a.slot_off[4];
    if (a.mask_stage) {
        const int off_m = (cap_c + 15) & ~15;
        if (need_c && ok_c) mc = stage_bytes(mc, cap_c, dyn);
        if (need_m && ok_m) mm = stage_bytes(mm, cap_m, dyn + off_m);
        __syncthreads();
    }
    build_mask_levels<UP_THREADS>(a, mc, mm, cap_c, cap_m, need_c, need_m, a.ws.bits + (int64_t)b * nwt, a.ws.prefix + (int64_t)b * nwt,
                                  a.ws.pop + b * 3);
}

// Re-assembly of 4 consecutive fine tokens of a row (w is a multiple of 4): they share one coarse
// cell, two medium cells and one bitmap word per level.  16-byte stores throughout.  bits / prefix / sym / cnt may live
// in global memory (two-kernel path) or in shared memory (fused small-grid kernel).  Returns true when an index beyond
// the table came up (sums of overlapping levels cannot occur with valid masks).
__device__ __forceinline__ bool assemble_quad_core(const UnpackArgs &a, int b, int64_t quad, const uint32_t *bits, const uint32_t *prefix,
                                                   const uint16_t *sym, int cnt0, int cnt1, int cnt2)
{
    const Geo &g = a.g;
    const int64_t p = quad * 4;
    if (p >= g.n4) return false;
    const int y = (int)(p / g.w), x = (int)(p - (int64_t)y * g.w);
    const int64_t p16 = (int64_t)(y >> 2) * g.w16 + (x >> 2);
    const int64_t p8 = (int64_t)(y >> 1) * g.w8 + (x >> 1);  // even: p8 and p8 + 1 share a word
    // bitmap words, popcount prefixes and counts are all independent loads: issued together, one round trip
    const uint32_t wc = bits[p16 >> 5], wm = bits[g.nw16 + (p8 >> 5)], wf = bits[g.nw16 + g.nw8 + (p >> 5)];
    const uint32_t pre_c = prefix[p16 >> 5], pre_m = prefix[g.nw16 + (p8 >> 5)], pre_f = prefix[g.nw16 + g.nw8 + (p >> 5)];
    const int sc = (int)(p16 & 31), sm = (int)(p8 & 31), sf = (int)(p & 31);
    const int cbit = (wc >> sc) & 1;
    int64_t base = 0;
    if (cbit && cnt0 > 0) {
        const int r = pre_c + __popc(wc & ((1u << sc) - 1u));
        if (r < cnt0) base = sym[r];
    }
    int mbit[2];
    int64_t mval[2] = {0, 0};
    {
        const uint32_t pre = pre_m;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            mbit[j] = (wm >> (sm + j)) & 1;
            if (mbit[j] && cnt1 > 0) {
                const int r = pre + __popc(wm & ((1u << (sm + j)) - 1u));
                if (r < cnt1) mval[j] = sym[g.n16 + r];
            }
        }
    }
    int fbit[4];
    int64_t ind[4];
    {
        const uint32_t pre = pre_f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            fbit[i] = (wf >> (sf + i)) & 1;
            int64_t v = base + mval[i >> 1];
            if (fbit[i] && cnt2 > 0) {
                const int r = pre + __popc(wf & ((1u << (sf + i)) - 1u));
                if (r < cnt2) v += sym[g.n16 + g.n8 + r];
            }
            ind[i] = v;
        }
    }
    longlong2 *ind_o = reinterpret_cast<longlong2 *>(a.ind_out + (int64_t)b * g.n4 + p);
    ind_o[0] = make_longlong2(ind[0], ind[1]);
    ind_o[1] = make_longlong2(ind[2], ind[3]);
    longlong2 *mf_o = reinterpret_cast<longlong2 *>(a.mf_out + (int64_t)b * g.n4 + p);
    mf_o[0] = make_longlong2(fbit[0], fbit[1]);
    mf_o[1] = make_longlong2(fbit[2], fbit[3]);
    if ((y & 1) == 0) *reinterpret_cast<longlong2 *>(a.mm_out + (int64_t)b * g.n8 + p8) = make_longlong2(mbit[0], mbit[1]);
    if ((y & 3) == 0) a.mc_out[(int64_t)b * g.n16 + p16] = cbit;
    bool bad = false;
    if (a.quant_out) {
        float4 e[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            bad |= ind[i] >= a.T.K;
            e[i] = __ldg(reinterpret_cast<const float4 *>(a.codebook) + (ind[i] < a.T.K ? ind[i] : 0));
        }
        float4 *q = reinterpret_cast<float4 *>(a.quant_out + (int64_t)b * 4 * g.n4 + p);
        const int64_t plane4 = g.n4 / 4;
        q[0] = make_float4(e[0].x, e[1].x, e[2].x, e[3].x);
        q[plane4] = make_float4(e[0].y, e[1].y, e[2].y, e[3].y);
        q[2 * plane4] = make_float4(e[0].z, e[1].z, e[2].z, e[3].z);
        q[3 * plane4] = make_float4(e[0].w, e[1].w, e[2].w, e[3].w);
    }
    return bad;
}

// Re-assembly of 4 consecutive fine tokens of a row for the fused kernel: 32-bit arithmetic (the grid has at most 4096
// cells), tables in shared memory.  Same outputs as assemble_quad_core.
__device__ __forceinline__ bool assemble_quad_small(const UnpackArgs &a, int b, int quad, const uint32_t *bits, const uint32_t *prefix,
                                                    const uint16_t *sym, int cnt0, int cnt1, int cnt2)
{
    const Geo &g = a.g;
    const int n4 = (int)g.n4, n8 = (int)g.n8, n16 = (int)g.n16;
    const int p = quad * 4;
    const int wq = g.w >> 2;
    const int y = quad / wq, x = (quad - y * wq) * 4;
    const int p16 = (y >> 2) * g.w16 + (x >> 2);
    const int p8 = (y >> 1) * g.w8 + (x >> 1);  // even: p8 and p8 + 1 share a word
    const uint32_t wc = bits[p16 >> 5], wm = bits[g.nw16 + (p8 >> 5)], wf = bits[g.nw16 + g.nw8 + (p >> 5)];
    const uint32_t pre_c = prefix[p16 >> 5], pre_m = prefix[g.nw16 + (p8 >> 5)], pre_f = prefix[g.nw16 + g.nw8 + (p >> 5)];
    const int sc = p16 & 31, sm = p8 & 31, sf = p & 31;
    const int cbit = (wc >> sc) & 1;
    int base = 0;
    if (cbit && cnt0 > 0) {
        const int r = pre_c + __popc(wc & ((1u << sc) - 1u));
        if (r < cnt0) base = sym[r];
    }
    int mbit[2], mval[2] = {0, 0};
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        mbit[j] = (wm >> (sm + j)) & 1;
        if (mbit[j] && cnt1 > 0) {
            const int r = pre_m + __popc(wm & ((1u << (sm + j)) - 1u));
            if (r < cnt1) mval[j] = sym[n16 + r];
        }
    }
    int fbit[4], ind[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        fbit[i] = (wf >> (sf + i)) & 1;
        int v = base + mval[i >> 1];
        if (fbit[i] && cnt2 > 0) {
            const int r = pre_f + __popc(wf & ((1u << (sf + i)) - 1u));
            if (r < cnt2) v += sym[n16 + n8 + r];
        }
        ind[i] = v;
    }
    const size_t o4 = (size_t)b * n4 + p;
    longlong2 *ind_o = reinterpret_cast<longlong2 *>(a.ind_out + o4);
    ind_o[0] = make_longlong2(ind[0], ind[1]);
    ind_o[1] = make_longlong2(ind[2], ind[3]);
    longlong2 *mf_o = reinterpret_cast<longlong2 *>(a.mf_out + o4);
    mf_o[0] = make_longlong2(fbit[0], fbit[1]);
    mf_o[1] = make_longlong2(fbit[2], fbit[3]);
    if ((y & 1) == 0) *reinterpret_cast<longlong2 *>(a.mm_out + (size_t)b * n8 + p8) = make_longlong2(mbit[0], mbit[1]);
    if ((y & 3) == 0) a.mc_out[(size_t)b * n16 + p16] = cbit;
    bool bad = false;
    if (a.quant_out) {
        float4 e[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            bad |= ind[i] >= a.T.K;  // sums of overlapping levels cannot occur with valid masks
            e[i] = __ldg(reinterpret_cast<const float4 *>(a.codebook) + (ind[i] < a.T.K ? ind[i] : 0));
        }
        float4 *q = reinterpret_cast<float4 *>(a.quant_out + (size_t)b * 4 * n4 + p);
        const int plane4 = n4 >> 2;
        q[0] = make_float4(e[0].x, e[1].x, e[2].x, e[3].x);
        q[plane4] = make_float4(e[0].y, e[1].y, e[2].y, e[3].y);
        q[2 * plane4] = make_float4(e[0].z, e[1].z, e[2].z, e[3].z);
        q[3 * plane4] = make_float4(e[0].w, e[1].w, e[2].w, e[3].w);
    }
    return bad;
}

// the reference's masked assignment raises unless #symbols == #set cells (an empty coarse / medium stream stands for
// zeros, model.py:284-290)
__device__ __forceinline__ bool counts_mismatch(int mode, const int32_t *cnt, const int32_t *pop)
{
    bool bad = false;
    for (int s = 0; s < 3; ++s) {
        if (!stream_present(mode, s)) continue;
        const bool empty_ok = cnt[s] == -1 && (s < 2 || pop[s] == 0);
        if (!empty_ok && cnt[s] != pop[s]) bad = true;
    }
    return bad;
}

__device__ __forceinline__ void assemble_quad(const UnpackArgs &a, int b, int64_t quad)
{
    const Geo &g = a.g;
    const int nwt = g.nw16 + g.nw8 + g.nw4;
    const int32_t *cntp = a.ws.count + b * 3;
    if (quad == 0) {
        bool bad = counts_mismatch(a.mode, cntp, a.ws.pop + b * 3);
        for (int s = 0; s < 5; ++s) bad |= a.ws.flag[b * 5 + s] != 0;
        if (bad) atomicExch(&a.status[b], CGIC_EFORMAT);
    }
    bool bad;
    if (g.n4 < ((int64_t)1 << 28)) {  // (always, in practice) 32-bit arithmetic
        bad = quad * 4 < g.n4 && assemble_quad_small(a, b, (int)quad, a.ws.bits + (int64_t)b * nwt, a.ws.prefix + (int64_t)b * nwt,
                                                     a.ws.sym + (int64_t)b * (g.n16 + g.n8 + g.n4), cntp[0], cntp[1], cntp[2]);
    } else {
        bad = assemble_quad_core(a, b, quad, a.ws.bits + (int64_t)b * nwt, a.ws.prefix + (int64_t)b * nwt,
                                 a.ws.sym + (int64_t)b * (g.n16 + g.n8 + g.n4), cntp[0], cntp[1], cntp[2]);
    }
    if (bad) atomicExch(&a.status[b], CGIC_EFORMAT);
}

// grid (4, B): one CTA per index stream, then the mask CTA
__global__ void __launch_bounds__(UP_THREADS) unpack_decode_kernel(const UnpackArgs a)
{
    extern __shared__ __align__(128) unsigned char dyn[];
    __shared__ __align__(8) unsigned long long mbar;
    unpack_decode_cta<false>(a, dyn, mbar);
}

// grid (3 * nslots + 1, B): nslots CTAs per index stream (consecutive block ids), then the mask CTA.  Large token grids
// only -- a separate kernel so that the hand-over code does not weigh on the small-grid kernel's instruction footprint
// (measured: +2 us on the 256 x 256 batch when both lived in one kernel).
__global__ void __launch_bounds__(UP_THREADS, 3) unpack_decode_chained_kernel(const UnpackArgs a)  // <= 42 registers: three CTAs per SM
{
    extern __shared__ __align__(128) unsigned char dyn[];
    __shared__ __align__(8) unsigned long long mbar;
    unpack_decode_cta<true>(a, dyn, mbar);
}

// re-assembly: one thread per quad, its own PDL-chained launch (see the note in cgic_unpack)
__global__ void __launch_bounds__(256) unpack_assemble_kernel(const UnpackArgs a)
{
    CGIC_STAMP(assemble, 0);
    pdl_launch_dependents();
    pdl_wait();
    CGIC_STAMP(assemble, 1);
    assemble_quad(a, blockIdx.y, (int64_t)blockIdx.x * blockDim.x + threadIdx.x);
    CGIC_STAMP(assemble, 2);
}

// ======================================================================================================================
// Small token grids (<= DS_MAX_N4 fine tokens) with codes of at most 32 bits: ONE CTA per image decodes the three index
// streams, builds the mask levels and re-assembles the image -- symbols, bitmaps and prefixes never leave shared memory
// and the step is one launch instead of two (decode -> hand-over -> re-assembly).
//
// Decoder: the same idea as the candidate decoder above (the first codeword of a subsequence can only start at one of
// D = max_len offsets), without its D-fold redundant chain walks.  Subsequences are single 32-bit stream words; after the
// pass that writes the code length at every bit position (len8[]),
//   DP    one thread per word sweeps its 32 positions from the last to the first:
//             fn[q] = (exit offset into the next word, codewords started) = q + len >= 32 ? (q + len - 32, 1) : fn[q + len] + (0, 1)
//         -- one table look-up per bit position, chains that merge share their tail.  Positions are handled in groups of
//         G = 4: a group's look-ups only reach positions >= q + min_len of EARLIER groups, so with min_len >= 4 they are
//         independent of the group's own stores and their latencies overlap (G = 1 for tables with shorter codes);
//   B     the true entry offset of every word by composing fn along the stream: blocks of 16 words for every
//         candidate entry (one thread each), a serial walk over the <= 32 blocks, blocks replayed;
//   C     every word decodes the codewords that start in it from its true entry offset and writes the symbols at its
//         prefix count into the image's symbol list in shared memory.
// The three streams of an image share every pass: a batch is up to 512 words cut into at most three segments (one per
// stream, block aligned, one look-ahead word each); streams longer than a batch continue in the next one, carrying
// (entry offset, symbols so far).  Greedy semantics of the reference (indices_coding.py:140-151) as everywhere: a position
// whose codeword would run past the payload ends the decoding of that stream.
#ifndef CGIC_DS_THREADS
#define CGIC_DS_THREADS 512
#endif
constexpr int DS_THREADS = CGIC_DS_THREADS;
// 448 words per batch and the code lengths overlaid on the fn[] rows keep the kernel's shared memory under 75 KB = three
// CTAs per SM (40 registers at 512 threads): 167 -> 160 us at 2048 images of 256 x 256; 512 words at two CTAs per SM are as
// fast at 512 images, smaller batches lose (an image of ~375 words must stay one batch).
#ifndef CGIC_DS_WORDS
#define CGIC_DS_WORDS 448
#endif
#ifndef CGIC_DS_CTAS
#define CGIC_DS_CTAS 3
#endif
constexpr int DS_WORDS = CGIC_DS_WORDS;    // stream words (= subsequences) per batch
constexpr int DS_BS = 16;                  // words per composition block
constexpr int DS_NBLK = DS_WORDS / DS_BS;  // 32
constexpr int DS_WCAP = DS_NBLK * 17 + 1;   // a CTA's blocks as 16 words + the word after them
constexpr int DS_MAX_D = 32;
constexpr int DS_MAX_N4 = 4096;
constexpr uint32_t DS_STOP = 0xFFu;

// len8[]: 36 bytes per word, fn[]: 34 u16 (68 bytes) per word -- word strides of 9 / 17 banks, so that the threads of a
// warp (one word each) never meet in a bank

struct DsSeg {
    int s, sub0, n, nb, w0, blk0;  // stream, first stream word, words, blocks, first batch word, first batch block
};

struct DsLayout {
    size_t words, len, fn, sym, bits, prefix, masks, total;
};
__host__ __device__ inline DsLayout ds_layout(uint32_t dec_stage_words, const Geo &g)
{
    auto up = [](size_t v) { return (v + 15) / 16 * 16; };
    DsLayout L;
    size_t o = up((size_t)dec_stage_words * 4);
    L.words = o;
    o += up((size_t)(DS_WCAP + 1) * 4);
    L.fn = o;  // (the code lengths of a word occupy the first 32 bytes of its fn[] row until the DP has read them)
    o += up((size_t)DS_WCAP * 68);
    L.len = L.fn;
    L.sym = o;
    o += up((size_t)(g.n16 + g.n8 + g.n4) * 2);
    L.bits = o;
    o += up((size_t)(g.nw16 + g.nw8 + g.nw4) * 4);
    L.prefix = o;
    o += up((size_t)(g.nw16 + g.nw8 + g.nw4) * 4);
    L.masks = o;
    o += up((size_t)(g.n16 / 8 + 2)) + up((size_t)(g.n8 / 8 + 2));
    L.total = o;
    return L;
}

// (sym << 8) | len of the codeword at the head of `win` (32 stream bits, MSB first); codes of at most 32 bits
__device__ __forceinline__ uint32_t ds_decode_win(uint32_t win, const uint32_t *s_lut, const uint32_t *lut2, const DevTable &T, int L)
{
    const uint32_t e = s_lut[win >> (32 - L)];
    const uint32_t f = e & 0xFFu;
    if (!(f & 0x80u)) return e;
    if (f != 0xFFu) {
        const uint32_t hgt = f & 0x7Fu;
        return lut2[(e >> 8) + ((win << L) >> (32 - hgt))] + (uint32_t)L;  // second-level entries hold the bits used beyond L
    }
    int node = (int)(e >> 8);
    uint32_t used = (uint32_t)L;
    while (node >= T.K && used < 32u) {
        node = __ldg(&T.child[2 * node + (int)((win >> (31u - used)) & 1u)]);
        ++used;
    }
    return node >= T.K ? 0u : (((uint32_t)node << 8) | used);  // (max_len <= 32: the walk always ends on a leaf)
}

// bitmap + exclusive popcount prefix + population of one level by ONE warp (the levels of a small grid have at most
// 128 words); wordfn(wi) = the 32 cells wi*32 .. wi*32+31 (bits past n are masked here)
template <typename F>
__device__ __forceinline__ void build_level_warp(int n, int nw, uint32_t *bits, uint32_t *prefix, int32_t *pop_out, int lane, F wordfn)
{
    int run = 0;
    for (int base = 0; base < nw; base += 32) {
        const int wi = base + lane;
        uint32_t word = 0;
        if (wi < nw) {
            word = wordfn(wi);
            const int left = n - wi * 32;
            if (left < 32) word &= (1u << left) - 1u;
            bits[wi] = word;
        }
        const int v = __popc(word);
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (wi < nw) prefix[wi] = (uint32_t)(run + inc - v);
        run += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) *pop_out = run;
}

// CL = CTAs per image (a thread-block cluster when > 1).  All CTAs of an image plan the same batches; the 16-word blocks of
// a batch are dealt round-robin over them (block k -> CTA k % CL): staging, len8[], DP, the block functions (B1), the replay
// (B3) and the symbol write-out (C) touch a CTA's own blocks only.  What every CTA needs from the others travels through
// distributed shared memory: the block functions (written into every CTA's table before the serial walk B2, which each CTA
// then does for itself) and the decoded symbols (written into every CTA's symbol list); the re-assembly is split by quads.
// Small batches thus spread one image over 2 or 4 SMs instead of leaving most of the machine idle.
template <int G, int CL>
__global__ void __launch_bounds__(DS_THREADS, DS_THREADS <= 512 ? CGIC_DS_CTAS : 1) unpack_small_kernel(const UnpackArgs a)
{
    namespace cgs = cooperative_groups;
    extern __shared__ __align__(128) unsigned char dyn[];
    __shared__ __align__(8) unsigned long long mbar;
    __shared__ uint32_t s_blockfn[DS_NBLK * 32];   // (symbols << 8) | exit offset of a block, per candidate entry offset
    __shared__ uint8_t s_blkstart[DS_NBLK];
    __shared__ uint32_t s_blkbase[DS_NBLK];
    __shared__ uint8_t s_substart[DS_WCAP];
    __shared__ uint32_t s_subbase[DS_WCAP];
    __shared__ int s_nbits[3], s_nwords[3], s_done[3], s_nsym[3], s_nbytes[3];
    __shared__ uint32_t s_entry[3];
    __shared__ int32_t s_pop[3], s_cnt[3];
    __shared__ int s_bad;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = CL == 1 ? (int)blockIdx.x : (int)blockIdx.x / CL, rank = CL == 1 ? 0 : (int)blockIdx.x % CL;
    const Geo &g = a.g;
    const DsLayout SL = ds_layout(a.T.dec_stage_words, g);
    uint32_t *s_dec = reinterpret_cast<uint32_t *>(dyn);
    uint32_t *s_w = reinterpret_cast<uint32_t *>(dyn + SL.words);
    uint8_t *s_len = dyn + SL.len;
    uint16_t *s_fn = reinterpret_cast<uint16_t *>(dyn + SL.fn);
    uint16_t *s_sym = reinterpret_cast<uint16_t *>(dyn + SL.sym);
    uint32_t *s_bits = reinterpret_cast<uint32_t *>(dyn + SL.bits);
    uint32_t *s_prefix = reinterpret_cast<uint32_t *>(dyn + SL.prefix);
    unsigned char *s_masks = dyn + SL.masks;
    const int D = a.T.max_len, L = a.T.lut_bits;
    auto cluster_sync = [&]() {
        if (CL == 1) __syncthreads();
        else cgs::this_cluster().sync();
    };
    // the same object in the shared memory of cluster rank rr
    auto remote = [&](auto *ptr, int rr) { return CL == 1 ? ptr : cgs::this_cluster().map_shared_rank(ptr, rr); };
    CGIC_STAMP(unpack, 0);
    pdl_trigger_step<4>();
    // the (immutable) decode tables are staged while the predecessor kernel may still be running
    if (tid == 0) {
        mbar_init(&mbar);
        s_bad = 0;
    }
    __syncthreads();
    if (tid == 0) tma_load_1d(s_dec, a.T.lut, a.T.dec_stage_words * 4u, &mbar);
    const bool lut2_shared = a.T.dec_stage_words > a.T.lut_pad;
    const uint32_t *lut2 = lut2_shared ? s_dec + a.T.lut_pad : a.T.lut2;
    pdl_wait();
    CGIC_STAMP(unpack, 1);
    const uint8_t *img = a.bytes + (int64_t)b * a.image_stride;
    const int32_t *sz = a.sizes + b * 5;
    // ---- ONE round trip to global memory for everything that does not depend on the stream sizes: the sizes themselves, the
    //      header bytes of the index streams, the mask streams (their length follows from the grid)
    const bool need_c = a.mode == 0 || a.mode == 2 || a.mode == 3;
    const bool need_m = a.mode == 0 || a.mode == 1;
    const int cap_c = (int)(g.n16 / 8 + 2), cap_m = (int)(g.n8 / 8 + 2);
    const int off_m = (cap_c + 15) & ~15;
    {
        const int nc16 = need_c ? (cap_c + 15) >> 4 : 0, nm16 = need_m ? (cap_m + 15) >> 4 : 0;
        for (int i = tid; i < nc16 + nm16; i += DS_THREADS) {
            const bool isc = i < nc16;
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(img + (isc ? a.slot_off[3] : a.slot_off[4])) + (isc ? i : i - nc16));
            reinterpret_cast<uint4 *>(s_masks + (isc ? 0 : off_m))[isc ? i : i - nc16] = v;
        }
    }
    if (tid < 3) {
        const int s = tid;
        int nbytes = stream_present(a.mode, s) ? __ldg(sz + s) : 0;
        const int pad = __ldg(img + a.slot_off[s]);  // (read whether or not the stream exists: the slot does)
        if (nbytes < 0 || nbytes > a.slot_cap[s]) {  // a size no packer can have produced: never read past the slot
            nbytes = 0;
            atomicOr(&s_bad, 1);
        }
        int nbits = 0;
        if (nbytes > 0) {
            nbits = (nbytes - 1) * 8 - pad;
            if (pad == 0 || nbits < 0) nbits = 0;  // text[:-0] is empty in the reference
        }
        s_nbytes[s] = nbytes;
        s_nbits[s] = nbits;
        s_nwords[s] = (nbits + 31) >> 5;
        s_done[s] = 0;
        s_nsym[s] = 0;
        s_entry[s] = 0;
    }
    if (tid == 32 && need_c && __ldg(sz + 3) != cap_c) atomicOr(&s_bad, 1);
    if (tid == 33 && need_m && __ldg(sz + 4) != cap_m) atomicOr(&s_bad, 1);
    cluster_sync();  // (also: every CTA of the cluster has started -- a precondition for writing into its shared memory)
    // ---- mask levels: one warp per level (framing of the mask streams checked on the staged bytes); meanwhile the other warps
    //      go on to the first batch, whose global loads thus overlap this.  Every CTA of a cluster builds its own copy.
    if (warp < 3) {
        const uint8_t *mc = s_masks, *mm = s_masks + off_m;
        if (warp == 0) {
            if (need_c && lane == 0 && mc[0] != 8 - (int)(g.n16 & 7)) atomicOr(&s_bad, 1);
            if (need_c) build_level_warp((int)g.n16, g.nw16, s_bits, s_prefix, s_pop + 0, lane, [&](int wi) { return word_from_stream(mc, cap_c, wi); });
            else build_level_warp((int)g.n16, g.nw16, s_bits, s_prefix, s_pop + 0, lane, [&](int) { return a.mode == 4 ? 0xFFFFFFFFu : 0u; });
        } else if (warp == 1) {
            uint32_t *bm = s_bits + g.nw16, *pm = s_prefix + g.nw16;
            if (need_m && lane == 0 && mm[0] != 8 - (int)(g.n8 & 7)) atomicOr(&s_bad, 1);
            if (need_m) build_level_warp((int)g.n8, g.nw8, bm, pm, s_pop + 1, lane, [&](int wi) { return word_from_stream(mm, cap_m, wi); });
            else if (a.mode == 3)
                build_level_warp((int)g.n8, g.nw8, bm, pm, s_pop + 1, lane, [&](int wi) {
                    return word_from_cells(wi, g.w8, g.n8, [&](int y, int x) { return medium_cell(a, mc, mm, y, x); });
                });
            else build_level_warp((int)g.n8, g.nw8, bm, pm, s_pop + 1, lane, [&](int) { return a.mode == 5 ? 0xFFFFFFFFu : 0u; });
        } else {
            uint32_t *fb = s_bits + g.nw16 + g.nw8, *fp = s_prefix + g.nw16 + g.nw8;
            if (a.mode > 2) {
                build_level_warp((int)g.n4, g.nw4, fb, fp, s_pop + 2, lane, [&](int) { return a.mode == 6 ? 0xFFFFFFFFu : 0u; });
            } else if (g.w % 32 == 0) {
                // a bitmap word = 32 cells of one row = 16 medium cells (2 stream bytes) and 8 coarse cells (1 byte)
                build_level_warp((int)g.n4, g.nw4, fb, fp, s_pop + 2, lane, [&](int wi) {
                    const int p0 = wi * 32;
                    const int y = p0 / g.w, x0 = p0 - y * g.w;
                    uint32_t m16 = 0, c8 = 0;
                    if (need_m) {
                        const int p8 = (y >> 1) * g.w8 + (x0 >> 1);
                        const uint8_t *q = mm + 1 + (p8 >> 3);
                        m16 = __brev(((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16));
                    }
                    if (need_c) {
                        const int p16 = (y >> 2) * g.w16 + (x0 >> 2);
                        c8 = __brev((uint32_t)mc[1 + (p16 >> 3)] << 24);
                    }
                    return ~(spread2(m16) | spread2(spread2(c8)));
                });
            } else {
                build_level_warp((int)g.n4, g.nw4, fb, fp, s_pop + 2, lane, [&](int wi) {
                    return word_from_cells(wi, g.w, g.n4, [&](int y, int x) { return fine_cell(a, mc, mm, y, x); });
                });
            }
        }
    }
    bool tables_ready = false;
    CGIC_STAMP(unpack, 2);

    auto symoff = [&](int s) { return s == 0 ? 0 : (s == 1 ? (int)g.n16 : (int)(g.n16 + g.n8)); };
    auto symcap = [&](int s) { return s == 0 ? (int)g.n16 : (s == 1 ? (int)g.n8 : (int)g.n4); };
    for (;;) {
        // ---- plan the batch: up to three segments, block aligned (every thread of every CTA computes the same plan from its
        //      copy of the stream state; scalars, not an array: a dynamically indexed array would live in local memory)
        int nseg = 0, nblk = 0;
        int g_s[3] = {0, 0, 0}, g_sub0[3] = {0, 0, 0}, g_n[3] = {0, 0, 0}, g_nb[3] = {0, 0, 0}, g_blk0[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff};
        {
            int free_blk = DS_NBLK;
#pragma unroll
            for (int s = 0; s < 3; ++s) {
                const int rem = s_nwords[s] - s_done[s];
                if (rem <= 0 || s_entry[s] == DS_STOP || free_blk == 0) continue;
                const int n = min(rem, free_blk * DS_BS), nb = (n + DS_BS - 1) / DS_BS;
#pragma unroll
                for (int k = 0; k < 3; ++k)
                    if (k == nseg) {
                        g_s[k] = s;
                        g_sub0[k] = s_done[s];
                        g_n[k] = n;
                        g_nb[k] = nb;
                        g_blk0[k] = DS_NBLK - free_blk;
                    }
                free_blk -= nb;
                ++nseg;
            }
            nblk = DS_NBLK - free_blk;
        }
        if (nseg == 0) break;
        // this CTA's blocks: k = rank, rank + CL, ...; local block j lives in words [17 j, 17 j + 17) of the CTA's buffers
        // (16 words + the word after them, which the last word's windows reach into)
        const int nown = nblk > rank ? (nblk - rank + CL - 1) / CL : 0;
        auto blk_stream = [&](int k) { return k >= g_blk0[2] ? g_s[2] : (k >= g_blk0[1] ? g_s[1] : g_s[0]); };
        auto blk_word0 = [&](int k) {  // first STREAM word of block k
            return k >= g_blk0[2] ? g_sub0[2] + (k - g_blk0[2]) * DS_BS : (k >= g_blk0[1] ? g_sub0[1] + (k - g_blk0[1]) * DS_BS : g_sub0[0] + (k - g_blk0[0]) * DS_BS);
        };
        // ---- stage the stream words of the own blocks: word i of a stream = bytes 1 + 4 i .. 4 + 4 i (header byte skipped), MSB first
        for (int it = tid; it < nown * 17; it += DS_THREADS) {
            const int j = it / 17, i = it - j * 17, k = rank + j * CL;
            const int ss = blk_stream(k), sw = blk_word0(k) + i;
            const uint32_t *W = reinterpret_cast<const uint32_t *>(img + a.slot_off[ss]);
            const int cap = (int)a.slot_cap[ss];
            const uint32_t lo = 4 * sw + 4 <= cap ? __ldg(W + sw) : 0u, hi = 4 * sw + 8 <= cap ? __ldg(W + sw + 1) : 0u;
            s_w[it] = __byte_perm(lo, hi, 0x1234);
        }
        if (!tables_ready) {
            tables_ready = true;
            mbar_wait(&mbar, 0);
        }
        __syncthreads();
        CGIC_STAMP(unpack, 3);
        // ---- A0: code length at every bit position (0 where the codeword would run past the payload).  The eight look-ups of an
        //      item are issued together; entries that need the second level (or the tree) are resolved afterwards.
        for (int it = tid; it < nown * 64; it += DS_THREADS) {
            const int j = it >> 6, i = (it >> 2) & 15, qtr = it & 3, k = rank + j * CL;
            const int wl = 17 * j + i;                                                  // local word
            const int lim = s_nbits[blk_stream(k)] - 32 * (blk_word0(k) + i) - qtr * 8;  // position p of this quarter is valid iff p + len <= lim
            const uint32_t w0 = s_w[wl], w1 = s_w[wl + 1];
            uint32_t e[8];
#pragma unroll
            for (int p8 = 0; p8 < 8; ++p8) e[p8] = s_dec[__funnelshift_l(w1, w0, qtr * 8 + p8) >> (32 - L)];
            if (lut2_shared) {
                // second-level codes WITHOUT a branch: a warp covers 256 positions per pass, so even a 1 % share of two-level
                // codes makes "some lane needs the second table" the normal case -- a divergent slow path would run every time.
                // The second look-up is always issued (entry 0 when unused); only codes beyond it (flag 0xFF) walk the tree.
                uint32_t e2[8], deep = 0;
#pragma unroll
                for (int p8 = 0; p8 < 8; ++p8) {
                    const uint32_t f = e[p8] & 0xFFu;
                    const bool two = f >= 0x80u;
                    const uint32_t hgt = two ? (f & 0x7Fu) : 1u;
                    const uint32_t win = __funnelshift_l(w1, w0, qtr * 8 + p8);
                    e2[p8] = lut2[two && f != 0xFFu ? (e[p8] >> 8) + ((win << L) >> (32 - hgt)) : 0u];
                    deep |= f == 0xFFu;
                }
#pragma unroll
                for (int p8 = 0; p8 < 8; ++p8) {
                    const uint32_t f = e[p8] & 0xFFu;
                    e[p8] = f >= 0x80u ? (e2[p8] & 0xFFu) + (uint32_t)L : f;
                }
                if (deep) {  // rare: a code deeper than the second level
#pragma unroll
                    for (int p8 = 0; p8 < 8; ++p8) {
                        const uint32_t win = __funnelshift_l(w1, w0, qtr * 8 + p8);
                        if ((s_dec[win >> (32 - L)] & 0xFFu) == 0xFFu) e[p8] = ds_decode_win(win, s_dec, lut2, a.T, L) & 0xFFu;
                    }
                }
            } else {
                uint32_t any = 0;
#pragma unroll
                for (int p8 = 0; p8 < 8; ++p8) {
                    any |= e[p8];
                    e[p8] &= 0xFFu;
                }
                if (any & 0x80u) {  // second-level table in global memory / tree walk
#pragma unroll
                    for (int p8 = 0; p8 < 8; ++p8)
                        if (e[p8] & 0x80u) e[p8] = ds_decode_win(__funnelshift_l(w1, w0, qtr * 8 + p8), s_dec, lut2, a.T, L) & 0xFFu;
                }
            }
#pragma unroll
            for (int p8 = 0; p8 < 8; ++p8)
                if (p8 + (int)e[p8] > lim) e[p8] = 0;
            uint32_t *dst = reinterpret_cast<uint32_t *>(s_len + 68 * wl + qtr * 8);
            dst[0] = e[0] | (e[1] << 8) | (e[2] << 16) | (e[3] << 24);
            dst[1] = e[4] | (e[5] << 8) | (e[6] << 16) | (e[7] << 24);
        }
        __syncthreads();
        CGIC_STAMP(unpack, 7);
        // ---- DP: fn[q] = (exit offset << 8) | codewords, from the last position of a word to the first
        for (int it = tid; it < nown * DS_BS; it += DS_THREADS) {
            const int wl = 17 * (it >> 4) + (it & 15);
            const uint32_t *lw = reinterpret_cast<const uint32_t *>(s_len + 68 * wl);  // all 32 lengths are read ...
            uint16_t *fw = s_fn + 34 * wl;                                            // ... before fn[] overwrites them (same thread)
            uint32_t l8[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) l8[i] = lw[i];
            if (G == 4) {
#pragma unroll
                for (int grp = 7; grp >= 0; --grp) {
                    uint32_t r[4], t[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const uint32_t len = (l8[grp] >> (8 * u)) & 0xFFu;
                        t[u] = (uint32_t)(4 * grp + u) + len;
                        r[u] = len == 0 ? (DS_STOP << 8) : (t[u] >= 32u ? (((t[u] - 32u) << 8) | 1u) : 0xFFFFFFFFu);
                    }
                    uint32_t ld[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) ld[u] = fw[r[u] == 0xFFFFFFFFu ? t[u] : 31u];  // (every look-up reaches an earlier group)
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (r[u] == 0xFFFFFFFFu) r[u] = ld[u] + 1u;  // same exit, one more codeword (<= 32: no carry into the exit byte)
                    uint32_t *dst = reinterpret_cast<uint32_t *>(fw + 4 * grp);
                    dst[0] = r[0] | (r[1] << 16);
                    dst[1] = r[2] | (r[3] << 16);
                }
            } else {
                for (int k = 31; k >= 0; --k) {
                    const uint32_t len = (l8[k >> 2] >> (8 * (k & 3))) & 0xFFu;
                    const uint32_t t = (uint32_t)k + len;
                    uint32_t r;
                    if (len == 0) r = DS_STOP << 8;
                    else if (t >= 32u) r = ((t - 32u) << 8) | 1u;
                    else r = (uint32_t)fw[t] + 1u;
                    fw[k] = (uint16_t)r;
                }
            }
        }
        __syncthreads();
        CGIC_STAMP(unpack, 4);
        // ---- B1: what an own block does to every candidate entry offset (items = blocks x D, no idle lanes); the result goes
        //      into the table of EVERY CTA of the cluster
        for (int it = tid; it < nown * D; it += DS_THREADS) {
            const int j = it / D, c = it - j * D;
            const uint16_t *fb = s_fn + 34 * 17 * j;
            uint32_t x = (uint32_t)c, n = 0;
#pragma unroll 4
            for (int i = 0; i < DS_BS; ++i) {
                if (x != DS_STOP) {
                    const uint32_t f = fb[34 * i + x];
                    n += f & 0xFFu;
                    x = f >> 8;
                }
            }
            const uint32_t v = (n << 8) | x;
            uint32_t *dst = &s_blockfn[(rank + j * CL) * 32 + c];
#pragma unroll
            for (int rr = 0; rr < CL; ++rr) *remote(dst, rr) = v;
        }
        cluster_sync();
        CGIC_STAMP(unpack, 10);
        // ---- B2: one thread per segment chains the blocks (every CTA for itself) and carries the stream's state into the next batch
        if (tid < nseg) {
            const int st = tid == 0 ? g_s[0] : (tid == 1 ? g_s[1] : g_s[2]), nb = tid == 0 ? g_nb[0] : (tid == 1 ? g_nb[1] : g_nb[2]);
            const int blk0 = tid == 0 ? g_blk0[0] : (tid == 1 ? g_blk0[1] : g_blk0[2]);
            const int done = tid == 0 ? g_sub0[0] + g_n[0] : (tid == 1 ? g_sub0[1] + g_n[1] : g_sub0[2] + g_n[2]);
            uint32_t x = s_entry[st], base = (uint32_t)s_nsym[st];
            for (int k = 0; k < nb; ++k) {
                const int blk = blk0 + k;
                s_blkstart[blk] = (uint8_t)x;
                s_blkbase[blk] = base;
                if (x != DS_STOP) {
                    const uint32_t f = s_blockfn[blk * 32 + x];
                    base += f >> 8;
                    x = f & 0xFFu;
                }
            }
            s_entry[st] = x;
            s_nsym[st] = (int)min(base, 0x7FFFFFFFu);
            s_done[st] = done;
        }
        __syncthreads();
        CGIC_STAMP(unpack, 11);
        // ---- B3: own blocks replayed from their true entry offset
        if (tid < nown) {
            const int k = rank + tid * CL;
            uint32_t x = s_blkstart[k], base = s_blkbase[k];
            const uint16_t *fb = s_fn + 34 * 17 * tid;
#pragma unroll 4
            for (int i = 0; i < DS_BS; ++i) {
                s_substart[17 * tid + i] = (uint8_t)x;
                s_subbase[17 * tid + i] = base;
                if (x != DS_STOP) {
                    const uint32_t f = fb[34 * i + x];
                    base += f & 0xFFu;
                    x = f >> 8;
                }
            }
        }
        __syncthreads();
        CGIC_STAMP(unpack, 5);
        // ---- C: every own word writes the symbols of the codewords that start in it -- into the symbol list of every CTA
        for (int it = tid; it < nown * DS_BS; it += DS_THREADS) {
            const int j = it >> 4, wl = 17 * j + (it & 15);
            uint32_t q = s_substart[wl];
            if (q >= 32u) continue;  // DS_STOP, or the codeword that straddles into this word ends beyond it
            const int s = blk_stream(rank + j * CL);
            uint16_t *out = s_sym + symoff(s);
            const int ocap = symcap(s);
            int o = (int)min(s_subbase[wl], 0x7FFFFFFFu);
            const uint32_t w0 = s_w[wl], w1 = s_w[wl + 1];
            const int k = rank + j * CL;
            const int lim = s_nbits[s] - 32 * (blk_word0(k) + (it & 15));  // a codeword at q is inside the payload iff q + len <= lim
            while (q < 32u) {
                const uint32_t e = ds_decode_win(__funnelshift_l(w1, w0, q), s_dec, lut2, a.T, L);
                const uint32_t len = e & 0xFFu;
                if (!len || (int)(q + len) > lim) break;
                if (o < ocap) {
#pragma unroll
                    for (int rr = 0; rr < CL; ++rr) remote(out, rr)[o] = (uint16_t)(e >> 8);
                }
                ++o;
                q += len;
            }
        }
        cluster_sync();  // symbols of the batch are in place everywhere; nobody reads the block functions any more
        CGIC_STAMP(unpack, 8);
    }
    if (!tables_ready) mbar_wait(&mbar, 0);  // never leave with the bulk copy in flight
    CGIC_STAMP(unpack, 6);
    // ---- symbol counts, status
    if (tid < 3) {
        int c = s_nbytes[tid] > 0 ? s_nsym[tid] : -1;
        if (c > symcap(tid)) {  // more symbols than the level has cells: only a corrupt stream can
            c = -2;
            atomicOr(&s_bad, 1);
        }
        s_cnt[tid] = c;
    }
    __syncthreads();
    const int cnt0 = s_cnt[0], cnt1 = s_cnt[1], cnt2 = s_cnt[2];
    bool bad = false;
    const int nquads = (int)(g.n4 >> 2);
    for (int quad = rank * DS_THREADS + tid; quad < nquads; quad += CL * DS_THREADS)
        bad |= assemble_quad_small(a, b, quad, s_bits, s_prefix, s_sym, cnt0, cnt1, cnt2);
    const int any_bad = __syncthreads_or(bad);
    CGIC_STAMP(unpack, 9);
    if (CL > 1) {
        // rank 0 reports for the image; nobody may leave while its shared memory can still be written from outside
        if (tid == 0 && rank != 0 && (any_bad || s_bad)) atomicOr(remote(&s_bad, 0), 1);
        cluster_sync();
    }
    if (tid == 0 && rank == 0) a.status[b] = (any_bad || s_bad || counts_mismatch(a.mode, s_cnt, s_pop)) ? CGIC_EFORMAT : 0;
}

size_t unpack_small_smem(const DevTable &T, const Geo &g) { return ds_layout(T.dec_stage_words, g).total; }

__global__ void __launch_bounds__(DEC_THREADS)
huff_decode_single_kernel(const uint8_t *bytes, int64_t nbytes, DevTable T, int32_t *out, int64_t cap, int32_t *count_out, int ch)
{
    extern __shared__ __align__(128) unsigned char dyn[];
    __shared__ __align__(8) unsigned long long mbar;
    uint32_t *s_dec = reinterpret_cast<uint32_t *>(dyn);
    const uint32_t *lut2 = stage_decode_tables(T, s_dec, &mbar);
    TableStage ts{&mbar, true};
    const int cnt = decode_stream_any<int32_t>(bytes, nbytes, T, s_dec, lut2, out, cap, ch, ts);
    ts.wait();
    if (threadIdx.x == 0) *count_out = cnt;
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

size_t decode_smem_bytes(const DevTable &T, int ch)
{
    size_t b = (size_t)T.dec_stage_words * 4;
    if (T.max_len <= DEC_MAX_D) {
        b += (size_t)ch * T.max_len * 2 + (size_t)cand_words(ch) * 4 + (size_t)cand_len_bytes(ch);
    }
    return b;
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
    for (int s = 0; s < 5; ++s) {
        a.slot_off[s] = L.off[s];
        a.slot_cap[s] = L.cap[s];
    }
    a.mode = mode;
    a.codebook = codebook;
    a.mc_out = mc_out;
    a.mm_out = mm_out;
    a.mf_out = mf_out;
    a.ind_out = ind_out;
    a.quant_out = quant_out;
    a.status = status_out;
    cudaStream_t stream = as_stream(stream_);
    for (const void *ptr : {(const void *)mm_out, (const void *)mf_out, (const void *)ind_out, (const void *)quant_out, (const void *)bytes})
        CGIC_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, CGIC_EINVAL, "cgic_unpack: buffers must be 16-byte aligned");
    // dynamic shared memory: decode tables for the stream CTAs, mask stream bytes for the mask CTAs
    a.ch = cand_chunk_subs(a.T.max_len, a.g.n4);
    // CTAs per index stream (large token grids, codes of at most DEC_CHAIN_D bits): a fine stream carries up to ~12 bits per
    // token; chunks of at most 128 subsequences keep five such CTAs on an SM.  The chunk size actually used is chosen per
    // stream inside the kernel (one chunk per CTA whenever the stream allows it).
    a.nslots = 1;
    if (a.g.n4 > DEC_SINGLE_N4 && a.T.max_len <= DEC_CHAIN_D) {
        static const int ch_env = getenv("CGIC_DEC_CH") ? atoi(getenv("CGIC_DEC_CH")) : 0;        // A-B runs only
        static const int slots_env = getenv("CGIC_DEC_SLOTS") ? atoi(getenv("CGIC_DEC_SLOTS")) : 0;
        const int ch_max = ch_env >= 32 && ch_env <= DEC_CHAIN_CH ? (ch_env & ~31) : DEC_CHAIN_CH;
        a.ch = std::min(a.ch, ch_max);
        const int slots_max = slots_env >= 1 && slots_env <= DEC_CHAIN_SLOTS ? slots_env : DEC_CHAIN_SLOTS;
        const int64_t chunks = (a.g.n4 * 12 / DEC_SUB_BITS + 31) / 32;   // 32 subsequences = the smallest chunk
        a.nslots = (int)std::min<int64_t>(slots_max, std::max<int64_t>(1, chunks));
    }
    // Small token grids, codes of at most 32 bits: decode + re-assembly fused, one CTA per image (unpack_small_kernel) -- once
    // the batch fills the machine.  Measured on B200 (256x256 images, decode + re-assembly per step): 2048 images 175 us
    // against 245 us for the two launches below, 512 images 58 us against 77 us; at 64 images the two launches, which put
    // four CTAs on every image, are the faster ones (15.8 us in the graph against 18.2 us for the fused kernel as clusters
    // of two CTAs per image, 20.2 us as one CTA per image), so batches that leave SMs idle keep them.
    // cgic_tune("fused_decode_ctas", 1 | 2 | 4) (or CGIC_DS_CLUSTER in the environment) forces the fused kernel with that many
    // CTAs per image -- a thread-block cluster when > 1 -- for A-B runs and the tests; -1 switches it off.
    const int force_cl = tune_fused_decode_ctas();
    const bool no_small = force_cl < 0;
    int n_sm = 0;
    rc = device_sm_count(&n_sm);
    if (rc) return rc;
    const bool forced = force_cl == 1 || force_cl == 2 || force_cl == 4;
    if (!no_small && a.g.n4 <= DS_MAX_N4 && a.T.max_len <= DS_MAX_D && (forced || B > n_sm)) {
        const size_t smem_small = unpack_small_smem(a.T, a.g);
        const bool g4 = a.T.min_len >= 4;
        const int cl = forced && g4 ? force_cl : 1;
        void (*kern)(const UnpackArgs) = !g4 ? unpack_small_kernel<1, 1> : (cl == 4 ? unpack_small_kernel<4, 4> : (cl == 2 ? unpack_small_kernel<4, 2> : unpack_small_kernel<4, 1>));
        rc = ensure_smem((const void *)kern, smem_small);
        if (rc) return rc;
        {
            CGIC_PROF("unpack_small_kernel", stream);
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3((unsigned)(B * cl));
            cfg.blockDim = dim3(DS_THREADS);
            cfg.dynamicSmemBytes = smem_small;
            cfg.stream = stream;
            static const bool no_pdl = getenv("CGIC_NO_PDL") != nullptr;
            cudaLaunchAttribute attr[2];
            int na = 0;
            if (!no_pdl) {
                attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                attr[na].val.programmaticStreamSerializationAllowed = 1;
                ++na;
            }
            if (cl > 1) {
                attr[na].id = cudaLaunchAttributeClusterDimension;
                attr[na].val.clusterDim.x = (unsigned)cl;
                attr[na].val.clusterDim.y = 1;
                attr[na].val.clusterDim.z = 1;
                ++na;
            }
            cfg.attrs = attr;
            cfg.numAttrs = na;
            CGIC_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, a));
        }
        CGIC_LAUNCH_CHECK();
        return CGIC_OK;
    }
    const size_t dec_bytes = decode_smem_bytes(a.T, a.ch);
    const size_t mask_bytes = (size_t)(((a.g.n16 / 8 + 2 + 15) & ~15) + ((a.g.n8 / 8 + 2 + 15) & ~15));
    a.mask_stage = mask_bytes <= 96 * 1024;
    const size_t smem = std::max(dec_bytes, a.mask_stage ? mask_bytes : (size_t)0);
    // static + dynamic shared memory may exceed the 48 KB default
    rc = ensure_smem((const void *)unpack_decode_kernel, 180 * 1024);
    if (!rc) rc = ensure_smem((const void *)unpack_decode_chained_kernel, 180 * 1024);
    if (rc) return rc;
    // Two PDL-chained launches.  Fusing the re-assembly into the decode kernel was measured on B200 (B = 64, 256^2,
    // decode + re-assembly per step) and lost both ways: as thread-block clusters of one image's CTAs with a cluster
    // barrier 36.5 us, as "the last CTA of an image to finish re-assembles it" (ticket) 30.2 us, against 27.9 us here --
    // the re-assembly wants many SMs per image, not the one that happens to finish last.
    {
        CGIC_PROF("unpack_decode_kernel", stream);
        if (a.nslots > 1) CGIC_CUDA_CHECK(launch_pdl(unpack_decode_chained_kernel, dim3(3 * a.nslots + 1, B), dim3(UP_THREADS), smem, stream, a));
        else CGIC_CUDA_CHECK(launch_pdl(unpack_decode_kernel, dim3(4, B), dim3(UP_THREADS), smem, stream, a));
    }
    CGIC_LAUNCH_CHECK();
    {
        CGIC_PROF("unpack_assemble_kernel", stream);
        CGIC_CUDA_CHECK(launch_pdl(unpack_assemble_kernel, dim3((unsigned)((a.g.n4 / 4 + 255) / 256), B), dim3(256), 0, stream, a));
    }
    CGIC_LAUNCH_CHECK();
    return CGIC_OK;
}

extern "C" int cgic_huff_decode(const uint8_t *bytes, int64_t nbytes, const cgic_table *t, int32_t *symbols_out, int64_t cap,
                                int32_t *count_out, cgic_stream_t stream)
{
    CGIC_REQUIRE(count_out && nbytes >= 0 && (bytes || nbytes == 0) && (symbols_out || cap == 0), CGIC_EINVAL,
                 "cgic_huff_decode: bad argument");
    CGIC_REQUIRE((reinterpret_cast<uintptr_t>(bytes) & 3) == 0, CGIC_EINVAL, "cgic_huff_decode: bytes must be 4-byte aligned");
    DevTable T;
    int rc = table_device_view(t, &T);
    if (rc) return rc;
    const int ch = cand_chunk_subs(T.max_len);
    const size_t smem = decode_smem_bytes(T, ch);
    rc = ensure_smem((const void *)huff_decode_single_kernel, 180 * 1024);
    if (rc) return rc;
    huff_decode_single_kernel<<<1, DEC_THREADS, smem, as_stream(stream)>>>(bytes, nbytes, T, symbols_out, cap, count_out, ch);
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
