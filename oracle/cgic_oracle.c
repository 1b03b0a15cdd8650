/*
 * cgic_oracle.c -- CPU restatement (plain C) of Control-GIC's VQ + entropy-coding hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under control-gic_b200/ may include, link or call this
 * file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs use it, and only as the checker / reported baseline.
 *
 * Parity pinning: the reference has no tests or golden vectors of its own, so
 * this oracle is pinned against outputs of the reference itself, generated in the build
 * container by tests/golden/make_golden.py (which imports /root/reference) and committed as
 * tests/golden/ (npz and json fixtures), plus the KAT1..KAT8 vectors of SURVEY.md 8(c).
 *
 * Every function cites the reference lines (relative to the reference repo root) it follows.
 * Build: oracle/Makefile (gcc -O2 -ffp-contract=off; contraction must stay off because the
 * rounding sequence below IS the specification).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_OK 0
#define ORC_EINVAL (-1)
#define ORC_ENOMEM (-2)
#define ORC_ESPACE (-3)

/* ------------------------------------------------------------------------------------------
 * a1  VectorQuantize2.forward         CGIC/modules/vqvae/quantize.py:69-98
 *
 * d = sum(z^2) + sum(e^2) - 2 * (z . e)      (quantize.py:73-75), argmin (quantize.py:78).
 * The reference evaluates this with torch CPU kernels; the rounding sequence that reproduces
 * torch's fp32 result bit for bit (verified against the imported reference, see
 * tests/golden/make_golden.py) is:
 *     z2  = ((z0^2 + z1^2) + z2^2) + z3^2           each square and each add rounded
 *     e2  likewise
 *     dot = fma(z3,e3, fma(z2,e2, fma(z1,e1, fl(z0*e0))))
 *     d   = fl( fl(z2 + e2) - 2*dot )
 * torch.argmin returns the lowest index among equal minima.
 * z_q (value) = fl(z + fl(e - z))  (straight-through, quantize.py:93), NCHW (quantize.py:96).
 * loss = mean((e-z)^2) + beta*mean((e-z)^2)  (legacy branch, quantize.py:89-90); we return the
 * sum of squared errors in double and let the caller form the mean (float tolerance only).
 * ------------------------------------------------------------------------------------------ */
static inline float sq4(const float *v)
{
    float a = v[0] * v[0];
    float b = v[1] * v[1];
    float s = a + b;
    b = v[2] * v[2];
    s = s + b;
    b = v[3] * v[3];
    s = s + b;
    return s;
}

int orc_vq_assign(const float *z_nchw, int B, int h, int w, const float *codebook, int K,
                  int64_t *idx_out, float *zq_nchw, double *sqerr_out)
{
    if (!z_nchw || !codebook || !idx_out || K < 1) return ORC_EINVAL;
    const int64_t plane = (int64_t)h * w;
    float *e2 = (float *)malloc(sizeof(float) * (size_t)K);
    if (!e2) return ORC_ENOMEM;
    for (int k = 0; k < K; ++k) e2[k] = sq4(codebook + 4 * (int64_t)k);
    double sq = 0.0;
    for (int b = 0; b < B; ++b) {
        const float *zb = z_nchw + (int64_t)b * 4 * plane;
        for (int64_t p = 0; p < plane; ++p) {
            float z[4] = {zb[p], zb[plane + p], zb[2 * plane + p], zb[3 * plane + p]};
            const float z2 = sq4(z);
            float best = 0.f;
            int bi = 0;
            for (int k = 0; k < K; ++k) {
                const float *e = codebook + 4 * (int64_t)k;
                float dot = z[0] * e[0];
                dot = fmaf(z[1], e[1], dot);
                dot = fmaf(z[2], e[2], dot);
                dot = fmaf(z[3], e[3], dot);
                float s = z2 + e2[k];
                float d = s - 2.0f * dot;
                /* torch.argmin: NaN counts as minimal, first NaN wins; otherwise strict <. */
                if (k == 0) { best = d; bi = 0; }
                else if (!(best != best) && (d < best || d != d)) { best = d; bi = k; }
            }
            idx_out[(int64_t)b * plane + p] = bi;
            const float *e = codebook + 4 * (int64_t)bi;
            for (int c = 0; c < 4; ++c) {
                float diff = e[c] - z[c];
                if (zq_nchw) zq_nchw[((int64_t)b * 4 + c) * plane + p] = z[c] + diff;
                sq += (double)diff * (double)diff;
            }
        }
    }
    if (sqerr_out) *sqerr_out = sq;
    free(e2);
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------
 * a8  HuffmanCoding.__init__ / make_heap / merge_nodes / make_codes
 *                                              CGIC/tools/indices_coding.py:10-17, 46-75
 *
 * The code table is defined by CPython's heapq on nodes ordered by freq ONLY
 * (indices_coding.py:26-27).  heapq (Lib/heapq.py, unchanged 3.10..3.12) restated on an array
 * of node ids: push = append + sift towards the root while strictly smaller than the parent;
 * pop = take the last element, put it at the root, walk down always to the smaller child
 * (the RIGHT child when the two are equal: "not left < right"), then sift back up.
 * Push order = iteration order of the `frequency` mapping (indices_coding.py:46-49), passed in
 * as `order` (order[i] = symbol pushed i-th; NULL = 0..K-1).  NOTE: inference.py:137-139 passes
 * the model's nn.ParameterDict, which torch builds from a plain dict via sorted(items()), so the
 * real order is the LEXICOGRAPHIC order of the decimal key strings ("0","1","10","100",...).
 * Merge: first pop is the left ('0') child, second pop the right ('1') child
 * (indices_coding.py:51-60, 62-71).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    int64_t *freq;  /* per node */
    int *heap;
    int n;
} orc_heap;

static void heap_sift_to_root(orc_heap *H, int startpos, int pos)
{
    int item = H->heap[pos];
    while (pos > startpos) {
        int parentpos = (pos - 1) >> 1;
        int parent = H->heap[parentpos];
        if (H->freq[item] < H->freq[parent]) {
            H->heap[pos] = parent;
            pos = parentpos;
            continue;
        }
        break;
    }
    H->heap[pos] = item;
}

static void heap_push(orc_heap *H, int node)
{
    H->heap[H->n++] = node;
    heap_sift_to_root(H, 0, H->n - 1);
}

static int heap_pop(orc_heap *H)
{
    int last = H->heap[--H->n];
    if (H->n == 0) return last;
    int ret = H->heap[0];
    H->heap[0] = last;
    int endpos = H->n, pos = 0, item = last;
    int child = 1;
    while (child < endpos) {
        int right = child + 1;
        if (right < endpos && !(H->freq[H->heap[child]] < H->freq[H->heap[right]])) child = right;
        H->heap[pos] = H->heap[child];
        pos = child;
        child = 2 * pos + 1;
    }
    H->heap[pos] = item;
    heap_sift_to_root(H, 0, pos);
    return ret;
}

/* Outputs: left/right child arrays for nodes K..2K-2 (leaf ids 0..K-1), root id, and per
 * symbol the code as '0'/'1' characters at codes + sym*K (NUL terminated, length < K). */
int orc_huff_build(const int64_t *freq, const int32_t *order, int K, int32_t *len_out, char *codes_out /* K*K */,
                   int32_t *left_out /* 2K */, int32_t *right_out /* 2K */, int32_t *root_out)
{
    if (!freq || K < 2 || !len_out) return ORC_EINVAL;
    const int nn = 2 * K - 1;
    orc_heap H;
    H.freq = (int64_t *)malloc(sizeof(int64_t) * (size_t)nn);
    H.heap = (int *)malloc(sizeof(int) * (size_t)nn);
    int *left = (int *)malloc(sizeof(int) * (size_t)nn);
    int *right = (int *)malloc(sizeof(int) * (size_t)nn);
    int *stack = (int *)malloc(sizeof(int) * (size_t)(2 * nn + 2));
    int *depth = (int *)malloc(sizeof(int) * (size_t)nn);
    int *parent = (int *)malloc(sizeof(int) * (size_t)nn);
    char *bit = (char *)malloc((size_t)nn);
    if (!H.freq || !H.heap || !left || !right || !stack || !depth || !parent || !bit) return ORC_ENOMEM;
    H.n = 0;
    for (int i = 0; i < K; ++i) { H.freq[i] = freq[i]; left[i] = right[i] = -1; }
    for (int i = 0; i < K; ++i) {
        int s = order ? order[i] : i;
        if (s < 0 || s >= K) return ORC_EINVAL;
        heap_push(&H, s);
    }
    int next = K;
    while (H.n > 1) {
        int a = heap_pop(&H);
        int b = heap_pop(&H);
        H.freq[next] = H.freq[a] + H.freq[b];
        left[next] = a;
        right[next] = b;
        heap_push(&H, next);
        ++next;
    }
    int root = heap_pop(&H);
    /* DFS from the root; code = path bits, left '0', right '1'. */
    int sp = 0;
    stack[sp++] = root;
    depth[root] = 0;
    parent[root] = -1;
    bit[root] = 0;
    while (sp) {
        int n = stack[--sp];
        if (n < K) {
            int L = depth[n];
            len_out[n] = L;
            if (codes_out) {
                char *dst = codes_out + (size_t)n * K;
                dst[L] = 0;
                int cur = n;
                for (int i = L - 1; i >= 0; --i) { dst[i] = bit[cur]; cur = parent[cur]; }
            }
        } else {
            int l = left[n], r = right[n];
            depth[l] = depth[r] = depth[n] + 1;
            parent[l] = parent[r] = n;
            bit[l] = '0';
            bit[r] = '1';
            stack[sp++] = r;
            stack[sp++] = l;
        }
    }
    if (left_out && right_out)
        for (int i = 0; i < nn; ++i) { left_out[i] = left[i]; right_out[i] = right[i]; }
    if (root_out) *root_out = root;
    free(H.freq); free(H.heap); free(left); free(right); free(stack); free(depth); free(parent); free(bit);
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------
 * Framing shared by HuffmanCoding and BinaryCoding
 *   pad_encoded_text   indices_coding.py:91-98   mask_coding.py:22-29
 *   get_byte_array     indices_coding.py:101-110 mask_coding.py:31-38
 *   remove_padding     indices_coding.py:131-138 mask_coding.py:61-68
 * stream = [8-bit pad count][payload bits MSB first][pad zero bits], pad = 8 - nbits%8 in 1..8.
 * Empty symbol list -> zero-byte file (indices_coding.py:116-118).
 * ------------------------------------------------------------------------------------------ */
typedef struct { uint8_t *p; int64_t cap; int64_t bitpos; int err; } bitw;

static inline void bw_put(bitw *w, int b)
{
    int64_t byte = w->bitpos >> 3;
    if (byte >= w->cap) { w->err = 1; return; }
    if (b) w->p[byte] |= (uint8_t)(0x80u >> (w->bitpos & 7));
    w->bitpos++;
}

/* a9  HuffmanCoding.compress   indices_coding.py:113-126 (78-82 concatenation of codes). */
int64_t orc_huff_encode(const int64_t *sym, int64_t n, int K, const int32_t *len, const char *codes,
                        uint8_t *out, int64_t cap)
{
    if (n == 0) return 0;
    int64_t nbits = 0;
    for (int64_t i = 0; i < n; ++i) {
        if (sym[i] < 0 || sym[i] >= K) return ORC_EINVAL;
        nbits += len[sym[i]];
    }
    int64_t total = nbits / 8 + 2;
    if (total > cap) return ORC_ESPACE;
    memset(out, 0, (size_t)total);
    int pad = 8 - (int)(nbits % 8);
    out[0] = (uint8_t)pad;
    bitw w = {out, cap, 8, 0};
    for (int64_t i = 0; i < n; ++i) {
        const char *c = codes + (size_t)sym[i] * K;
        for (int j = 0; j < len[sym[i]]; ++j) bw_put(&w, c[j] == '1');
    }
    return w.err ? ORC_ESPACE : total;
}

/* a10 HuffmanCoding.decompress_string   indices_coding.py:153-168, decode_text 140-151:
 * greedy prefix decode over the payload bits; a trailing incomplete code is dropped.
 * Returns the symbol count, or -100 for the empty file (reference returns None). */
int64_t orc_huff_decode(const uint8_t *in, int64_t nbytes, int K, const int32_t *left,
                        const int32_t *right, int root, int64_t *sym_out, int64_t cap)
{
    if (nbytes == 0) return -100;
    int pad = in[0];
    int64_t nbits = (nbytes - 1) * 8 - pad;
    /* remove_padding slices text[:-pad]; pad == 0 would yield an empty text (text[:-0]). */
    if (pad == 0 || nbits < 0) nbits = 0;
    int64_t cnt = 0;
    int node = root;
    for (int64_t i = 0; i < nbits; ++i) {
        int64_t bp = 8 + i;
        int b = (in[bp >> 3] >> (7 - (bp & 7))) & 1;
        node = b ? right[node] : left[node];
        if (node < K) {
            if (cnt >= cap) return ORC_ESPACE;
            sym_out[cnt++] = node;
            node = root;
        }
    }
    return cnt;
}

/* a11 BinaryCoding.compress / decompress_string   mask_coding.py:40-55, 81-96. */
int64_t orc_bits_encode(const int32_t *v, int64_t n, uint8_t *out, int64_t cap)
{
    if (n == 0) return 0;
    int64_t total = n / 8 + 2;
    if (total > cap) return ORC_ESPACE;
    memset(out, 0, (size_t)total);
    out[0] = (uint8_t)(8 - (int)(n % 8));
    bitw w = {out, cap, 8, 0};
    for (int64_t i = 0; i < n; ++i) {
        if (v[i] != 0 && v[i] != 1) return ORC_EINVAL; /* KeyError in the reference */
        bw_put(&w, v[i]);
    }
    return total;
}

int64_t orc_bits_decode(const uint8_t *in, int64_t nbytes, int64_t *out, int64_t cap)
{
    if (nbytes == 0) return -100;
    int pad = in[0];
    int64_t nbits = (nbytes - 1) * 8 - pad;
    if (pad == 0 || nbits < 0) nbits = 0;
    if (nbits > cap) return ORC_ESPACE;
    for (int64_t i = 0; i < nbits; ++i) {
        int64_t bp = 8 + i;
        out[i] = (in[bp >> 3] >> (7 - (bp & 7))) & 1;
    }
    return nbits;
}

/* ------------------------------------------------------------------------------------------
 * a7 + a12  index selection and 5-stream pack for ONE image   CGIC/models/model.py:217-260
 *   coarse symbols: ind[::4, ::4][m_c == 1], medium: ind[::2, ::2][m_m == 1], fine: ind[m_f == 1]
 *   (row-major compaction).  Streams by mode (model.py:225-260):
 *   m0 {ic,im,if,mc,mm}  m1 {im,if,mm}  m2 {ic,if,mc}  m3 {ic,im,mc}  m4 {ic}  m5 {im}  m6 {if}
 *   Stream order in out/sizes: 0 ic, 1 im, 2 if, 3 mc, 4 mm.  Absent stream -> size 0.
 * ------------------------------------------------------------------------------------------ */
static const int ORC_STREAMS[7][5] = {
    {1, 1, 1, 1, 1}, {0, 1, 1, 0, 1}, {1, 0, 1, 1, 0}, {1, 1, 0, 1, 0},
    {1, 0, 0, 0, 0}, {0, 1, 0, 0, 0}, {0, 0, 1, 0, 0}};

int orc_pack_image(const int64_t *ind /* h*w */, const int32_t *mc, const int32_t *mm, const int32_t *mf,
                   int h, int w, int mode, int K, const int32_t *len, const char *codes,
                   uint8_t *out, const int64_t *slot_off /* 5 */, const int64_t *slot_cap /* 5 */,
                   int32_t *sizes /* 5 */)
{
    if (mode < 0 || mode > 6) return ORC_EINVAL;
    const int h8 = h / 2, w8 = w / 2, h16 = h / 4, w16 = w / 4;
    int64_t *tmp = (int64_t *)malloc(sizeof(int64_t) * (size_t)h * w + 8);
    if (!tmp) return ORC_ENOMEM;
    for (int s = 0; s < 5; ++s) sizes[s] = 0;
    int rc = ORC_OK;
    for (int s = 0; s < 3 && rc == ORC_OK; ++s) {
        if (!ORC_STREAMS[mode][s]) continue;
        const int step = s == 0 ? 4 : (s == 1 ? 2 : 1);
        const int gh = s == 0 ? h16 : (s == 1 ? h8 : h), gw = s == 0 ? w16 : (s == 1 ? w8 : w);
        const int32_t *m = s == 0 ? mc : (s == 1 ? mm : mf);
        int64_t n = 0;
        for (int y = 0; y < gh; ++y)
            for (int x = 0; x < gw; ++x)
                if (m[(int64_t)y * gw + x] == 1) tmp[n++] = ind[(int64_t)(y * step) * w + x * step];
        int64_t r = orc_huff_encode(tmp, n, K, len, codes, out + slot_off[s], slot_cap[s]);
        if (r < 0) rc = (int)r; else sizes[s] = (int32_t)r;
    }
    if (rc == ORC_OK && ORC_STREAMS[mode][3]) {
        int64_t r = orc_bits_encode(mc, (int64_t)h16 * w16, out + slot_off[3], slot_cap[3]);
        if (r < 0) rc = (int)r; else sizes[3] = (int32_t)r;
    }
    if (rc == ORC_OK && ORC_STREAMS[mode][4]) {
        int64_t r = orc_bits_encode(mm, (int64_t)h8 * w8, out + slot_off[4], slot_cap[4]);
        if (r < 0) rc = (int)r; else sizes[4] = (int32_t)r;
    }
    free(tmp);
    return rc;
}

/* ------------------------------------------------------------------------------------------
 * a13 + a14  unpack, mask / index re-assembly and codebook gather for ONE image
 *                                                       CGIC/models/model.py:269-392
 *   mode 0 (model.py:269-293): masks from bits; fine = 1 - up2(m) - up4(c); decoded symbols
 *   land on the mask==1 positions in row-major order; ind = fine + up2(medium) + up4(coarse);
 *   an empty index stream contributes zeros (model.py:284-290).  Modes 1..6: model.py:296-389.
 *   quant = codebook[ind] laid out NCHW (model.py:391-392).
 *   Returns ORC_EINVAL if a stream's symbol count differs from its mask population (the
 *   reference raises a shape error from the masked assignment in that case).
 * ------------------------------------------------------------------------------------------ */
int orc_unpack_image(const uint8_t *in, const int64_t *slot_off, const int32_t *sizes, int h, int w, int mode,
                     int K, const int32_t *left, const int32_t *right, int root, const float *codebook,
                     int64_t *mc_out, int64_t *mm_out, int64_t *mf_out, int64_t *ind_out, float *quant_nchw)
{
    if (mode < 0 || mode > 6) return ORC_EINVAL;
    const int h8 = h / 2, w8 = w / 2, h16 = h / 4, w16 = w / 4;
    const int64_t n16 = (int64_t)h16 * w16, n8 = (int64_t)h8 * w8, n4 = (int64_t)h * w;
    int64_t *sym = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n4 + 8));
    if (!sym) return ORC_ENOMEM;
    int rc = ORC_OK;
    /* masks */
    for (int64_t i = 0; i < n16; ++i) mc_out[i] = (mode == 4);
    for (int64_t i = 0; i < n8; ++i) mm_out[i] = (mode == 5);
    for (int64_t i = 0; i < n4; ++i) mf_out[i] = (mode == 6);
    if (mode == 0 || mode == 2 || mode == 3) {
        int64_t r = orc_bits_decode(in + slot_off[3], sizes[3], mc_out, n16);
        if (r != n16) rc = ORC_EINVAL;
    }
    if (rc == ORC_OK && (mode == 0 || mode == 1)) {
        int64_t r = orc_bits_decode(in + slot_off[4], sizes[4], mm_out, n8);
        if (r != n8) rc = ORC_EINVAL;
    }
    if (rc == ORC_OK && mode == 3)
        for (int y = 0; y < h8; ++y)
            for (int x = 0; x < w8; ++x) mm_out[(int64_t)y * w8 + x] = 1 - mc_out[(int64_t)(y / 2) * w16 + x / 2];
    if (rc == ORC_OK && mode <= 2)
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x)
                mf_out[(int64_t)y * w + x] = 1 - mm_out[(int64_t)(y / 2) * w8 + x / 2] - mc_out[(int64_t)(y / 4) * w16 + x / 4];
    for (int64_t i = 0; i < n4; ++i) ind_out[i] = 0;
    for (int s = 0; s < 3 && rc == ORC_OK; ++s) {
        if (!ORC_STREAMS[mode][s]) continue;
        const int rep = s == 0 ? 4 : (s == 1 ? 2 : 1);
        const int gh = h / rep, gw = w / rep;
        const int64_t *m = s == 0 ? mc_out : (s == 1 ? mm_out : mf_out);
        int64_t cnt = orc_huff_decode(in + slot_off[s], sizes[s], K, left, right, root, sym, n4);
        if (cnt == -100) {
            /* empty stream: zeros for coarse/medium (model.py:284-290); for the fine stream the
             * reference would fail on torch.tensor(None) unless the mask is empty as well. */
            cnt = 0;
            if (s < 2) continue;
        }
        if (cnt < 0) { rc = (int)cnt; break; }
        int64_t pop = 0;
        for (int64_t i = 0; i < (int64_t)gh * gw; ++i) pop += (m[i] == 1);
        if (pop != cnt) { rc = ORC_EINVAL; break; }
        int64_t j = 0;
        for (int y = 0; y < gh; ++y)
            for (int x = 0; x < gw; ++x)
                if (m[(int64_t)y * gw + x] == 1) {
                    int64_t v = sym[j++];
                    for (int dy = 0; dy < rep; ++dy)
                        for (int dx = 0; dx < rep; ++dx) ind_out[(int64_t)(y * rep + dy) * w + x * rep + dx] += v;
                }
    }
    if (rc == ORC_OK && quant_nchw)
        for (int64_t p = 0; p < n4; ++p) {
            int64_t k = ind_out[p];
            if (k < 0 || k >= K) { rc = ORC_EINVAL; break; }
            for (int c = 0; c < 4; ++c) quant_nchw[c * n4 + p] = codebook[4 * k + c];
        }
    free(sym);
    return rc;
}

/* ------------------------------------------------------------------------------------------
 * a5  TripleGrainFixedEntropyRouter.forward        CGIC/modules/vqvae/RouterTriple.py:15-96
 * The thresholds are taken over ALL B images of the call (flatten(), RouterTriple.py:21,27);
 * per-image behaviour = call with B == 1.  mode and the two ranks k_c, k_m come from the host
 * (Python round() on doubles, RouterTriple.py:23,30,42,54,66).
 * ------------------------------------------------------------------------------------------ */
static int cmp_f32(const void *a, const void *b)
{
    float x = *(const float *)a, y = *(const float *)b;
    return (x > y) - (x < y);
}

static float kth_smallest(const float *v, int64_t n, int64_t k /* rank, 0 -> index 0 */)
{
    float *s = (float *)malloc(sizeof(float) * (size_t)n);
    memcpy(s, v, sizeof(float) * (size_t)n);
    qsort(s, (size_t)n, sizeof(float), cmp_f32);
    float t = s[k != 0 ? k - 1 : 0];
    free(s);
    return t;
}

int orc_router(const float *e16, const float *e8, int B, int h16, int w16, int mode, int64_t k_c, int64_t k_m,
               int32_t *mc, int32_t *mm, int32_t *mf)
{
    const int h8 = 2 * h16, w8 = 2 * w16, h = 4 * h16, w = 4 * w16;
    const int64_t n16 = (int64_t)B * h16 * w16, n8 = (int64_t)B * h8 * w8, n4 = (int64_t)B * h * w;
    if (mode < 0 || mode > 6) return ORC_EINVAL;
    for (int64_t i = 0; i < n16; ++i) mc[i] = (mode == 4);
    for (int64_t i = 0; i < n8; ++i) mm[i] = (mode == 5);
    for (int64_t i = 0; i < n4; ++i) mf[i] = (mode == 6);
    if (mode >= 4) return ORC_OK;
    if (mode == 0 || mode == 2 || mode == 3) {
        float thr = kth_smallest(e16, n16, k_c);
        for (int64_t i = 0; i < n16; ++i) mc[i] = e16[i] < thr;
    }
    if (mode == 0) {
        float *z8 = (float *)malloc(sizeof(float) * (size_t)n8);
        for (int b = 0; b < B; ++b)
            for (int y = 0; y < h8; ++y)
                for (int x = 0; x < w8; ++x) {
                    int64_t i = ((int64_t)b * h8 + y) * w8 + x;
                    float g = (float)mc[((int64_t)b * h16 + y / 2) * w16 + x / 2];
                    z8[i] = e8[i] * (1.0f - g);
                }
        float thr = kth_smallest(z8, n8, k_m);
        for (int b = 0; b < B; ++b)
            for (int y = 0; y < h8; ++y)
                for (int x = 0; x < w8; ++x) {
                    int64_t i = ((int64_t)b * h8 + y) * w8 + x;
                    mm[i] = (e8[i] < thr) && !mc[((int64_t)b * h16 + y / 2) * w16 + x / 2];
                }
        free(z8);
    } else if (mode == 1) {
        float thr = kth_smallest(e8, n8, k_m);
        for (int64_t i = 0; i < n8; ++i) mm[i] = e8[i] < thr;
    } else if (mode == 3) {
        for (int b = 0; b < B; ++b)
            for (int y = 0; y < h8; ++y)
                for (int x = 0; x < w8; ++x)
                    mm[((int64_t)b * h8 + y) * w8 + x] = 1 - mc[((int64_t)b * h16 + y / 2) * w16 + x / 2];
    }
    if (mode <= 2)
        for (int b = 0; b < B; ++b)
            for (int y = 0; y < h; ++y)
                for (int x = 0; x < w; ++x)
                    mf[((int64_t)b * h + y) * w + x] =
                        (1 - mc[((int64_t)b * h16 + y / 4) * w16 + x / 4] - mm[((int64_t)b * h8 + y / 2) * w8 + x / 2]) != 0;
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------
 * a6  mask-mix tail of Encoder.forward        CGIC/modules/vqvae/vqvae_blocks.py:361-366
 *   h = up4(h_c)*up4(m0) + up2(h_m)*up2(m1) + h_f*m2       (fp32, left to right)
 * ------------------------------------------------------------------------------------------ */
int orc_mask_mix(const float *hc, const float *hm, const float *hf, const int32_t *mc, const int32_t *mm,
                 const int32_t *mf, int B, int C, int h, int w, float *out)
{
    const int h8 = h / 2, w8 = w / 2, h16 = h / 4, w16 = w / 4;
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < C; ++c)
            for (int y = 0; y < h; ++y)
                for (int x = 0; x < w; ++x) {
                    float a = hc[(((int64_t)b * C + c) * h16 + y / 4) * w16 + x / 4] *
                              (float)mc[((int64_t)b * h16 + y / 4) * w16 + x / 4];
                    float m = hm[(((int64_t)b * C + c) * h8 + y / 2) * w8 + x / 2] *
                              (float)mm[((int64_t)b * h8 + y / 2) * w8 + x / 2];
                    float f = hf[(((int64_t)b * C + c) * h + y) * w + x] * (float)mf[((int64_t)b * h + y) * w + x];
                    float s = a + m;
                    out[(((int64_t)b * C + c) * h + y) * w + x] = s + f;
                }
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------
 * a4  Entropy.forward / Entropy.entropy                 CGIC/models/model.py:440-483
 *   gray = .2989 R + .5870 G + .1140 B; per p x p patch and 32 bins (passed in by the caller as
 *   torch.linspace(-1,1,32) values): pdf_j = mean_i exp(-0.5*((v_i-bin_j)/0.01)^2);
 *   pdf = pdf/(sum+1e-40) + 1e-40; H = -sum pdf*log(pdf).  Float tolerance only (the
 *   reference's reduction order is torch's); sums here are sequential fp32.
 * ------------------------------------------------------------------------------------------ */
int orc_entropy(const float *x /* B,3,H,W */, int B, int H, int W, int psize, const float *bins /* 32 */,
                float *out /* B, H/p, W/p */)
{
    const int hn = H / psize, wn = W / psize;
    const float sigma = 0.01f, eps = 1e-40f;
    const int64_t plane = (int64_t)H * W;
    for (int b = 0; b < B; ++b)
        for (int py = 0; py < hn; ++py)
            for (int px = 0; px < wn; ++px) {
                float pdf[32];
                for (int j = 0; j < 32; ++j) pdf[j] = 0.f;
                for (int dy = 0; dy < psize; ++dy)
                    for (int dx = 0; dx < psize; ++dx) {
                        int64_t o = (int64_t)(py * psize + dy) * W + px * psize + dx;
                        const float *xb = x + (int64_t)b * 3 * plane;
                        float g = 0.2989f * xb[o];
                        float t = 0.5870f * xb[plane + o];
                        g = g + t;
                        t = 0.1140f * xb[2 * plane + o];
                        g = g + t;
                        for (int j = 0; j < 32; ++j) {
                            float r = (g - bins[j]) / sigma;
                            float q = r * r;
                            pdf[j] += expf(-0.5f * q);
                        }
                    }
                float norm = 0.f;
                for (int j = 0; j < 32; ++j) { pdf[j] = pdf[j] / (float)(psize * psize); norm += pdf[j]; }
                norm += eps;
                float ent = 0.f;
                for (int j = 0; j < 32; ++j) {
                    float p = pdf[j] / norm + eps;
                    ent += p * logf(p);
                }
                out[((int64_t)b * hn + py) * wn + px] = -ent;
            }
    return ORC_OK;
}
