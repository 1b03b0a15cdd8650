// Static Huffman code table: host builder (bit-compatible with the reference's heapq-driven
// construction) + device upload + decode LUT.
//
// Replaces HuffmanCoding.__init__/make_heap/merge_nodes/make_codes
// (CGIC/tools/indices_coding.py:10-17, 46-75).  The reference's codes are whatever CPython's
// heapq does with nodes compared on frequency alone, so the builder below re-implements the
// binary heap of Lib/heapq.py operation for operation:
//   push : append, then move the new node towards the root while it is strictly smaller than
//          its parent;
//   pop  : remove the last node, place it at the root, descend along the smaller child (the
//          right one when the children compare equal), then move it back up as in push.
// Nodes are pushed in the caller-given order; the two nodes popped for a merge become the
// '0' and '1' child in pop order.
#include <algorithm>
#include <atomic>
#include <map>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"

namespace cgic {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ---- per-kernel event timing ---------------------------------------------------------------
namespace {
struct ProfRecord {
    const char *name;
    cudaEvent_t e0, e1;
};
std::mutex g_prof_mu;
std::vector<ProfRecord> g_prof;
std::atomic<int> g_prof_on{0};
}  // namespace

ProfScope::ProfScope(const char *name, cudaStream_t st) : slot(-1), stream(st)
{
    if (!g_prof_on.load(std::memory_order_relaxed)) return;
    ProfRecord r{name, nullptr, nullptr};
    if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return;
    cudaEventRecord(r.e0, st);
    std::lock_guard<std::mutex> lock(g_prof_mu);
    g_prof.push_back(r);
    slot = (int)g_prof.size() - 1;
}

ProfScope::~ProfScope()
{
    if (slot < 0) return;
    std::lock_guard<std::mutex> lock(g_prof_mu);
    cudaEventRecord(g_prof[slot].e1, stream);
}

int ensure_smem(const void *fn, size_t bytes)
{
    static std::mutex mu;
    static std::map<std::pair<int, const void *>, size_t> granted;
    if (bytes <= 48 * 1024) return CGIC_OK;
    int dev = -1;
    CGIC_CUDA_CHECK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    size_t &g = granted[{dev, fn}];
    if (bytes > g) {
        CGIC_CUDA_CHECK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        g = bytes;
    }
    return CGIC_OK;
}

int device_sm_count(int *n_sm)
{
    static std::mutex mu;
    static std::map<int, int> count;
    int dev = -1;
    CGIC_CUDA_CHECK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    int &c = count[dev];
    if (!c) CGIC_CUDA_CHECK(cudaDeviceGetAttribute(&c, cudaDevAttrMultiProcessorCount, dev));
    *n_sm = c;
    return CGIC_OK;
}

namespace {
std::atomic<int> g_tune_decode{-100}, g_tune_encode{-100}, g_tune_pack{-100};
}
int tune_pack_image()
{
    int v = g_tune_pack.load(std::memory_order_relaxed);
    if (v == -100) {
        v = getenv("CGIC_NO_SMALL_KERNELS") || getenv("CGIC_NO_PACK_IMAGE") ? -1 : 0;
        g_tune_pack.store(v);
    }
    return v;
}
int tune_fused_decode_ctas()
{
    int v = g_tune_decode.load(std::memory_order_relaxed);
    if (v == -100) {
        v = 0;
        if (getenv("CGIC_NO_SMALL_KERNELS")) v = -1;
        else if (const char *e = getenv("CGIC_DS_CLUSTER")) {
            const int c = atoi(e);
            if (c == 1 || c == 2 || c == 4) v = c;
        }
        g_tune_decode.store(v);
    }
    return v;
}
int tune_fused_encode()
{
    int v = g_tune_encode.load(std::memory_order_relaxed);
    if (v == -100) {
        v = getenv("CGIC_FUSED_ENCODE") && !getenv("CGIC_NO_SMALL_KERNELS") ? 1 : 0;
        g_tune_encode.store(v);
    }
    return v;
}

PackLayout make_pack_layout(int max_len, int h, int w)
{
    PackLayout L;
    const int64_t n4 = (int64_t)h * w, n8 = (int64_t)(h / 2) * (w / 2), n16 = (int64_t)(h / 4) * (w / 4);
    const int64_t raw[5] = {n16 * max_len / 8 + 2, n8 * max_len / 8 + 2, n4 * max_len / 8 + 2, n16 / 8 + 2, n8 / 8 + 2};
    int64_t o = 0;
    for (int s = 0; s < 5; ++s) {
        L.off[s] = o;
        L.cap[s] = (raw[s] + 15) / 16 * 16;
        o += L.cap[s];
    }
    L.stride = o;
    return L;
}

}  // namespace cgic

struct cgic_table {
    int K = 0;
    int max_len = 0;
    int min_len = 0;
    int lut_bits = 0;
    int root = 0;
    std::vector<uint16_t> len;
    std::vector<uint32_t> off;
    std::vector<uint32_t> pool;
    std::vector<uint32_t> lut;
    std::vector<uint32_t> lut2;  // second-level decode tables (codes a little longer than lut_bits)
    std::vector<uint32_t> dec;   // lut ++ lut2, each padded to 16 bytes: one bulk copy stages both in shared memory
    std::vector<uint2> enc;      // [K] (code bits left aligned, length) when max_len <= 32, else empty
    size_t lut_pad = 0;
    std::vector<int32_t> child;  // 2 * (2K-1)
    std::vector<std::string> code;
    // device copies, one per device the table was uploaded to (immutable once made)
    struct DevCopy {
        void *blob = nullptr;
        cgic::DevTable view{};
    };
    mutable std::mutex mu;
    std::map<int, DevCopy> dev;
};

namespace {

struct FreqHeap {
    const std::vector<int64_t> &freq;
    std::vector<int> a;
    explicit FreqHeap(const std::vector<int64_t> &f) : freq(f) {}
    bool less(int x, int y) const { return freq[x] < freq[y]; }
    void rise(size_t pos)
    {
        const int item = a[pos];
        while (pos > 0) {
            const size_t up = (pos - 1) / 2;
            if (!less(item, a[up])) break;
            a[pos] = a[up];
            pos = up;
        }
        a[pos] = item;
    }
    void push(int node)
    {
        a.push_back(node);
        rise(a.size() - 1);
    }
    int pop()
    {
        const int tail = a.back();
        a.pop_back();
        if (a.empty()) return tail;
        const int top = a[0];
        size_t pos = 0;
        const size_t n = a.size();
        for (size_t kid = 1; kid < n; kid = 2 * pos + 1) {
            if (kid + 1 < n && !less(a[kid], a[kid + 1])) ++kid;
            a[pos] = a[kid];
            pos = kid;
        }
        a[pos] = tail;
        rise(pos);
        return top;
    }
};

}  // namespace

extern "C" {

int cgic_abi_version(void) { return CGIC_ABI_VERSION; }

int cgic_tune(const char *key, int value)
{
    CGIC_REQUIRE(key, CGIC_EINVAL, "cgic_tune: null key");
    if (!strcmp(key, "fused_decode_ctas")) {
        CGIC_REQUIRE(value == -1 || value == 0 || value == 1 || value == 2 || value == 4, CGIC_EINVAL, "cgic_tune: fused_decode_ctas must be -1, 0, 1, 2 or 4");
        cgic::g_tune_decode.store(value);
        return CGIC_OK;
    }
    if (!strcmp(key, "fused_encode")) {
        CGIC_REQUIRE(value == 0 || value == 1, CGIC_EINVAL, "cgic_tune: fused_encode must be 0 or 1");
        cgic::g_tune_encode.store(value);
        return CGIC_OK;
    }
    if (!strcmp(key, "pack_image")) {
        CGIC_REQUIRE(value == 0 || value == -1 || value == 1, CGIC_EINVAL, "cgic_tune: pack_image must be -1, 0 or 1");
        cgic::g_tune_pack.store(value);
        return CGIC_OK;
    }
    cgic::set_error("cgic_tune: unknown key '%s'", key);
    return CGIC_EINVAL;
}

const char *cgic_last_error(void) { return cgic::g_err; }

int cgic_prof_enable(int on)
{
    std::lock_guard<std::mutex> lock(cgic::g_prof_mu);
    for (auto &r : cgic::g_prof) {
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
    }
    cgic::g_prof.clear();
    cgic::g_prof_on.store(on ? 1 : 0);
    return CGIC_OK;
}

int cgic_prof_report(char *buf, int cap)
{
    CGIC_REQUIRE(buf && cap > 0, CGIC_EINVAL, "cgic_prof_report: bad buffer");
    CGIC_CUDA_CHECK(cudaDeviceSynchronize());
    std::lock_guard<std::mutex> lock(cgic::g_prof_mu);
    std::map<std::string, std::pair<long, double>> acc;
    for (auto &r : cgic::g_prof) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.e0, r.e1) != cudaSuccess) continue;
        auto &a = acc[r.name];
        a.first += 1;
        a.second += ms;
    }
    std::string out;
    char line[160];
    for (auto &kv : acc) {
        snprintf(line, sizeof(line), "%s %ld %.6f\n", kv.first.c_str(), kv.second.first, kv.second.second);
        out += line;
    }
    CGIC_REQUIRE((int)out.size() + 1 <= cap, CGIC_ESPACE, "cgic_prof_report: buffer of %d too small for %zu", cap, out.size());
    std::memcpy(buf, out.c_str(), out.size() + 1);
    return (int)out.size();
}

int cgic_huff_build(const int64_t *freq, const int32_t *order, int K, cgic_table **out)
{
    CGIC_REQUIRE(freq && out, CGIC_EINVAL, "cgic_huff_build: null argument");
    CGIC_REQUIRE(K >= 2 && K <= 65536, CGIC_EINVAL, "cgic_huff_build: K=%d outside [2, 65536]", K);
    const int nn = 2 * K - 1;
    std::vector<int64_t> f(nn, 0);
    std::vector<char> seen(K, 0);
    for (int s = 0; s < K; ++s) f[s] = freq[s];
    FreqHeap heap(f);
    heap.a.reserve(K);
    for (int i = 0; i < K; ++i) {
        const int s = order ? order[i] : i;
        CGIC_REQUIRE(s >= 0 && s < K && !seen[s], CGIC_EINVAL, "cgic_huff_build: order[%d]=%d is not a permutation", i, s);
        seen[s] = 1;
        heap.push(s);
    }
    auto *t = new (std::nothrow) cgic_table();
    CGIC_REQUIRE(t, CGIC_ENOMEM, "cgic_huff_build: out of memory");
    t->K = K;
    t->child.assign(2 * (size_t)nn, -1);
    int next = K;
    while (heap.a.size() > 1) {
        const int zero_kid = heap.pop();
        const int one_kid = heap.pop();
        f[next] = f[zero_kid] + f[one_kid];
        t->child[2 * (size_t)next] = zero_kid;
        t->child[2 * (size_t)next + 1] = one_kid;
        heap.push(next++);
    }
    t->root = heap.pop();

    // codes by iterative descent
    t->code.assign(K, std::string());
    std::vector<std::pair<int, std::string>> todo;
    todo.emplace_back(t->root, std::string());
    while (!todo.empty()) {
        auto [node, prefix] = std::move(todo.back());
        todo.pop_back();
        if (node < K) {
            t->code[node] = std::move(prefix);
            continue;
        }
        todo.emplace_back(t->child[2 * (size_t)node + 1], prefix + '1');
        todo.emplace_back(t->child[2 * (size_t)node], prefix + '0');
    }
    t->len.resize(K);
    t->off.resize(K);
    t->max_len = 0;
    for (int s = 0; s < K; ++s) {
        const std::string &c = t->code[s];
        t->len[s] = (uint16_t)c.size();
        t->max_len = std::max<int>(t->max_len, (int)c.size());
        t->min_len = s == 0 ? (int)c.size() : std::min<int>(t->min_len, (int)c.size());
        t->off[s] = (uint32_t)t->pool.size();
        const size_t nw = (c.size() + 31) / 32;
        for (size_t j = 0; j < nw; ++j) {
            uint32_t word = 0;
            for (size_t b = 0; b < 32 && 32 * j + b < c.size(); ++b)
                if (c[32 * j + b] == '1') word |= 0x80000000u >> b;
            t->pool.push_back(word);
        }
    }
    if (t->pool.empty()) t->pool.push_back(0);
    // decode LUT on the first lut_bits bits
#ifndef CGIC_LUT_BITS
#define CGIC_LUT_BITS 12
#endif
    t->lut_bits = std::min(t->max_len, CGIC_LUT_BITS);
    t->lut.resize((size_t)1 << t->lut_bits);
    for (uint32_t v = 0; v < t->lut.size(); ++v) {
        int node = t->root, used = 0;
        while (node >= K && used < t->lut_bits) {
            const int bit = (v >> (t->lut_bits - 1 - used)) & 1;
            node = t->child[2 * (size_t)node + bit];
            ++used;
        }
        t->lut[v] = node < K ? ((uint32_t)node << 8) | (uint32_t)used : ((uint32_t)node << 8) | 0xFFu;
    }
    // second level: an internal node reached after lut_bits bits whose subtree is at most 8 deep
    // gets a 2^height table indexed by the next `height` bits; entry = (sym << 8) | bits used.
    // First-level entry becomes (offset << 8) | 0x80 | height.  Deeper subtrees keep the tree walk.
    {
        std::vector<int> height(nn, 0);
        for (int node = K; node < nn; ++node)  // children are always created before their parent
            height[node] = 1 + std::max(height[t->child[2 * (size_t)node]], height[t->child[2 * (size_t)node + 1]]);
        for (uint32_t v = 0; v < t->lut.size(); ++v) {
            if ((t->lut[v] & 0xFFu) != 0xFFu) continue;
            const int top = (int)(t->lut[v] >> 8);
            const int hgt = height[top];
            if (hgt > 8 || t->lut2.size() + ((size_t)1 << hgt) >= ((size_t)1 << 24)) continue;
            const uint32_t base = (uint32_t)t->lut2.size();
            for (uint32_t u = 0; u < (1u << hgt); ++u) {
                int node = top, used = 0;
                while (node >= K) {
                    node = t->child[2 * (size_t)node + ((u >> (hgt - 1 - used)) & 1)];
                    ++used;
                }
                t->lut2.push_back(((uint32_t)node << 8) | (uint32_t)used);
            }
            t->lut[v] = (base << 8) | 0x80u | (uint32_t)hgt;
        }
        if (t->lut2.empty()) t->lut2.push_back(0);
    }
    t->lut_pad = (t->lut.size() + 3) / 4 * 4;
    t->dec.assign(t->lut_pad + (t->lut2.size() + 3) / 4 * 4, 0u);
    std::copy(t->lut.begin(), t->lut.end(), t->dec.begin());
    std::copy(t->lut2.begin(), t->lut2.end(), t->dec.begin() + t->lut_pad);
    if (t->max_len <= 32) {
        t->enc.resize((size_t)(K + 1) / 2 * 2);
        for (int s = 0; s < K; ++s) t->enc[s] = make_uint2(t->pool[t->off[s]], t->len[s]);
    }
    *out = t;
    return CGIC_OK;
}

void cgic_huff_free(cgic_table *t)
{
    if (!t) return;
    int cur = -1;
    const bool have_cur = cudaGetDevice(&cur) == cudaSuccess;
    for (auto &kv : t->dev) {
        if (!kv.second.blob) continue;
        if (cudaSetDevice(kv.first) == cudaSuccess) cudaFree(kv.second.blob);
    }
    if (have_cur && !t->dev.empty()) cudaSetDevice(cur);
    delete t;
}

int cgic_huff_num_symbols(const cgic_table *t) { return t ? t->K : CGIC_EINVAL; }
int cgic_huff_max_len(const cgic_table *t) { return t ? t->max_len : CGIC_EINVAL; }

int cgic_huff_code_len(const cgic_table *t, int sym)
{
    CGIC_REQUIRE(t && sym >= 0 && sym < t->K, CGIC_EINVAL, "cgic_huff_code_len: bad symbol %d", sym);
    return t->len[sym];
}

int cgic_huff_code(const cgic_table *t, int sym, char *buf, int cap)
{
    CGIC_REQUIRE(t && buf && sym >= 0 && sym < t->K, CGIC_EINVAL, "cgic_huff_code: bad argument");
    const std::string &c = t->code[sym];
    CGIC_REQUIRE((int)c.size() + 1 <= cap, CGIC_ESPACE, "cgic_huff_code: buffer of %d too small for %zu bits", cap, c.size());
    std::memcpy(buf, c.c_str(), c.size() + 1);
    return (int)c.size();
}

int cgic_huff_upload(cgic_table *t)
{
    CGIC_REQUIRE(t, CGIC_EINVAL, "cgic_huff_upload: null table");
    std::lock_guard<std::mutex> lock(t->mu);
    int dev = -1;
    CGIC_CUDA_CHECK(cudaGetDevice(&dev));
    if (t->dev.count(dev)) return CGIC_OK;
    auto pad = [](size_t n) { return (n + 255) / 256 * 256; };
    const size_t b_len = pad(t->len.size() * 2), b_off = pad(t->off.size() * 4), b_pool = pad(t->pool.size() * 4),
                 b_lut = pad(t->dec.size() * 4), b_child = pad(t->child.size() * 4), b_lut2 = pad(t->enc.size() * 8);
    const size_t total = b_len + b_off + b_pool + b_lut + b_child + b_lut2;
    std::vector<unsigned char> host(total, 0);
    size_t o = 0;
    std::memcpy(host.data() + o, t->len.data(), t->len.size() * 2);
    const size_t o_len = o;
    o += b_len;
    std::memcpy(host.data() + o, t->off.data(), t->off.size() * 4);
    const size_t o_off = o;
    o += b_off;
    std::memcpy(host.data() + o, t->pool.data(), t->pool.size() * 4);
    const size_t o_pool = o;
    o += b_pool;
    std::memcpy(host.data() + o, t->dec.data(), t->dec.size() * 4);
    const size_t o_lut = o;
    o += b_lut;
    std::memcpy(host.data() + o, t->child.data(), t->child.size() * 4);
    const size_t o_child = o;
    o += b_child;
    if (!t->enc.empty()) std::memcpy(host.data() + o, t->enc.data(), t->enc.size() * 8);
    const size_t o_lut2 = o;
    void *blob = nullptr;
    CGIC_CUDA_CHECK(cudaMalloc(&blob, total));
    cudaError_t e = cudaMemcpy(blob, host.data(), total, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        cudaFree(blob);
        cgic::set_error("cgic_huff_upload: cudaMemcpy failed: %s", cudaGetErrorString(e));
        return CGIC_ECUDA;
    }
    auto *base = static_cast<unsigned char *>(blob);
    cgic::DevTable v{};
    v.K = t->K;
    v.max_len = t->max_len;
    v.min_len = t->min_len;
    v.lut_bits = t->lut_bits;
    v.root = t->root;
    v.len = reinterpret_cast<const uint16_t *>(base + o_len);
    v.off = reinterpret_cast<const uint32_t *>(base + o_off);
    v.pool = reinterpret_cast<const uint32_t *>(base + o_pool);
    v.lut = reinterpret_cast<const uint32_t *>(base + o_lut);
    v.child = reinterpret_cast<const int32_t *>(base + o_child);
    v.lut2 = v.lut + t->lut_pad;
    v.lut_pad = (uint32_t)t->lut_pad;
    // stage lut + lut2 together when that stays small, else only the first level
    v.dec_stage_words = (uint32_t)(t->dec.size() * 4 <= 64 * 1024 ? t->dec.size() : t->lut_pad);
    v.enc = t->enc.empty() ? nullptr : reinterpret_cast<const uint2 *>(base + o_lut2);
    cgic_table::DevCopy copy;
    copy.blob = blob;
    copy.view = v;
    t->dev[dev] = copy;
    return CGIC_OK;
}

int64_t cgic_huff_stream_capacity(const cgic_table *t, int64_t n_symbols)
{
    if (!t || n_symbols < 0) return CGIC_EINVAL;
    return n_symbols * t->max_len / 8 + 2;
}

int cgic_pack_layout(const cgic_table *t, int h, int w, int64_t slot_off[5], int64_t slot_cap[5], int64_t *image_stride)
{
    CGIC_REQUIRE(t && h > 0 && w > 0 && h % 4 == 0 && w % 4 == 0, CGIC_EINVAL,
                 "cgic_pack_layout: token grid %dx%d must be positive multiples of 4", h, w);
    const cgic::PackLayout L = cgic::make_pack_layout(t->max_len, h, w);
    for (int s = 0; s < 5; ++s) {
        if (slot_off) slot_off[s] = L.off[s];
        if (slot_cap) slot_cap[s] = L.cap[s];
    }
    if (image_stride) *image_stride = L.stride;
    return CGIC_OK;
}

}  // extern "C"

namespace cgic {

int table_device_view(const cgic_table *t, DevTable *out)
{
    CGIC_REQUIRE(t, CGIC_EINVAL, "null Huffman table");
    int dev = -1;
    CGIC_CUDA_CHECK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(t->mu);
    const auto it = t->dev.find(dev);
    CGIC_REQUIRE(it != t->dev.end(), CGIC_EINVAL, "Huffman table is not uploaded to device %d (call cgic_huff_upload first)", dev);
    *out = it->second.view;
    return CGIC_OK;
}

int table_max_len(const cgic_table *t) { return t->max_len; }

}  // namespace cgic
