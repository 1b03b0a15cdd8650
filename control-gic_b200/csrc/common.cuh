// Shared internals of libcgic_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "../../include/cgic_b200.h"

namespace cgic {

void set_error(const char *fmt, ...);

#define CGIC_CUDA_CHECK(expr)                                                                  \
    do {                                                                                       \
        cudaError_t err__ = (expr);                                                            \
        if (err__ != cudaSuccess) {                                                            \
            ::cgic::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__), __FILE__, __LINE__); \
            return CGIC_ECUDA;                                                                 \
        }                                                                                      \
    } while (0)

#define CGIC_LAUNCH_CHECK() CGIC_CUDA_CHECK(cudaGetLastError())

#define CGIC_REQUIRE(cond, code, ...)                                                          \
    do {                                                                                       \
        if (!(cond)) {                                                                         \
            ::cgic::set_error(__VA_ARGS__);                                                    \
            return (code);                                                                     \
        }                                                                                      \
    } while (0)

// Device view of a Huffman table (passed by value to kernels).
struct DevTable {
    int K;
    int max_len;
    int min_len;
    int lut_bits;
    int root;
    const uint16_t *len;   // [K] code length in bits
    const uint32_t *off;   // [K] first word of the code in `pool`
    const uint32_t *pool;  // code bits, MSB first, left aligned, ceil(len/32) words per symbol
    const uint32_t *lut;   // [1 << lut_bits]: (sym << 8) | len, (lut2 offset << 8) | 0x80 | height, or (node << 8) | 0xFF to keep walking
    const uint32_t *lut2;  // second-level tables: (sym << 8) | bits used beyond lut_bits; == lut + lut_pad
    uint32_t lut_pad;          // words from lut to lut2 (16-byte multiple)
    uint32_t dec_stage_words;  // words a decoder CTA stages in shared memory: lut (+ lut2 when small)
    const uint2 *enc;          // [K] (code left aligned, length) when max_len <= 32, else null
    const int32_t *child;  // [2 * (2K-1)]: child[2*node + bit]; ids < K are leaves (= symbols)
};

int table_device_view(const cgic_table *t, DevTable *out);  // CGIC_EINVAL if not uploaded here
int table_max_len(const cgic_table *t);

// which of the five streams exist per compression mode (CGIC/models/model.py:225-260)
__host__ __device__ inline bool stream_present(int mode, int s)
{
    // byte `mode` of the constant, bit s: 0 ic, 1 im, 2 if, 3 mc, 4 mm  (modes 0..6: 1F 16 0D 0B 01 02 04)
    return (0x0402010B0D161FULL >> (8 * mode + s)) & 1;
}

#ifdef __CUDACC__
// ---- programmatic dependent launch (PDL): the hot kernels are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, so a kernel may start while its predecessor
// in the stream is still running.  Contract inside every such kernel: before pdl_wait() it may only
// touch its parameters, shared memory and IMMUTABLE global tables (the uploaded Huffman table);
// every read of data a predecessor may have produced and EVERY global write come after pdl_wait(),
// which returns once the predecessor grid has completed and its memory is visible.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// The step kernels choose WHEN their dependents may be launched (CGIC_PDL_EARLY: bit 0 VQ, bit 1 pack, bit 2 decode).
// Early = at the kernel's start: the dependent's CTAs become resident at once, stage their tables and sit in
// griddepcontrol.wait -- but they are then placed on whatever SMs have room while the primary still runs, not in
// linear block order on an empty machine.  Late = no explicit trigger: the dependent grid is launched as the primary's
// CTAs exit (its launch latency still overlaps the primary's drain).
#ifndef CGIC_PDL_EARLY
#define CGIC_PDL_EARLY 0
#endif
template <int BIT>
__device__ __forceinline__ void pdl_trigger_step()
{
    if (CGIC_PDL_EARLY & BIT) pdl_launch_dependents();
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args)
{
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    static const bool no_pdl = getenv("CGIC_NO_PDL") != nullptr;  // diagnosis only: plain stream-ordered launches
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = no_pdl ? 0 : 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- TMA bulk copy (cp.async.bulk, SASS UBLKCP) global -> shared with mbarrier completion ----
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// one thread: init the barrier for a single arrival
__device__ __forceinline__ void mbar_init(unsigned long long *mbar)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// one thread: announce `bytes` and start the copy; dst / src 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, unsigned long long *mbar)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(mbar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_addr(mbar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *mbar, uint32_t parity)
{
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(smem_addr(mbar)), "r"(parity)
                     : "memory");
}
#endif

struct PackLayout {
    int64_t off[5];
    int64_t cap[5];
    int64_t stride;
};
PackLayout make_pack_layout(int max_len, int h, int w);

// Phase tracing (debug builds only, -DCGIC_TRACE): thread 0 of a CTA stamps %globaltimer into a
// per-translation-unit device array; cgic_trace_<unit>() copies it out (profiles/trace_*.py).
#ifdef CGIC_TRACE
#define CGIC_TRACE_SLOTS 16
#define CGIC_TRACE_CTAS 1024
#define CGIC_TRACE_DECL(unit)                                                                        \
    static __device__ unsigned long long g_trace_##unit[CGIC_TRACE_CTAS * CGIC_TRACE_SLOTS];         \
    extern "C" __attribute__((visibility("default"))) int cgic_trace_##unit(unsigned long long *host)  \
    {                                                                                                \
        return (int)cudaMemcpyFromSymbol(host, g_trace_##unit, sizeof(g_trace_##unit));              \
    }
#define CGIC_STAMP(unit, k)                                                                          \
    do {                                                                                             \
        const unsigned cta__ = blockIdx.y * gridDim.x + blockIdx.x;                                  \
        if (threadIdx.x == 0 && cta__ < CGIC_TRACE_CTAS) {                                           \
            unsigned long long t__;                                                                  \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t__));                                  \
            g_trace_##unit[cta__ * CGIC_TRACE_SLOTS + (k)] = t__;                                    \
        }                                                                                            \
    } while (0)
#else
#define CGIC_TRACE_DECL(unit)
#define CGIC_STAMP(unit, k)
#endif

static inline cudaStream_t as_stream(cgic_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// Per-DEVICE one-time set-up (cudaFuncSetAttribute and the SM count belong to a device, not to the process):
// ensure_smem opts kernel `fn` in to `bytes` of dynamic shared memory on the current device (no-op at <= 48 KB or when
// already granted there); device_sm_count caches cudaDevAttrMultiProcessorCount of the current device.
int ensure_smem(const void *fn, size_t bytes);
int device_sm_count(int *n_sm);

// Dispatch knobs (cgic_tune): which kernel variant serves a call.  Defaults come from the environment once
// (CGIC_DS_CLUSTER, CGIC_FUSED_ENCODE, CGIC_NO_SMALL_KERNELS), cgic_tune() changes them at run time (tests, A-B runs).
int tune_fused_decode_ctas();  // 0 = automatic (by batch size), 1 / 2 / 4 = fused small-grid decoder with that many CTAs per image, -1 = never
int tune_fused_encode();       // 0 = two launches (default), 1 = one-CTA-per-image encoder on small grids
int tune_pack_image();         // 0 = automatic (one CTA per image on small grids once the batch exceeds the SM count), 1 = always, -1 = never

// Per-kernel device timing (cgic_prof_*): while enabled, every kernel launch of the library is
// bracketed by two CUDA events recorded on the launching stream.  Off by default; costs one
// relaxed load per launch when off.
struct ProfScope {
    ProfScope(const char *name, cudaStream_t stream);
    ~ProfScope();
    int slot;
    cudaStream_t stream;
};
#define CGIC_PROF(name, stream) ::cgic::ProfScope prof_scope__(name, stream)

}  // namespace cgic
