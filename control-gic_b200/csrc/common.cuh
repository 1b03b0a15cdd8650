// Shared internals of libcgic_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>

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
    int lut_bits;
    int root;
    const uint16_t *len;   // [K] code length in bits
    const uint32_t *off;   // [K] first word of the code in `pool`
    const uint32_t *pool;  // code bits, MSB first, left aligned, ceil(len/32) words per symbol
    const uint32_t *lut;   // [1 << lut_bits]: (sym << 8) | len, or (node << 8) | 0xFF to keep walking
    const int32_t *child;  // [2 * (2K-1)]: child[2*node + bit]; ids < K are leaves (= symbols)
};

int table_device_view(const cgic_table *t, DevTable *out);  // CGIC_EINVAL if not uploaded here
int table_max_len(const cgic_table *t);

// which of the five streams exist per compression mode (CGIC/models/model.py:225-260)
__host__ __device__ inline bool stream_present(int mode, int s)
{
    // bit s of the entry: 0 ic, 1 im, 2 if, 3 mc, 4 mm
    const unsigned char tab[7] = {0x1F, 0x16, 0x0D, 0x0B, 0x01, 0x02, 0x04};
    return (tab[mode] >> s) & 1;
}

struct PackLayout {
    int64_t off[5];
    int64_t cap[5];
    int64_t stride;
};
PackLayout make_pack_layout(int max_len, int h, int w);

static inline cudaStream_t as_stream(cgic_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

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
