// Prepared-codebook layout shared by the index builder (codebook.cu) and the search (vq_assign.cu).
#pragma once
#include "common.cuh"

namespace cgic {

constexpr int CB_NB = 256;        // fine bins per dimension
constexpr int CB_PAD = 48;        // bins beyond the codebook's min / max on each side
constexpr int CB_G = 12;          // cells per dimension
constexpr int CB_RW = 64;         // u16 words per cell record: count + up to 63 candidates
constexpr int CB_MAX_K = 4096;
constexpr int CB_NDOM = 16;       // dominators per stage: 8 nearest-to-centre (one per codebook slice), then the nearest survivor to each corner
constexpr int CB_HDR = 512;       // header bytes
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


}  // namespace cgic
