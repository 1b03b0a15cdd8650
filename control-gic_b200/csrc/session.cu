// Host-buffer session: what a reference-side plugin calls per batch when its tensors live in
// host memory.  Device buffers, workspaces, the stream and the uploaded codebook are created
// once; compress / decompress only enqueue H2D copies, the kernels and D2H copies, then
// synchronise.  Mirrors CGIC.compress (CGIC/models/model.py:206-401) minus the CNNs:
//   compress   = VectorQuantize2.forward (a1) + selection (a7) + 5-stream pack (a9/a11/a12)
//   decompress = decompress_string x5 (a10/a11) + re-assembly (a13) + codebook gather (a14)
#include <new>

#include "common.cuh"

struct cgic_session {
    int B = 0, h = 0, w = 0, mode = 0, K = 0;
    const cgic_table *table = nullptr;
    cgic::PackLayout L{};
    cudaStream_t stream = nullptr;
    unsigned char *arena = nullptr;
    // carved device buffers
    float *codebook = nullptr, *z = nullptr, *zq = nullptr, *quant = nullptr;
    int32_t *mc = nullptr, *mm = nullptr, *mf = nullptr, *sizes = nullptr, *status = nullptr;
    int64_t *idx = nullptr, *dmc = nullptr, *dmm = nullptr, *dmf = nullptr, *ind = nullptr;
    uint8_t *bytes = nullptr;
    double *sqerr = nullptr;
    void *ws_vq = nullptr, *ws_un = nullptr;
    size_t ws_vq_bytes = 0, ws_un_bytes = 0;
};

extern "C" int cgic_session_create(int B, int h, int w, int mode, const cgic_table *t, const float *codebook_host, int K,
                                   cgic_session **out)
{
    CGIC_REQUIRE(out && t && codebook_host, CGIC_EINVAL, "cgic_session_create: null argument");
    CGIC_REQUIRE(B > 0 && h > 0 && w > 0 && h % 4 == 0 && w % 4 == 0 && mode >= 0 && mode <= 6 && K >= 2, CGIC_EINVAL,
                 "cgic_session_create: bad argument B=%d h=%d w=%d mode=%d K=%d", B, h, w, mode, K);
    CGIC_REQUIRE(K == cgic_huff_num_symbols(t), CGIC_EINVAL, "cgic_session_create: codebook has %d rows, table %d symbols", K,
                 cgic_huff_num_symbols(t));
    int rc = cgic_huff_upload(const_cast<cgic_table *>(t));
    if (rc) return rc;
    auto *s = new (std::nothrow) cgic_session();
    CGIC_REQUIRE(s, CGIC_ENOMEM, "cgic_session_create: out of memory");
    s->B = B;
    s->h = h;
    s->w = w;
    s->mode = mode;
    s->K = K;
    s->table = t;
    s->L = cgic::make_pack_layout(cgic::table_max_len(t), h, w);
    const size_t n4 = (size_t)B * h * w, n8 = n4 / 4, n16 = n4 / 16;
    s->ws_vq_bytes = cgic_vq_workspace_bytes((int64_t)n4);
    s->ws_un_bytes = cgic_unpack_workspace_bytes(B, h, w);
    size_t o = 0;
    auto take = [&](size_t bytes) {
        const size_t at = o;
        o += (bytes + 255) / 256 * 256;
        return at;
    };
    const size_t o_cb = take((size_t)K * 16), o_z = take(n4 * 16), o_zq = take(n4 * 16), o_quant = take(n4 * 16),
                 o_mc = take(n16 * 4), o_mm = take(n8 * 4), o_mf = take(n4 * 4), o_sizes = take((size_t)B * 5 * 4),
                 o_status = take((size_t)B * 4), o_idx = take(n4 * 8), o_dmc = take(n16 * 8), o_dmm = take(n8 * 8),
                 o_dmf = take(n4 * 8), o_ind = take(n4 * 8), o_bytes = take((size_t)B * s->L.stride), o_sq = take(8),
                 o_wv = take(s->ws_vq_bytes), o_wu = take(s->ws_un_bytes);
    cudaError_t e = cudaMalloc(&s->arena, o);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMemcpy(s->arena + o_cb, codebook_host, (size_t)K * 16, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        cgic::set_error("cgic_session_create: %s", cudaGetErrorString(e));
        cgic_session_destroy(s);
        return e == cudaErrorMemoryAllocation ? CGIC_ENOMEM : CGIC_ECUDA;
    }
    unsigned char *a = s->arena;
    s->codebook = reinterpret_cast<float *>(a + o_cb);
    s->z = reinterpret_cast<float *>(a + o_z);
    s->zq = reinterpret_cast<float *>(a + o_zq);
    s->quant = reinterpret_cast<float *>(a + o_quant);
    s->mc = reinterpret_cast<int32_t *>(a + o_mc);
    s->mm = reinterpret_cast<int32_t *>(a + o_mm);
    s->mf = reinterpret_cast<int32_t *>(a + o_mf);
    s->sizes = reinterpret_cast<int32_t *>(a + o_sizes);
    s->status = reinterpret_cast<int32_t *>(a + o_status);
    s->idx = reinterpret_cast<int64_t *>(a + o_idx);
    s->dmc = reinterpret_cast<int64_t *>(a + o_dmc);
    s->dmm = reinterpret_cast<int64_t *>(a + o_dmm);
    s->dmf = reinterpret_cast<int64_t *>(a + o_dmf);
    s->ind = reinterpret_cast<int64_t *>(a + o_ind);
    s->bytes = a + o_bytes;
    s->sqerr = reinterpret_cast<double *>(a + o_sq);
    s->ws_vq = a + o_wv;
    s->ws_un = a + o_wu;
    *out = s;
    return CGIC_OK;
}

extern "C" void cgic_session_destroy(cgic_session *s)
{
    if (!s) return;
    if (s->stream) cudaStreamDestroy(s->stream);
    if (s->arena) cudaFree(s->arena);
    delete s;
}

extern "C" int64_t cgic_session_image_stride(const cgic_session *s) { return s ? s->L.stride : CGIC_EINVAL; }

extern "C" int cgic_session_compress_host(cgic_session *s, const float *z, const int32_t *m_c, const int32_t *m_m,
                                          const int32_t *m_f, uint8_t *bytes_out, int32_t *sizes_out, int64_t *idx_out,
                                          float *zq_out, double *sqerr_out)
{
    CGIC_REQUIRE(s && z && m_c && m_m && m_f && bytes_out && sizes_out, CGIC_EINVAL, "cgic_session_compress_host: null argument");
    const size_t n4 = (size_t)s->B * s->h * s->w, n8 = n4 / 4, n16 = n4 / 16;
    cudaStream_t st = s->stream;
    CGIC_CUDA_CHECK(cudaMemcpyAsync(s->z, z, n4 * 16, cudaMemcpyHostToDevice, st));
    CGIC_CUDA_CHECK(cudaMemcpyAsync(s->mc, m_c, n16 * 4, cudaMemcpyHostToDevice, st));
    CGIC_CUDA_CHECK(cudaMemcpyAsync(s->mm, m_m, n8 * 4, cudaMemcpyHostToDevice, st));
    CGIC_CUDA_CHECK(cudaMemcpyAsync(s->mf, m_f, n4 * 4, cudaMemcpyHostToDevice, st));
    int rc = cgic_vq_assign(s->z, s->B, s->h, s->w, s->codebook, s->K, s->idx, zq_out ? s->zq : nullptr,
                            sqerr_out ? s->sqerr : nullptr, s->ws_vq, s->ws_vq_bytes, st);
    if (rc) return rc;
    rc = cgic_pack(s->idx, s->mc, s->mm, s->mf, s->B, s->h, s->w, s->mode, s->table, s->bytes, s->sizes, st);
    if (rc) return rc;
    CGIC_CUDA_CHECK(cudaMemcpyAsync(bytes_out, s->bytes, (size_t)s->B * s->L.stride, cudaMemcpyDeviceToHost, st));
    CGIC_CUDA_CHECK(cudaMemcpyAsync(sizes_out, s->sizes, (size_t)s->B * 5 * 4, cudaMemcpyDeviceToHost, st));
    if (idx_out) CGIC_CUDA_CHECK(cudaMemcpyAsync(idx_out, s->idx, n4 * 8, cudaMemcpyDeviceToHost, st));
    if (zq_out) CGIC_CUDA_CHECK(cudaMemcpyAsync(zq_out, s->zq, n4 * 16, cudaMemcpyDeviceToHost, st));
    if (sqerr_out) CGIC_CUDA_CHECK(cudaMemcpyAsync(sqerr_out, s->sqerr, 8, cudaMemcpyDeviceToHost, st));
    CGIC_CUDA_CHECK(cudaStreamSynchronize(st));
    return CGIC_OK;
}

extern "C" int cgic_session_decompress_host(cgic_session *s, const uint8_t *bytes, const int32_t *sizes, int64_t *mc_out,
                                            int64_t *mm_out, int64_t *mf_out, int64_t *ind_out, float *quant_out,
                                            int32_t *status_out)
{
    CGIC_REQUIRE(s && bytes && sizes && ind_out && status_out, CGIC_EINVAL, "cgic_session_decompress_host: null argument");
    const size_t n4 = (size_t)s->B * s->h * s->w, n8 = n4 / 4, n16 = n4 / 16;
    cudaStream_t st = s->stream;
    CGIC_CUDA_CHECK(cudaMemcpyAsync(s->bytes, bytes, (size_t)s->B * s->L.stride, cudaMemcpyHostToDevice, st));
    CGIC_CUDA_CHECK(cudaMemcpyAsync(s->sizes, sizes, (size_t)s->B * 5 * 4, cudaMemcpyHostToDevice, st));
    int rc = cgic_unpack(s->bytes, s->sizes, s->B, s->h, s->w, s->mode, s->table, s->codebook, s->dmc, s->dmm, s->dmf, s->ind,
                         quant_out ? s->quant : nullptr, s->status, s->ws_un, s->ws_un_bytes, st);
    if (rc) return rc;
    CGIC_CUDA_CHECK(cudaMemcpyAsync(ind_out, s->ind, n4 * 8, cudaMemcpyDeviceToHost, st));
    if (quant_out) CGIC_CUDA_CHECK(cudaMemcpyAsync(quant_out, s->quant, n4 * 16, cudaMemcpyDeviceToHost, st));
    if (mc_out) CGIC_CUDA_CHECK(cudaMemcpyAsync(mc_out, s->dmc, n16 * 8, cudaMemcpyDeviceToHost, st));
    if (mm_out) CGIC_CUDA_CHECK(cudaMemcpyAsync(mm_out, s->dmm, n8 * 8, cudaMemcpyDeviceToHost, st));
    if (mf_out) CGIC_CUDA_CHECK(cudaMemcpyAsync(mf_out, s->dmf, n4 * 8, cudaMemcpyDeviceToHost, st));
    CGIC_CUDA_CHECK(cudaMemcpyAsync(status_out, s->status, (size_t)s->B * 4, cudaMemcpyDeviceToHost, st));
    CGIC_CUDA_CHECK(cudaStreamSynchronize(st));
    return CGIC_OK;
}
