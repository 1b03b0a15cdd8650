/*
 * cgic_b200.h -- C ABI of the B200-native VQ + entropy-coding hot path of Control-GIC.
 *
 * This is the drop-in boundary: every entry point replaces one function (or one inlined block)
 * of the reference's Python hot path and is what a binding on the reference side would call
 * (ctypes stub: INTEGRATION.md).  Reference citations are relative to the reference repo root.
 *
 * Conventions
 *   - plain C: pointers, sizes, a stream handle; no torch / C++ types.
 *   - every `*_dev` / unqualified data pointer is DEVICE memory owned by the caller (torch
 *     allocates); the library never allocates or frees on the hot path and never synchronises
 *     unless the function name says `_host`.  Work is enqueued on `stream` (a cudaStream_t,
 *     e.g. torch.cuda.current_stream().cuda_stream).  `cgic_table` and `cgic_session` are the
 *     only library-owned handles.
 *   - return value: CGIC_OK (0) or a negative CGIC_E* code; cgic_last_error() returns a
 *     thread-local description.  Nothing ever calls exit() (the reference does, at
 *     CGIC/tools/indices_coding.py:102-104).
 *   - layouts are the reference's: latents / images NCHW fp32, indices int64, router masks
 *     int32 [B,1,h,w], decoded masks int64.  h, w below are the FINE token grid (image H/4, W/4)
 *     and must be multiples of 4 for the pack / unpack / router entry points.
 *   - stream order inside a packed image: 0 indices_coarse, 1 indices_medium, 2 indices_fine,
 *     3 mask_coarse, 4 mask_medium (the five files of CGIC/models/model.py:226-232).
 */
#ifndef CGIC_B200_H
#define CGIC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CGIC_ABI_VERSION 1

#if defined(__GNUC__)
#define CGIC_API __attribute__((visibility("default")))
#else
#define CGIC_API
#endif

#define CGIC_OK 0
#define CGIC_EINVAL (-1)  /* bad argument */
#define CGIC_ENOMEM (-2)  /* host or device allocation failed (init-time calls only) */
#define CGIC_ESPACE (-3)  /* caller buffer / workspace too small */
#define CGIC_ECUDA (-4)   /* CUDA runtime error, see cgic_last_error() */
#define CGIC_EFORMAT (-5) /* corrupt bitstream */

typedef void *cgic_stream_t;          /* cudaStream_t */
typedef struct cgic_table cgic_table; /* static Huffman code table (host + device copy) */
typedef struct cgic_session cgic_session;
typedef struct cgic_codebook cgic_codebook; /* prepared codebook: device copy + e^2 + cell index */

CGIC_API int cgic_abi_version(void);
/* Dispatch knobs for A-B runs and tests (process-wide; defaults: environment CGIC_DS_CLUSTER / CGIC_FUSED_ENCODE /
 * CGIC_NO_SMALL_KERNELS, else automatic).  Results never depend on them, only which kernels produce them:
 *   "fused_decode_ctas"  0 automatic (the fused small-grid decoder serves batches larger than the SM count), 1 | 2 | 4 always,
 *                        with that many CTAs per image (a thread-block cluster when > 1), -1 never;
 *   "fused_encode"       0 (default) cgic_encode = two launches, 1 = one CTA per image on grids of at most 4096 cells;
 *   "pack_image"         0 (default) the packer uses one CTA per image on small token grids for batches larger than the SM
 *                        count, one CTA per stream otherwise; 1 = always one per image (where eligible), -1 = never. */
CGIC_API int cgic_tune(const char *key, int value);
CGIC_API const char *cgic_last_error(void);

/* Per-kernel device timing for bench.py's roofline: while enabled, every kernel the library
 * launches is bracketed by CUDA events on the launching stream.  cgic_prof_report synchronises
 * the device and writes one "kernel_name launches total_ms" line per kernel; returns the text
 * length.  (No reference counterpart: the reference has no timing code) */
CGIC_API int cgic_prof_enable(int on);
CGIC_API int cgic_prof_report(char *buf, int cap);

/* ------------------------------------------------------------------------------------------
 * a8  HuffmanCoding.__init__ / make_heap / merge_nodes / make_codes
 *     CGIC/tools/indices_coding.py:10-17, 46-75.  Host-side, init time, heapq-exact.
 *     freq[s]  = int(counter) of symbol s;   order[i] = symbol pushed i-th, i.e. the iteration
 *     order of the reference's `frequency` mapping (NULL = 0..K-1).  The model's
 *     nn.ParameterDict iterates its keys in lexicographic string order (quantize.py:28).
 *     The table is immutable after build and may be shared between threads.
 * ------------------------------------------------------------------------------------------ */
CGIC_API int cgic_huff_build(const int64_t *freq, const int32_t *order, int K, cgic_table **out);
CGIC_API void cgic_huff_free(cgic_table *t);
CGIC_API int cgic_huff_num_symbols(const cgic_table *t);
CGIC_API int cgic_huff_max_len(const cgic_table *t);
CGIC_API int cgic_huff_code_len(const cgic_table *t, int sym);
/* writes the code of `sym` as a NUL-terminated '0'/'1' string (HuffmanCoding.codes[sym]);
 * returns its length or CGIC_ESPACE. */
CGIC_API int cgic_huff_code(const cgic_table *t, int sym, char *buf, int cap);
/* copies the table to the CURRENT device (once; synchronous; init time). */
CGIC_API int cgic_huff_upload(cgic_table *t);

/* ------------------------------------------------------------------------------------------
 * a1  VectorQuantize2.forward      CGIC/modules/vqvae/quantize.py:69-98   (e_dim == 4)
 *     z [B,4,h,w] fp32 NCHW, codebook [K,4] fp32 (16-byte aligned)
 *     idx_out  int64 [B*h*w]   nearest code, reference rounding sequence, first index on ties
 *     zq_out   fp32 [B,4,h,w]  fl(z + fl(e - z))            (nullable)
 *     sqerr_out double[1]      sum over all elements of (e - z)^2; loss = (1+beta)*sqerr/numel
 *                              (nullable)
 *     workspace: cgic_vq_workspace_bytes() bytes whose first 64 bytes are ZERO before the first call
 *     (the kernel leaves them zero, except bytes 8..11: a running int32 count of the latents the indexed
 *     search had to hand to its exhaustive path); one workspace per concurrently running call.
 *     Tokens whose latent is bit-identical to the top-left token of their 4x4 / 2x2 block (the
 *     structure the mask-mix of vqvae_blocks.py:364-366 creates) share that token's search.
 * ------------------------------------------------------------------------------------------ */
CGIC_API size_t cgic_vq_workspace_bytes(int64_t n_tokens);
CGIC_API int cgic_vq_assign(const float *z, int B, int h, int w, const float *codebook, int K, int64_t *idx_out,
                   float *zq_out, double *sqerr_out, void *workspace, size_t workspace_bytes,
                   cgic_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * a1/a2  prepared codebook + indexed search.   State of VectorQuantize2 (`embedding.weight`,
 *     quantize.py:25-26) turned into a search index: a 4-D grid over the codebook's bounding box
 *     whose cells list every code that can be the reference's fp32 argmin for some latent in the
 *     cell (all codes not dominated by more than the rounding slack of quantize.py:73-75, see
 *     csrc/codebook.cu).  cgic_vq_assign_indexed evaluates the reference's rounding sequence on the
 *     cell's list only; latents outside the grid or in an overflowing cell are searched
 *     exhaustively in the same kernel, so results are bit-identical to cgic_vq_assign for ANY
 *     input.  create allocates (init time); update (re)builds the index from a DEVICE codebook
 *     [K,4] on `stream` (call it again whenever the weights change; no allocation, no sync);
 *     stats_host synchronises and returns {valid, cells, longest list, overflowing cells}.
 *     Arguments of cgic_vq_assign_indexed are those of cgic_vq_assign.
 * ------------------------------------------------------------------------------------------ */
CGIC_API int cgic_codebook_create(int K, cgic_codebook **out);
CGIC_API void cgic_codebook_free(cgic_codebook *cb);
CGIC_API int cgic_codebook_update(cgic_codebook *cb, const float *codebook, cgic_stream_t stream);
CGIC_API int cgic_codebook_stats_host(const cgic_codebook *cb, int32_t out[4]);
/* Staleness guard for weights written behind torch's back (`weight.data.copy_`, as the reference's LitEma.copy_to /
 * restore do, CGIC/models/ema.py:51,76): check enqueues one small kernel that compares the live DEVICE codebook with the
 * copy the index was built from; on a mismatch it refreshes the copy, switches cgic_vq_assign_indexed to its exhaustive
 * path (results stay those of the live weights) and raises a host-visible flag.  is_stale polls that flag without
 * synchronising (1 = a check since the last update found a difference: call cgic_codebook_update), 0 otherwise. */
CGIC_API int cgic_codebook_check(cgic_codebook *cb, const float *codebook, cgic_stream_t stream);
CGIC_API int cgic_codebook_is_stale(const cgic_codebook *cb);
CGIC_API int cgic_vq_assign_indexed(const float *z, int B, int h, int w, const cgic_codebook *cb, int64_t *idx_out,
                           float *zq_out, double *sqerr_out, void *workspace, size_t workspace_bytes,
                           cgic_stream_t stream);

/* a3  training-mode counter update, quantize.py:79-81: counters[idx[i]] += 1 (fp32 counters). */
CGIC_API int cgic_vq_count(const int64_t *idx, int64_t n, float *counters, int K, cgic_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * a4  Entropy.forward              CGIC/models/model.py:440-483
 *     x [B,3,H,W] fp32; bins32_host = the 32 values of torch.linspace(-1,1,32) (HOST pointer);
 *     e8_out [B,H/8,W/8], e16_out [B,H/16,W/16] fp32 (either nullable).  H, W multiples of 16.
 * ------------------------------------------------------------------------------------------ */
CGIC_API int cgic_entropy_maps(const float *x, int B, int H, int W, const float *bins32_host, float *e8_out,
                      float *e16_out, cgic_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * f1  The encode tail in two launches that read the image once:
 *   cgic_entropy_route = Entropy(8) + Entropy(16) (model.py:433-483) AND TripleGrainFixedEntropyRouter.forward
 *     (RouterTriple.py:15-96) with per-image thresholds (the reference's B == 1 call for every image): the last CTA of an
 *     image to finish its entropy maps selects the thresholds and writes m_c int32 [B,1,H/16,W/16], m_m int32 [B,1,H/8,W/8].
 *     mode / k_c / k_m as for cgic_router (host doubles, banker's rounding; k for ONE image).  near_out int32 [B,2]
 *     (nullable): how many coarse / medium entropies lie within rtol*|thr| + atol of their threshold -- the cells a
 *     float-tolerance difference between this Entropy and the reference's could flip (0 = masks provably identical).
 *     workspace: cgic_entropy_route_workspace_bytes(B) bytes, ZERO before the first use (left zero).
 *   cgic_route_mix = fine mask (RouterTriple.py:34) + gate (nullable, fp32 [B,1,h,3w]) + mask-mix (vqvae_blocks.py:361-366)
 *     of the three encoder heads: m_f_out int32 [B,1,h,w], out fp32 [B,C,h,w].  The CNN encoder runs between the two.
 * ------------------------------------------------------------------------------------------ */
CGIC_API size_t cgic_entropy_route_workspace_bytes(int B);
CGIC_API int cgic_entropy_route(const float *x, int B, int H, int W, const float *bins32_host, float *e8_out, float *e16_out,
                       int mode, int64_t k_c, int64_t k_m, float rtol, float atol, int32_t *m_c, int32_t *m_m,
                       int32_t *near_out, void *workspace, size_t workspace_bytes, cgic_stream_t stream);
CGIC_API int cgic_route_mix(const float *h_c, const float *h_m, const float *h_f, const int32_t *m_c, const int32_t *m_m, int mode,
                   int B, int C, int h, int w, int32_t *m_f_out, float *gate_out, float *out, cgic_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * a5  TripleGrainFixedEntropyRouter.forward   CGIC/modules/vqvae/RouterTriple.py:15-96
 *     e16 [B,h16,w16], e8 [B,2*h16,2*w16] fp32.  mode 0..6 and the ranks k_c, k_m are computed
 *     by the caller in Python doubles with round() exactly as RouterTriple.py:19-30,36-90 does.
 *     per_image = 0: thresholds over the whole batch (what the reference does for B > 1);
 *     per_image = 1: B independent B == 1 calls (k_c, k_m then refer to ONE image).
 *     m_c int32 [B,1,h16,w16], m_m int32 [B,1,2h16,2w16], m_f int32 [B,1,4h16,4w16];
 *     gate_out fp32 [B,1,4h16,3*4w16] = cat(up4(c), up2(m), f) on the last dim (nullable).
 * ------------------------------------------------------------------------------------------ */
CGIC_API size_t cgic_router_workspace_bytes(int B, int h16, int w16);
CGIC_API int cgic_router(const float *e16, const float *e8, int B, int h16, int w16, int mode, int64_t k_c,
                int64_t k_m, int per_image, int32_t *m_c, int32_t *m_m, int32_t *m_f, float *gate_out,
                void *workspace, size_t workspace_bytes, cgic_stream_t stream);

/* a6  mask-mix, CGIC/modules/vqvae/vqvae_blocks.py:361-366:
 *     out = up4(h_c)*up4(m_c) + up2(h_m)*up2(m_m) + h_f*m_f       [B,C,h,w] fp32 */
CGIC_API int cgic_mask_mix(const float *h_c, const float *h_m, const float *h_f, const int32_t *m_c,
                  const int32_t *m_m, const int32_t *m_f, int B, int C, int h, int w, float *out,
                  cgic_stream_t stream);

/* f4  decoder entry, CGIC/modules/vqvae/decoder.py:373-382: the mask-gated merge of the decoder's branches
 *     level 2: out = h*up2(m_c) + other*m_m                     h, other, out [B,C,hh,ww]; m_c [B,1,hh/2,ww/2], m_m [B,1,hh,ww]
 *     level 3: out = h*up4(m_c) + h*up2(m_m) + other*m_f        m_c [B,1,hh/4,ww/4], m_m [B,1,hh/2,ww/2], m_f [B,1,hh,ww]
 *     (products and sums rounded in torch's order: bit-identical to the eager expression)
 *     mask_elem: 4 = int32 masks (router output), 8 = int64 (decoded masks), -4 = float32. */
CGIC_API int cgic_decoder_merge(const float *h, const float *other, const void *m_c, const void *m_m, const void *m_f,
                       int mask_elem, int level, int B, int C, int hh, int ww, float *out, cgic_stream_t stream);

/* f4  SpatialNorm, CGIC/modules/vqvae/decoder.py:34-53 (the conditioning of every decoder block on the decoded latents):
 *     new_f = GroupNorm(f; groups, eps, gn_weight, gn_bias) * conv_y(zq_up) + conv_b(zq_up),  zq_up = nearest(zq -> H x W)
 *     f, out [B,C,H,W] fp32; zq [B,Cz,hz,wz] fp32 (Cz <= 8; any hz, wz: torch's nearest index rule); conv_y / conv_b are the
 *     1x1 convolutions' weights wy, wb [C,Cz] and biases by, bb [C] (nullable = no bias); gn_weight / gn_bias [C] nullable
 *     (affine=False).  Neither zq_up nor the two convolution outputs are materialised.  Floating point: group statistics
 *     are accumulated in a different order than torch's -- tolerance, not bit parity (tests: rtol 1e-5, atol 1e-5).
 *     workspace: cgic_spatial_norm_workspace_bytes(B, groups) bytes, 16-byte aligned, any content. */
CGIC_API size_t cgic_spatial_norm_workspace_bytes(int B, int groups);
CGIC_API int cgic_spatial_norm(const float *f, const float *zq, const float *gn_weight, const float *gn_bias, const float *wy,
                      const float *by, const float *wb, const float *bb, int B, int C, int H, int W, int Cz, int hz, int wz,
                      int groups, float eps, float *out, void *workspace, size_t workspace_bytes, cgic_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * a7 + a9 + a11 + a12  index selection + 5-stream pack     CGIC/models/model.py:217-260,
 *     HuffmanCoding.compress indices_coding.py:113-126, BinaryCoding.compress mask_coding.py:40-55.
 *     Per image: slot s of image b starts at bytes_out + b*image_stride + slot_off[s] and holds
 *     sizes_out[b*5+s] bytes (0 = stream absent in this mode, or empty: the reference's 0-byte
 *     file).  bpp = 8 * sum(sizes) / (H*W).  sizes_out < 0 flags an out-of-range symbol.
 *     cgic_pack_layout fills the worst-case slot offsets / capacities for a table and grid.
 * ------------------------------------------------------------------------------------------ */
CGIC_API int cgic_pack_layout(const cgic_table *t, int h, int w, int64_t slot_off[5], int64_t slot_cap[5],
                     int64_t *image_stride);
CGIC_API int cgic_pack(const int64_t *idx, const int32_t *m_c, const int32_t *m_m, const int32_t *m_f, int B, int h,
              int w, int mode, const cgic_table *t, uint8_t *bytes_out, int32_t *sizes_out,
              cgic_stream_t stream);
/* The same with a workspace (cgic_pack_workspace_bytes() bytes, ZERO-filled before the first use; the kernels leave
 * it zeroed): on token grids with more than one 4096-position tile per stream the tiles of a stream are packed by
 * several CTAs, chained through hand-over records in the workspace.  Identical bytes. */
CGIC_API size_t cgic_pack_workspace_bytes(int B, int h, int w);
CGIC_API int cgic_pack_ws(const int64_t *idx, const int32_t *m_c, const int32_t *m_m, const int32_t *m_f, int B, int h,
                 int w, int mode, const cgic_table *t, uint8_t *bytes_out, int32_t *sizes_out, void *workspace,
                 size_t workspace_bytes, cgic_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * a1 + a7 + a9 + a11 + a12  the encoder half in one call: VectorQuantize2.forward (quantize.py:69-98) on a prepared
 *     codebook, then selection + the five-stream pack (model.py:217-260) -- i.e. cgic_vq_assign_indexed followed by
 *     cgic_pack_ws, with identical outputs (idx_out, zq_out (nullable), sqerr_out (nullable), bytes_out, sizes_out).
 *     On token grids of at most 4096 cells with codes of at most 32 bits both steps run in ONE launch, one CTA per
 *     image (the indices reach the packer through shared memory); larger grids take the two launches.
 *     workspace: cgic_encode_workspace_bytes() bytes, ZERO-filled before the first use (bytes 8..11 hold a running
 *     count of latents that took the exhaustive search, everything else is left zeroed).
 * ------------------------------------------------------------------------------------------ */
CGIC_API size_t cgic_encode_workspace_bytes(int B, int h, int w);
CGIC_API int cgic_encode(const float *z, const int32_t *m_c, const int32_t *m_m, const int32_t *m_f, int B, int h, int w,
                int mode, const cgic_codebook *cb, const cgic_table *t, int64_t *idx_out, float *zq_out,
                double *sqerr_out, uint8_t *bytes_out, int32_t *sizes_out, void *workspace, size_t workspace_bytes,
                cgic_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * a10 + a11 + a13 + a14  unpack + mask / index re-assembly + codebook gather
 *     CGIC/models/model.py:269-392, decompress_string indices_coding.py:153-168, mask_coding.py:81-96.
 *     bytes / sizes use the cgic_pack layout.  Outputs (all device, caller-allocated):
 *     mc_out int64 [B,h/4,w/4], mm_out int64 [B,h/2,w/2], mf_out int64 [B,h,w]   (grain masks)
 *     ind_out int64 [B,h,w], quant_out fp32 [B,4,h,w] = codebook[ind] (exact rows, NCHW)
 *     status_out int32 [B]: 0 ok, else CGIC_EFORMAT (symbol count != mask population, bad
 *     framing, ...) -- the cases in which the reference raises.
 *     workspace: cgic_unpack_workspace_bytes() bytes, ZERO-filled before the first use (the kernels
 *     leave the chunk hand-over tables in it zeroed); one workspace per concurrently running call.
 *     A workspace belongs to ONE geometry (B, h, w): its carve-up depends on them, so zero-fill it again
 *     before using it with another geometry (the same holds for cgic_pack / cgic_encode workspaces).
 * ------------------------------------------------------------------------------------------ */
CGIC_API size_t cgic_unpack_workspace_bytes(int B, int h, int w);
CGIC_API int cgic_unpack(const uint8_t *bytes, const int32_t *sizes, int B, int h, int w, int mode,
                const cgic_table *t, const float *codebook, int64_t *mc_out, int64_t *mm_out,
                int64_t *mf_out, int64_t *ind_out, float *quant_out, int32_t *status_out, void *workspace,
                size_t workspace_bytes, cgic_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Single-stream codec entry points (what HuffmanCoding / BinaryCoding objects bind to).
 *     a9  HuffmanCoding.compress           indices_coding.py:113-126
 *     a10 HuffmanCoding.decompress_string  indices_coding.py:153-168
 *     a11 BinaryCoding.compress / decompress_string   mask_coding.py:40-55, 81-96
 *     size_out / count_out are device int32[1].  Decoders: nbytes == 0 -> count -1 (the
 *     reference returns None).  cap is in bytes (encode) or symbols (decode).
 * ------------------------------------------------------------------------------------------ */
CGIC_API int64_t cgic_huff_stream_capacity(const cgic_table *t, int64_t n_symbols);
CGIC_API int cgic_huff_encode(const int64_t *symbols, int64_t n, const cgic_table *t, uint8_t *out, int64_t cap,
                     int32_t *size_out, cgic_stream_t stream);
CGIC_API int cgic_huff_decode(const uint8_t *bytes, int64_t nbytes, const cgic_table *t, int32_t *symbols_out,
                     int64_t cap, int32_t *count_out, cgic_stream_t stream);
CGIC_API int cgic_bits_encode(const int32_t *values, int64_t n, uint8_t *out, int64_t cap, int32_t *size_out,
                     cgic_stream_t stream);
CGIC_API int cgic_bits_decode(const uint8_t *bytes, int64_t nbytes, int32_t *values_out, int64_t cap,
                     int32_t *count_out, cgic_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Host-buffer session: the call a reference-side plugin makes per batch when its tensors live
 * in host memory (pinned for full PCIe speed).  Owns its device buffers, workspace and stream
 * (allocated at create time, never on the hot path).  compress = a1 + a7 + a12 for B
 * independent images; decompress = a13 + a14.  Both copy host->device, run, copy device->host
 * and synchronise before returning; the batch is processed in `parts` image ranges on separate
 * streams so that copies in both directions overlap the kernels.
 * ------------------------------------------------------------------------------------------ */
CGIC_API int cgic_session_create(int B, int h, int w, int mode, const cgic_table *t, const float *codebook_host, int K,
                        cgic_session **out);
CGIC_API void cgic_session_destroy(cgic_session *s);
CGIC_API int64_t cgic_session_image_stride(const cgic_session *s);
/* number of contiguous image ranges the batch is pipelined in (own stream each; default min(B, 4)) */
CGIC_API int cgic_session_set_pipeline(cgic_session *s, int parts);
CGIC_API int cgic_session_compress_host(cgic_session *s, const float *z, const int32_t *m_c, const int32_t *m_m,
                               const int32_t *m_f, uint8_t *bytes_out, int32_t *sizes_out, int64_t *idx_out,
                               float *zq_out, double *sqerr_out);
CGIC_API int cgic_session_decompress_host(cgic_session *s, const uint8_t *bytes, const int32_t *sizes, int64_t *mc_out,
                                 int64_t *mm_out, int64_t *mf_out, int64_t *ind_out, float *quant_out,
                                 int32_t *status_out);

/* CGIC.compress (CGIC/models/model.py:206-401) in one call: compress_host followed by the decode
 * of the device-resident streams (no H2D of the streams; the two D2H groups overlap). */
CGIC_API int cgic_session_roundtrip_host(cgic_session *s, const float *z, const int32_t *m_c, const int32_t *m_m,
                                const int32_t *m_f, uint8_t *bytes_out, int32_t *sizes_out, int64_t *idx_out,
                                float *zq_out, double *sqerr_out, int64_t *mc_out, int64_t *mm_out, int64_t *mf_out,
                                int64_t *ind_out, float *quant_out, int32_t *status_out);

/* Pinned-arena round trip: the same call with the host buffers owned by the session.  The batch is cut
 * into `parts` contiguous image ranges; every range has one contiguous input block and one contiguous
 * output block in pinned host memory (and a mirror on the device), so a range moves with ONE copy per
 * direction and the ranges pipeline: H2D of range p+1, the kernels of range p and D2H of range p-1 run
 * at the same time.  After the first (eager) call the whole round trip is replayed as one CUDA graph.
 *   cgic_session_arena        fixes `parts` (<= 8) and (re)allocates the arenas (init time);
 *   cgic_session_arena_tensor host pointer of tensor `what` of range `part` (+ its image range);
 *                             inputs are written there by the caller before the call, outputs read after;
 *   cgic_session_roundtrip_arena  flags: bit 0 also return idx (VQ indices), bit 1 also z_q; bit 2 (alone): the
 *                             decoded tensors (IND, QUANT, DMC, DMM, DMF) stay in HBM for the decoder CNN, as
 *                             model.py:391-399 hands them over -- only BYTES, SIZES, STATUS, SQERR come back to the host;
 *                             bit 3 (excludes bits 0, 1; K <= 32768): NARROW WIRE -- the masks are read from MC8 / MM8 / MF8
 *                             (one byte per cell) instead of MC / MM / MF and widened on the device; the decoded tensors
 *                             come back as IND16 (int16) and DMC8 / DMM8 / DMF8 (uint8) next to QUANT, BYTES, SIZES,
 *                             STATUS, SQERR -- the reference's int64 tensors (IND, DMC, DMM, DMF) stay on the device
 *                             (cgic_session_arena_gather_device).  9.9 MB instead of 15.4 MB per 64 images of 256 x 256;
 *   cgic_session_arena_gather_device  copies output tensor `what` of all ranges, in image order, into one contiguous
 *                             DEVICE buffer of the caller (device to device, then synchronises).
 * Tensor shapes per range of nb images: Z/QUANT/ZQ fp32 [nb,4,h,w]; MC/MM/MF int32 [nb,1,.,.]; BYTES
 * uint8 [nb,image_stride]; SIZES int32 [nb,5]; STATUS int32 [nb]; SQERR double[1]; IND/IDX int64 [nb,h,w];
 * DMC/DMM/DMF int64 [nb,.,.]; MC8/MM8/MF8/DMC8/DMM8/DMF8 uint8, IND16 int16, same shapes as their wide twins. */
enum {
    CGIC_ARENA_Z = 0, CGIC_ARENA_MC, CGIC_ARENA_MM, CGIC_ARENA_MF,          /* inputs */
    CGIC_ARENA_BYTES, CGIC_ARENA_SIZES, CGIC_ARENA_STATUS, CGIC_ARENA_SQERR, /* outputs */
    CGIC_ARENA_IND, CGIC_ARENA_QUANT, CGIC_ARENA_DMC, CGIC_ARENA_DMM, CGIC_ARENA_DMF,
    CGIC_ARENA_IDX, CGIC_ARENA_ZQ,
    CGIC_ARENA_MC8, CGIC_ARENA_MM8, CGIC_ARENA_MF8,                          /* narrow wire: masks in, one byte per cell */
    CGIC_ARENA_IND16, CGIC_ARENA_DMC8, CGIC_ARENA_DMM8, CGIC_ARENA_DMF8,     /* narrow wire: int16 indices, u8 masks out */
    CGIC_ARENA_COUNT
};
CGIC_API int cgic_session_arena(cgic_session *s, int parts);
CGIC_API int cgic_session_arena_tensor(const cgic_session *s, int what, int part, void **host_ptr, int *first_image,
                              int *n_images);
CGIC_API int cgic_session_roundtrip_arena(cgic_session *s, int flags, double *sqerr_out);
CGIC_API int cgic_session_arena_gather_device(cgic_session *s, int what, void *dst_device);
/* Two round trips in flight.  The session owns two independent arena sets ("slots" 0 and 1; slot 1 is allocated on first
 * use; the calls above work on slot 0).  _submit enqueues the round trip of a slot and returns at once; _wait blocks
 * until its results are in the slot's host arena.  Alternating the slots overlaps the D2H copies of one batch with the
 * H2D copies and kernels of the next:
 *     fill(slot 0); submit(0);  loop { fill(slot 1); submit(1); wait(0); use(slot 0);  fill(slot 0); submit(0); wait(1); use(slot 1); }
 * A slot's tensors must not be touched between its _submit and its _wait; one thread drives a session. */
CGIC_API int cgic_session_arena_slot_tensor(cgic_session *s, int slot, int what, int part, void **host_ptr, int *first_image,
                                   int *n_images);
CGIC_API int cgic_session_roundtrip_arena_submit(cgic_session *s, int slot, int flags);
CGIC_API int cgic_session_roundtrip_arena_wait(cgic_session *s, int slot, double *sqerr_out);
CGIC_API int cgic_session_arena_slot_gather_device(cgic_session *s, int slot, int what, void *dst_device);

#ifdef __cplusplus
}
#endif
#endif /* CGIC_B200_H */
