/*
 * recode_b200.h -- C ABI of librecode_b200.so, the sm_100a replacement for pyReCoDe's
 * native extension `c_recode` (reference: pyrecode/pyrecode.cpp + pyrecode/c_extensions/reader.h)
 * and for the numba / numpy / scipy / zlib steps of the per-frame hot path
 * (reference: pyrecode/recode_writer.py:430-557, pyrecode/recode_reader.py:379-462).
 *
 * Conventions
 *   - plain pointers and sizes only; every data pointer is a DEVICE pointer owned by the caller
 *     (PyTorch on the Python side) unless the name ends in `_host`;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*);
 *   - return 0 on success, < 0 on error; rc_last_error(ctx) gives the text.  Never throws, never exits;
 *   - no global mutable state: one rc_ctx per (GPU, host thread);
 *   - data-dependent failures (an output buffer too small, a corrupt deflate stream) cannot be known
 *     at launch time: they set bits in the caller-supplied device status word(s) documented per call.
 *
 * Frame geometry: ny rows x nx cols, row-major, P = ny*nx pixels.  Pixels are uint16 (itemsize 2,
 * source_bit_depth 9..16) or uint8 (itemsize 1, source_bit_depth 1..8) -- misc.py:41-71 (map_dtype).
 * Binary map: ceil(P/8) bytes, pixel i at byte i>>3 bit i&7 (recode_writer.py:622-634).  On the device a
 * map is held as uint32 words with a stride of rc_map_stride_words(P) words per frame.
 */
#ifndef RECODE_B200_H
#define RECODE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rc_ctx rc_ctx;

/* run configuration == the header / params fields the hot path depends on
 * (pyrecode/params.py:204-211, pyrecode/recode_header.py:128-163) */
typedef struct rc_config {
    int32_t ny, nx;              /* num_rows, num_cols                                   */
    int32_t itemsize;            /* 1 or 2: numpy itemsize of the source dtype            */
    int32_t bit_depth;           /* source_bit_depth == target_bit_depth (SURVEY B-10)    */
    int32_t reduction_level;     /* 1..4                                                  */
    int32_t rc_operation_mode;   /* 0 reduce only, 1 reduce + deflate                     */
    int32_t l2_statistics;       /* 0/1 max, 2 sum   (recode_writer.py:358-365)           */
    int32_t l4_centroiding;      /* 0/1 weighted, 2 max pixel, 3 unweighted (:367-378)    */
    int32_t compression_level;   /* 0 = stored blocks, 1..9 = RLE + dynamic Huffman       */
    int32_t max_frames;          /* frames per call the workspace is sized for            */
} rc_config;

/* status bits written to device status words */
#define RC_STATUS_OK               0u
#define RC_STATUS_RECORDS_OVERFLOW 1u   /* records buffer too small                       */
#define RC_STATUS_BAD_STREAM       2u   /* inflate: corrupt / unsupported deflate data    */
#define RC_STATUS_OUT_OVERFLOW     4u   /* inflate: output larger than the given capacity */
#define RC_STATUS_SIZE_MISMATCH    8u   /* inflate/unpack: stream shorter than required   */

/* ---- lifecycle ------------------------------------------------------------------------- */
int          rc_create(rc_ctx **ctx, int device);      /* replaces c_recode.Reader() (pyrecode.cpp:41-55) */
void         rc_destroy(rc_ctx *ctx);
const char  *rc_last_error(const rc_ctx *ctx);
int          rc_version(void);
int          rc_sm_count(const rc_ctx *ctx);

/* instrumentation (replaces the datetime.now() stage timers of recode_writer.py:433-555): CUDA-event stage
 * timing of the last rc_reduce_compress call and a count of kernels launched through the context */
int                 rc_profile_enable(rc_ctx *ctx, int on);
int                 rc_profile_read(rc_ctx *ctx, float *ms, int capacity);   /* returns number of stages (4) */
/* rc_profile_enable(ctx, 2) additionally records one event after every kernel (group) of the reduction's second stage
 * (labelling, cross-tile links, root compaction, scan, bit packing ...); this returns the elapsed milliseconds between
 * them for the most recent call (a profiling aid of bench.py; no reference counterpart). */
int                 rc_profile_read_detail(rc_ctx *ctx, float *ms, int capacity);
unsigned long long  rc_launch_count(const rc_ctx *ctx);
/* Tell the context that several contexts keep batches in flight on this GPU (one stream each).  rc_reduce_compress
 * then runs everything after the streaming kernel on context-owned high-priority streams and labels puddles with a
 * few persistent CTAs per SM, so that the batches share the SMs instead of queueing behind each other.  Off
 * (the default) every kernel gets the whole GPU, which is what a single stream wants. */
int                 rc_set_pipelined(rc_ctx *ctx, int on);

/* ---- sizes ----------------------------------------------------------------------------- */
size_t rc_map_stride_words(size_t n_pixels);                 /* uint32 words per frame map on device      */
size_t rc_packed_stride_bytes(const rc_config *cfg);         /* bytes per frame of packed values (worst)  */
size_t rc_workspace_bytes(const rc_config *cfg);             /* device workspace for rc_reduce_compress   */
size_t rc_stage_workspace_bytes(const rc_config *cfg);       /* ... for rc_ccl_label / rc_l4_centroids    */
size_t rc_records_capacity(const rc_config *cfg);            /* worst-case record bytes for max_frames    */
size_t rc_read_workspace_bytes(const rc_config *cfg);             /* device workspace for rc_unpack_*          */

/* ---- write side ------------------------------------------------------------------------ */

/* The whole per-frame hot path for a batch of frames: replaces ReCoDeWriter._reduce_compress
 * (recode_writer.py:430-557) x n_frames.  Produces, for frame i, the part-file record
 *   mode 1, L1/L2: [frame_id u32][n_comp_map u32][n_comp_vals u32][n_packed u32][zlib(map)][zlib(vals)]
 *   mode 1, L3/L4: [frame_id u32][n_comp_map u32][zlib(map)]
 *   mode 0, L1/L2: [frame_id u32][n_packed u32][map][packed vals]
 *   mode 0, L3/L4: [frame_id u32][map]
 * at d_records + d_record_offsets[i]; d_record_offsets has n_frames + 1 entries (last = total bytes).
 * d_counts[i] = foreground pixels (L1/L3) or puddles (L2/L4).  d_status: one uint32 (RC_STATUS_*).
 * frame_id of frame i = first_frame_id + i (recode_writer.py:385). */
int rc_reduce_compress(rc_ctx *ctx, const rc_config *cfg,
                       const void *d_frames, int n_frames, const void *d_thr, uint32_t first_frame_id,
                       void *d_workspace, size_t workspace_bytes,
                       uint8_t *d_records, size_t records_capacity,
                       uint64_t *d_record_offsets, uint32_t *d_counts, uint32_t *d_status,
                       void *stream);

/* thr = dark + eps in the source dtype, wrapping (recode_writer.py:126-127,132-137) */
int rc_make_threshold(rc_ctx *ctx, const rc_config *cfg, const void *d_dark, uint64_t eps, void *d_thr,
                      void *stream);

/* Stage entry points (same kernels rc_reduce_compress chains; exposed for parity tests and for callers
 * that want the reduced streams without the container).
 *
 * rc_reduce: threshold + binary map (recode_writer.py:437,456) and, per level, the second stream:
 *   L1 packed (frame - thr) of foreground pixels (:440,:461-477); L2 packed per-puddle max/sum (:443-446);
 *   L3 nothing; L4 the map becomes the centroid map (:448-449).
 * d_maps: n_frames x rc_map_stride_words(P) uint32, 16-byte aligned (the labelling kernels fetch it with bulk async
 * copies).  d_packed: n_frames x rc_packed_stride_bytes(cfg).
 * d_packed_bytes[i] = ceil(count*b/8).  d_counts as above. */
int rc_reduce(rc_ctx *ctx, const rc_config *cfg, const void *d_frames, int n_frames, const void *d_thr,
              void *d_workspace, size_t workspace_bytes,
              uint32_t *d_maps, uint8_t *d_packed, uint32_t *d_packed_bytes, uint32_t *d_counts,
              void *stream);

/* 8-connected labelling of binary maps; labels 1..k in raster order of first pixel, 0 = background:
 * replaces scipy.ndimage.label(binary, structure=3x3) at recode_writer.py:443.
 * d_labels: n_frames x P int32.  d_counts[i] = k. */
int rc_ccl_label(rc_ctx *ctx, const rc_config *cfg, const uint32_t *d_maps, int n_frames,
                 void *d_workspace, size_t workspace_bytes, int32_t *d_labels, uint32_t *d_counts,
                 void *stream);

/* L4 centroid list (row, col) float32 per puddle in label order: replaces get_centroids_2D_nb
 * (pyrecode/utils/converters.py:157-259).  d_centroids: n_frames x centroid_capacity x 2 float32. */
int rc_l4_centroids(rc_ctx *ctx, const rc_config *cfg, const void *d_frames, int n_frames, const void *d_thr,
                    void *d_workspace, size_t workspace_bytes, float *d_centroids, size_t centroid_capacity,
                    uint32_t *d_counts, void *stream);

/* Batched zlib-format deflate: replaces zlib.compress (recode_compressors.py:84-85).
 * Stream s is d_in + in_offsets[s], in_bytes[s] long (both device arrays).  Output stream s is written at
 * d_out + s * out_stride; d_out_bytes[s] = its length.  out_stride must be >= rc_deflate_bound(max in_bytes). */
size_t rc_deflate_bound(size_t in_bytes);
size_t rc_deflate_workspace_bytes(int n_streams, size_t max_in_bytes);
int rc_deflate_zlib(rc_ctx *ctx, int compression_level, const uint8_t *d_in, const uint64_t *d_in_offsets,
                    const uint32_t *d_in_bytes, int n_streams, size_t max_in_bytes,
                    void *d_workspace, size_t workspace_bytes,
                    uint8_t *d_out, size_t out_stride, uint32_t *d_out_bytes, void *stream);

/* ---- read side ------------------------------------------------------------------------- */

/* Batched zlib-format inflate: replaces zlib.decompress (recode_compressors.py:42-43).  Accepts stored,
 * fixed and dynamic blocks (reference-written files are multi-block dynamic streams).
 * Output stream s at d_out + s * out_stride (capacity out_stride); d_out_bytes[s] = inflated length;
 * d_status[s] = RC_STATUS_* per stream.  The input is read in aligned 16-byte units: d_in must be readable from
 * the 16-byte boundary below each stream's first byte to the one above its last byte.  The bytes of d_out past a
 * stream's inflated length are unspecified (zero when d_out and out_stride are multiples of 4). */
size_t rc_inflate_workspace_bytes(int n_streams, size_t out_stride);
int rc_inflate_zlib(rc_ctx *ctx, const uint8_t *d_in, const uint64_t *d_in_offsets, const uint32_t *d_in_bytes,
                    int n_streams, void *d_workspace, size_t workspace_bytes,
                    uint8_t *d_out, size_t out_stride, uint32_t *d_out_bytes,
                    uint32_t *d_status, void *stream);

/* (row, col, value) uint64 triples in raster order: replaces c_recode.Reader.get_frame_sparse
 * (pyrecode.cpp:95-119 -> reader.h:10-68).  Level 1: value = b bits at stream bit rank*b; other levels: 1.
 * d_triples: n_frames x triple_capacity x 3 uint64; d_counts[i] = n foreground pixels of frame i. */
int rc_unpack_sparse(rc_ctx *ctx, const rc_config *cfg, const uint32_t *d_maps, const uint8_t *d_packed,
                     size_t packed_stride, int n_frames, void *d_workspace, size_t workspace_bytes,
                     uint64_t *d_triples, size_t triple_capacity, uint32_t *d_counts, void *stream);

/* dense reconstruction (what coo_matrix(...).todense() gives, recode_reader.py:464-471): d_dense is
 * n_frames x P of the source dtype.  If d_sum != NULL the frames are also accumulated into the uint32
 * live-view image d_sum[P] (examples/ReCoDe_Live_View_MT.ipynb cell 1); d_dense may then be NULL. */
int rc_unpack_dense(rc_ctx *ctx, const rc_config *cfg, const uint32_t *d_maps, const uint8_t *d_packed,
                    size_t packed_stride, int n_frames, void *d_workspace, size_t workspace_bytes,
                    void *d_dense, uint32_t *d_sum, uint32_t *d_counts, void *stream);

/* inverse of the variable-bit-depth packing: replaces c_recode.Reader.bit_unpack_pixel_intensities
 * (pyrecode.cpp:74-93 -> reader.h:74-99, loop bug fixed).  d_out: n_values uint64. */
int rc_bit_unpack(rc_ctx *ctx, int bit_depth, const uint8_t *d_packed, uint64_t n_values, uint64_t *d_out,
                  void *stream);

/* variable-bit-depth packing of a plain value array: replaces c_recode.Reader.bit_pack_pixel_intensities
 * (pyrecode.cpp:121-141 -> reader.h:105-140) and numba _bit_pack (recode_writer.py:637-652).
 * d_vals: n_values uint16.  d_packed: ceil(n_values*b/8) bytes (fully written). */
int rc_bit_pack(rc_ctx *ctx, int bit_depth, const uint16_t *d_vals, uint64_t n_values, uint8_t *d_packed,
                void *stream);

/* offline recalibration of reconstructed L1 frames: replaces the per-frame numpy arithmetic of recalibrate_l1
 * (pyrecode/utils/converters.py:15-57): d_out = (T) clamp(float64(frame) + d_diff, 0, max(T)), d_diff[n_pixels] =
 * original_calibration - (new_calibration + epsilon) as float64.  Buffers 16-byte aligned, n_pixels * itemsize a
 * multiple of 16.  d_out must not overlap d_frames. */
int rc_recalibrate(rc_ctx *ctx, int itemsize, const void *d_frames, const double *d_diff, size_t n_pixels,
                   int n_frames, void *d_out, void *stream);

/* per-pixel median and population standard deviation over a stack of n_frames frames (d_stack: n_frames x n_pixels of
 * the source dtype), as float32: replaces _median_std_nb (pyrecode/utils/calibration.py:48-57), the heavy step of
 * make_calibration_frames (:87-138).  The median is exact (np.median); the standard deviation comes from exact
 * integer sums, rounded once in float64 and once to float32. */
size_t rc_median_std_workspace_bytes(size_t n_pixels);
int rc_median_std(rc_ctx *ctx, int itemsize, const void *d_stack, int n_frames, size_t n_pixels, float *d_median,
                  float *d_std, void *d_workspace, size_t workspace_bytes, void *stream);

/* per-pixel "accurate" thresholds: replaces _get_pixel_thresh_2 (pyrecode/utils/calibration.py:27-45).  For every
 * pixel, of the stack values above d_thr[pixel] the expected_n_events + 1 largest are kept (missing ones count as the
 * float32 minimum, as in the reference); d_out = mean of the two smallest of them, float32.  as_run != 0: what the
 * reference computes when actually run (its removal of a found maximum is an out-of-range float -> unsigned store that
 * changes nothing): the largest value above d_thr[pixel], whatever expected_n_events. */
int rc_pixel_thresholds(rc_ctx *ctx, int itemsize, const void *d_stack, int n_frames, size_t n_pixels,
                        const float *d_thr, int expected_n_events, int as_run, float *d_out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* RECODE_B200_H */
