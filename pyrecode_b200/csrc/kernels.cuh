// kernels.cuh -- launcher declarations shared by the translation units of librecode_b200.
#pragma once
#include "common.cuh"

// reduce.cu
int launch_reduce_tiles(rc_ctx *ctx, const Geom &g, int itemsize, int valmode, const void *frames, const void *thr,
                        int F, uint32_t *maps, uint32_t *tilecnt, uint16_t *wordpre, void *vals, cudaStream_t st);
int launch_map_counts(rc_ctx *ctx, const Geom &g, const uint32_t *maps, int F, uint32_t *tilecnt, uint16_t *wordpre,
                      cudaStream_t st);
int launch_scan_tiles(rc_ctx *ctx, const Geom &g, const uint32_t *tilecnt, int F, uint32_t *tilepre,
                      uint32_t *counts, uint32_t *packed_bytes, int b, cudaStream_t st);
int launch_bitpack(rc_ctx *ctx, const Geom &g, int val_itemsize, const void *vals, const uint32_t *tilepre, int F,
                   int b, uint8_t *packed, size_t packed_stride, cudaStream_t st);
int launch_bitpack_flat(rc_ctx *ctx, int b, const uint16_t *vals, uint64_t n, uint8_t *packed, cudaStream_t st);
int launch_bitunpack_flat(rc_ctx *ctx, int b, const uint8_t *packed, uint64_t n, uint64_t *out, cudaStream_t st);
int launch_make_threshold(rc_ctx *ctx, int itemsize, const void *dark, uint64_t eps, void *thr, size_t n,
                          cudaStream_t st);

// ccl.cu
int launch_ccl_init(rc_ctx *ctx, const Geom &g, const uint32_t *tilecnt, uint32_t *parent, int F, cudaStream_t st);
int launch_ccl_union(rc_ctx *ctx, const Geom &g, const uint32_t *maps, const uint16_t *wordpre, uint32_t *parent,
                     int F, cudaStream_t st);
int launch_ccl_tiles(rc_ctx *ctx, const Geom &g, int fold, const uint32_t *maps, const uint16_t *wordpre,
                     const uint32_t *tilecnt, const uint32_t *vp, uint8_t *tileovf, uint32_t *xcount, void *xlinks,
                     uint32_t *parent, uint32_t *acc, int l4mode, uint32_t *bbox, uint32_t *map2, uint64_t *cent,
                     uint32_t *rootcnt, int F, cudaStream_t st);
size_t ccl_xlinks_bytes(const Geom &g, size_t F);
int launch_ccl_border(rc_ctx *ctx, const Geom &g, int fold, const uint32_t *maps, const uint16_t *wordpre,
                      const uint8_t *tileovf, const uint32_t *xcount, const void *xlinks, uint32_t *parent,
                      uint32_t *acc, int F, cudaStream_t st);
int launch_l4_open(rc_ctx *ctx, const Geom &g, int mode, const uint32_t *maps, const uint16_t *wordpre,
                   const uint32_t *tilecnt, const uint8_t *tileovf, const uint32_t *xcount, const void *xlinks,
                   const uint32_t *parent, uint32_t *claim, const uint32_t *bbox, const uint32_t *vp, uint32_t *map2,
                   uint64_t *cent, uint32_t *rootcnt, int F, cudaStream_t st);
int launch_ccl_flatten(rc_ctx *ctx, const Geom &g, const uint32_t *maps, const uint16_t *wordpre, uint32_t *parent,
                       int F, cudaStream_t st);
int launch_ccl_roots(rc_ctx *ctx, const Geom &g, int payload, const uint32_t *tilecnt, const uint32_t *parent,
                     const uint32_t *acc, const uint64_t *cent, uint32_t *rootcnt, uint32_t *ord, uint16_t *out16,
                     uint64_t *out64, int F, cudaStream_t st);
int launch_ccl_label_image(rc_ctx *ctx, const Geom &g, const uint32_t *maps, const uint16_t *wordpre,
                           const uint32_t *parent, const uint32_t *ord, const uint32_t *rootpre, int32_t *labels,
                           int F, cudaStream_t st);
int launch_gather_centroids(rc_ctx *ctx, const Geom &g, const uint64_t *cent_tiles, const uint32_t *rootpre, int F,
                            float *out, size_t capacity, cudaStream_t st);

// deflate.cu
struct DeflateWs {
    uint32_t *chunk_base;     // [S+1]
    uint32_t *counters;       // [8]
    uint32_t *chunk_bytes;    // [max_chunks]
    uint32_t *chunk_rel;      // [max_chunks]
    uint2 *chunk_adler;       // [max_chunks]
    uint32_t *stream_bytes;   // [S]
    uint32_t *stream_adler;   // [S]
    uint64_t *stream_dst;     // [S]
    uint32_t *ghist;          // [S][288] token histogram per stream
    void *tables;             // [S] DeflateTable
    uint8_t *scratch;         // [max_chunks * slot]
    size_t max_chunks;
};
size_t deflate_max_chunks(int n_streams, size_t max_in_bytes);
DeflateWs carve_deflate_ws(Carver &c, int n_streams, size_t max_chunks, bool need_scratch);
int launch_deflate_streams(rc_ctx *ctx, int level, int wrap, int shared_table, void *kept_table, int build_table,
                           const uint8_t *in, const uint64_t *in_off, const uint32_t *in_bytes, int n_streams,
                           const DeflateWs &w, cudaStream_t st);
size_t deflate_table_bytes();
int launch_layout_strided(rc_ctx *ctx, const DeflateWs &w, int n_streams, size_t stride, uint32_t *out_bytes,
                          cudaStream_t st);
int launch_layout_records(rc_ctx *ctx, const DeflateWs &wm, const DeflateWs &wv, const uint32_t *packed_bytes,
                          int n_frames, int spf, int mode, uint32_t first_frame_id, uint8_t *records, size_t capacity,
                          uint64_t *record_off, uint32_t *status, cudaStream_t st);
int launch_copy_pieces(rc_ctx *ctx, const DeflateWs &w, int wrap, const uint8_t *raw_in, const uint64_t *in_off,
                       const uint32_t *in_bytes, int n_streams, uint8_t *out, size_t capacity, uint32_t *status,
                       cudaStream_t st);

// inflate.cu
size_t inflate_workspace_bytes(int n_streams, size_t out_stride);
int launch_inflate(rc_ctx *ctx, const uint8_t *in, const uint64_t *in_off, const uint32_t *in_bytes, int n_streams,
                   void *ws, uint8_t *out, size_t out_stride, uint32_t *out_bytes, uint32_t *status, cudaStream_t st);

// unpack.cu
int launch_unpack_sparse(rc_ctx *ctx, const Geom &g, int level, int b, const uint32_t *maps, const uint8_t *packed,
                         size_t packed_stride, const uint16_t *wordpre, const uint32_t *tilepre, int F,
                         uint64_t *triples, size_t capacity, cudaStream_t st);
int launch_unpack_dense(rc_ctx *ctx, const Geom &g, int itemsize, int level, int b, const uint32_t *maps,
                        const uint8_t *packed, size_t packed_stride, const uint16_t *wordpre, const uint32_t *tilepre,
                        int F, void *dense, uint32_t *sum, cudaStream_t st);
int launch_recalibrate(rc_ctx *ctx, int itemsize, const void *frames, const double *diff, size_t P, int F, void *out,
                       cudaStream_t st);
size_t median_std_workspace_bytes(size_t P);
int launch_median_std(rc_ctx *ctx, int itemsize, const void *stack, size_t P, int N, float *med, float *sd, void *ws,
                      cudaStream_t st);
int launch_cal_topk(rc_ctx *ctx, int itemsize, const void *stack, size_t P, int N, const float *thr, int k, int as_run,
                    float *out, cudaStream_t st);
