// api.cu -- extern "C" entry points of librecode_b200.so (see include/recode_b200.h).
#include "common.cuh"
#include "kernels.cuh"
#include "deflate_chunk.cuh"

// ---- helpers -----------------------------------------------------------------------------------
static int check_cfg(rc_ctx *ctx, const rc_config *c)
{
    if (!c) RC_FAIL(ctx, -1, "config is null");
    if (c->ny < 1 || c->nx < 1 || c->ny > 65535 || c->nx > 65535)
        RC_FAIL(ctx, -1, "ny/nx out of range [1, 65535]: %d x %d", c->ny, c->nx);
    if ((size_t)c->ny * (size_t)c->nx > 0x7fffffffu - TILE_PX)
        RC_FAIL(ctx, -1, "frame has too many pixels for 32-bit indexing");
    if (c->itemsize != 1 && c->itemsize != 2) RC_FAIL(ctx, -1, "itemsize must be 1 or 2 (got %d)", c->itemsize);
    if (c->bit_depth < 1 || c->bit_depth > 8 * c->itemsize)
        RC_FAIL(ctx, -1, "bit_depth %d does not fit itemsize %d", c->bit_depth, c->itemsize);
    if (c->reduction_level < 1 || c->reduction_level > 4) RC_FAIL(ctx, -1, "reduction_level must be 1..4");
    if (c->rc_operation_mode != 0 && c->rc_operation_mode != 1) RC_FAIL(ctx, -1, "rc_operation_mode must be 0 or 1");
    if (c->compression_level < 0 || c->compression_level > 9) RC_FAIL(ctx, -1, "compression_level must be 0..9");
    if (c->l2_statistics < 0 || c->l2_statistics > 2) RC_FAIL(ctx, -1, "l2_statistics must be 0..2");
    if (c->l4_centroiding < 0 || c->l4_centroiding > 3) RC_FAIL(ctx, -1, "l4_centroiding must be 0..3");
    if (c->max_frames < 1) RC_FAIL(ctx, -1, "max_frames must be >= 1");
    return 0;
}

// bytes per frame of the packed value stream, worst case (every pixel foreground), 16-byte multiple + slack
static size_t packed_stride_of(const rc_config *c)
{
    const size_t P = (size_t)c->ny * c->nx;
    if (c->reduction_level == 3 || c->reduction_level == 4) return 16;
    // level 2 packs one value per puddle: at most ceil(ny / 2) * ceil(nx / 2) puddles are pairwise non-adjacent under
    // 8-connectivity
    const size_t n = c->reduction_level == 2 ? ((size_t)c->ny + 1) / 2 * (((size_t)c->nx + 1) / 2) : P;
    return round_up((n * (size_t)c->bit_depth + 7) / 8, 16) + 16;
}

struct ReduceWs {
    uint32_t *tilecnt, *tilepre, *rootcnt, *rootpre;
    uint16_t *wordpre;
    uint8_t *tileovf;
    uint32_t *xcount;
    void *xlinks;
    void *vals;
    uint32_t *parent, *acc, *bbox, *ord;
    uint16_t *stats16;
    uint64_t *cent, *cent_tiles;
    uint32_t *map1;
};

// what: 0 = rc_reduce, 1 = rc_ccl_label, 2 = rc_l4_centroids
static ReduceWs carve_reduce(Carver &c, const rc_config *cfg, const Geom &g, int what)
{
    ReduceWs w;
    memset(&w, 0, sizeof(w));
    const size_t F = (size_t)cfg->max_frames;
    const int level = cfg->reduction_level;
    w.tilecnt = c.take<uint32_t>(F * g.NT);
    w.tilepre = c.take<uint32_t>(F * (g.NT + 1));
    w.wordpre = c.take<uint16_t>(F * g.MS);
    w.tileovf = c.take<uint8_t>(F * g.NT);
    const bool ccl = what != 0 || level == 2 || level == 4;
    // L1: foreground values in the source dtype.  L2 / L4: (value << 16) | position words
    if (what == 0 && level == 1) w.vals = c.take<uint8_t>(F * g.slots * cfg->itemsize);
    if ((what == 0 && (level == 2 || level == 4)) || what == 2) w.vals = c.take<uint32_t>(F * g.slots);
    if (ccl) {
        w.xcount = c.take<uint32_t>(F * g.NT);
        w.xlinks = c.take<uint8_t>(ccl_xlinks_bytes(g, F));
        w.parent = c.take<uint32_t>(F * g.slots);
        w.rootcnt = c.take<uint32_t>(F * g.NT);
        w.rootpre = c.take<uint32_t>(F * (g.NT + 1));
    }
    if (what == 0 && level == 2) {
        w.acc = c.take<uint32_t>(F * g.slots);
        // the compacted statistics reuse the (value, position) words: nothing reads those after the tile labelling, and
        // k_ccl_roots only reads parent / acc while it writes here
        w.stats16 = (uint16_t *)w.vals;
    }
    if ((what == 0 && level == 4) || what == 2) w.acc = c.take<uint32_t>(F * g.slots);    // L4: claim flags
    if ((what == 0 && level == 4) || what == 2) {
        w.bbox = c.take<uint32_t>(F * g.slots * 4);
        w.map1 = c.take<uint32_t>(F * g.MS + 16);
    }
    if (what == 1) w.ord = c.take<uint32_t>(F * g.slots);
    if (what == 2) {
        w.cent = c.take<uint64_t>(F * g.slots);
        w.cent_tiles = c.take<uint64_t>(F * g.slots);
    }
    return w;
}

struct CompressWs {
    uint32_t *maps;
    uint8_t *packed;
    uint32_t *packed_bytes;
    uint64_t *map_off, *val_off;
    uint32_t *map_len, *val_len;
    DeflateWs dm, dv;              // deflate groups: map streams, value streams (one stream per frame each)
    size_t packed_stride;
    int spf;
};

static CompressWs carve_compress(Carver &c, const rc_config *cfg, const Geom &g)
{
    CompressWs w;
    const size_t F = (size_t)cfg->max_frames;
    w.spf = (cfg->reduction_level <= 2) ? 2 : 1;
    w.packed_stride = packed_stride_of(cfg);
    w.maps = c.take<uint32_t>(F * g.MS + 16);
    w.packed = c.take<uint8_t>(F * w.packed_stride + 16);
    w.packed_bytes = c.take<uint32_t>(F + 1);
    w.map_off = c.take<uint64_t>(F + 1);
    w.val_off = c.take<uint64_t>(F + 1);
    w.map_len = c.take<uint32_t>(F + 1);
    w.val_len = c.take<uint32_t>(F + 1);
    const bool scratch = cfg->rc_operation_mode == 1;
    w.dm = carve_deflate_ws(c, (int)F, F * ((g.map_bytes + DF_CHUNK - 1) / DF_CHUNK), scratch);
    w.dv = carve_deflate_ws(c, (int)F, w.spf == 2 ? F * ((w.packed_stride + DF_CHUNK - 1) / DF_CHUNK) : 1,
                            scratch && w.spf == 2);
    return w;
}

// stream f of a deflate group: base + f * stride, len[f] bytes (uniform when len == nullptr)
__global__ void k_stream_desc(const uint8_t *base, const uint8_t *first, size_t stride, const uint32_t *len,
                              uint32_t uniform_len, int F, uint64_t *in_off, uint32_t *in_bytes)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    in_off[f] = (uint64_t)(first + (size_t)f * stride - base);
    in_bytes[f] = len ? len[f] : uniform_len;
}

// ---- lifecycle ---------------------------------------------------------------------------------
extern "C" int rc_create(rc_ctx **out, int device)
{
    if (!out) return -1;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) return -2;     // no CUDA device: there is no fallback
    if (device < 0 || device >= n) return -1;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -2;
    if (prop.major < 10) return -4;                                      // built for sm_100a only
    rc_ctx *c = (rc_ctx *)calloc(1, sizeof(rc_ctx));
    if (!c) return -5;
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    const char *e = getenv("RECODE_B200_PRIORITY");
    c->use_priority = e ? (atoi(e) != 0) : -1;
    e = getenv("RECODE_B200_CCL_CTAS");
    c->ccl_ctas_per_sm = e ? atoi(e) : -1;
    *out = c;
    return 0;
}

extern "C" void rc_destroy(rc_ctx *ctx)
{
    if (!ctx) return;
    if (ctx->profile) {
        for (int i = 0; i < RC_MAX_MARKS; i++) cudaEventDestroy(ctx->marks[i]);
        for (int i = 0; i < RC_MAX_DMARKS; i++) cudaEventDestroy(ctx->dmarks[i]);
    }
    if (ctx->kept_tables) cudaFree(ctx->kept_tables);
    if (ctx->side_ready) {
        cudaStreamDestroy(ctx->side);
        cudaStreamDestroy(ctx->post);
        cudaStreamDestroy(ctx->side_hi);
        cudaEventDestroy(ctx->ev_fork);
        cudaEventDestroy(ctx->ev_join);
        cudaEventDestroy(ctx->ev_post);
    }
    free(ctx);
}

// stage timing of rc_reduce_compress: marks = start | threshold+map+compaction | rest of the reduction
// (scan, CCL, bit packing) | deflate | record assembly.  rc_profile_read synchronizes on the last mark and
// returns the elapsed milliseconds between consecutive marks of the most recent call.
extern "C" int rc_profile_enable(rc_ctx *ctx, int on)
{
    if (!ctx) return -1;
    if (on && !ctx->profile) {
        for (int i = 0; i < RC_MAX_MARKS; i++) RC_CUDA(ctx, cudaEventCreate(&ctx->marks[i]));
        for (int i = 0; i < RC_MAX_DMARKS; i++) RC_CUDA(ctx, cudaEventCreate(&ctx->dmarks[i]));
        ctx->n_marks = 0;
        ctx->n_dmarks = 0;
    } else if (!on && ctx->profile) {
        for (int i = 0; i < RC_MAX_MARKS; i++) cudaEventDestroy(ctx->marks[i]);
        for (int i = 0; i < RC_MAX_DMARKS; i++) cudaEventDestroy(ctx->dmarks[i]);
    }
    ctx->profile = on < 0 ? 0 : (on > 2 ? 2 : on);      // 2 = also the per-kernel marks of stage 2 (rc_profile_read_detail)
    return 0;
}

// profile level 2: elapsed ms between the per-kernel marks of the most recent rc_reduce_compress call's second stage:
// L2: labelling | cross-tile links + folds | root compaction | scan | bit packing;  L4: clear | labelling | cross-tile |
// open puddles | scan;  L1: scan | bit packing
extern "C" int rc_profile_read_detail(rc_ctx *ctx, float *ms, int capacity)
{
    if (!ctx || ctx->profile < 2 || ctx->n_dmarks < 2) return 0;
    RC_CUDA(ctx, cudaEventSynchronize(ctx->dmarks[ctx->n_dmarks - 1]));
    int n = 0;
    for (int i = 0; i + 1 < ctx->n_dmarks && n < capacity; i++, n++)
        RC_CUDA(ctx, cudaEventElapsedTime(&ms[n], ctx->dmarks[i], ctx->dmarks[i + 1]));
    return n;
}

extern "C" int rc_profile_read(rc_ctx *ctx, float *ms, int capacity)
{
    if (!ctx || !ctx->profile || ctx->n_marks < 2) return 0;
    RC_CUDA(ctx, cudaEventSynchronize(ctx->marks[ctx->n_marks - 1]));
    int n = 0;
    for (int i = 0; i + 1 < ctx->n_marks && n < capacity; i++, n++)
        RC_CUDA(ctx, cudaEventElapsedTime(&ms[n], ctx->marks[i], ctx->marks[i + 1]));
    return n;
}

extern "C" int rc_set_pipelined(rc_ctx *ctx, int on)
{
    if (!ctx) return -1;
    ctx->pipelined = on != 0;
    return 0;
}

extern "C" unsigned long long rc_launch_count(const rc_ctx *ctx) { return ctx ? ctx->launches : 0; }
extern "C" const char *rc_last_error(const rc_ctx *ctx) { return ctx ? ctx->err : "null context"; }
extern "C" int rc_version(void) { return RC_VERSION; }
extern "C" int rc_sm_count(const rc_ctx *ctx) { return ctx ? ctx->sm_count : 0; }

// ---- sizes -------------------------------------------------------------------------------------
extern "C" size_t rc_map_stride_words(size_t n_pixels)
{
    return (n_pixels + TILE_PX - 1) / TILE_PX * TILE_WORDS;
}

extern "C" size_t rc_packed_stride_bytes(const rc_config *cfg) { return packed_stride_of(cfg); }

extern "C" size_t rc_deflate_bound(size_t in_bytes)
{
    return in_bytes + 10 * ((in_bytes + DF_CHUNK - 1) / DF_CHUNK) + 8;
}

extern "C" size_t rc_workspace_bytes(const rc_config *cfg)
{
    const Geom g = make_geom(cfg->ny, cfg->nx);
    rc_config c2 = *cfg;
    Carver cc(nullptr);
    carve_reduce(cc, &c2, g, 0);
    carve_compress(cc, &c2, g);
    return cc.used() + 256;
}

// rc_ccl_label / rc_l4_centroids (label images, centroid lists: per-slot ordinals, 64-bit centroids, boxes) need more
// than the write path does; they take their own workspace so that the hot path's stays small
extern "C" size_t rc_stage_workspace_bytes(const rc_config *cfg)
{
    const Geom g = make_geom(cfg->ny, cfg->nx);
    rc_config c2 = *cfg;
    size_t best = 0;
    for (int what = 1; what < 3; what++) {
        Carver cc(nullptr);
        carve_reduce(cc, &c2, g, what);
        if (cc.used() > best) best = cc.used();
    }
    return best + 256;
}

extern "C" size_t rc_records_capacity(const rc_config *cfg)
{
    const Geom g = make_geom(cfg->ny, cfg->nx);
    size_t per = 16 + rc_deflate_bound(g.map_bytes);
    if (cfg->reduction_level <= 2) per += rc_deflate_bound(packed_stride_of(cfg));
    return per * (size_t)cfg->max_frames + 64;
}

extern "C" size_t rc_deflate_workspace_bytes(int n_streams, size_t max_in_bytes)
{
    Carver c(nullptr);
    carve_deflate_ws(c, n_streams, deflate_max_chunks(n_streams, max_in_bytes), true);
    return c.used() + 256;
}

extern "C" size_t rc_read_workspace_bytes(const rc_config *cfg)
{
    const Geom g = make_geom(cfg->ny, cfg->nx);
    const size_t F = (size_t)cfg->max_frames;
    Carver c(nullptr);
    c.take<uint32_t>(F * g.NT);
    c.take<uint32_t>(F * (g.NT + 1));
    c.take<uint16_t>(F * g.MS);
    return c.used() + 256;
}

// ---- write side --------------------------------------------------------------------------------
extern "C" int rc_make_threshold(rc_ctx *ctx, const rc_config *cfg, const void *d_dark, uint64_t eps, void *d_thr,
                                 void *stream)
{
    if (!ctx) return -1;
    if (check_cfg(ctx, cfg)) return -1;
    return launch_make_threshold(ctx, cfg->itemsize, d_dark, eps, d_thr, (size_t)cfg->ny * cfg->nx, (cudaStream_t)stream);
}

// The reduction in two stages.  Stage 1 is the streaming kernel (threshold, binary map, value compaction); for
// L1 / L2 / L3 the binary map is final after it.  Stage 2 is everything that follows on the compact data.
static int reduce_stage1(rc_ctx *ctx, const rc_config *cfg, const Geom &g, const ReduceWs &w, const void *frames, int F,
                         const void *thr, uint32_t *maps, cudaStream_t st)
{
    const int level = cfg->reduction_level, isz = cfg->itemsize;
    if (level == 1) return launch_reduce_tiles(ctx, g, isz, 1, frames, thr, F, maps, w.tilecnt, w.wordpre, w.vals, st);
    if (level == 3) return launch_reduce_tiles(ctx, g, isz, 0, frames, thr, F, maps, w.tilecnt, w.wordpre, nullptr, st);
    if (level == 2) return launch_reduce_tiles(ctx, g, isz, 2, frames, thr, F, maps, w.tilecnt, w.wordpre, w.vals, st);
    // level 4: the threshold map goes to map1, the centroid map (stage 2) to maps
    return launch_reduce_tiles(ctx, g, isz, 2, frames, thr, F, w.map1, w.tilecnt, w.wordpre, w.vals, st);
}

static int reduce_stage2(rc_ctx *ctx, const rc_config *cfg, const Geom &g, const ReduceWs &w, int F, uint32_t *maps,
                         uint8_t *packed, size_t packed_stride, uint32_t *packed_bytes, uint32_t *counts,
                         cudaStream_t st)
{
    const int level = cfg->reduction_level, b = cfg->bit_depth, isz = cfg->itemsize;
    int rc;
    ctx->n_dmarks = 0;
    rc_dmark(ctx, 0, st);
    if (level == 1) {
        if ((rc = launch_scan_tiles(ctx, g, w.tilecnt, F, w.tilepre, counts, packed_bytes, b, st))) return rc;
        rc_dmark(ctx, 1, st);
        rc = launch_bitpack(ctx, g, isz, w.vals, w.tilepre, F, b, packed, packed_stride, st);
        rc_dmark(ctx, 2, st);
        return rc;
    }
    if (level == 3) return launch_scan_tiles(ctx, g, w.tilecnt, F, w.tilepre, counts, nullptr, 0, st);
    if (level == 2) {
        const int sum = cfg->l2_statistics == 2;
        // RC_ABLATE (timing experiments only, results are not valid records): bit 0 = no labelling stage at all
        static const int ablate = getenv("RC_ABLATE") ? atoi(getenv("RC_ABLATE")) : 0;
        if (ablate & 1) {
            RC_CUDA(ctx, cudaMemsetAsync(counts, 0, (size_t)F * sizeof(uint32_t), st));
            RC_CUDA(ctx, cudaMemsetAsync(packed_bytes, 0, (size_t)F * sizeof(uint32_t), st));
            return 0;
        }
        if ((rc = launch_ccl_tiles(ctx, g, sum ? 2 : 1, maps, w.wordpre, w.tilecnt, (const uint32_t *)w.vals, w.tileovf,
                                   w.xcount, w.xlinks, w.parent, w.acc, 0, nullptr, nullptr, nullptr, nullptr, F,
                                   st))) return rc;
        rc_dmark(ctx, 1, st);
        if ((rc = launch_ccl_border(ctx, g, sum ? 2 : 1, maps, w.wordpre, w.tileovf, w.xcount, w.xlinks, w.parent, w.acc,
                                    F, st))) return rc;
        rc_dmark(ctx, 2, st);
        if ((rc = launch_ccl_roots(ctx, g, 1, w.tilecnt, w.parent, w.acc, nullptr, w.rootcnt, nullptr, w.stats16,
                                   nullptr, F, st))) return rc;
        rc_dmark(ctx, 3, st);
        if ((rc = launch_scan_tiles(ctx, g, w.rootcnt, F, w.rootpre, counts, packed_bytes, b, st))) return rc;
        rc_dmark(ctx, 4, st);
        rc = launch_bitpack(ctx, g, 2, w.stats16, w.rootpre, F, b, packed, packed_stride, st);
        rc_dmark(ctx, 5, st);
        return rc;
    }
    // level 4: puddles inside one tile are finished by k_ccl_tiles, the ones that cross tiles by k_l4_open
    RC_CUDA(ctx, cudaMemsetAsync(maps, 0, (size_t)F * g.MS * sizeof(uint32_t), st));
    rc_dmark(ctx, 1, st);
    if ((rc = launch_ccl_tiles(ctx, g, 3, w.map1, w.wordpre, w.tilecnt, (const uint32_t *)w.vals, w.tileovf, w.xcount,
                               w.xlinks, w.parent, w.acc, cfg->l4_centroiding, w.bbox, maps, nullptr, w.rootcnt, F,
                               st))) return rc;
    rc_dmark(ctx, 2, st);
    if ((rc = launch_ccl_border(ctx, g, 3, w.map1, w.wordpre, w.tileovf, w.xcount, w.xlinks, w.parent, w.bbox, F,
                                st))) return rc;
    rc_dmark(ctx, 3, st);
    if ((rc = launch_l4_open(ctx, g, cfg->l4_centroiding, w.map1, w.wordpre, w.tilecnt, w.tileovf, w.xcount, w.xlinks,
                             w.parent, w.acc, w.bbox, (const uint32_t *)w.vals, maps, nullptr, w.rootcnt, F, st)))
        return rc;
    rc_dmark(ctx, 4, st);
    rc = launch_scan_tiles(ctx, g, w.rootcnt, F, w.rootpre, counts, nullptr, 0, st);
    rc_dmark(ctx, 5, st);
    return rc;
}

// one deflate group (F streams; group 0 = maps, 1 = values): descriptors, then encode (wrap = 1) or size
// (wrap = 0) its chunks
static int deflate_group(rc_ctx *ctx, const rc_config *cfg, int group, const uint8_t *base, const uint8_t *first,
                         size_t stride, const uint32_t *len, uint32_t uniform_len, int F, uint64_t *in_off,
                         uint32_t *in_bytes, const DeflateWs &d, cudaStream_t st)
{
    k_stream_desc<<<(F + 127) / 128, 128, 0, st>>>(base, first, stride, len, uniform_len, F, in_off, in_bytes);
    RC_LAUNCH_CHECK(ctx, "k_stream_desc");
    // The frames of one acquisition are statistically alike: levels 1..5 use one sampled code per group, kept in
    // the context and rebuilt every RC_TABLE_REFRESH calls.  Any code encodes any data (all symbols are
    // smoothed to non-zero counts, a chunk that would grow is stored), so staleness only costs ratio.
    const int shared = cfg->compression_level >= 1 && cfg->compression_level < 6 && cfg->rc_operation_mode == 1;
    void *kept = nullptr;
    int build = 1;
    if (shared) {
        kept = (uint8_t *)ctx->kept_tables + (size_t)group * deflate_table_bytes();
        build = ctx->table_age[group] == 0;
        ctx->table_age[group] = (ctx->table_age[group] + 1) % RC_TABLE_REFRESH;
    }
    return launch_deflate_streams(ctx, cfg->compression_level, cfg->rc_operation_mode == 1, shared, kept, build, base,
                                  in_off, in_bytes, F, d, st);
}

// the kept codes belong to one configuration: any change starts them afresh
static int prepare_kept_tables(rc_ctx *ctx, const rc_config *cfg)
{
    if (!ctx->kept_tables) RC_CUDA(ctx, cudaMalloc(&ctx->kept_tables, 2 * deflate_table_bytes()));
    const unsigned long long key = ((unsigned long long)cfg->ny << 40) ^ ((unsigned long long)cfg->nx << 20) ^
                                   ((unsigned long long)cfg->bit_depth << 12) ^ ((unsigned long long)cfg->itemsize << 10) ^
                                   ((unsigned long long)cfg->reduction_level << 7) ^
                                   ((unsigned long long)cfg->l2_statistics << 5) ^
                                   ((unsigned long long)cfg->l4_centroiding << 3) ^ (unsigned long long)cfg->compression_level;
    if (key != ctx->table_key) {
        ctx->table_key = key;
        ctx->table_age[0] = ctx->table_age[1] = 0;
    }
    return 0;
}

static int ensure_side_stream(rc_ctx *ctx)
{
    if (ctx->side_ready) return 0;
    int lo = 0, hi = 0;                                   // numerically lower = higher priority
    RC_CUDA(ctx, cudaDeviceGetStreamPriorityRange(&lo, &hi));
    RC_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->side, cudaStreamNonBlocking, lo));
    RC_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->side_hi, cudaStreamNonBlocking, hi));
    RC_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->post, cudaStreamNonBlocking, hi));
    RC_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
    RC_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
    RC_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_post, cudaEventDisableTiming));
    ctx->side_ready = 1;
    return 0;
}

extern "C" int rc_reduce(rc_ctx *ctx, const rc_config *cfg, const void *d_frames, int n_frames, const void *d_thr,
                         void *d_workspace, size_t workspace_bytes, uint32_t *d_maps, uint8_t *d_packed,
                         uint32_t *d_packed_bytes, uint32_t *d_counts, void *stream)
{
    if (!ctx) return -1;
    if (check_cfg(ctx, cfg)) return -1;
    if (n_frames < 0 || n_frames > cfg->max_frames) RC_FAIL(ctx, -1, "n_frames %d exceeds max_frames %d", n_frames, cfg->max_frames);
    if (workspace_bytes < rc_workspace_bytes(cfg)) RC_FAIL(ctx, -1, "workspace too small");
    if (((uintptr_t)d_maps | (uintptr_t)d_workspace) % 16) RC_FAIL(ctx, -1, "d_maps and d_workspace must be 16-byte aligned");
    const Geom g = make_geom(cfg->ny, cfg->nx);
    Carver c(d_workspace);
    const ReduceWs w = carve_reduce(c, cfg, g, 0);
    int rc;
    if ((rc = reduce_stage1(ctx, cfg, g, w, d_frames, n_frames, d_thr, d_maps, (cudaStream_t)stream))) return rc;
    return reduce_stage2(ctx, cfg, g, w, n_frames, d_maps, d_packed, packed_stride_of(cfg), d_packed_bytes, d_counts,
                         (cudaStream_t)stream);
}

extern "C" int rc_reduce_compress(rc_ctx *ctx, const rc_config *cfg, const void *d_frames, int n_frames,
                                  const void *d_thr, uint32_t first_frame_id, void *d_workspace, size_t workspace_bytes,
                                  uint8_t *d_records, size_t records_capacity, uint64_t *d_record_offsets,
                                  uint32_t *d_counts, uint32_t *d_status, void *stream)
{
    if (!ctx) return -1;
    if (check_cfg(ctx, cfg)) return -1;
    if (n_frames < 0 || n_frames > cfg->max_frames) RC_FAIL(ctx, -1, "n_frames %d exceeds max_frames %d", n_frames, cfg->max_frames);
    if (workspace_bytes < rc_workspace_bytes(cfg)) RC_FAIL(ctx, -1, "workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const Geom g = make_geom(cfg->ny, cfg->nx);
    Carver c(d_workspace);
    const ReduceWs w = carve_reduce(c, cfg, g, 0);
    const CompressWs cw = carve_compress(c, cfg, g);
    const int F = n_frames;
    int rc;
    RC_CUDA(ctx, cudaMemsetAsync(d_status, 0, sizeof(uint32_t), st));
    if (F == 0) {
        RC_CUDA(ctx, cudaMemsetAsync(d_record_offsets, 0, sizeof(uint64_t), st));
        return 0;
    }
    const uint8_t *base = (const uint8_t *)d_workspace;
    const int level = cfg->reduction_level;
    if ((rc = prepare_kept_tables(ctx, cfg))) return rc;
    // The map streams of L1 / L2 / L3 are final after stage 1: they are deflated on the side stream while the main
    // stream labels puddles and packs the values.  (L4's map is the last product of stage 2.)
    const bool fork = level != 4 && F > 0;
    if ((rc = ensure_side_stream(ctx))) return rc;
    rc_mark(ctx, 0, st);
    if ((rc = reduce_stage1(ctx, cfg, g, w, d_frames, F, d_thr, cw.maps, st))) return rc;
    rc_mark(ctx, 1, st);
    // everything else on the context's high-priority stream(s); the caller's stream joins at the end
    const bool prio = ctx->use_priority < 0 ? ctx->pipelined != 0 : ctx->use_priority != 0;
    cudaStream_t sp = prio ? ctx->post : st;
    RC_CUDA(ctx, cudaEventRecord(ctx->ev_fork, st));
    if (sp != st) RC_CUDA(ctx, cudaStreamWaitEvent(sp, ctx->ev_fork, 0));
    cudaStream_t sm = sp;
    if (fork) {
        static const bool side_low = getenv("RECODE_B200_SIDE_LOW") && atoi(getenv("RECODE_B200_SIDE_LOW")) != 0;
        sm = (prio && !side_low) ? ctx->side_hi : ctx->side;
        RC_CUDA(ctx, cudaStreamWaitEvent(sm, ctx->ev_fork, 0));
        if ((rc = deflate_group(ctx, cfg, 0, base, (const uint8_t *)cw.maps, g.MS * 4, nullptr, (uint32_t)g.map_bytes, F,
                                cw.map_off, cw.map_len, cw.dm, sm))) return rc;
        RC_CUDA(ctx, cudaEventRecord(ctx->ev_join, sm));
    }
    if ((rc = reduce_stage2(ctx, cfg, g, w, F, cw.maps, cw.packed, cw.packed_stride, cw.packed_bytes, d_counts, sp)))
        return rc;
    rc_mark(ctx, 2, sp);
    if (!fork && (rc = deflate_group(ctx, cfg, 0, base, (const uint8_t *)cw.maps, g.MS * 4, nullptr,
                                     (uint32_t)g.map_bytes, F, cw.map_off, cw.map_len, cw.dm, sp))) return rc;
    if (cw.spf == 2 && (rc = deflate_group(ctx, cfg, 1, base, cw.packed, cw.packed_stride, cw.packed_bytes, 0, F,
                                           cw.val_off, cw.val_len, cw.dv, sp))) return rc;
    if (fork) RC_CUDA(ctx, cudaStreamWaitEvent(sp, ctx->ev_join, 0));
    rc_mark(ctx, 3, sp);
    const int wrap = cfg->rc_operation_mode == 1;
    if ((rc = launch_layout_records(ctx, cw.dm, cw.dv, cw.packed_bytes, F, cw.spf, cfg->rc_operation_mode, first_frame_id,
                                    d_records, records_capacity, d_record_offsets, d_status, sp))) return rc;
    if ((rc = launch_copy_pieces(ctx, cw.dm, wrap, base, cw.map_off, cw.map_len, F, d_records, records_capacity, d_status,
                                 sp))) return rc;
    if (cw.spf == 2 && (rc = launch_copy_pieces(ctx, cw.dv, wrap, base, cw.val_off, cw.val_len, F, d_records,
                                                records_capacity, d_status, sp))) return rc;
    rc_mark(ctx, 4, sp);
    if (sp != st) {
        RC_CUDA(ctx, cudaEventRecord(ctx->ev_post, sp));
        RC_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_post, 0));
    }
    return 0;
}

extern "C" int rc_ccl_label(rc_ctx *ctx, const rc_config *cfg, const uint32_t *d_maps, int n_frames, void *d_workspace,
                            size_t workspace_bytes, int32_t *d_labels, uint32_t *d_counts, void *stream)
{
    if (!ctx) return -1;
    if (check_cfg(ctx, cfg)) return -1;
    if (n_frames < 0 || n_frames > cfg->max_frames) RC_FAIL(ctx, -1, "n_frames exceeds max_frames");
    if (workspace_bytes < rc_stage_workspace_bytes(cfg)) RC_FAIL(ctx, -1, "workspace too small (rc_stage_workspace_bytes)");
    cudaStream_t st = (cudaStream_t)stream;
    const Geom g = make_geom(cfg->ny, cfg->nx);
    Carver c(d_workspace);
    const ReduceWs w = carve_reduce(c, cfg, g, 1);
    const int F = n_frames;
    int rc;
    if ((uintptr_t)d_maps % 16) RC_FAIL(ctx, -1, "d_maps must be 16-byte aligned");
    if ((rc = launch_map_counts(ctx, g, d_maps, F, w.tilecnt, w.wordpre, st))) return rc;
    if ((rc = launch_ccl_init(ctx, g, w.tilecnt, w.parent, F, st))) return rc;
    if ((rc = launch_ccl_union(ctx, g, d_maps, w.wordpre, w.parent, F, st))) return rc;
    if ((rc = launch_ccl_flatten(ctx, g, d_maps, w.wordpre, w.parent, F, st))) return rc;
    if ((rc = launch_ccl_roots(ctx, g, 0, w.tilecnt, w.parent, nullptr, nullptr, w.rootcnt, w.ord, nullptr, nullptr, F, st))) return rc;
    if ((rc = launch_scan_tiles(ctx, g, w.rootcnt, F, w.rootpre, d_counts, nullptr, 0, st))) return rc;
    return launch_ccl_label_image(ctx, g, d_maps, w.wordpre, w.parent, w.ord, w.rootpre, d_labels, F, st);
}

extern "C" int rc_l4_centroids(rc_ctx *ctx, const rc_config *cfg, const void *d_frames, int n_frames, const void *d_thr,
                               void *d_workspace, size_t workspace_bytes, float *d_centroids, size_t centroid_capacity,
                               uint32_t *d_counts, void *stream)
{
    if (!ctx) return -1;
    if (check_cfg(ctx, cfg)) return -1;
    if (n_frames < 0 || n_frames > cfg->max_frames) RC_FAIL(ctx, -1, "n_frames exceeds max_frames");
    if (workspace_bytes < rc_stage_workspace_bytes(cfg)) RC_FAIL(ctx, -1, "workspace too small (rc_stage_workspace_bytes)");
    cudaStream_t st = (cudaStream_t)stream;
    const Geom g = make_geom(cfg->ny, cfg->nx);
    Carver c(d_workspace);
    const ReduceWs w = carve_reduce(c, cfg, g, 2);
    const int F = n_frames, isz = cfg->itemsize;
    int rc;
    if ((rc = launch_reduce_tiles(ctx, g, isz, 2, d_frames, d_thr, F, w.map1, w.tilecnt, w.wordpre, w.vals, st))) return rc;
    if ((rc = launch_ccl_tiles(ctx, g, 3, w.map1, w.wordpre, w.tilecnt, (const uint32_t *)w.vals, w.tileovf, w.xcount,
                               w.xlinks, w.parent, w.acc, cfg->l4_centroiding, w.bbox, nullptr, w.cent, w.rootcnt, F,
                               st))) return rc;
    if ((rc = launch_ccl_border(ctx, g, 3, w.map1, w.wordpre, w.tileovf, w.xcount, w.xlinks, w.parent, w.bbox, F,
                                st))) return rc;
    if ((rc = launch_l4_open(ctx, g, cfg->l4_centroiding, w.map1, w.wordpre, w.tilecnt, w.tileovf, w.xcount, w.xlinks,
                             w.parent, w.acc, w.bbox, (const uint32_t *)w.vals, nullptr, w.cent, w.rootcnt, F, st)))
        return rc;
    if ((rc = launch_ccl_roots(ctx, g, 2, w.tilecnt, w.parent, nullptr, w.cent, w.rootcnt, nullptr, nullptr, w.cent_tiles,
                               F, st))) return rc;
    if ((rc = launch_scan_tiles(ctx, g, w.rootcnt, F, w.rootpre, d_counts, nullptr, 0, st))) return rc;
    return launch_gather_centroids(ctx, g, w.cent_tiles, w.rootpre, F, d_centroids, centroid_capacity, st);
}

extern "C" int rc_deflate_zlib(rc_ctx *ctx, int compression_level, const uint8_t *d_in, const uint64_t *d_in_offsets,
                               const uint32_t *d_in_bytes, int n_streams, size_t max_in_bytes, void *d_workspace,
                               size_t workspace_bytes, uint8_t *d_out, size_t out_stride, uint32_t *d_out_bytes,
                               void *stream)
{
    if (!ctx) return -1;
    if (compression_level < 0 || compression_level > 9) RC_FAIL(ctx, -1, "compression_level must be 0..9");
    if (n_streams <= 0) return 0;
    if (out_stride < rc_deflate_bound(max_in_bytes)) RC_FAIL(ctx, -1, "out_stride smaller than rc_deflate_bound");
    if (workspace_bytes < rc_deflate_workspace_bytes(n_streams, max_in_bytes)) RC_FAIL(ctx, -1, "workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    Carver c(d_workspace);
    DeflateWs w = carve_deflate_ws(c, n_streams, deflate_max_chunks(n_streams, max_in_bytes), true);
    uint32_t *status = w.counters + 4;
    int rc;
    if ((rc = launch_deflate_streams(ctx, compression_level, 1, 0, nullptr, 1, d_in, d_in_offsets, d_in_bytes, n_streams, w, st))) return rc;
    if ((rc = launch_layout_strided(ctx, w, n_streams, out_stride, d_out_bytes, st))) return rc;
    return launch_copy_pieces(ctx, w, 1, d_in, d_in_offsets, d_in_bytes, n_streams, d_out, (size_t)n_streams * out_stride,
                              status, st);
}

// ---- read side ---------------------------------------------------------------------------------
extern "C" size_t rc_inflate_workspace_bytes(int n_streams, size_t out_stride)
{
    return inflate_workspace_bytes(n_streams, out_stride) + 256;
}

extern "C" int rc_inflate_zlib(rc_ctx *ctx, const uint8_t *d_in, const uint64_t *d_in_offsets, const uint32_t *d_in_bytes,
                               int n_streams, void *d_workspace, size_t workspace_bytes, uint8_t *d_out,
                               size_t out_stride, uint32_t *d_out_bytes, uint32_t *d_status, void *stream)
{
    if (!ctx) return -1;
    if (n_streams <= 0) return 0;
    if (out_stride % 16) RC_FAIL(ctx, -1, "out_stride must be a multiple of 16");
    if (workspace_bytes < rc_inflate_workspace_bytes(n_streams, out_stride)) RC_FAIL(ctx, -1, "workspace too small");
    return launch_inflate(ctx, d_in, d_in_offsets, d_in_bytes, n_streams, d_workspace, d_out, out_stride, d_out_bytes,
                          d_status, (cudaStream_t)stream);
}

struct ReadWs {
    uint32_t *tilecnt, *tilepre;
    uint16_t *wordpre;
};

static ReadWs carve_read(Carver &c, const rc_config *cfg, const Geom &g)
{
    ReadWs w;
    const size_t F = (size_t)cfg->max_frames;
    w.tilecnt = c.take<uint32_t>(F * g.NT);
    w.tilepre = c.take<uint32_t>(F * (g.NT + 1));
    w.wordpre = c.take<uint16_t>(F * g.MS);
    return w;
}

extern "C" int rc_unpack_sparse(rc_ctx *ctx, const rc_config *cfg, const uint32_t *d_maps, const uint8_t *d_packed,
                                size_t packed_stride, int n_frames, void *d_workspace, size_t workspace_bytes,
                                uint64_t *d_triples, size_t triple_capacity, uint32_t *d_counts, void *stream)
{
    if (!ctx) return -1;
    if (check_cfg(ctx, cfg)) return -1;
    if (n_frames < 0 || n_frames > cfg->max_frames) RC_FAIL(ctx, -1, "n_frames exceeds max_frames");
    if (workspace_bytes < rc_read_workspace_bytes(cfg)) RC_FAIL(ctx, -1, "workspace too small");
    if (packed_stride % 4) RC_FAIL(ctx, -1, "packed_stride must be a multiple of 4");
    if ((uintptr_t)d_maps % 16) RC_FAIL(ctx, -1, "d_maps must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const Geom g = make_geom(cfg->ny, cfg->nx);
    Carver c(d_workspace);
    const ReadWs w = carve_read(c, cfg, g);
    int rc;
    if ((rc = launch_map_counts(ctx, g, d_maps, n_frames, w.tilecnt, w.wordpre, st))) return rc;
    if ((rc = launch_scan_tiles(ctx, g, w.tilecnt, n_frames, w.tilepre, d_counts, nullptr, 0, st))) return rc;
    return launch_unpack_sparse(ctx, g, cfg->reduction_level, cfg->bit_depth, d_maps, d_packed, packed_stride, w.wordpre,
                                w.tilepre, n_frames, d_triples, triple_capacity, st);
}

extern "C" int rc_unpack_dense(rc_ctx *ctx, const rc_config *cfg, const uint32_t *d_maps, const uint8_t *d_packed,
                               size_t packed_stride, int n_frames, void *d_workspace, size_t workspace_bytes,
                               void *d_dense, uint32_t *d_sum, uint32_t *d_counts, void *stream)
{
    if (!ctx) return -1;
    if (check_cfg(ctx, cfg)) return -1;
    if (n_frames < 0 || n_frames > cfg->max_frames) RC_FAIL(ctx, -1, "n_frames exceeds max_frames");
    if (workspace_bytes < rc_read_workspace_bytes(cfg)) RC_FAIL(ctx, -1, "workspace too small");
    if (packed_stride % 4) RC_FAIL(ctx, -1, "packed_stride must be a multiple of 4");
    if ((uintptr_t)d_maps % 16) RC_FAIL(ctx, -1, "d_maps must be 16-byte aligned");
    if (!d_dense && !d_sum) RC_FAIL(ctx, -1, "need d_dense and/or d_sum");
    cudaStream_t st = (cudaStream_t)stream;
    const Geom g = make_geom(cfg->ny, cfg->nx);
    Carver c(d_workspace);
    const ReadWs w = carve_read(c, cfg, g);
    int rc;
    if ((rc = launch_map_counts(ctx, g, d_maps, n_frames, w.tilecnt, w.wordpre, st))) return rc;
    if ((rc = launch_scan_tiles(ctx, g, w.tilecnt, n_frames, w.tilepre, d_counts, nullptr, 0, st))) return rc;
    return launch_unpack_dense(ctx, g, cfg->itemsize, cfg->reduction_level, cfg->bit_depth, d_maps, d_packed,
                               packed_stride, w.wordpre, w.tilepre, n_frames, d_dense, d_sum, st);
}

extern "C" int rc_bit_unpack(rc_ctx *ctx, int bit_depth, const uint8_t *d_packed, uint64_t n_values, uint64_t *d_out,
                             void *stream)
{
    if (!ctx) return -1;
    if (bit_depth < 1 || bit_depth > 16) RC_FAIL(ctx, -1, "bit_depth must be 1..16");
    return launch_bitunpack_flat(ctx, bit_depth, d_packed, n_values, d_out, (cudaStream_t)stream);
}

extern "C" int rc_bit_pack(rc_ctx *ctx, int bit_depth, const uint16_t *d_vals, uint64_t n_values, uint8_t *d_packed,
                           void *stream)
{
    if (!ctx) return -1;
    if (bit_depth < 1 || bit_depth > 16) RC_FAIL(ctx, -1, "bit_depth must be 1..16");
    if ((uintptr_t)d_packed % 4) RC_FAIL(ctx, -1, "d_packed must be 4-byte aligned");
    return launch_bitpack_flat(ctx, bit_depth, d_vals, n_values, d_packed, (cudaStream_t)stream);
}

extern "C" int rc_recalibrate(rc_ctx *ctx, int itemsize, const void *d_frames, const double *d_diff, size_t n_pixels,
                              int n_frames, void *d_out, void *stream)
{
    if (!ctx) return -1;
    if (itemsize != 1 && itemsize != 2) RC_FAIL(ctx, -1, "itemsize must be 1 or 2 (got %d)", itemsize);
    if (n_frames < 0) RC_FAIL(ctx, -1, "n_frames must be >= 0");
    return launch_recalibrate(ctx, itemsize, d_frames, d_diff, n_pixels, n_frames, d_out, (cudaStream_t)stream);
}

extern "C" size_t rc_median_std_workspace_bytes(size_t n_pixels) { return median_std_workspace_bytes(n_pixels); }

extern "C" int rc_median_std(rc_ctx *ctx, int itemsize, const void *d_stack, int n_frames, size_t n_pixels,
                             float *d_median, float *d_std, void *d_workspace, size_t workspace_bytes, void *stream)
{
    if (!ctx) return -1;
    if (itemsize != 1 && itemsize != 2) RC_FAIL(ctx, -1, "itemsize must be 1 or 2 (got %d)", itemsize);
    if (n_frames < 1 || n_frames > (1 << 20)) RC_FAIL(ctx, -1, "n_frames must be 1..2^20");
    if (workspace_bytes < median_std_workspace_bytes(n_pixels)) RC_FAIL(ctx, -1, "workspace too small");
    return launch_median_std(ctx, itemsize, d_stack, n_pixels, n_frames, d_median, d_std, d_workspace,
                             (cudaStream_t)stream);
}

extern "C" int rc_pixel_thresholds(rc_ctx *ctx, int itemsize, const void *d_stack, int n_frames, size_t n_pixels,
                                   const float *d_thr, int expected_n_events, int as_run, float *d_out, void *stream)
{
    if (!ctx) return -1;
    if (itemsize != 1 && itemsize != 2) RC_FAIL(ctx, -1, "itemsize must be 1 or 2 (got %d)", itemsize);
    if (n_frames < 1) RC_FAIL(ctx, -1, "n_frames must be >= 1");
    return launch_cal_topk(ctx, itemsize, d_stack, n_pixels, n_frames, d_thr, expected_n_events, as_run, d_out,
                           (cudaStream_t)stream);
}
