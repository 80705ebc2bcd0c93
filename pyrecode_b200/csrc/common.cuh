// common.cuh -- shared definitions of librecode_b200 (sm_100a only).
//
// Device data layout (all per batch of F frames; P = ny*nx pixels per frame):
//   map      uint32 [F][MS]        binary map words, pixel i -> word i>>5, bit i&31 (== byte i>>3, bit i&7
//                                  on a little-endian host: recode_writer.py:622-634).  MS = MW rounded up to 8.
//   tile     32768 consecutive pixels (1024 map words) in linear (raster) order; NT tiles per frame.  For
//            nx = 4096 a tile is a strip of 8 whole rows: most puddles are labelled entirely inside one tile.
//   segment  256 consecutive pixels (8 map words = one 32-byte sector); 128 segments per tile.
//   slot     index of a foreground pixel in the "tile-compacted" space: tile*32768 + rank of the pixel
//            among the foreground pixels of its tile.  Monotone in raster order.  All per-foreground
//            arrays (vals, parent, acc ...) are indexed by slot, so a tile never needs another tile's
//            prefix to place its data, and only tilecnt[tile] entries per tile are ever touched.
//   tilecnt  uint32 [F][NT]        foreground pixels per tile
//   tilepre  uint32 [F][NT+1]      exclusive scan of tilecnt (tilepre[NT] = n foreground pixels of the frame)
//   wordpre  uint16 [F][MS]        foreground pixels of the tile before each map word (exclusive, per tile)
//   tileovf  uint8  [F][NT]        1 when a tile had more foreground pixels than the shared-memory labelling
//                                  handles (CCL_CAP): its pixels are then linked by the global kernels instead
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/recode_b200.h"

#define RC_VERSION 100

constexpr int TILE_LOG2 = 15;
constexpr int TILE_PX = 1 << TILE_LOG2;      // 32768
constexpr int TILE_WORDS = TILE_PX / 32;     // 1024
constexpr int TILE_WORDS_LOG2 = TILE_LOG2 - 5;
constexpr int SEG_PX = 256;
constexpr int SEG_WORDS = SEG_PX / 32;       // 8
constexpr int SEGS_PER_TILE = TILE_PX / SEG_PX;  // 128
constexpr int SEGS_PER_TILE_LOG2 = TILE_LOG2 - 8;

// union-find parent encoding: bit 31 set = "not a tile-local root; low bits = slot of my tile-local root"
// (written by k_reduce_tiles after the tile-local labelling).  Roots and re-parented tile-local roots are plain.
constexpr uint32_t UF_FLAG = 0x80000000u;

constexpr int RC_MAX_MARKS = 8;
constexpr int RC_MAX_DMARKS = 12;

struct rc_ctx {
    int device;
    int sm_count;
    char err[512];
    unsigned long long launches;       // kernels launched through this context (bench.py's gpu_launches)
    int profile;                       // when set, stage boundaries of rc_reduce_compress record events
    int n_marks;
    cudaEvent_t marks[RC_MAX_MARKS];
    // profile == 2: additionally one event after every kernel (group) of the reduction's second stage
    int n_dmarks;
    cudaEvent_t dmarks[RC_MAX_DMARKS];
    bool deflate_attr_set, inflate_attr_set;
    // side stream: the map streams are deflated while the main stream still labels puddles / packs values
    cudaStream_t side;
    cudaEvent_t ev_fork, ev_join;
    int side_ready;
    // Everything after the streaming kernel runs on a context-owned HIGH-priority stream (`post`; `side` has the
    // same priority): with several batches in flight (one context each) the small latency-bound kernels of one
    // batch are then dispatched in front of the not yet resident CTAs of another batch's streaming kernel and
    // share the SMs with it, instead of waiting for its last wave.
    // (Measured, 32 frames per batch, 3 batches in flight: L2 56.0 -> 60.7 k frames/s, L4 50.4 -> 53.4 k, L1 82.6 ->
    // 84.4 k.)
    cudaStream_t post, side_hi;
    cudaEvent_t ev_post;
    int pipelined;                     // rc_set_pipelined: several contexts keep batches in flight on this GPU
    int use_priority;                  // RECODE_B200_PRIORITY: 0 = never, 1 = always, unset = when pipelined
    int ccl_ctas_per_sm;               // RECODE_B200_CCL_CTAS: n > 0 = persistent k_ccl_tiles with n CTAs per SM,
                                       // 0 = one CTA per tile, unset = 2 (L2) / 3 (L4) when pipelined, else 0
    // Huffman codes kept across rc_reduce_compress calls (compression levels 1..5): [0] map streams, [1] value
    // streams.  Frames of one acquisition share their statistics, so a code is rebuilt only every
    // RC_TABLE_REFRESH calls (or when the configuration changes) instead of once per batch.
    void *kept_tables;
    int table_age[2];
    unsigned long long table_key;
};

constexpr int RC_TABLE_REFRESH = 16;

static inline void rc_mark(rc_ctx *ctx, int idx, cudaStream_t st)
{
    if (ctx->profile && idx < RC_MAX_MARKS) {
        cudaEventRecord(ctx->marks[idx], st);
        if (idx + 1 > ctx->n_marks) ctx->n_marks = idx + 1;
    }
}

static inline void rc_dmark(rc_ctx *ctx, int idx, cudaStream_t st)
{
    if (ctx->profile >= 2 && idx < RC_MAX_DMARKS) {
        cudaEventRecord(ctx->dmarks[idx], st);
        if (idx + 1 > ctx->n_dmarks) ctx->n_dmarks = idx + 1;
    }
}

#define RC_FAIL(ctx, code, ...)                                   \
    do {                                                          \
        snprintf((ctx)->err, sizeof((ctx)->err), __VA_ARGS__);    \
        return (code);                                            \
    } while (0)

#define RC_CUDA(ctx, call)                                                                        \
    do {                                                                                          \
        cudaError_t e__ = (call);                                                                 \
        if (e__ != cudaSuccess)                                                                   \
            RC_FAIL(ctx, -2, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__,                \
                    cudaGetErrorString(e__));                                                     \
    } while (0)

#define RC_LAUNCH_CHECK(ctx, name)                                                                \
    do {                                                                                          \
        cudaError_t e__ = cudaGetLastError();                                                     \
        if (e__ != cudaSuccess)                                                                   \
            RC_FAIL(ctx, -3, "launch of %s failed: %s", name, cudaGetErrorString(e__));          \
        (ctx)->launches++;                                                                        \
    } while (0)

// ---- geometry ------------------------------------------------------------------------------
struct Geom {
    int ny, nx;
    size_t P;        // pixels per frame
    size_t MW;       // map words actually holding pixels: ceil(P/32)
    size_t MS;       // map stride in words (multiple of 8 -> 32-byte aligned frames)
    int NT;          // tiles per frame
    size_t slots;    // NT * TILE_PX
    size_t map_bytes;  // ceil(P/8): the on-disk size of a map
};

static inline __host__ __device__ size_t round_up(size_t x, size_t m) { return (x + m - 1) / m * m; }

static inline Geom make_geom(int ny, int nx)
{
    Geom g;
    g.ny = ny; g.nx = nx;
    g.P = (size_t)ny * (size_t)nx;
    g.MW = (g.P + 31) / 32;
    g.NT = (int)((g.P + TILE_PX - 1) / TILE_PX);
    g.MS = (size_t)g.NT * TILE_WORDS;   // whole tiles: kernels may write full tile rows of words
    g.slots = (size_t)g.NT * TILE_PX;
    g.map_bytes = (g.P + 7) / 8;
    return g;
}

// ---- bump allocator over the caller-supplied workspace ---------------------------------------
struct Carver {
    uint8_t *base;
    size_t off;
    explicit Carver(void *p) : base((uint8_t *)p), off(0) {}
    template <typename T>
    T *take(size_t n)
    {
        off = round_up(off, 256);
        T *r = base ? (T *)(base + off) : nullptr;
        off += n * sizeof(T);
        return r;
    }
    size_t used() const { return round_up(off, 256); }
};

// ---- warp / block primitives -----------------------------------------------------------------
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v)
{
    // shfl.up hands back its own predicate "the source lane exists": a predicated add instead of a lane compare and a
    // select per step (two instructions a step instead of four)
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        asm volatile("{\n.reg .u32 r0;\n.reg .pred p;\n"
                     "shfl.sync.up.b32 r0|p, %0, %1, 0, 0xffffffff;\n"
                     "@p add.u32 %0, %0, r0;\n}"
                     : "+r"(v) : "r"(d));
    }
    return v;
}

// exclusive scan over a block of NW warps; returns exclusive prefix, total in *total.
// s_warp must hold NW+1 uint32.  Contains two __syncthreads().
template <int NW>
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t *s_warp, uint32_t *total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = warp_incl_scan(v);
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < NW ? s_warp[lane] : 0;
        uint32_t wi = warp_incl_scan(w);
        if (lane < NW) s_warp[lane] = wi - w;
        if (lane == NW - 1) s_warp[NW] = wi;
    }
    __syncthreads();
    *total = s_warp[NW];
    return incl - v + s_warp[warp];
}

__device__ __forceinline__ uint32_t ld_nc_u32(const uint32_t *p)
{
    return __ldg(p);
}

// streaming 128-bit load: read-only path, do not allocate in L1 (frames are read exactly once)
__device__ __forceinline__ uint4 ld_stream_u4(const void *p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ uint2 ld_stream_u2(const void *p)
{
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}

// slot of pixel q (which must be foreground) of one frame.  map/wordpre point at the frame's arrays.
__device__ __forceinline__ uint32_t slot_of(const uint32_t *__restrict__ map, const uint16_t *__restrict__ wordpre,
                                            uint32_t q)
{
    const uint32_t w = q >> 5;
    return (q & ~(uint32_t)(TILE_PX - 1)) + wordpre[w] + __popc(map[w] & ((1u << (q & 31)) - 1u));
}

// slot of the first pixel of word w (whether or not it is set)
__device__ __forceinline__ uint32_t word_slot_base(const uint16_t *__restrict__ wordpre, uint32_t w)
{
    return ((w >> TILE_WORDS_LOG2) << TILE_LOG2) + wordpre[w];
}

// ---- bulk async copies (cp.async.bulk, the 1-D form of TMA) completing on a shared-memory mbarrier ---------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    // try_wait suspends the warp in hardware; the bound turns a programming error into a trap, not a hang
    for (uint32_t spin = 0; !mbar_try_wait(bar, parity); spin++)
        if (spin > (1u << 26)) __trap();
}

