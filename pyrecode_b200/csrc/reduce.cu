// reduce.cu -- streaming reduction kernels of the write path.
//
//   k_reduce_tiles   fused dark-threshold compare + binary-map bit packing + per-tile compaction of the
//                    foreground values (reference: recode_writer.py:437 compare, :440 gather/subtract,
//                    :456 + :622-634 _pack_binary_frame).  One CTA per (frame, tile of 8192 pixels); the
//                    frame is read exactly once with 128-bit streaming loads.  HBM-bound:
//                    algorithmic bytes per tile = 8192 * itemsize (frame) [+ the same for the threshold
//                    tile, which stays L2-resident across the frames of a batch].
//   k_scan_tiles     per-frame exclusive scan of the tile counts (tiny).
//   k_bitpack        variable-bit-depth packing of tile-compacted values into the LSB-first bit stream
//                    (reference: recode_writer.py:637-652 _bit_pack == reader.h:105-140); owner-computes
//                    per 32-bit output word, no atomics.
#include "common.cuh"
#include "kernels.cuh"
#include <stdlib.h>

// ---- pixel-type helpers --------------------------------------------------------------------
// 8 pixels = W 32-bit words.  fg_words: d = max(f, t) - t per lane (= f - t where f > t, else 0; no borrow
// crosses a lane because max >= t), so the L1 value and the foreground test come out of the same two
// instructions; the mask bit of a lane is min(d, 1).
template <typename T> struct Px;

template <> struct Px<uint16_t> {
    static constexpr int W = 4;   // 32-bit words per group of 8 pixels
    static __device__ __forceinline__ void load_stream(const uint16_t *p, uint32_t (&w)[4])
    {
        uint4 v = ld_stream_u4(p);
        w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
    }
    static __device__ __forceinline__ void load_cached(const uint16_t *p, uint32_t (&w)[4])
    {
        uint4 v = __ldg(reinterpret_cast<const uint4 *>(p));
        w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
    }
    static __device__ __forceinline__ void load_shared(const uint16_t *p, uint32_t (&w)[4])
    {
        const uint4 v = *reinterpret_cast<const uint4 *>(p);
        w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
    }
    static __device__ __forceinline__ void set(uint32_t (&w)[4], int k, uint32_t v)
    {
        w[k >> 1] |= v << ((k & 1) * 16);
    }
    // d = frame - threshold where frame > threshold else 0; returns the 8-bit foreground mask
    static __device__ __forceinline__ uint32_t fg_words(const uint32_t (&f)[4], const uint32_t (&t)[4], uint32_t (&d)[4])
    {
        uint32_t e[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            d[i] = __vmaxu2(f[i], t[i]) - t[i];
            e[i] = __vminu2(d[i], 0x00010001u);          // bit 0 / bit 16 = lane is foreground
        }
        // even pixels at bits 0..6, odd at 16..22 (disjoint bits: the multiply-adds are ORs on the FMA pipe)
        const uint32_t c = e[0] + 4u * e[1] + 16u * e[2] + 64u * e[3];
        return (c | (c >> 15)) & 0xffu;
    }
    static __device__ __forceinline__ void store_raw(uint16_t *dst, const uint32_t (&w)[4])
    {
        *reinterpret_cast<uint4 *>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
    }
};

template <> struct Px<uint8_t> {
    static constexpr int W = 2;
    static __device__ __forceinline__ void load_stream(const uint8_t *p, uint32_t (&w)[2])
    {
        uint2 v = ld_stream_u2(p);
        w[0] = v.x; w[1] = v.y;
    }
    static __device__ __forceinline__ void load_cached(const uint8_t *p, uint32_t (&w)[2])
    {
        uint2 v = __ldg(reinterpret_cast<const uint2 *>(p));
        w[0] = v.x; w[1] = v.y;
    }
    static __device__ __forceinline__ void load_shared(const uint8_t *p, uint32_t (&w)[2])
    {
        const uint2 v = *reinterpret_cast<const uint2 *>(p);
        w[0] = v.x; w[1] = v.y;
    }
    static __device__ __forceinline__ void set(uint32_t (&w)[2], int k, uint32_t v)
    {
        w[k >> 2] |= v << ((k & 3) * 8);
    }
    static __device__ __forceinline__ uint32_t fg_words(const uint32_t (&f)[2], const uint32_t (&t)[2], uint32_t (&d)[2])
    {
        uint32_t m = 0;
#pragma unroll
        for (int i = 0; i < 2; i++) {
            d[i] = __vmaxu4(f[i], t[i]) - t[i];
            uint32_t r = __vminu4(d[i], 0x01010101u);    // bit 0 of each byte
            r = (r | (r >> 7) | (r >> 14) | (r >> 21)) & 0xfu;
            m |= r << (4 * i);
        }
        return m;
    }
    static __device__ __forceinline__ void store_raw(uint8_t *dst, const uint32_t (&w)[2])
    {
        *reinterpret_cast<uint2 *>(dst) = make_uint2(w[0], w[1]);
    }
};

// ---- K1 ------------------------------------------------------------------------------------
// One CTA (256 threads = 8 warps) per (frame, tile of 32768 pixels), processed as 4 sub-tiles of 8192
// pixels.  Within a sub-tile warp w owns the 1024 consecutive pixels [1024 w, 1024 w + 1024): four
// 128-bit loads per lane (512 contiguous bytes per warp-level load), and lane l then owns map word l of
// those 32 words.  Mask bytes and raw values are staged in warp-private shared memory, so the only
// block-level barrier per sub-tile is the one for the cross-warp prefix of the foreground counts.
//
// VALMODE: 0 = no value stream (L3), 1 = frame - thr as T (L1), 2 = (raw frame value << 16) | pixel index
//          within the tile, as uint32 (L2 / L4)
// This kernel only streams: the latency-bound puddle labelling of L2 / L4 runs afterwards on the compact
// per-tile data (k_ccl_tiles, ccl.cu).
constexpr int SUB_PX = 8192;
constexpr int SUB_WORDS = SUB_PX / 32;            // 256 = threads per CTA
constexpr int NSUB = TILE_PX / SUB_PX;            // 4

// position (in 8-pixel granules) of granule g inside the sub-tile's raw staging: the XOR spreads the word
// owners' 16-bit reads (stride 64 bytes) over the banks; 128-bit writes stay conflict-free
__device__ __forceinline__ uint32_t raw_pos(uint32_t g) { return g ^ ((g >> 3) & 7u); }

template <typename T, int VALMODE>
__global__ void __launch_bounds__(256, 5)
k_reduce_tiles(const T *__restrict__ frames, const T *__restrict__ thr, size_t P, int NT, size_t MS,
               uint32_t *__restrict__ maps, uint32_t *__restrict__ tilecnt, uint16_t *__restrict__ wordpre,
               void *__restrict__ vals_out, int vec_ok)
{
    constexpr int W = Px<T>::W;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int f = blockIdx.x;                 // frame fastest: CTAs running together share the threshold tile
    const int tile = blockIdx.y;
    const size_t base = (size_t)tile * TILE_PX;
    const size_t left = P - base;
    const int npx = left < (size_t)TILE_PX ? (int)left : TILE_PX;
    const T *fr = frames + (size_t)f * P + base;
    const T *th = thr + base;
    const size_t sbase = (size_t)f * ((size_t)NT * TILE_PX) + base;      // first slot of this tile

    __shared__ __align__(16) uint32_t s_mask[TILE_WORDS];
    __shared__ __align__(16) uint16_t s_wpre[TILE_WORDS];
    __shared__ __align__(16) T s_raw[VALMODE ? SUB_PX : 8];
    __shared__ __align__(16) uint32_t s_wsum[2][8];

    uint32_t run = 0;                         // foreground pixels of the tile before the current sub-tile
#pragma unroll 1
    for (int sub = 0; sub < NSUB; sub++) {
        const int sub_px0 = sub * SUB_PX;
        const int wpx0 = sub_px0 + warp * 1024;                 // first pixel of this warp's region
        uint32_t fw[4][W], tw[4][W];
        const bool fast = vec_ok && sub_px0 + SUB_PX <= npx;
        if (fast) {
#pragma unroll
            for (int j = 0; j < 4; j++) Px<T>::load_stream(fr + wpx0 + j * 256 + lane * 8, fw[j]);
#pragma unroll
            for (int j = 0; j < 4; j++) Px<T>::load_cached(th + wpx0 + j * 256 + lane * 8, tw[j]);
            // next sub-tile's threshold lines into L1 (see k_reduce_tiles_bulk)
            if (vec_ok > 1 && sub_px0 + 2 * SUB_PX <= npx && lane * (128 / (int)sizeof(T)) < 1024)
                asm volatile("prefetch.global.L1 [%0];" ::"l"(th + wpx0 + SUB_PX + lane * (128 / (int)sizeof(T))));
        } else {
#pragma unroll
            for (int j = 0; j < 4; j++) {
#pragma unroll
                for (int i = 0; i < W; i++) { fw[j][i] = 0; tw[j][i] = 0; }
                const int q0 = wpx0 + j * 256 + lane * 8;
                if (q0 < npx) {
#pragma unroll
                    for (int k = 0; k < 8; k++) {
                        if (q0 + k < npx) {
                            Px<T>::set(fw[j], k, fr[q0 + k]);
                            Px<T>::set(tw[j], k, th[q0 + k]);
                        }
                    }
                }
            }
        }
        uint8_t *mb = reinterpret_cast<uint8_t *>(s_mask + sub * SUB_WORDS + warp * 32);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            uint32_t d[W];
            const uint32_t mj = Px<T>::fg_words(fw[j], tw[j], d);
            mb[j * 32 + lane] = (uint8_t)mj;
            if (VALMODE) {
                const uint32_t g = (uint32_t)(warp * 128 + j * 32 + lane);
                Px<T>::store_raw(s_raw + raw_pos(g) * 8, VALMODE == 1 ? d : fw[j]);
            }
        }
        __syncwarp();
        const uint32_t word = s_mask[sub * SUB_WORDS + t];      // pixels [32 t, 32 t + 32) of the sub-tile
        const uint32_t pc = __popc(word);
        const uint32_t incl = warp_incl_scan(pc);
        if (lane == 31) s_wsum[sub & 1][warp] = incl;
        __syncthreads();
        uint32_t before = 0, total = 0;
        {
            const uint4 a = *reinterpret_cast<const uint4 *>(&s_wsum[sub & 1][0]);
            const uint4 b = *reinterpret_cast<const uint4 *>(&s_wsum[sub & 1][4]);
            const uint32_t ws[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (i < warp) before += ws[i];
                total += ws[i];
            }
        }
        uint32_t rank = run + before + incl - pc;               // tile-local slot of this word's first pixel
        s_wpre[sub * SUB_WORDS + t] = (uint16_t)rank;
        if (VALMODE) {
            uint32_t bits = word;
            while (bits) {
                const uint32_t k = __ffs(bits) - 1;
                bits &= bits - 1;
                const uint32_t q = (uint32_t)(t * 32) + k;       // pixel within the sub-tile
                const T v = s_raw[raw_pos(q >> 3) * 8 + (q & 7)];
                // L1: the value stream itself.  L2 / L4: value and tile-local pixel position in one word, the
                // input of the per-tile labelling (k_ccl_tiles)
                if (VALMODE == 1) reinterpret_cast<T *>(vals_out)[sbase + rank] = v;
                else reinterpret_cast<uint32_t *>(vals_out)[sbase + rank] = ((uint32_t)v << 16) | (uint32_t)(sub * SUB_PX) | q;
                rank++;
            }
            __syncwarp();                                       // s_raw of this warp is rewritten next sub-tile
        }
        run += total;
    }
    __syncthreads();

    // whole-tile outputs: map words, per-word prefixes, tile count
    {
        const uint4 mw = *reinterpret_cast<const uint4 *>(&s_mask[t * 4]);
        *reinterpret_cast<uint4 *>(&maps[(size_t)f * MS + (size_t)tile * TILE_WORDS + t * 4]) = mw;
        const uint2 wp = *reinterpret_cast<const uint2 *>(&s_wpre[t * 4]);
        *reinterpret_cast<uint2 *>(&wordpre[(size_t)f * MS + (size_t)tile * TILE_WORDS + t * 4]) = wp;
        if (t == 0) tilecnt[(size_t)f * NT + tile] = run;
    }
}

// ---- K1, bulk-copy variant (full tiles, 16-byte aligned frames) ---------------------------------------
// Same outputs as k_reduce_tiles.  The whole 32768-pixel frame tile is brought into shared memory by the bulk
// async-copy engine (cp.async.bulk, the 1-D form of TMA): lane 0 of every warp issues the four 2 KiB copies of
// own pixel regions into a warp-private ring of BULK_STAGES stages, each stage signalling its own mbarrier, and
// refills a stage as soon as the warp has consumed it -- no registers are spent on covering DRAM latency, the
// raw values need no second staging copy (the compaction reads them where the copy engine put them), and no
// load instruction is issued for the frame at all.
constexpr int BULK_STAGES = 2;                 // sub-tiles in flight per warp (ring)
template <typename T>
constexpr size_t bulk_smem_bytes(bool thr_bulk) { return (size_t)BULK_STAGES * SUB_PX * sizeof(T) * (thr_bulk ? 2 : 1); }

// THRB: the threshold tile travels through the ring as well (second region of every stage) instead of being
// loaded just in time with cached 128-bit loads (which wait for L2 once per sub-tile).
template <typename T, int VALMODE, bool THRB>
__global__ void __launch_bounds__(256)
k_reduce_tiles_bulk(const T *__restrict__ frames, const T *__restrict__ thr, size_t P, int NT, size_t MS,
                    uint32_t *__restrict__ maps, uint32_t *__restrict__ tilecnt, uint16_t *__restrict__ wordpre,
                    void *__restrict__ vals_out, int k1_prefetch)
{
    constexpr int W = Px<T>::W;
    constexpr uint32_t REGION_BYTES = 1024 * sizeof(T);         // one warp's pixels of one sub-tile
    extern __shared__ __align__(128) uint8_t s_dyn[];
    T *s_ring = reinterpret_cast<T *>(s_dyn);                   // [BULK_STAGES][8 warps][1024 pixels] (x 2: frame, thr)
    constexpr int RS = THRB ? 2048 : 1024;                      // ring pixels per (stage, warp)
    __shared__ __align__(16) uint32_t s_mask[TILE_WORDS];
    __shared__ __align__(16) uint16_t s_wpre[TILE_WORDS];
    __shared__ __align__(16) uint32_t s_wsum[2][8];
    __shared__ __align__(8) uint64_t s_bar[BULK_STAGES][8];

    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int f = blockIdx.x;                 // frame fastest: CTAs running together share the threshold tile
    const int tile = blockIdx.y;
    const size_t base = (size_t)tile * TILE_PX;
    const T *fr = frames + (size_t)f * P + base;
    const T *th = thr + base;
    const size_t sbase = (size_t)f * ((size_t)NT * TILE_PX) + base;

    if (t < BULK_STAGES * 8) mbar_init(smem_u32(&s_bar[t >> 3][t & 7]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    // the ring is warp-private: warp w owns [stage][w], its lane 0 is the producer, all its lanes consume
    auto issue = [&](int sub) {
        const int stg = sub % BULK_STAGES;
        const uint32_t bar = smem_u32(&s_bar[stg][warp]);
        mbar_expect_tx(bar, THRB ? 2 * REGION_BYTES : REGION_BYTES);
        bulk_g2s(smem_u32(s_ring + (stg * 8 + warp) * RS), fr + sub * SUB_PX + warp * 1024, REGION_BYTES, bar);
        if (THRB)
            bulk_g2s(smem_u32(s_ring + (stg * 8 + warp) * RS + 1024), th + sub * SUB_PX + warp * 1024, REGION_BYTES, bar);
    };
    if (lane == 0) {
#pragma unroll
        for (int sub = 0; sub < BULK_STAGES; sub++) issue(sub);
    }

    uint32_t run = 0;
#pragma unroll 2
    for (int sub = 0; sub < NSUB; sub++) {
        const int wpx0 = sub * SUB_PX + warp * 1024;
        uint32_t fw[4][W], tw[4][W];
        if (!THRB) {
#pragma unroll
            for (int j = 0; j < 4; j++) Px<T>::load_cached(th + wpx0 + j * 256 + lane * 8, tw[j]);
            // the threshold pixels of the NEXT sub-tile: one 128-byte line per lane into L1, so that the loads above
            // find them there one iteration later instead of waiting for L2
            if ((k1_prefetch & 1) && sub + 1 < NSUB && lane * (128 / (int)sizeof(T)) < 1024)
                asm volatile("prefetch.global.L1 [%0];" ::"l"(th + wpx0 + SUB_PX + lane * (128 / (int)sizeof(T))));
        }
        const int stg = sub % BULK_STAGES;
        T *region = s_ring + (stg * 8 + warp) * RS;              // this warp's 1024 pixels of the sub-tile
        mbar_wait(smem_u32(&s_bar[stg][warp]), (uint32_t)(sub / BULK_STAGES) & 1u);
#pragma unroll
        for (int j = 0; j < 4; j++) Px<T>::load_shared(region + j * 256 + lane * 8, fw[j]);
        if (THRB) {
#pragma unroll
            for (int j = 0; j < 4; j++) Px<T>::load_shared(region + 1024 + j * 256 + lane * 8, tw[j]);
        }
        uint8_t *mb = reinterpret_cast<uint8_t *>(s_mask + sub * SUB_WORDS + warp * 32);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            uint32_t d[W];
            const uint32_t mj = Px<T>::fg_words(fw[j], tw[j], d);
            mb[j * 32 + lane] = (uint8_t)mj;
            if (VALMODE == 1) Px<T>::store_raw(region + j * 256 + lane * 8, d);           // frame - thr, in place
        }
        __syncwarp();
        const uint32_t word = s_mask[sub * SUB_WORDS + t];
        const uint32_t pc = __popc(word);
        // inclusive prefix of the 32 counts.  Bit-sliced with ballots (k1_prefetch bit 1, the default; RC_K1_BALLOT_SCAN=0
        // selects the shuffle scan; L1 +1.9 %, L2 +0.4 % frames/s): the three low bits of the counts
        // cost three INDEPENDENT vote + popc pairs (the shuffle scan is five dependent round trips); counts of 8 or more in
        // any lane (rare at a few per cent occupancy) add the three high bits.
        uint32_t incl;
        if (k1_prefetch & 2) {
            const uint32_t le = 0xffffffffu >> (31 - lane);
            incl = __popc(__ballot_sync(0xffffffffu, pc & 1u) & le) + (__popc(__ballot_sync(0xffffffffu, pc & 2u) & le) << 1) +
                   (__popc(__ballot_sync(0xffffffffu, pc & 4u) & le) << 2);
            if (__any_sync(0xffffffffu, pc >> 3))
                incl += (__popc(__ballot_sync(0xffffffffu, pc & 8u) & le) << 3) +
                        (__popc(__ballot_sync(0xffffffffu, pc & 16u) & le) << 4) +
                        (__popc(__ballot_sync(0xffffffffu, pc & 32u) & le) << 5);
        } else {
            incl = warp_incl_scan(pc);
        }
        if (lane == 31) s_wsum[sub & 1][warp] = incl;
        __syncthreads();
        // foreground pixels of the warps before this one / of all eight: two warp-wide integer reductions (REDUX) over
        // the eight sums instead of eight selects and adds per thread
        const uint32_t wsv = s_wsum[sub & 1][lane & 7];
        const uint32_t before = __reduce_add_sync(0xffffffffu, lane < warp ? wsv : 0u);
        const uint32_t total = __reduce_add_sync(0xffffffffu, lane < 8 ? wsv : 0u);
        const uint32_t rank = run + before + incl - pc;
        s_wpre[sub * SUB_WORDS + t] = (uint16_t)rank;
        if (VALMODE) {
            const T *src = region + lane * 32;                  // this thread's word = 32 consecutive pixels
            uint32_t bits = word;
            if (VALMODE == 1) {
                T *o = reinterpret_cast<T *>(vals_out) + sbase + rank;
                while (bits) {
                    const uint32_t k = __ffs(bits) - 1;
                    bits &= bits - 1;
                    *o++ = src[k];
                }
            } else {
                // one 64-bit multiply-add per store from a 32-bit index (the compiler otherwise carries a 64-bit
                // pointer through the loop: four instructions per pixel)
                uint32_t *const ob = reinterpret_cast<uint32_t *>(vals_out) + sbase;
                uint32_t idx = rank;
                const uint32_t posbase = (uint32_t)(sub * SUB_PX + t * 32);
                while (bits) {
                    const uint32_t k = __ffs(bits) - 1;
                    bits &= bits - 1;
                    asm volatile("" : "+r"(idx));
                    ob[idx] = ((uint32_t)src[k] << 16) | posbase | k;
                    idx++;
                }
            }
        }
        // the warp is done with this stage: refill it with the sub-tile BULK_STAGES ahead
        __syncwarp();
        if (lane == 0 && sub + BULK_STAGES < NSUB) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic accesses before the async write
            issue(sub + BULK_STAGES);
        }
        run += total;
    }
    __syncthreads();
    {
        const uint4 mw = *reinterpret_cast<const uint4 *>(&s_mask[t * 4]);
        *reinterpret_cast<uint4 *>(&maps[(size_t)f * MS + (size_t)tile * TILE_WORDS + t * 4]) = mw;
        const uint2 wp = *reinterpret_cast<const uint2 *>(&s_wpre[t * 4]);
        *reinterpret_cast<uint2 *>(&wordpre[(size_t)f * MS + (size_t)tile * TILE_WORDS + t * 4]) = wp;
        if (t == 0) tilecnt[(size_t)f * NT + tile] = run;
    }
}

template <typename T>
static int launch_reduce_tiles_t(rc_ctx *ctx, const Geom &g, int valmode, const void *frames, const void *thr, int F,
                                 uint32_t *maps, uint32_t *tilecnt, uint16_t *wordpre, void *vals, cudaStream_t st)
{
    const int vec_ok = ((g.P * sizeof(T)) % 16 == 0) && ((uintptr_t)frames % 16 == 0) && ((uintptr_t)thr % 16 == 0);
    dim3 grid(F, g.NT), block(256);
    // full tiles + aligned frames: the bulk-copy variant (RC_K1_GENERIC=1 in the environment forces the other).
    // L1 writes frame - thr back into the staged tile; alone it is as fast as the register path (0.224 vs 0.228 ms
    // per 32 frames), inside the pipeline it leaves more of every SM to the other kernels (81.0 vs 76.9 k frames/s).
    static const bool force_generic = getenv("RC_K1_GENERIC") != nullptr;
    static const bool l1_bulk = getenv("RC_K1_L1_BULK") ? atoi(getenv("RC_K1_L1_BULK")) != 0 : true;
    const bool bulk = vec_ok && g.P % TILE_PX == 0 && !force_generic && (valmode != 1 || l1_bulk);
    // unused dynamic shared memory requested on top of the ring: fewer K1 CTAs per SM (5 without), i.e. room for the
    // labelling and encoder CTAs of the batches in flight on the other streams
    static const size_t k1_pad = getenv("RC_K1_SMEM_PAD") ? (size_t)atoi(getenv("RC_K1_SMEM_PAD")) : 0;
    static const bool thr_bulk = getenv("RC_K1_THR_BULK") ? atoi(getenv("RC_K1_THR_BULK")) != 0 : false;
#define RC_K1B(VM, TB)                                                                                     \
    {                                                                                                      \
        cudaFuncSetAttribute(k_reduce_tiles_bulk<T, VM, TB>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                             (int)(bulk_smem_bytes<T>(TB) + k1_pad));                                      \
        k_reduce_tiles_bulk<T, VM, TB><<<grid, block, bulk_smem_bytes<T>(TB) + k1_pad, st>>>(              \
            (const T *)frames, (const T *)thr, g.P, g.NT, g.MS, maps, tilecnt, wordpre, vals, k1_prefetch); \
    }
    static const int k1_prefetch = (getenv("RC_K1_PREFETCH") ? (atoi(getenv("RC_K1_PREFETCH")) & 1) : 1) |
                                   ((getenv("RC_K1_BALLOT_SCAN") ? atoi(getenv("RC_K1_BALLOT_SCAN")) != 0 : true) ? 2 : 0);
#define RC_K1(VM)                                                                                          \
    if (bulk) {                                                                                            \
        if (thr_bulk) RC_K1B(VM, true) else RC_K1B(VM, false)                                              \
    } else {                                                                                               \
        k_reduce_tiles<T, VM><<<grid, block, 0, st>>>((const T *)frames, (const T *)thr, g.P, g.NT, g.MS,  \
                                                      maps, tilecnt, wordpre, vals,                        \
                                                      vec_ok ? ((k1_prefetch & 1) ? 2 : 1) : 0);           \
    }
    if (valmode == 0) { RC_K1(0) }
    else if (valmode == 1) { RC_K1(1) }
    else { RC_K1(2) }
#undef RC_K1
#undef RC_K1B
    RC_LAUNCH_CHECK(ctx, "k_reduce_tiles");
    return 0;
}

int launch_reduce_tiles(rc_ctx *ctx, const Geom &g, int itemsize, int valmode, const void *frames, const void *thr,
                        int F, uint32_t *maps, uint32_t *tilecnt, uint16_t *wordpre, void *vals, cudaStream_t st)
{
    if (F <= 0) return 0;
    if (itemsize == 2)
        return launch_reduce_tiles_t<uint16_t>(ctx, g, valmode, frames, thr, F, maps, tilecnt, wordpre, vals, st);
    return launch_reduce_tiles_t<uint8_t>(ctx, g, valmode, frames, thr, F, maps, tilecnt, wordpre, vals, st);
}

// ---- map-only tile counts (read side: a map came out of inflate) ------------------------------
// Computes tilecnt / wordpre from existing maps so the unpack kernels can rank pixels.
__global__ void __launch_bounds__(256)
k_map_counts(const uint32_t *__restrict__ maps, size_t MS, int NT, uint32_t *__restrict__ tilecnt,
             uint16_t *__restrict__ wordpre)
{
    __shared__ uint32_t s_warp[9];
    const int t = threadIdx.x, f = blockIdx.x, tile = blockIdx.y;
    const size_t o = (size_t)f * MS + (size_t)tile * TILE_WORDS + t * 4;      // thread t: words 4t .. 4t+3
    const uint4 w = *reinterpret_cast<const uint4 *>(maps + o);
    const uint32_t p0 = __popc(w.x), p1 = __popc(w.y), p2 = __popc(w.z), p3 = __popc(w.w);
    uint32_t total;
    const uint32_t e = block_excl_scan<8>(p0 + p1 + p2 + p3, s_warp, &total);
    ushort4 o4;
    o4.x = (uint16_t)e; o4.y = (uint16_t)(e + p0); o4.z = (uint16_t)(e + p0 + p1); o4.w = (uint16_t)(e + p0 + p1 + p2);
    *reinterpret_cast<ushort4 *>(wordpre + o) = o4;
    if (t == 0) tilecnt[(size_t)f * NT + tile] = total;
}

int launch_map_counts(rc_ctx *ctx, const Geom &g, const uint32_t *maps, int F, uint32_t *tilecnt, uint16_t *wordpre,
                      cudaStream_t st)
{
    if (F <= 0) return 0;
    k_map_counts<<<dim3(F, g.NT), 256, 0, st>>>(maps, g.MS, g.NT, tilecnt, wordpre);
    RC_LAUNCH_CHECK(ctx, "k_map_counts");
    return 0;
}

// ---- K2: per-frame exclusive scan of tile counts ----------------------------------------------
// counts[f] = n; packed_bytes[f] = ceil(n*b/8) when b > 0 (either may be null).
__global__ void __launch_bounds__(256)
k_scan_tiles(const uint32_t *__restrict__ tilecnt, int NT, uint32_t *__restrict__ tilepre,
             uint32_t *__restrict__ counts, uint32_t *__restrict__ packed_bytes, int b)
{
    __shared__ uint32_t s_warp[9];
    const int f = blockIdx.x, t = threadIdx.x;
    const uint32_t *c = tilecnt + (size_t)f * NT;
    uint32_t *p = tilepre + (size_t)f * (NT + 1);
    uint32_t carry = 0;
    for (int i0 = 0; i0 < NT; i0 += 256) {
        const int i = i0 + t;
        const uint32_t v = i < NT ? c[i] : 0;
        uint32_t total;
        const uint32_t e = block_excl_scan<8>(v, s_warp, &total);
        if (i < NT) p[i] = carry + e;
        carry += total;
        __syncthreads();
    }
    if (t == 0) {
        p[NT] = carry;
        if (counts) counts[f] = carry;
        if (packed_bytes) packed_bytes[f] = (uint32_t)(((uint64_t)carry * (uint32_t)b + 7) / 8);
    }
}

int launch_scan_tiles(rc_ctx *ctx, const Geom &g, const uint32_t *tilecnt, int F, uint32_t *tilepre,
                      uint32_t *counts, uint32_t *packed_bytes, int b, cudaStream_t st)
{
    if (F <= 0) return 0;
    k_scan_tiles<<<F, 256, 0, st>>>(tilecnt, g.NT, tilepre, counts, packed_bytes, b);
    RC_LAUNCH_CHECK(ctx, "k_scan_tiles");
    return 0;
}

// ---- K3: bit packing -------------------------------------------------------------------------
// Output word w of frame f gathers every value whose bit range [j*b, (j+1)*b) overlaps [32w, 32w+32).
// The tile prefix table of the frame is staged in shared memory (when it fits) so that the per-word binary
// search runs at shared-memory latency.
constexpr int BP_SMEM_TILES = 4096;

template <typename T>
__global__ void __launch_bounds__(256)
k_bitpack(const T *__restrict__ vals, const uint32_t *__restrict__ tilepre, int NT, int b,
          uint8_t *__restrict__ packed, size_t packed_stride)
{
    __shared__ uint32_t s_pre[BP_SMEM_TILES + 1];
    const int f = blockIdx.y;
    const uint32_t *gpre = tilepre + (size_t)f * (NT + 1);
    const bool staged = NT <= BP_SMEM_TILES;
    if (staged) {
        for (int i = threadIdx.x; i <= NT; i += 256) s_pre[i] = gpre[i];
        __syncthreads();
    }
    const uint32_t *pre = staged ? s_pre : gpre;
    const T *v = vals + (size_t)f * ((size_t)NT * TILE_PX);
    uint32_t *out = reinterpret_cast<uint32_t *>(packed + (size_t)f * packed_stride);
    const uint64_t n = pre[NT];
    const uint64_t nbits = n * (uint64_t)b;
    const uint64_t nwords = (nbits + 31) / 32;
    const uint32_t vmask = b >= 32 ? 0xffffffffu : ((1u << b) - 1u);
    if (nbits < 0xffffffe0ull) {
        // 32-bit indices (any frame up to 2^27 pixels): no 64-bit divisions in the loop
        const uint32_t n32 = (uint32_t)n, nw32 = (uint32_t)nwords, ub = (uint32_t)b;
        for (uint32_t w = blockIdx.x * 256 + threadIdx.x; w < nw32; w += gridDim.x * 256) {
            const uint32_t bit0 = w * 32;
            uint32_t j = bit0 / ub;
            uint32_t jl = (bit0 + 31) / ub;
            if (jl >= n32) jl = n32 - 1;
            int lo = 0, hi = NT;                    // invariant: pre[lo] <= j < pre[hi]
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (pre[mid] <= j) lo = mid; else hi = mid;
            }
            int tt = lo;
            uint32_t acc = 0;
            int sh = (int)(j * ub) - (int)bit0;     // bit position of value j relative to the word: (-b, 32)
            for (; j <= jl; j++, sh += b) {
                while (pre[tt + 1] <= j) tt++;
                const uint32_t val = (uint32_t)v[(size_t)tt * TILE_PX + (j - pre[tt])] & vmask;
                acc |= sh >= 0 ? (val << sh) : (val >> (-sh));
            }
            out[w] = acc;
        }
        return;
    }
    for (uint64_t w = (uint64_t)blockIdx.x * 256 + threadIdx.x; w < nwords; w += (uint64_t)gridDim.x * 256) {
        const uint64_t bit0 = w * 32;
        uint64_t j = bit0 / (uint32_t)b;
        uint64_t jl = (bit0 + 31) / (uint32_t)b;
        if (jl >= n) jl = n - 1;
        // tile holding rank j: largest tt with pre[tt] <= j
        int lo = 0, hi = NT;                    // invariant: pre[lo] <= j < pre[hi]
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (pre[mid] <= j) lo = mid; else hi = mid;
        }
        int tt = lo;
        uint32_t acc = 0;
        for (; j <= jl; j++) {
            while (pre[tt + 1] <= j) tt++;
            const uint32_t val = (uint32_t)v[(size_t)tt * TILE_PX + (size_t)(j - pre[tt])] & vmask;
            const int64_t sh = (int64_t)(j * (uint32_t)b) - (int64_t)bit0;
            acc |= sh >= 0 ? (val << sh) : (val >> (-sh));
        }
        out[w] = acc;
    }
}

// Same result, organised by tile: one warp per (tile, frame) packs the output words whose FIRST bit belongs to one of
// the tile's values -- words [ceil(pre[tile] b / 32), ceil(pre[tile + 1] b / 32)) -- so every word is still written by
// exactly one thread (no atomics, no zero-initialised output), but nobody searches the prefix table: a word's
// values start in the warp's own tile and at most spill into the following ones.  (k_bitpack spent 350 instructions per
// output word, most of them in the binary search.)  For frames of fewer than 2^32 value bits.
template <typename T>
__global__ void __launch_bounds__(256)
k_bitpack_tiles(const T *__restrict__ vals, const uint32_t *__restrict__ tilepre, int NT, int n_tiles_total, int b,
                uint8_t *__restrict__ packed, size_t packed_stride)
{
    const int gt = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (gt >= n_tiles_total) return;
    const int lane = threadIdx.x & 31;
    const int f = gt / NT, tile = gt - f * NT;
    const uint32_t *pre = tilepre + (size_t)f * (NT + 1);
    const uint32_t j0 = pre[tile], j1 = pre[tile + 1];
    if (j1 == j0) return;
    const uint32_t n = pre[NT], ub = (uint32_t)b;
    const T *v = vals + (size_t)f * ((size_t)NT * TILE_PX);
    uint32_t *out = reinterpret_cast<uint32_t *>(packed + (size_t)f * packed_stride);
    const uint32_t vmask = b >= 32 ? 0xffffffffu : ((1u << b) - 1u);
    const uint32_t w_lo = (j0 * ub + 31u) >> 5, w_hi = (j1 * ub + 31u) >> 5;
    // floor(x / b) for x < 2^32 without a division: umulhi by ceil(2^32 / b) is exact or one too large
    const uint32_t magic = ub > 1 ? (uint32_t)((0x100000000ull + ub - 1) / ub) : 0u;
    const uint32_t max_cnt = (32u + ub - 1u) / ub + 1u;           // values that can start before a word ends
    for (uint32_t w = w_lo + lane; w < w_hi; w += 32) {
        const uint32_t bit0 = w << 5;
        uint32_t j = bit0;                                        // >= j0: the word's first bit is one of this tile's
        if (ub > 1) {
            j = __umulhi(bit0, magic);
            if (j * ub > bit0) j--;
        }
        uint32_t acc = 0;
        int sh = (int)(j * ub) - (int)bit0;                       // bit position of value j relative to the word: (-b, 32)
        if (j + max_cnt <= j1) {
            // every value that starts before the word ends is this tile's (all words but the tile's last ones):
            // consecutive loads, no prefix lookups
            const T *vt = v + (size_t)tile * TILE_PX + (j - j0);
            for (; sh < 32; sh += b) {
                const uint32_t val = (uint32_t)(*vt++) & vmask;
                acc |= sh >= 0 ? (val << sh) : (val >> (-sh));
            }
        } else {
            int tt = tile;
            for (; j < n && sh < 32; j++, sh += b) {
                while (pre[tt + 1] <= j) tt++;
                const uint32_t val = (uint32_t)v[(size_t)tt * TILE_PX + (j - pre[tt])] & vmask;
                acc |= sh >= 0 ? (val << sh) : (val >> (-sh));
            }
        }
        out[w] = acc;
    }
}

int launch_bitpack(rc_ctx *ctx, const Geom &g, int val_itemsize, const void *vals, const uint32_t *tilepre, int F,
                   int b, uint8_t *packed, size_t packed_stride, cudaStream_t st)
{
    if (F <= 0) return 0;
    if ((uint64_t)g.P * (uint64_t)b < 0xffffffe0ull) {
        const int nt = F * g.NT;
        const unsigned blocks = (unsigned)((nt + 7) / 8);
        if (val_itemsize == 2)
            k_bitpack_tiles<uint16_t><<<blocks, 256, 0, st>>>((const uint16_t *)vals, tilepre, g.NT, nt, b, packed, packed_stride);
        else
            k_bitpack_tiles<uint8_t><<<blocks, 256, 0, st>>>((const uint8_t *)vals, tilepre, g.NT, nt, b, packed, packed_stride);
        RC_LAUNCH_CHECK(ctx, "k_bitpack_tiles");
        return 0;
    }
    dim3 grid(128, F);
    if (val_itemsize == 2)
        k_bitpack<uint16_t><<<grid, 256, 0, st>>>((const uint16_t *)vals, tilepre, g.NT, b, packed, packed_stride);
    else
        k_bitpack<uint8_t><<<grid, 256, 0, st>>>((const uint8_t *)vals, tilepre, g.NT, b, packed, packed_stride);
    RC_LAUNCH_CHECK(ctx, "k_bitpack");
    return 0;
}

// ---- plain-array packers (c_recode.Reader.bit_pack / bit_unpack replacements) --------------------
__global__ void k_bitpack_flat(const uint16_t *__restrict__ v, uint64_t n, int b, uint32_t *__restrict__ out)
{
    const uint64_t nwords = (n * (uint64_t)b + 31) / 32;
    const uint32_t vmask = (1u << b) - 1u;
    for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < nwords; w += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t bit0 = w * 32;
        uint64_t j = bit0 / (uint32_t)b, jl = (bit0 + 31) / (uint32_t)b;
        if (jl >= n) jl = n - 1;
        uint32_t acc = 0;
        for (; j <= jl; j++) {
            const uint32_t val = (uint32_t)v[j] & vmask;
            const int64_t sh = (int64_t)(j * (uint32_t)b) - (int64_t)bit0;
            acc |= sh >= 0 ? (val << sh) : (val >> (-sh));
        }
        out[w] = acc;
    }
}

__global__ void k_bitunpack_flat(const uint8_t *__restrict__ packed, uint64_t n, int b, uint64_t *__restrict__ out)
{
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t bit = j * (uint32_t)b;
        const uint64_t by = bit >> 3;
        const int sh = (int)(bit & 7);
        // b <= 16 and sh <= 7 -> at most 3 bytes
        uint32_t x = packed[by];
        if (sh + b > 8) x |= (uint32_t)packed[by + 1] << 8;
        if (sh + b > 16) x |= (uint32_t)packed[by + 2] << 16;
        out[j] = (x >> sh) & ((1u << b) - 1u);
    }
}

int launch_bitpack_flat(rc_ctx *ctx, int b, const uint16_t *vals, uint64_t n, uint8_t *packed, cudaStream_t st)
{
    if (n == 0) return 0;
    const uint64_t nwords = (n * (uint64_t)b + 31) / 32;
    const int blocks = (int)((nwords + 255) / 256 > 4096 ? 4096 : (nwords + 255) / 256);
    k_bitpack_flat<<<blocks, 256, 0, st>>>(vals, n, b, reinterpret_cast<uint32_t *>(packed));
    RC_LAUNCH_CHECK(ctx, "k_bitpack_flat");
    return 0;
}

int launch_bitunpack_flat(rc_ctx *ctx, int b, const uint8_t *packed, uint64_t n, uint64_t *out, cudaStream_t st)
{
    if (n == 0) return 0;
    const int blocks = (int)((n + 255) / 256 > 4096 ? 4096 : (n + 255) / 256);
    k_bitunpack_flat<<<blocks, 256, 0, st>>>(packed, n, b, out);
    RC_LAUNCH_CHECK(ctx, "k_bitunpack_flat");
    return 0;
}

// ---- threshold frame -------------------------------------------------------------------------
template <typename T>
__global__ void k_make_threshold(const T *__restrict__ dark, T eps, T *__restrict__ thr, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        thr[i] = (T)(dark[i] + eps);
}

int launch_make_threshold(rc_ctx *ctx, int itemsize, const void *dark, uint64_t eps, void *thr, size_t n,
                          cudaStream_t st)
{
    const int blocks = (int)((n + 255) / 256 > 2048 ? 2048 : (n + 255) / 256);
    if (itemsize == 2)
        k_make_threshold<uint16_t><<<blocks, 256, 0, st>>>((const uint16_t *)dark, (uint16_t)eps, (uint16_t *)thr, n);
    else
        k_make_threshold<uint8_t><<<blocks, 256, 0, st>>>((const uint8_t *)dark, (uint8_t)eps, (uint8_t *)thr, n);
    RC_LAUNCH_CHECK(ctx, "k_make_threshold");
    return 0;
}
