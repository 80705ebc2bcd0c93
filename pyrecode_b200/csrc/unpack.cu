// unpack.cu -- read-side unpack kernels.
//
// Replaces c_recode.Reader.get_frame_sparse (pyrecode/pyrecode.cpp:95-119 -> c_extensions/reader.h:10-68) and the
// scipy COO -> dense step user code performs (recode_reader.py:464-471, tests/minimal_read_write_test.py:91),
// plus the live-view accumulation of examples/ReCoDe_Live_View_MT.ipynb cell 1.
//
// The reference walks all ny*nx pixels with a per-bit inner loop.  Here a foreground pixel's rank comes from
// popcounts (tile prefix + segment prefix + in-segment popc), its value is one funnel-shifted load at bit
// rank*b, and the dense frame is written with full 128-bit coalesced stores (one warp = one 256-pixel
// segment = 512 contiguous output bytes for uint16).  HBM-bound on the dense write: ny*nx*itemsize bytes/frame.
#include "common.cuh"
#include "kernels.cuh"

__device__ __forceinline__ uint32_t fetch_bits(const uint32_t *__restrict__ packed32, uint64_t bit, int b)
{
    const uint64_t w = bit >> 5;
    const uint32_t sh = (uint32_t)(bit & 31);
    const uint32_t lo = packed32[w];
    const uint32_t hi = (sh + b > 32) ? packed32[w + 1] : 0;
    return __funnelshift_r(lo, hi, sh) & ((1u << b) - 1u);
}

__global__ void __launch_bounds__(256)
k_unpack_sparse(const uint32_t *__restrict__ maps, size_t MS, const uint8_t *__restrict__ packed, size_t packed_stride,
                const uint16_t *__restrict__ wordpre_all, const uint32_t *__restrict__ tilepre_all, int NT, int nx,
                uint32_t MW, int level, int b, uint64_t *__restrict__ triples, size_t capacity)
{
    const int f = blockIdx.y;
    const uint32_t w = blockIdx.x * 256 + threadIdx.x;
    if (w >= MW) return;
    const uint32_t *map = maps + (size_t)f * MS;
    uint32_t bits = map[w];
    if (!bits) return;
    const uint32_t *tilepre = tilepre_all + (size_t)f * (NT + 1);
    const uint32_t *pk = reinterpret_cast<const uint32_t *>(packed + (size_t)f * packed_stride);
    // global rank of the first set bit of this word
    uint64_t rank = (uint64_t)tilepre[w >> TILE_WORDS_LOG2] + wordpre_all[(size_t)f * MS + w];
    uint64_t *out = triples + (size_t)f * capacity * 3;
    const uint32_t p0 = w << 5;
    while (bits) {
        const uint32_t k = __ffs(bits) - 1;
        bits &= bits - 1;
        const uint32_t p = p0 + k;
        if (rank < capacity) {
            const uint32_t r = p / (uint32_t)nx;
            out[rank * 3 + 0] = r;
            out[rank * 3 + 1] = p - r * (uint32_t)nx;
            out[rank * 3 + 2] = level == 1 ? fetch_bits(pk, rank * (uint32_t)b, b) : 1u;
        }
        rank++;
    }
}

// One warp per UD_SEGS consecutive 256-pixel segments; lane l owns pixels [8l, 8l+8) of each, so every store
// instruction of the warp writes 512 contiguous bytes (uint16).  The rank of a lane's first foreground pixel comes
// from the per-word prefix that k_map_counts left (tile prefix + word prefix + popc of the word's lower bytes): no
// scan, and the UD_SEGS map words of a lane are loaded up front, so their latencies overlap.
constexpr int UD_SEGS = 8;

template <typename T>
__global__ void __launch_bounds__(256)
k_unpack_dense(const uint32_t *__restrict__ maps, size_t MS, const uint8_t *__restrict__ packed, size_t packed_stride,
               const uint16_t *__restrict__ wordpre_all, const uint32_t *__restrict__ tilepre_all, int NT, size_t P,
               int level, int b, T *__restrict__ dense, uint32_t *__restrict__ sum, int vec_ok)
{
    const int f = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const uint32_t seg0 = (blockIdx.x * 8 + (threadIdx.x >> 5)) * UD_SEGS;
    const uint32_t nseg = (uint32_t)NT * SEGS_PER_TILE;
    if (seg0 >= nseg) return;
    const uint32_t *map = maps + (size_t)f * MS;
    const uint32_t *pk = reinterpret_cast<const uint32_t *>(packed + (size_t)f * packed_stride);
    uint32_t mw[UD_SEGS];
#pragma unroll
    for (int u = 0; u < UD_SEGS; u++) {
        const uint32_t seg = seg0 + u;
        mw[u] = seg < nseg ? map[(size_t)seg * SEG_WORDS + (lane >> 2)] : 0u;    // 4 lanes share a word
    }
    const uint32_t bsh = (uint32_t)(lane & 3) * 8u;
#pragma unroll
    for (int u = 0; u < UD_SEGS; u++) {
        const uint32_t seg = seg0 + u;
        const size_t p_base = (size_t)seg * SEG_PX + (size_t)lane * 8;
        if (seg >= nseg || (size_t)seg * SEG_PX >= P) break;
        const uint32_t m = (mw[u] >> bsh) & 0xffu;
        // 8 pixel values as two 64-bit halves of 16-bit slots (pixels 0..3, 4..7); levels 2..4: the map bit itself
        uint64_t qlo = 0, qhi = 0;
        if (level != 1) {
            const uint32_t a = m & 0xfu, c = m >> 4;
            qlo = (uint64_t)((a & 1u) | ((a & 2u) << 15)) | ((uint64_t)(((a >> 2) & 1u) | ((a & 8u) << 13)) << 32);
            qhi = (uint64_t)((c & 1u) | ((c & 2u) << 15)) | ((uint64_t)(((c >> 2) & 1u) | ((c & 8u) << 13)) << 32);
        }
        if (m && (level == 1 || sum)) {
            // one iteration per foreground pixel (a few per hundred pixels)
            uint64_t rank = 0;
            if (level == 1)
                rank = (uint64_t)tilepre_all[(size_t)f * (NT + 1) + (seg >> SEGS_PER_TILE_LOG2)] +
                       wordpre_all[(size_t)f * MS + (size_t)seg * SEG_WORDS + (lane >> 2)] +
                       __popc(mw[u] & ((1u << bsh) - 1u));
            uint32_t todo = m;
            while (todo) {
                const uint32_t k = __ffs(todo) - 1;
                todo &= todo - 1;
                uint32_t val = 1u;
                if (level == 1) {
                    val = fetch_bits(pk, rank * (uint32_t)b, b);
                    rank++;
                    const uint64_t x = (uint64_t)val << ((k & 3u) * 16u);
                    if (k & 4u) qhi |= x; else qlo |= x;
                }
                if (sum) atomicAdd(&sum[p_base + k], val);
            }
        }
        if (dense) {
            T *o = dense + (size_t)f * P + p_base;
            if (vec_ok && p_base + 8 <= P) {
                if (sizeof(T) == 2) {
                    uint4 q;
                    q.x = (uint32_t)qlo; q.y = (uint32_t)(qlo >> 32); q.z = (uint32_t)qhi; q.w = (uint32_t)(qhi >> 32);
                    __stcs(reinterpret_cast<uint4 *>(o), q);
                } else {
                    // 8-bit targets: squeeze the 16-bit slots to bytes
                    const uint32_t l0 = (uint32_t)qlo, l1 = (uint32_t)(qlo >> 32), h0 = (uint32_t)qhi, h1 = (uint32_t)(qhi >> 32);
                    uint2 q;
                    q.x = (l0 & 0xffu) | ((l0 >> 8) & 0xff00u) | ((l1 & 0xffu) << 16) | ((l1 & 0xff0000u) << 8);
                    q.y = (h0 & 0xffu) | ((h0 >> 8) & 0xff00u) | ((h1 & 0xffu) << 16) | ((h1 & 0xff0000u) << 8);
                    __stcs(reinterpret_cast<uint2 *>(o), q);
                }
            } else {
#pragma unroll
                for (int k = 0; k < 8; k++)
                    if (p_base + k < P) o[k] = (T)(((k < 4 ? qlo : qhi) >> ((k & 3) * 16)) & 0xffffu);
            }
        }
    }
}

int launch_unpack_sparse(rc_ctx *ctx, const Geom &g, int level, int b, const uint32_t *maps, const uint8_t *packed,
                         size_t packed_stride, const uint16_t *wordpre, const uint32_t *tilepre, int F,
                         uint64_t *triples, size_t capacity, cudaStream_t st)
{
    if (F <= 0) return 0;
    dim3 grid((unsigned)((g.MW + 255) / 256), F);
    k_unpack_sparse<<<grid, 256, 0, st>>>(maps, g.MS, packed, packed_stride, wordpre, tilepre, g.NT, g.nx,
                                          (uint32_t)g.MW, level, b, triples, capacity);
    RC_LAUNCH_CHECK(ctx, "k_unpack_sparse");
    return 0;
}

int launch_unpack_dense(rc_ctx *ctx, const Geom &g, int itemsize, int level, int b, const uint32_t *maps,
                        const uint8_t *packed, size_t packed_stride, const uint16_t *wordpre, const uint32_t *tilepre,
                        int F, void *dense, uint32_t *sum, cudaStream_t st)
{
    if (F <= 0) return 0;
    const uint32_t nseg = (uint32_t)g.NT * SEGS_PER_TILE;
    dim3 grid((nseg + 8 * UD_SEGS - 1) / (8 * UD_SEGS), F);
    const int vec_ok = dense && ((g.P * itemsize) % 16 == 0) && ((uintptr_t)dense % 16 == 0);
    if (itemsize == 2)
        k_unpack_dense<uint16_t><<<grid, 256, 0, st>>>(maps, g.MS, packed, packed_stride, wordpre, tilepre, g.NT, g.P,
                                                       level, b, (uint16_t *)dense, sum, vec_ok);
    else
        k_unpack_dense<uint8_t><<<grid, 256, 0, st>>>(maps, g.MS, packed, packed_stride, wordpre, tilepre, g.NT, g.P,
                                                      level, b, (uint8_t *)dense, sum, vec_ok);
    RC_LAUNCH_CHECK(ctx, "k_unpack_dense");
    return 0;
}

// ---- offline recalibration of L1 frames (pyrecode/utils/converters.py:15-57, recalibrate_l1) --------------------
// out = (T) clamp(float64(frame) + diff, 0, max(T)) per pixel, diff = original_calibration - (new_calibration + eps)
// as float64 [P] (shared by all frames, L2-resident).  float64 -> integer conversion truncates, like numpy's astype.
// HBM-bound: itemsize read + itemsize written per pixel.
template <typename T>
__global__ void __launch_bounds__(256)
k_recalibrate(const T *__restrict__ frames, const double *__restrict__ diff, size_t P, int F, T *__restrict__ out)
{
    constexpr int V = 16 / sizeof(T);                       // pixels per 128-bit access
    const double tmax = (double)((1u << (8 * sizeof(T))) - 1u);
    const size_t nvec = P / V;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
        double d[V];
#pragma unroll
        for (int k = 0; k < V; k += 2) {
            const double2 dd = reinterpret_cast<const double2 *>(diff)[(i * V + k) / 2];
            d[k] = dd.x; d[k + 1] = dd.y;
        }
        for (int f = 0; f < F; f++) {
            const uint4 q = ld_stream_u4(frames + (size_t)f * P + i * V);
            const T *x = reinterpret_cast<const T *>(&q);
            uint4 r;
            T *y = reinterpret_cast<T *>(&r);
#pragma unroll
            for (int k = 0; k < V; k++) {
                double v = (double)x[k] + d[k];
                v = v < 0.0 ? 0.0 : (v > tmax ? tmax : v);
                y[k] = (T)v;
            }
            __stcs(reinterpret_cast<uint4 *>(out + (size_t)f * P + i * V), r);
        }
    }
    // ragged tail (P not a multiple of the vector width)
    const size_t t0 = nvec * V;
    for (size_t i = t0 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (size_t)gridDim.x * blockDim.x)
        for (int f = 0; f < F; f++) {
            double v = (double)frames[(size_t)f * P + i] + diff[i];
            v = v < 0.0 ? 0.0 : (v > tmax ? tmax : v);
            out[(size_t)f * P + i] = (T)v;
        }
}

int launch_recalibrate(rc_ctx *ctx, int itemsize, const void *frames, const double *diff, size_t P, int F, void *out,
                       cudaStream_t st)
{
    if (F <= 0 || P == 0) return 0;
    const int vec_ok = ((P * itemsize) % 16 == 0) && ((uintptr_t)frames % 16 == 0) && ((uintptr_t)out % 16 == 0) &&
                       ((uintptr_t)diff % 16 == 0);
    if (!vec_ok) RC_FAIL(ctx, -1, "rc_recalibrate needs 16-byte aligned buffers and frames of a multiple of 16 bytes");
    const unsigned grid = (unsigned)ctx->sm_count * 8;
    if (itemsize == 2) k_recalibrate<uint16_t><<<grid, 256, 0, st>>>((const uint16_t *)frames, diff, P, F, (uint16_t *)out);
    else k_recalibrate<uint8_t><<<grid, 256, 0, st>>>((const uint8_t *)frames, diff, P, F, (uint8_t *)out);
    RC_LAUNCH_CHECK(ctx, "k_recalibrate");
    return 0;
}
