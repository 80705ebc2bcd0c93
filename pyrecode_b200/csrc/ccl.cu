// ccl.cu -- 8-connected puddle labelling on the bit-packed binary map and the per-puddle reductions
// of reduction levels 2 and 4.
//
// Replaces scipy.ndimage.label(binary, 3x3) (recode_writer.py:443), get_summary_stats_nb
// (pyrecode/utils/converters.py:262-297), get_centroids_2D_nb (:157-259) and make_binary_map (:300-309).
//
// Representation: union-find over foreground SLOTS (common.cuh).  parent[slot] <= slot always, so the root
// of a puddle is its first pixel in raster order -- exactly scipy's label order (labels numbered by the
// raster order of each component's first pixel).  Only foreground pixels are ever touched.
//
// All kernels run one thread per 32-pixel map word.  When nx is a multiple of 32 (every shipped geometry)
// the neighbourhood of the 32 pixels is evaluated bit-parallel from the 3 x 3 surrounding words, so a word
// whose pixels have no neighbour at all (the common case in electron-counting frames) costs a handful of
// coalesced loads and logic ops and exits; other geometries take a per-pixel path with the same results.
//
//   (k_reduce_tiles labels every 32768-pixel tile in shared memory and folds the L2 statistic there.)
//   k_ccl_border    PASS 0 links pixels across tile boundaries with atomicMin unions; PASS 1 folds the L2
//                   statistic of every tile-local root that was merged into an earlier tile's puddle.
//   k_ccl_union     full-frame linking for maps that did not come from k_reduce_tiles (rc_ccl_label).
//   k_ccl_flatten   parent[slot] = root; L4: grows the root's bounding box (row extent, left / right columns).
//   k_ccl_roots     one warp per tile: compacts per-root payloads (L2 statistics, centroids, ordinals) in
//                   slot order == label order.
//   k_l4_centroids  one thread per root: replays the puddle's pixels in raster order inside its bounding
//                   box with the reference's float32-after-every-add accumulation, divides, rounds half to
//                   even and sets the centroid bit.  Single-pixel puddles are their own centroid.
#include "common.cuh"
#include "kernels.cuh"
#include "ccl_core.cuh"

// ---- tile-local labelling ----------------------------------------------------------------------
// One CTA per (tile, frame).  A tile of an electron-counting frame holds a few hundred foreground pixels in
// puddles of a few pixels, so the work is organised per foreground pixel, not per map word:
//   phase 1  every foreground pixel (position from k_reduce_tiles) tests its W / NW / N / NE neighbours on
//            the shared-memory copy of the tile's map and appends the links it finds to a dense list
//            (warp-aggregated).  Links to pixels of earlier tiles go to the tile's global cross-link list.
//   phase 2  the dense list is processed with atomicMin unions in shared memory (all lanes busy).
//   phase 3  flatten; L2 folds each member's value into its root (max or sum).
//   phase 4  parent[slot] = slot of the tile-local root (| UF_FLAG for non-roots), acc[slot].
// A tile with more than CCL_CAP foreground pixels (> 6 % occupancy), or whose link lists overflow, is not
// labelled here: it gets parent[slot] = slot, acc[slot] = value, tileovf = 1 and k_ccl_border links all of
// its pixels with the global word-parallel path.
constexpr int CCL_CAP = 2048;          // foreground pixels per tile handled in shared memory
constexpr int CCL_LINKS = 2560;        // tile-local links
constexpr int CCL_XCAP = 256;          // cross-tile links per tile (global list)
constexpr int CCL_HALO = 264;          // map words kept in front of the tile (multiple of 4): nx <= 8447
constexpr int CCL_THREADS = 256;

// FOLD: 0 = labels only (L4), 1 = L2 max, 2 = L2 sum
template <int FOLD>
__global__ void __launch_bounds__(CCL_THREADS)
k_ccl_tiles(const uint32_t *__restrict__ maps, size_t MS, const uint16_t *__restrict__ wordpre_all, int NT,
            const uint32_t *__restrict__ tilecnt, const uint32_t *__restrict__ vp_all, uint8_t *__restrict__ tileovf,
            uint32_t *__restrict__ xcount, uint2 *__restrict__ xlinks, uint32_t *__restrict__ parent_all,
            uint32_t *__restrict__ acc_all, int ny, int nx)
{
    // map words of the tile preceded by a halo: the CCL_HALO words before the tile (zeros before the frame),
    // so that the W / NW / N / NE probes of every pixel are plain shared-memory reads
    __shared__ __align__(16) uint32_t s_maskx[CCL_HALO + TILE_WORDS];
    __shared__ __align__(16) uint16_t s_wpre[TILE_WORDS];
    __shared__ uint32_t s_parent[CCL_CAP];
    __shared__ uint32_t s_acc[FOLD ? CCL_CAP : 1];
    __shared__ uint16_t s_pos[CCL_CAP];
    __shared__ uint32_t s_links[CCL_LINKS];            // (a << 16) | b, tile-local slots
    __shared__ uint32_t s_nlinks, s_nx, s_bad;
    const int tile = blockIdx.x, f = blockIdx.y, t = threadIdx.x, lane = t & 31;
    const size_t ti = (size_t)f * NT + tile;
    const uint32_t base = (uint32_t)tile << TILE_LOG2;
    const size_t sbase = (size_t)f * ((size_t)NT * TILE_PX) + base;
    const size_t wo = (size_t)f * MS + (size_t)tile * TILE_WORDS;
    uint32_t *parent = parent_all + sbase;
    const uint32_t *vp = vp_all + sbase;
    const uint32_t *s_mask = s_maskx + CCL_HALO;
    // issue the tile's loads before the (dependent) per-pixel ones
    const uint4 r_map = reinterpret_cast<const uint4 *>(maps + wo)[t];
    uint4 r_wpre = make_uint4(0, 0, 0, 0), r_halo = make_uint4(0, 0, 0, 0);
    if (t < TILE_WORDS / 8) r_wpre = reinterpret_cast<const uint4 *>(wordpre_all + wo)[t];
    if (t < CCL_HALO / 4 && tile > 0) r_halo = reinterpret_cast<const uint4 *>(maps + wo - CCL_HALO)[t];
    const uint32_t total = tilecnt[ti];
    if (total == 0) {
        if (t == 0) { tileovf[ti] = 0; xcount[ti] = 0; }
        return;
    }
    const uint32_t unx = (uint32_t)nx;
    // the probes reach nx + 1 pixels back; wider frames than the halo covers take the global path
    bool overflow = total > (uint32_t)CCL_CAP || unx + 1 > (uint32_t)CCL_HALO * 32;
    if (!overflow) {
        reinterpret_cast<uint4 *>(s_maskx + CCL_HALO)[t] = r_map;
        if (t < TILE_WORDS / 8) reinterpret_cast<uint4 *>(s_wpre)[t] = r_wpre;
        if (t < CCL_HALO / 4) reinterpret_cast<uint4 *>(s_maskx)[t] = r_halo;
        for (uint32_t i = t; i < total; i += CCL_THREADS) {
            const uint32_t v = vp[i];
            s_pos[i] = (uint16_t)v;
            if (FOLD) s_acc[i] = v >> 16;
            s_parent[i] = i;
        }
        if (t == 0) { s_nlinks = 0; s_nx = 0; s_bad = 0; }
        __syncthreads();

        // ---- phase 1: link detection
        const bool pow2 = (unx & (unx - 1u)) == 0;
        uint2 *xl = xlinks + ti * CCL_XCAP;
        constexpr uint32_t HP = CCL_HALO * 32;         // halo pixels
        for (uint32_t i0 = 0; i0 < total; i0 += CCL_THREADS) {
            const uint32_t i = i0 + t;
            uint32_t l0 = 0, l1 = 0, l2 = 0;           // local links found by this pixel: (i << 16) | other
            uint32_t n = 0;
            if (i < total) {
                const uint32_t p = s_pos[i];
                const uint32_t gp = base + p;
                const uint32_t col = pow2 ? (gp & (unx - 1u)) : (gp % unx);
                const bool hl = col > 0, hr = col + 1 < unx, up = gp >= unx;
                const uint32_t e = p + HP;             // pixel index in the halo-extended map
                const uint32_t q = e - unx;            // >= 1: the halo covers nx + 1 pixels
                const bool bw = (s_maskx[(e - 1) >> 5] >> ((e - 1) & 31)) & 1u;
                const bool bnw = (s_maskx[(q - 1) >> 5] >> ((q - 1) & 31)) & 1u;
                const bool bn = (s_maskx[q >> 5] >> (q & 31)) & 1u;
                const bool bne = (s_maskx[(q + 1) >> 5] >> ((q + 1) & 31)) & 1u;
                // up to three links: W, and N or (NW, NE) -- NW / NE are implied when N is set
                constexpr uint32_t NONE = 0xffffffffu;
                const uint32_t c0 = (bw && hl) ? e - 1 : NONE;
                const uint32_t c1 = !up ? NONE : (bn ? q : ((bnw && hl) ? q - 1 : NONE));
                const uint32_t c2 = (up && !bn && bne && hr) ? q + 1 : NONE;
                // candidates at or above HP are in this tile (branch-free slot lookups); the others are rare
                const bool k0 = c0 != NONE && c0 >= HP, k1 = c1 != NONE && c1 >= HP, k2 = c2 != NONE && c2 >= HP;
                const uint32_t q1 = k1 ? c1 - HP : 0u, q2 = k2 ? c2 - HP : 0u;
                const uint32_t e0 = (i << 16) | (i - 1);
                const uint32_t e1 = (i << 16) | (s_wpre[q1 >> 5] + __popc(s_mask[q1 >> 5] & ((1u << (q1 & 31)) - 1u)));
                const uint32_t e2 = (i << 16) | (s_wpre[q2 >> 5] + __popc(s_mask[q2 >> 5] & ((1u << (q2 & 31)) - 1u)));
                n = (uint32_t)k0 + (uint32_t)k1 + (uint32_t)k2;
                l0 = k0 ? e0 : (k1 ? e1 : e2);
                l1 = (k0 && k1) ? e1 : e2;
                l2 = e2;
                if ((c0 < HP) | (c1 < HP) | (c2 < HP)) {
                    // neighbour in an earlier tile: (slot, neighbour PIXEL); k_ccl_border resolves its slot
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        const uint32_t ce = c == 0 ? c0 : (c == 1 ? c1 : c2);
                        if (ce < HP) {
                            const uint32_t k = atomicAdd(&s_nx, 1u);
                            if (k < (uint32_t)CCL_XCAP) xl[k] = make_uint2(base + i, base - (HP - ce));
                            else s_bad = 1;
                        }
                    }
                }
            }
            // warp-aggregated append of n in {0..3} entries per lane
            const uint32_t b0 = __ballot_sync(0xffffffffu, n & 1u), b1 = __ballot_sync(0xffffffffu, n & 2u);
            const uint32_t lt = (1u << lane) - 1u;
            const uint32_t pre = __popc(b0 & lt) + 2u * __popc(b1 & lt);
            const uint32_t tot = __popc(b0) + 2u * __popc(b1);
            uint32_t wbase = 0;
            if (lane == 0 && tot) wbase = atomicAdd(&s_nlinks, tot);
            wbase = __shfl_sync(0xffffffffu, wbase, 0);
            const uint32_t o = wbase + pre;
            if (n > 0 && o < (uint32_t)CCL_LINKS) s_links[o] = l0;
            if (n > 1 && o + 1 < (uint32_t)CCL_LINKS) s_links[o + 1] = l1;
            if (n > 2 && o + 2 < (uint32_t)CCL_LINKS) s_links[o + 2] = l2;
        }
        __syncthreads();
        overflow = s_bad || s_nlinks > (uint32_t)CCL_LINKS;
    }
    if (overflow) {
        for (uint32_t i = t; i < total; i += CCL_THREADS) {
            parent[i] = base + i;
            if (FOLD) acc_all[sbase + i] = vp[i] >> 16;
        }
        if (t == 0) { tileovf[ti] = 1; xcount[ti] = 0; }
        return;
    }
    if (t == 0) { tileovf[ti] = 0; xcount[ti] = s_nx; }

    // ---- phase 2: unions
    const uint32_t nl = s_nlinks;
    for (uint32_t j = t; j < nl; j += CCL_THREADS) {
        const uint32_t e = s_links[j];
        uf_union(s_parent, e >> 16, e & 0xffffu);
    }
    __syncthreads();
    // ---- phase 3: flatten; L2 folds every member's value into its root (non-roots are never written again)
    for (uint32_t i = t; i < total; i += CCL_THREADS) {
        const uint32_t r = uf_find_ro(s_parent, i);
        if (r != i) {
            s_parent[i] = r;
            if (FOLD == 1) atomicMax(&s_acc[r], s_acc[i]);
            if (FOLD == 2) atomicAdd(&s_acc[r], s_acc[i]);
        }
    }
    __syncthreads();
    // ---- phase 4
    for (uint32_t i = t; i < total; i += CCL_THREADS) {
        const uint32_t r = s_parent[i];
        parent[i] = r == i ? base + i : ((base + r) | UF_FLAG);
        if (FOLD) acc_all[sbase + i] = s_acc[i];
    }
}

// full-frame union (maps that did not come from k_reduce_tiles: rc_ccl_label)
__global__ void __launch_bounds__(256)
k_ccl_union(const uint32_t *__restrict__ maps, size_t MS, const uint16_t *__restrict__ wordpre_all, int NT,
            uint32_t *__restrict__ parent_all, int ny, int nx, uint32_t MW)
{
    const int f = blockIdx.y;
    const uint32_t w = blockIdx.x * 256 + threadIdx.x;
    if (w >= MW) return;
    const uint32_t *map = maps + (size_t)f * MS;
    const uint32_t bits = map[w];
    if (!bits) return;
    GlobalSpace sp{map, wordpre_all + (size_t)f * MS, parent_all + (size_t)f * ((size_t)NT * TILE_PX)};
    link_word(sp, UnionAct{sp.parent}, w, bits, ny, nx, 0xffffffffu);
}

// L2 fold of a tile-local root that lost its root status in k_ccl_border<0>: exactly once (claimed by setting
// UF_FLAG on its parent entry), its accumulated statistic goes to the final root of its puddle.
struct FoldAct {
    uint32_t *parent, *acc;
    int sum;
    __device__ __forceinline__ void one(uint32_t x) const
    {
        const uint32_t px = parent[x];
        const uint32_t r = (px & UF_FLAG) ? (px & ~UF_FLAG) : x;     // tile-local root of x
        const uint32_t pr = parent[r];
        if (pr == r || (pr & UF_FLAG)) return;                       // still a root, or already folded
        if (atomicOr(&parent[r], UF_FLAG) & UF_FLAG) return;
        const uint32_t g = uf_find_ro(parent, r);
        if (sum) atomicAdd(&acc[g], acc[r]);
        else atomicMax(&acc[g], acc[r]);
    }
    __device__ __forceinline__ void operator()(uint32_t a, uint32_t b) const { one(a); one(b); }
};

// Links across tile boundaries.  k_ccl_tiles labelled every tile on its own and listed the links from its
// first rows to pixels of earlier tiles; a tile that overflowed the shared-memory labelling (tileovf) has all
// of its links made here instead, with the word-parallel global path.
// PASS 0: union.  PASS 1: L2 fold of the re-parented tile-local roots (separate launch: needs final roots).
// One warp per tile.
template <int PASS>
__global__ void __launch_bounds__(256)
k_ccl_border(const uint32_t *__restrict__ maps, size_t MS, const uint16_t *__restrict__ wordpre_all, int NT,
             const uint8_t *__restrict__ tileovf, const uint32_t *__restrict__ xcount,
             const uint2 *__restrict__ xlinks, uint32_t *__restrict__ parent_all, uint32_t *__restrict__ acc_all,
             int ny, int nx, uint32_t MW, int sum)
{
    const int tile = blockIdx.x * 8 + (threadIdx.x >> 5), f = blockIdx.y, lane = threadIdx.x & 31;
    if (tile >= NT) return;
    const size_t ti = (size_t)f * NT + tile;
    const size_t slots = (size_t)NT * TILE_PX;
    uint32_t *parent = parent_all + (size_t)f * slots;
    uint32_t *acc = acc_all + (size_t)f * slots;
    if (!tileovf[ti]) {
        const uint32_t n = xcount[ti];
        const uint2 *xl = xlinks + ti * CCL_XCAP;
        const uint32_t *map = maps + (size_t)f * MS;
        const uint16_t *wordpre = wordpre_all + (size_t)f * MS;
        for (uint32_t j = lane; j < n; j += 32) {
            const uint2 e = xl[j];                       // (slot in this tile, neighbour pixel in an earlier tile)
            const uint32_t sq = slot_of(map, wordpre, e.y);
            if (PASS == 0) uf_union(parent, e.x, sq);
            else FoldAct{parent, acc, sum}(e.x, sq);
        }
        return;
    }
    const uint32_t w0 = (uint32_t)tile * TILE_WORDS;
    uint32_t w1 = w0 + TILE_WORDS;
    if (w1 > MW) w1 = MW;
    const uint32_t *map = maps + (size_t)f * MS;
    GlobalSpace sp{map, wordpre_all + (size_t)f * MS, parent};
    for (uint32_t w = w0 + lane; w < w1; w += 32) {
        const uint32_t bits = map[w];
        if (!bits) continue;
        if (PASS == 0) link_word(sp, UnionAct{parent}, w, bits, ny, nx, 0xffffffffu);
        else link_word(sp, FoldAct{parent, acc, sum}, w, bits, ny, nx, 0xffffffffu);
    }
}

// MODE 0: labels only.  MODE 3: L4 bounding boxes.  Afterwards every parent entry is the plain root slot.
template <int MODE>
__global__ void __launch_bounds__(256)
k_ccl_flatten(const uint32_t *__restrict__ maps, size_t MS, const uint16_t *__restrict__ wordpre_all, int NT,
              uint32_t *__restrict__ parent_all, uint32_t *__restrict__ bbox_all, int ny, int nx, uint32_t MW)
{
    const int f = blockIdx.y;
    const uint32_t w = blockIdx.x * 256 + threadIdx.x;
    if (w >= MW) return;
    const uint32_t *map = maps + (size_t)f * MS;
    const uint32_t bits = map[w];
    if (!bits) return;
    const size_t slots = (size_t)NT * TILE_PX;
    uint32_t *parent = parent_all + (size_t)f * slots;
    uint32_t todo = bits;
    if ((nx & 31) == 0) {
        // a pixel without any of its 8 neighbours set is a single-pixel puddle: already its own (plain) root
        GlobalSpace sp{map, wordpre_all + (size_t)f * MS, parent};
        const Nbr m = neighbour_masks<true>(sp, w, bits, (uint32_t)nx >> 5, (uint32_t)ny);
        todo = bits & (m.west | m.east | m.n | m.nw | m.ne | m.s | m.sw | m.se);
        if (!todo) return;
    }
    const uint32_t sb = word_slot_base(wordpre_all + (size_t)f * MS, w);
    const uint32_t p0 = w << 5;
    while (todo) {
        const uint32_t k = __ffs(todo) - 1;
        todo &= todo - 1;
        const uint32_t s = sb + __popc(bits & ((1u << k) - 1u));
        const uint32_t root = uf_find_ro(parent, s);
        if (root != s) {
            parent[s] = root;
            if (MODE == 3) {
                // the root's pixel index was stored in bbox[3] by k_l4_init_bbox
                const uint32_t p = p0 + k;
                uint32_t *bb = bbox_all + ((size_t)f * slots + root) * 4;
                const uint32_t rp = bb[3];
                const uint32_t r = p / (uint32_t)nx, c = p - r * (uint32_t)nx;
                const uint32_t rr = rp / (uint32_t)nx, rc = rp - rr * (uint32_t)nx;
                atomicMax(&bb[0], r - rr);                       // rows below the root (root is the top row)
                if (c < rc) atomicMax(&bb[1], rc - c);           // columns left of the root
                if (c > rc) atomicMax(&bb[2], c - rc);           // columns right of the root
            }
        }
    }
}

// L4: bbox[slot] = {0, 0, 0, pixel index} for every foreground slot (must precede k_ccl_flatten<3>)
__global__ void __launch_bounds__(256)
k_l4_init_bbox(const uint32_t *__restrict__ maps, size_t MS, const uint16_t *__restrict__ wordpre_all, int NT,
               uint32_t *__restrict__ bbox_all, uint32_t MW)
{
    const int f = blockIdx.y;
    const uint32_t w = blockIdx.x * 256 + threadIdx.x;
    if (w >= MW) return;
    uint32_t bits = maps[(size_t)f * MS + w];
    if (!bits) return;
    const size_t slots = (size_t)NT * TILE_PX;
    uint32_t s = word_slot_base(wordpre_all + (size_t)f * MS, w);
    while (bits) {
        const uint32_t k = __ffs(bits) - 1;
        bits &= bits - 1;
        reinterpret_cast<uint4 *>(bbox_all)[(size_t)f * slots + s] = make_uint4(0, 0, 0, (w << 5) + k);
        s++;
    }
}

// parent[slot] = slot for every foreground slot (when the map did not come from k_reduce_tiles)
__global__ void __launch_bounds__(256)
k_ccl_init(const uint32_t *__restrict__ tilecnt, int NT, uint32_t *__restrict__ parent_all)
{
    const int f = blockIdx.x, tile = blockIdx.y;
    const uint32_t cnt = tilecnt[(size_t)f * NT + tile];
    const uint32_t tb = (uint32_t)tile * TILE_PX;
    uint32_t *parent = parent_all + (size_t)f * ((size_t)NT * TILE_PX);
    for (uint32_t i = threadIdx.x; i < cnt; i += 256) parent[tb + i] = tb + i;
}

// ---- root compaction ---------------------------------------------------------------------------
// One warp per tile: roots in slot order get local ordinals 0..; rootcnt[f][tile] = number of roots.
// PAYLOAD 0: ord[slot] = local ordinal (for label images)
// PAYLOAD 1: out16[tile-compacted] = (uint16) acc[slot]                (L2 statistics)
// PAYLOAD 2: out64[tile-compacted] = cent[slot] (float2 as uint64)     (L4 centroid lists)
// PAYLOAD 3: count only
template <int PAYLOAD>
__global__ void __launch_bounds__(256)
k_ccl_roots(const uint32_t *__restrict__ tilecnt, int NT, int n_tiles_total, const uint32_t *__restrict__ parent_all,
            const uint32_t *__restrict__ acc_all, const uint64_t *__restrict__ cent_all,
            uint32_t *__restrict__ rootcnt, uint32_t *__restrict__ ord_all, uint16_t *__restrict__ out16,
            uint64_t *__restrict__ out64)
{
    const int gt = blockIdx.x * 8 + (threadIdx.x >> 5);       // global tile index = f * NT + tile
    if (gt >= n_tiles_total) return;
    const int lane = threadIdx.x & 31;
    const int f = gt / NT, tile = gt - f * NT;
    const size_t fs = (size_t)f * ((size_t)NT * TILE_PX);
    const uint32_t tb = (uint32_t)tile * TILE_PX;
    const uint32_t cnt = tilecnt[gt];
    uint32_t carry = 0;
    for (uint32_t i0 = 0; i0 < cnt; i0 += 128) {
        // four independent parent loads in flight per lane
        bool is_root[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const uint32_t i = i0 + 32 * u + lane;
            is_root[u] = i < cnt && parent_all[fs + tb + i] == tb + i;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const uint32_t s = tb + i0 + 32 * u + lane;
            const uint32_t b = __ballot_sync(0xffffffffu, is_root[u]);
            if (is_root[u]) {
                const uint32_t j = carry + __popc(b & ((1u << lane) - 1u));
                if (PAYLOAD == 0) ord_all[fs + s] = j;
                if (PAYLOAD == 1) out16[fs + tb + j] = (uint16_t)acc_all[fs + s];
                if (PAYLOAD == 2) out64[fs + tb + j] = cent_all[fs + s];
            }
            carry += __popc(b);
        }
    }
    if (lane == 0) rootcnt[gt] = carry;
}

// dense label image: label = global ordinal of the root + 1 (scipy numbering), 0 = background
__global__ void __launch_bounds__(256)
k_ccl_label_image(const uint32_t *__restrict__ maps, size_t MS, const uint16_t *__restrict__ wordpre_all, int NT,
                  const uint32_t *__restrict__ parent_all, const uint32_t *__restrict__ ord_all,
                  const uint32_t *__restrict__ rootpre_all, int32_t *__restrict__ labels, size_t P, uint32_t MW)
{
    const int f = blockIdx.y;
    const uint32_t w = blockIdx.x * 256 + threadIdx.x;
    if (w >= MW) return;
    const uint32_t bits = maps[(size_t)f * MS + w];
    const size_t slots = (size_t)NT * TILE_PX;
    const uint32_t *parent = parent_all + (size_t)f * slots;
    const uint32_t *ord = ord_all + (size_t)f * slots;
    const uint32_t *rootpre = rootpre_all + (size_t)f * (NT + 1);
    uint32_t s = bits ? word_slot_base(wordpre_all + (size_t)f * MS, w) : 0;
    int32_t *out = labels + (size_t)f * P;
    const size_t p0 = (size_t)w << 5;
    for (int k = 0; k < 32; k++) {
        if (p0 + k >= P) break;
        int32_t L = 0;
        if (bits & (1u << k)) {
            const uint32_t root = parent[s];
            L = (int32_t)(rootpre[root >> TILE_LOG2] + ord[root]) + 1;
            s++;
        }
        out[p0 + k] = L;
    }
}

// centroid lists in label order from the tile-compacted per-root payloads
__global__ void __launch_bounds__(256)
k_gather_centroids(const uint64_t *__restrict__ cent_tiles, const uint32_t *__restrict__ rootpre_all, int NT,
                   float *__restrict__ out, size_t capacity)
{
    const int f = blockIdx.x, tile = blockIdx.y;
    const uint32_t *rootpre = rootpre_all + (size_t)f * (NT + 1);
    const uint32_t r0 = rootpre[tile], r1 = rootpre[tile + 1];
    const uint64_t *src = cent_tiles + (size_t)f * ((size_t)NT * TILE_PX) + (size_t)tile * TILE_PX;
    for (uint32_t j = threadIdx.x; j < r1 - r0; j += 256) {
        if ((size_t)(r0 + j) >= capacity) break;
        const uint64_t pk = src[j];
        float *o = out + ((size_t)f * capacity + r0 + j) * 2;
        o[0] = __uint_as_float((uint32_t)pk);
        o[1] = __uint_as_float((uint32_t)(pk >> 32));
    }
}

// ---- L4 centroids ------------------------------------------------------------------------------
// One thread per root.  mode: 0/1 weighted (converters.py:167-197), 2 max pixel (:229-259), 3 unweighted (:200-226).
// Each += of the reference is float64 arithmetic rounded to float32 (numba: float32 element += float64 value).
__global__ void __launch_bounds__(128)
k_l4_centroids(const uint32_t *__restrict__ maps, size_t MS, const uint16_t *__restrict__ wordpre_all, int NT,
               const uint32_t *__restrict__ parent_all, const uint32_t *__restrict__ bbox_all,
               const uint32_t *__restrict__ vp_all, int ny, int nx, uint32_t MW, int mode,
               uint32_t *__restrict__ map2_all, uint64_t *__restrict__ cent_all)
{
    const int f = blockIdx.y;
    const uint32_t w = blockIdx.x * 128 + threadIdx.x;
    if (w >= MW) return;
    const uint32_t *map = maps + (size_t)f * MS;
    uint32_t bits = map[w];
    if (!bits) return;
    const size_t slots = (size_t)NT * TILE_PX;
    const uint16_t *wordpre = wordpre_all + (size_t)f * MS;
    const uint32_t *parent = parent_all + (size_t)f * slots;
    const uint32_t *vp = vp_all + (size_t)f * slots;            // (value << 16) | position, from k_reduce_tiles
    uint32_t *map2 = map2_all + (size_t)f * MS;
    uint32_t s = word_slot_base(wordpre, w);
    const uint32_t p0 = w << 5;
    for (; bits; s++) {
        const uint32_t k = __ffs(bits) - 1;
        bits &= bits - 1;
        if (parent[s] != s) continue;
        const uint32_t p = p0 + k;
        const uint32_t r0 = p / (uint32_t)nx, c0 = p - r0 * (uint32_t)nx;
        const uint4 bb = reinterpret_cast<const uint4 *>(bbox_all)[(size_t)f * slots + s];
        float fr, fc;
        if ((bb.x | bb.y | bb.z) == 0) {
            // single-pixel puddle: the reference computes RN32(v*r) / RN32(v)
            const float v = (float)(vp[s] >> 16);
            if (mode == 2 || mode == 3) { fr = (float)r0; fc = (float)c0; }
            else {
                fr = __fdiv_rn((float)((double)v * (double)r0), v);
                fc = __fdiv_rn((float)((double)v * (double)c0), v);
            }
        } else {
            float a0 = 0.f, a1 = 0.f, a2 = 0.f;
            bool first = true;
            const uint32_t cl = c0 - bb.y, cr = c0 + bb.z;
            for (uint32_t r = r0; r <= r0 + bb.x; r++) {
                const uint32_t q0 = r * (uint32_t)nx + cl, q1 = r * (uint32_t)nx + cr;   // inclusive pixel range
                for (uint32_t ww = q0 >> 5; ww <= (q1 >> 5); ww++) {
                    uint32_t mb = map[ww];
                    if (ww == (q0 >> 5)) mb &= 0xffffffffu << (q0 & 31);
                    if (ww == (q1 >> 5)) mb &= 0xffffffffu >> (31 - (q1 & 31));
                    while (mb) {
                        const uint32_t kk = __ffs(mb) - 1;
                        mb &= mb - 1;
                        const uint32_t q = (ww << 5) + kk;
                        const uint32_t sq = slot_of(map, wordpre, q);
                        if (parent[sq] != s) continue;
                        const double v = (double)(vp[sq] >> 16);
                        const uint32_t c = q - r * (uint32_t)nx;
                        if (mode == 2) {
                            if (first || v > (double)a2) { a0 = (float)r; a1 = (float)c; a2 = (float)v; }
                        } else if (mode == 3) {
                            a0 = (float)((double)a0 + (double)r);
                            a1 = (float)((double)a1 + (double)c);
                            a2 = (float)((double)a2 + 1.0);
                        } else {
                            a0 = (float)((double)a0 + v * (double)r);
                            a1 = (float)((double)a1 + v * (double)c);
                            a2 = (float)((double)a2 + v);
                        }
                        first = false;
                    }
                }
            }
            if (mode == 2) { fr = a0; fc = a1; }
            else { fr = __fdiv_rn(a0, a2); fc = __fdiv_rn(a1, a2); }
        }
        if (cent_all) {
            const uint64_t pk = (uint64_t)__float_as_uint(fr) | ((uint64_t)__float_as_uint(fc) << 32);
            cent_all[(size_t)f * slots + s] = pk;
        }
        if (map2_all) {
            const long rr = (long)rintf(fr), cc = (long)rintf(fc);     // round half to even (np.round)
            if (rr >= 0 && rr < ny && cc >= 0 && cc < nx) {
                const uint32_t q = (uint32_t)rr * (uint32_t)nx + (uint32_t)cc;
                atomicOr(&map2[q >> 5], 1u << (q & 31));
            }
        }
    }
}

// ---- launchers ---------------------------------------------------------------------------------
int launch_ccl_init(rc_ctx *ctx, const Geom &g, const uint32_t *tilecnt, uint32_t *parent, int F, cudaStream_t st)
{
    if (F <= 0) return 0;
    k_ccl_init<<<dim3(F, g.NT), 256, 0, st>>>(tilecnt, g.NT, parent);
    RC_LAUNCH_CHECK(ctx, "k_ccl_init");
    return 0;
}

int launch_gather_centroids(rc_ctx *ctx, const Geom &g, const uint64_t *cent_tiles, const uint32_t *rootpre, int F,
                            float *out, size_t capacity, cudaStream_t st)
{
    if (F <= 0) return 0;
    k_gather_centroids<<<dim3(F, g.NT), 256, 0, st>>>(cent_tiles, rootpre, g.NT, out, capacity);
    RC_LAUNCH_CHECK(ctx, "k_gather_centroids");
    return 0;
}

// fold: 0 = labels only (L4), 1 = L2 max, 2 = L2 sum
int launch_ccl_tiles(rc_ctx *ctx, const Geom &g, int fold, const uint32_t *maps, const uint16_t *wordpre,
                     const uint32_t *tilecnt, const uint32_t *vp, uint8_t *tileovf, uint32_t *xcount, void *xlinks,
                     uint32_t *parent, uint32_t *acc, int F, cudaStream_t st)
{
    if (F <= 0) return 0;
    dim3 grid((unsigned)g.NT, F);
#define RC_CT(FO)                                                                                              \
    k_ccl_tiles<FO><<<grid, CCL_THREADS, 0, st>>>(maps, g.MS, wordpre, g.NT, tilecnt, vp, tileovf, xcount,       \
                                                  (uint2 *)xlinks, parent, acc, g.ny, g.nx)
    if (fold == 0) RC_CT(0);
    else if (fold == 1) RC_CT(1);
    else RC_CT(2);
#undef RC_CT
    RC_LAUNCH_CHECK(ctx, "k_ccl_tiles");
    return 0;
}

size_t ccl_xlinks_bytes(const Geom &g, size_t F) { return F * (size_t)g.NT * CCL_XCAP * sizeof(uint2); }

int launch_ccl_union(rc_ctx *ctx, const Geom &g, const uint32_t *maps, const uint16_t *wordpre, uint32_t *parent,
                     int F, cudaStream_t st)
{
    if (F <= 0) return 0;
    dim3 grid((unsigned)((g.MW + 255) / 256), F);
    k_ccl_union<<<grid, 256, 0, st>>>(maps, g.MS, wordpre, g.NT, parent, g.ny, g.nx, (uint32_t)g.MW);
    RC_LAUNCH_CHECK(ctx, "k_ccl_union");
    return 0;
}

// fold: 0 = links only (L4), 1 = + L2 max fold, 2 = + L2 sum fold
int launch_ccl_border(rc_ctx *ctx, const Geom &g, int fold, const uint32_t *maps, const uint16_t *wordpre,
                      const uint8_t *tileovf, const uint32_t *xcount, const void *xlinks, uint32_t *parent,
                      uint32_t *acc, int F, cudaStream_t st)
{
    if (F <= 0) return 0;
    dim3 grid((unsigned)((g.NT + 7) / 8), F);
    k_ccl_border<0><<<grid, 256, 0, st>>>(maps, g.MS, wordpre, g.NT, tileovf, xcount, (const uint2 *)xlinks, parent, acc,
                                          g.ny, g.nx, (uint32_t)g.MW, 0);
    RC_LAUNCH_CHECK(ctx, "k_ccl_border<0>");
    if (fold) {
        k_ccl_border<1><<<grid, 256, 0, st>>>(maps, g.MS, wordpre, g.NT, tileovf, xcount, (const uint2 *)xlinks, parent,
                                              acc, g.ny, g.nx, (uint32_t)g.MW, fold == 2);
        RC_LAUNCH_CHECK(ctx, "k_ccl_border<1>");
    }
    return 0;
}

int launch_ccl_flatten(rc_ctx *ctx, const Geom &g, int mode, const uint32_t *maps, const uint16_t *wordpre,
                       uint32_t *parent, uint32_t *bbox, int F, cudaStream_t st)
{
    if (F <= 0) return 0;
    dim3 grid((unsigned)((g.MW + 255) / 256), F);
    const uint32_t MW = (uint32_t)g.MW;
    if (mode == 3) {
        k_l4_init_bbox<<<grid, 256, 0, st>>>(maps, g.MS, wordpre, g.NT, bbox, MW);
        RC_LAUNCH_CHECK(ctx, "k_l4_init_bbox");
        k_ccl_flatten<3><<<grid, 256, 0, st>>>(maps, g.MS, wordpre, g.NT, parent, bbox, g.ny, g.nx, MW);
    } else {
        k_ccl_flatten<0><<<grid, 256, 0, st>>>(maps, g.MS, wordpre, g.NT, parent, bbox, g.ny, g.nx, MW);
    }
    RC_LAUNCH_CHECK(ctx, "k_ccl_flatten");
    return 0;
}

int launch_ccl_roots(rc_ctx *ctx, const Geom &g, int payload, const uint32_t *tilecnt, const uint32_t *parent,
                     const uint32_t *acc, const uint64_t *cent, uint32_t *rootcnt, uint32_t *ord, uint16_t *out16,
                     uint64_t *out64, int F, cudaStream_t st)
{
    if (F <= 0) return 0;
    const int nt = F * g.NT;
    const unsigned grid = (unsigned)((nt + 7) / 8);
    if (payload == 0)
        k_ccl_roots<0><<<grid, 256, 0, st>>>(tilecnt, g.NT, nt, parent, acc, cent, rootcnt, ord, out16, out64);
    else if (payload == 1)
        k_ccl_roots<1><<<grid, 256, 0, st>>>(tilecnt, g.NT, nt, parent, acc, cent, rootcnt, ord, out16, out64);
    else if (payload == 2)
        k_ccl_roots<2><<<grid, 256, 0, st>>>(tilecnt, g.NT, nt, parent, acc, cent, rootcnt, ord, out16, out64);
    else
        k_ccl_roots<3><<<grid, 256, 0, st>>>(tilecnt, g.NT, nt, parent, acc, cent, rootcnt, ord, out16, out64);
    RC_LAUNCH_CHECK(ctx, "k_ccl_roots");
    return 0;
}

int launch_ccl_label_image(rc_ctx *ctx, const Geom &g, const uint32_t *maps, const uint16_t *wordpre,
                           const uint32_t *parent, const uint32_t *ord, const uint32_t *rootpre, int32_t *labels,
                           int F, cudaStream_t st)
{
    if (F <= 0) return 0;
    dim3 grid((unsigned)((g.MW + 255) / 256), F);
    k_ccl_label_image<<<grid, 256, 0, st>>>(maps, g.MS, wordpre, g.NT, parent, ord, rootpre, labels, g.P,
                                            (uint32_t)g.MW);
    RC_LAUNCH_CHECK(ctx, "k_ccl_label_image");
    return 0;
}

int launch_l4_centroids(rc_ctx *ctx, const Geom &g, int mode, const uint32_t *maps, const uint16_t *wordpre,
                        const uint32_t *parent, const uint32_t *bbox, const uint32_t *vp, uint32_t *map2,
                        uint64_t *cent, int F, cudaStream_t st)
{
    if (F <= 0) return 0;
    dim3 grid((unsigned)((g.MW + 127) / 128), F);
    k_l4_centroids<<<grid, 128, 0, st>>>(maps, g.MS, wordpre, g.NT, parent, bbox, vp, g.ny, g.nx, (uint32_t)g.MW, mode,
                                         map2, cent);
    RC_LAUNCH_CHECK(ctx, "k_l4_centroids");
    return 0;
}
