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
//   k_ccl_flatten   parent[slot] = root (rc_ccl_label only).
//   k_ccl_roots     one warp per tile: compacts per-root payloads (L2 statistics, centroids, ordinals) in
//                   slot order == label order.
//   k_l4_open       L4 puddles that cross tiles: one thread per puddle replays its pixels in raster order inside
//                   its bounding box with the reference's float32-after-every-add accumulation, divides, rounds
//                   half to even and sets the centroid bit.  (Puddles inside one tile: k_ccl_tiles<3>.)
#include "common.cuh"
#include "kernels.cuh"
#include "ccl_core.cuh"

// ---- tile-local labelling ----------------------------------------------------------------------
// One CTA per (tile, frame).  A tile of an electron-counting frame holds a few hundred foreground pixels in
// puddles of a few pixels, so the work is organised per foreground pixel, not per map word:
//   phase 1  every foreground pixel (value and position from k_reduce_tiles) probes its W / NW / N / NE
//            neighbours on the shared-memory copy of the tile's map and takes ONE of them as its parent with a
//            plain store (all of them precede it in raster order, so the parent entries form a forest).  The four
//            earlier neighbours of a pixel are pairwise adjacent except NE with NW / W: only a pixel that has
//            NE and (NW or W) but not N can join two trees, and only such pixels append a link to a short list.
//            Links to pixels of earlier tiles go to the tile's global cross-link list.
//   phase 2  the (few) listed links are united with atomicMin on the roots.
//   phase 3  flatten; L2 folds each member's value into its root (max or sum).
//   phase 4  parent[slot] = slot of the tile-local root (| UF_FLAG for non-roots), acc[slot].
// A tile with more than CCL_CAP foreground pixels (> 6 % occupancy), or whose link lists overflow, is not
// labelled here: it gets parent[slot] = slot, acc[slot] = value, tileovf = 1 and k_ccl_border links all of
// its pixels with the global word-parallel path.
constexpr int CCL_CAP = 2048;          // foreground pixels per tile handled in shared memory
constexpr int CCL_LINKS = 512;         // tile-local tree-joining links
constexpr int CCL_XCAP = 256;          // cross-tile links per tile (global list)
constexpr int CCL_HALO = 264;          // map words kept in front of the tile (multiple of 4): nx <= 8447

// shared-memory atomic add issued as is: around an atomicAdd() that only one lane executes the compiler still emits
// its warp-aggregation sequence (vote, leader election, shuffle)
__device__ __forceinline__ uint32_t atom_add_shared(uint32_t *p, uint32_t v)
{
    uint32_t old;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"((uint32_t)__cvta_generic_to_shared(p)), "r"(v) : "memory");
    return old;
}

// root of x in a shared-memory forest whose entries only ever decrease (any value read is a valid ancestor)
__device__ __forceinline__ uint32_t find_shared(const uint32_t *parent, uint32_t x)
{
    uint32_t p = ((const volatile uint32_t *)parent)[x];
    while (p != x) {
        x = p;
        p = ((const volatile uint32_t *)parent)[x];
    }
    return x;
}

// Centroid of one puddle with the reference's arithmetic (pyrecode/utils/converters.py): members are added in
// raster order, every += is float64 arithmetic rounded to float32 (numba: float32 element += float64 value).
// mode: 0/1 weighted average (:167-197), 2 maximum pixel (:229-259), 3 unweighted (:200-226).
struct CentAcc {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    bool first = true;
    __device__ __forceinline__ void add(int mode, uint32_t r, uint32_t c, uint32_t val)
    {
        const double v = (double)val;
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
    __device__ __forceinline__ void finish(int mode, float &fr, float &fc) const
    {
        if (mode == 2) { fr = a0; fc = a1; }
        else { fr = __fdiv_rn(a0, a2); fc = __fdiv_rn(a1, a2); }
    }
};

// centroid -> bit of the centroid map (make_binary_map intent, converters.py:300-309: round half to even)
__device__ __forceinline__ bool centroid_pixel(float fr, float fc, int ny, int nx, uint32_t &q)
{
    const long rr = (long)rintf(fr), cc = (long)rintf(fc);
    if (rr < 0 || rr >= ny || cc < 0 || cc >= nx) return false;
    q = (uint32_t)rr * (uint32_t)nx + (uint32_t)cc;
    return true;
}

// FOLD: 1 = L2 max, 2 = L2 sum, 3 = L4: centroids of the puddles that lie entirely inside the tile ("closed");
// puddles that continue in another tile ("open") get a bounding box and are finished by k_l4_open.
// One tile's inputs as the bulk-copy engine delivers them (cp.async.bulk, 16-byte granules): the map words of the
// tile preceded by a halo of CCL_HALO words, the per-word prefixes, the first CCL_VPS (value, position) words of the
// tile's foreground pixels and, for L4, the first CCL_HALO map words of the next tile.
constexpr int CCL_VPS = 1024;
static_assert(CCL_VPS * 2 >= CCL_CAP, "the staged (value, position) words double as the L4 root list (uint16 per root)");
template <bool L4>
struct __align__(128) CclStage {
    uint32_t maskx[CCL_HALO + TILE_WORDS];
    uint16_t wpre[TILE_WORDS];
    uint32_t vp[CCL_VPS];
    uint32_t bot[L4 ? CCL_HALO : 4];
};

template <int FOLD, int CCL_THREADS>
__device__ __forceinline__ void
ccl_tile(const int tile, const int f, CclStage<FOLD == 3> &S, const uint32_t total, int NT, size_t MS,
         const uint32_t *__restrict__ vp_all, uint8_t *__restrict__ tileovf,
         uint32_t *__restrict__ xcount, uint2 *__restrict__ xlinks, uint32_t *__restrict__ parent_all,
         uint32_t *__restrict__ acc_all, int ny, int nx, int l4mode, uint32_t *__restrict__ bbox_all,
         uint32_t *__restrict__ map2_all, uint64_t *__restrict__ cent_all, uint32_t *__restrict__ rootcnt)
{
    constexpr bool L4 = FOLD == 3;
    constexpr int CCL_WARPS = CCL_THREADS / 32;
    // S.maskx: map words of the tile preceded by a halo (zeros before the frame), so that the W / NW / N / NE probes
    // of every pixel are plain shared-memory reads
    uint32_t *s_maskx = S.maskx;
    uint16_t *s_wpre = S.wpre;
    const uint32_t *s_bot = S.bot;
    __shared__ uint32_t s_parent[CCL_CAP];
    __shared__ uint16_t s_pos[L4 ? CCL_CAP : 4];
    __shared__ uint32_t s_acc[CCL_CAP];                // L2: statistic per slot; L4: member-list heads
    __shared__ uint32_t s_links[CCL_LINKS];            // (a << 16) | b, tile-local slots
    __shared__ uint8_t s_open[L4 ? CCL_CAP : 4];       // pixel, then root: its puddle continues in another tile
    __shared__ uint32_t s_cmap[L4 ? TILE_WORDS : 4];   // centroid bits that fall inside the tile
    __shared__ uint32_t s_nlinks, s_nx, s_bad, s_nlist, s_nclosed;
    const int t = threadIdx.x, lane = t & 31;
    const size_t ti = (size_t)f * NT + tile;
    const uint32_t base = (uint32_t)tile << TILE_LOG2;
    const size_t slots = (size_t)NT * TILE_PX;
    const size_t sbase = (size_t)f * slots + base;
    uint32_t *parent = parent_all + sbase;
    const uint32_t *vp = vp_all + sbase;
    const uint32_t *s_mask = s_maskx + CCL_HALO;
    if (total == 0) {
        if (t == 0) { tileovf[ti] = 0; xcount[ti] = 0; if (L4) rootcnt[ti] = 0; }
        return;
    }
    const uint32_t unx = (uint32_t)nx;
    // the probes reach nx + 1 pixels back; wider frames than the halo covers take the global path
    bool overflow = total > (uint32_t)CCL_CAP || unx + 1 > (uint32_t)CCL_HALO * 32;
    const bool pow2 = (unx & (unx - 1u)) == 0;
    const uint32_t lg = 31 - __clz(unx);
    if (!overflow) {
        if (L4) for (int i = t; i < TILE_WORDS; i += CCL_THREADS) s_cmap[i] = 0;
        if (t == 0) { s_nlinks = 0; s_nx = 0; s_bad = 0; s_nlist = 0; s_nclosed = 0; }
        __syncthreads();

        // ---- phase 1: one parent per pixel, tree-joining links to the list
        uint2 *xl = xlinks + ti * CCL_XCAP;
        constexpr uint32_t HP = CCL_HALO * 32;         // halo pixels
        constexpr uint32_t NONE = 0xffffffffu;
        for (uint32_t i = t; i < total; i += CCL_THREADS) {
            const uint32_t v = i < (uint32_t)CCL_VPS ? S.vp[i] : vp[i];
            const uint32_t p = v & 0xffffu;
            if (L4) { s_pos[i] = (uint16_t)p; s_open[i] = 0; }
            else s_acc[i] = v >> 16;
            const uint32_t gp = base + p;
            const uint32_t col = pow2 ? (gp & (unx - 1u)) : (gp % unx);
            const bool hl = col > 0, hr = col + 1 < unx, up = gp >= unx;
            const uint32_t e = p + HP;             // pixel index in the halo-extended map
            const uint32_t q = e - unx;            // >= 1: the halo covers nx + 1 pixels
            const bool bw = (s_maskx[(e - 1) >> 5] >> ((e - 1) & 31)) & 1u;
            // NW, N, NE = three consecutive map bits from q - 1: one funnel shift over two words
            const uint32_t uw = (q - 1) >> 5;
            const uint32_t up3 = __funnelshift_r(s_maskx[uw], s_maskx[uw + 1], (q - 1) & 31);
            const bool bnw = up3 & 1u, bn = up3 & 2u, bne = up3 & 4u;
            // candidates: W; N, else NW; NE when N is clear (NW / NE next to a set N are linked through N's own W link)
            const uint32_t c0 = (bw && hl) ? e - 1 : NONE;
            const uint32_t c1 = !up ? NONE : (bn ? q : ((bnw && hl) ? q - 1 : NONE));
            const uint32_t c2 = (up && !bn && bne && hr) ? q + 1 : NONE;
            // candidates at or above HP are in this tile; the others (rare) lie in an earlier tile
            const bool k0 = c0 != NONE && c0 >= HP, k1 = c1 != NONE && c1 >= HP, k2 = c2 != NONE && c2 >= HP;
            uint32_t par = k0 ? i - 1 : i;
            if (k1 | k2) {
                const uint32_t qq = (k1 ? c1 : c2) - HP;
                par = s_wpre[qq >> 5] + __popc(s_mask[qq >> 5] & ((1u << (qq & 31)) - 1u));
            }
            s_parent[i] = par;
            // NE joins another tree than the parent's: with NW as parent, or as parent itself next to W
            if (k2 & (k1 | k0)) {
                uint32_t other = i - 1;
                if (k1) {
                    const uint32_t qq = c2 - HP;
                    other = s_wpre[qq >> 5] + __popc(s_mask[qq >> 5] & ((1u << (qq & 31)) - 1u));
                }
                const uint32_t k = atom_add_shared(&s_nlinks, 1u);
                if (k < (uint32_t)CCL_LINKS) s_links[k] = (i << 16) | other;
            }
            if ((c0 < HP) | (c1 < HP) | (c2 < HP)) {
                // neighbour in an earlier tile: (slot, neighbour PIXEL); k_ccl_border resolves its slot
                if (L4) s_open[i] = 1;
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    const uint32_t ce = c == 0 ? c0 : (c == 1 ? c1 : c2);
                    if (ce < HP) {
                        const uint32_t k = atom_add_shared(&s_nx, 1u);
                        if (k < (uint32_t)CCL_XCAP) xl[k] = make_uint2(base + i, base - (HP - ce));
                        else s_bad = 1;
                    }
                }
            }
            if (L4 && p + unx + 1 >= (uint32_t)TILE_PX && gp + unx < (uint32_t)ny * unx) {
                // last rows of the tile: a SW / S / SE neighbour in the next tile also opens the puddle
                bool any = false;
#pragma unroll
                for (int d = -1; d <= 1; d++) {
                    if ((d < 0 && !hl) || (d > 0 && !hr)) continue;
                    const uint32_t tq = p + unx + (uint32_t)d;
                    if (tq >= (uint32_t)TILE_PX) {
                        const uint32_t o = tq - TILE_PX;
                        any |= (s_bot[o >> 5] >> (o & 31)) & 1u;
                    }
                }
                if (any) s_open[i] = 1;
            }
        }
        __syncthreads();
        overflow = s_bad || s_nlinks > (uint32_t)CCL_LINKS;
    }
    if (overflow) {
        for (uint32_t i = t; i < total; i += CCL_THREADS) {
            const uint32_t v = vp[i];
            parent[i] = base + i;
            if (L4) {
                // every pixel is its own puddle until k_ccl_border links the tile: box = the pixel, not claimed
                const uint32_t gp = base + (v & 0x7fffu), r = gp / unx, c = gp - r * unx;
                reinterpret_cast<uint4 *>(bbox_all)[sbase + i] = make_uint4(r, r, c, c);
                acc_all[sbase + i] = 0;
            } else {
                acc_all[sbase + i] = v >> 16;
            }
        }
        if (t == 0) { tileovf[ti] = 1; xcount[ti] = 0; if (L4) rootcnt[ti] = 0; }
        return;
    }
    if (t == 0) { tileovf[ti] = 0; xcount[ti] = s_nx; }

    // ---- phase 2: the tree-joining links (block-uniform count: no barrier when there is none)
    const uint32_t nl = s_nlinks;
    if (nl) {
        for (uint32_t j = t; j < nl; j += CCL_THREADS) {
            const uint32_t e = s_links[j];
            uf_union(s_parent, e >> 16, e & 0xffffu);
        }
        __syncthreads();
    }
    // L4: members of each root as a linked list (head per root, next per member); the map words are consumed,
    // their shared memory is reused
    constexpr uint32_t NIL = 0xffffu;
    uint32_t *s_head = s_acc;
    uint16_t *s_next = reinterpret_cast<uint16_t *>(s_maskx);
    if (L4) {
        for (uint32_t i = t; i < total; i += CCL_THREADS) s_head[i] = NIL;
        __syncthreads();
    }
    // ---- phase 3: flatten; L2 folds every member's value into its root (non-roots are never written again)
    for (uint32_t i = t; i < total; i += CCL_THREADS) {
        const uint32_t r = find_shared(s_parent, i);
        if (r != i) {
            s_parent[i] = r;
            if (FOLD == 1) atomicMax(&s_acc[r], s_acc[i]);
            if (FOLD == 2) atomicAdd(&s_acc[r], s_acc[i]);
            if (L4) {
                s_next[i] = (uint16_t)atomicExch(&s_head[r], i);
                if (s_open[i]) s_open[r] = 1;
            }
        }
    }
    __syncthreads();
    // ---- phase 4
    for (uint32_t i = t; i < total; i += CCL_THREADS) {
        const uint32_t r = s_parent[i];
        parent[i] = r == i ? base + i : ((base + r) | UF_FLAG);
        if (!L4) acc_all[sbase + i] = s_acc[i];
    }
    if (!L4) return;

    // ---- phase 5 (L4): one thread per root.  The roots go to a dense list first, so that the warps that replay
    // puddles have every lane busy and the others skip the (long, inlined) replay code altogether: the path is bound
    // by instruction issue, and a warp pays for the whole body however few of its lanes have a puddle.  Members are
    // visited in slot order = raster order.
    uint32_t *cmap_g = map2_all ? map2_all + (size_t)f * MS : nullptr;
    uint16_t *s_list = reinterpret_cast<uint16_t *>(S.vp);   // up to CCL_CAP roots; the staged (value, position) words are consumed
    uint32_t nclosed = 0;
    auto finish_root = [&](uint32_t i, bool open, const CentAcc &ca, uint4 box) {
        if (open) {
            reinterpret_cast<uint4 *>(bbox_all)[sbase + i] = box;
            acc_all[sbase + i] = 0;                   // not yet claimed by k_l4_open
            return;
        }
        float fr, fc;
        ca.finish(l4mode, fr, fc);
        nclosed++;
        if (cent_all) cent_all[sbase + i] = (uint64_t)__float_as_uint(fr) | ((uint64_t)__float_as_uint(fc) << 32);
        uint32_t q;
        if (cmap_g && centroid_pixel(fr, fc, ny, nx, q)) {
            const uint32_t ql = q - base;                             // wraps when q < base
            if (ql < (uint32_t)TILE_PX) atomicOr(&s_cmap[ql >> 5], 1u << (ql & 31));
            else atomicOr(&cmap_g[q >> 5], 1u << (q & 31));           // unaligned tiles only
        }
    };
    for (uint32_t i0 = 0; i0 < total; i0 += CCL_THREADS) {
        const uint32_t i = i0 + t;
        const bool root = i < total && s_parent[i] == i;
        const uint32_t bm = __ballot_sync(0xffffffffu, root);
        uint32_t wbase = 0;
        if (lane == 0 && bm) wbase = atom_add_shared(&s_nlist, __popc(bm));
        wbase = __shfl_sync(0xffffffffu, wbase, 0);
        if (root) s_list[wbase + __popc(bm & ((1u << lane) - 1u))] = (uint16_t)i;
    }
    __syncthreads();
    const uint32_t nlist = s_nlist;
    // list entry k -> thread k: the first warps are full, the rest have nothing to do
    for (uint32_t k = (uint32_t)t; k < nlist; k += CCL_THREADS) {
        const uint32_t i = s_list[k];
        const bool open = s_open[i];
        CentAcc ca;
        uint32_t rmin = 0xffffffffu, rmax = 0, cmin = 0xffffffffu, cmax = 0;
        auto visit = [&](uint32_t j) {
            const uint32_t gp = base + s_pos[j];
            const uint32_t r = pow2 ? gp >> lg : gp / unx, c = gp - r * unx;
            if (open) {
                rmin = min(rmin, r); rmax = max(rmax, r); cmin = min(cmin, c); cmax = max(cmax, c);
            } else {
                ca.add(l4mode, r, c, vp[j] >> 16);       // L1 / L2-cache hit: loaded in phase 0
            }
        };
        // up to 8 members besides the root: collect, sort (the list is in arrival order), replay in slot order
        uint32_t m[8];
        uint32_t p = s_head[i];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            m[u] = p;
            if (p != NIL) p = s_next[p];
        }
        if (p == NIL) {
#define RC_CSWAP(a, b) { const uint32_t lo_ = min(m[a], m[b]), hi_ = max(m[a], m[b]); m[a] = lo_; m[b] = hi_; }
            // 19-comparator sorting network for 8 keys (NIL sorts last)
            RC_CSWAP(0, 1) RC_CSWAP(2, 3) RC_CSWAP(4, 5) RC_CSWAP(6, 7)
            RC_CSWAP(0, 2) RC_CSWAP(1, 3) RC_CSWAP(4, 6) RC_CSWAP(5, 7)
            RC_CSWAP(1, 2) RC_CSWAP(5, 6) RC_CSWAP(0, 4) RC_CSWAP(3, 7)
            RC_CSWAP(1, 5) RC_CSWAP(2, 6)
            RC_CSWAP(1, 4) RC_CSWAP(3, 6)
            RC_CSWAP(2, 4) RC_CSWAP(3, 5)
            RC_CSWAP(3, 4)
#undef RC_CSWAP
            visit(i);
#pragma unroll
            for (int u = 0; u < 8; u++) if (m[u] != NIL) visit(m[u]);
        } else {
            // larger puddle: scan the slots after the root (members have larger slots than their root)
            for (uint32_t j = i; j < total; j++) if (s_parent[j] == i) visit(j);
        }
        finish_root(i, open, ca, make_uint4(rmin, rmax, cmin, cmax));
    }
    if (nclosed) atomicAdd(&s_nclosed, nclosed);
    __syncthreads();
    if (cmap_g)
        for (int i = t; i < TILE_WORDS; i += CCL_THREADS) {
            const uint32_t w = s_cmap[i];
            if (w) atomicOr(&cmap_g[(size_t)tile * TILE_WORDS + i], w);
        }
    if (t == 0) rootcnt[ti] = s_nclosed;
}

// Persistent grid: every CTA walks the (tile, frame) pairs with a grid stride.  The inputs of the NEXT pair are
// brought into the other half of a two-stage shared-memory ring by bulk async copies (one elected thread, one
// mbarrier per stage) while the current pair is labelled, so no thread ever waits for a dependent global load: the
// kernel runs at the speed of its instruction stream instead of that of four DRAM round trips per tile.  With a few
// CTAs per SM it shares every SM with the streaming kernel of the next batch (memory-bound next to issue-bound work).
template <int FOLD, int CCL_THREADS>
__global__ void __launch_bounds__(CCL_THREADS)
k_ccl_tiles(const uint32_t *__restrict__ maps, size_t MS, const uint16_t *__restrict__ wordpre_all, int NT, int n_tiles_total,
            const uint32_t *__restrict__ tilecnt, const uint32_t *__restrict__ vp_all, uint8_t *__restrict__ tileovf,
            uint32_t *__restrict__ xcount, uint2 *__restrict__ xlinks, uint32_t *__restrict__ parent_all,
            uint32_t *__restrict__ acc_all, int ny, int nx, int l4mode, uint32_t *__restrict__ bbox_all,
            uint32_t *__restrict__ map2_all, uint64_t *__restrict__ cent_all, uint32_t *__restrict__ rootcnt)
{
    constexpr bool L4 = FOLD == 3;
    extern __shared__ __align__(128) uint8_t s_dyn[];
    CclStage<L4> *stage = reinterpret_cast<CclStage<L4> *>(s_dyn);        // [2]
    __shared__ __align__(8) uint64_t s_bar[2];
    const int t = threadIdx.x;
    if (t < 2) mbar_init(smem_u32(&s_bar[t]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const size_t slots = (size_t)NT * TILE_PX;
    auto issue = [&](int gt, int stg) {                 // thread 0: all copies of one (tile, frame) pair
        const int f = gt / NT, tile = gt - f * NT;
        const size_t wo = (size_t)f * MS + (size_t)tile * TILE_WORDS;
        const uint32_t bar = smem_u32(&s_bar[stg]);
        CclStage<L4> &S = stage[stg];
        const bool halo = tile > 0, bot = L4 && tile + 1 < NT;
        mbar_expect_tx(bar, (uint32_t)(TILE_WORDS * 4 + TILE_WORDS * 2 + CCL_VPS * 4 + (halo ? CCL_HALO * 4 : 0) +
                                       (bot ? CCL_HALO * 4 : 0)));
        if (halo) bulk_g2s(smem_u32(S.maskx), maps + wo - CCL_HALO, CCL_HALO * 4, bar);
        bulk_g2s(smem_u32(S.maskx + CCL_HALO), maps + wo, TILE_WORDS * 4, bar);
        bulk_g2s(smem_u32(S.wpre), wordpre_all + wo, TILE_WORDS * 2, bar);
        bulk_g2s(smem_u32(S.vp), vp_all + (size_t)f * slots + ((size_t)tile << TILE_LOG2), CCL_VPS * 4, bar);
        if (bot) bulk_g2s(smem_u32(S.bot), maps + wo + TILE_WORDS, CCL_HALO * 4, bar);
    };
    int gt = blockIdx.x;
    if (gt >= n_tiles_total) return;
    if (t == 0) issue(gt, 0);
    uint32_t total = tilecnt[gt];
    // (frame, tile) of gt, advanced by the grid stride without a division per tile
    int f = gt / NT, tile = gt - f * NT;
    const int step_f = (int)gridDim.x / NT, step_t = (int)gridDim.x - step_f * NT;
    for (int j = 0; gt < n_tiles_total; gt += gridDim.x, j++) {
        const int stg = j & 1;
        const int nxt = gt + (int)gridDim.x;
        uint32_t total_next = 0;
        if (nxt < n_tiles_total) {
            // the other stage was consumed by the previous iteration (barrier at the end of the loop body)
            if (t == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue(nxt, stg ^ 1);
            }
            total_next = tilecnt[nxt];
        }
        mbar_wait(smem_u32(&s_bar[stg]), (uint32_t)(j >> 1) & 1u);
        CclStage<L4> &S = stage[stg];
        // rows before the frame / after it read as background
        if (tile == 0 && t < CCL_HALO / 4) reinterpret_cast<uint4 *>(S.maskx)[t] = make_uint4(0, 0, 0, 0);
        if (L4 && tile + 1 >= NT && t < CCL_HALO / 4) reinterpret_cast<uint4 *>(S.bot)[t] = make_uint4(0, 0, 0, 0);
        ccl_tile<FOLD, CCL_THREADS>(tile, f, S, total, NT, MS, vp_all, tileovf, xcount, xlinks, parent_all, acc_all,
                       ny, nx, l4mode, bbox_all, map2_all, cent_all, rootcnt);
        __syncthreads();                               // the next tile reuses the shared arrays
        total = total_next;
        tile += step_t;
        f += step_f;
        if (tile >= NT) { tile -= NT; f++; }
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

// L4: bounding box of a tile-local root that lost its root status -> merged into the box of its final root,
// exactly once (same claim as FoldAct).  Boxes are {row min, row max, col min, col max}.
struct BboxFoldAct {
    uint32_t *parent, *bbox;
    __device__ __forceinline__ void one(uint32_t x) const
    {
        const uint32_t px = parent[x];
        const uint32_t r = (px & UF_FLAG) ? (px & ~UF_FLAG) : x;     // tile-local root of x
        const uint32_t pr = parent[r];
        if (pr == r || (pr & UF_FLAG)) return;                       // still a root, or already merged
        if (atomicOr(&parent[r], UF_FLAG) & UF_FLAG) return;
        const uint32_t g = uf_find_ro(parent, r);
        const uint4 b = reinterpret_cast<const uint4 *>(bbox)[r];
        atomicMin(&bbox[4 * (size_t)g + 0], b.x);
        atomicMax(&bbox[4 * (size_t)g + 1], b.y);
        atomicMin(&bbox[4 * (size_t)g + 2], b.z);
        atomicMax(&bbox[4 * (size_t)g + 3], b.w);
    }
    __device__ __forceinline__ void operator()(uint32_t a, uint32_t b) const { one(a); one(b); }
};

// Links across tile boundaries.  k_ccl_tiles labelled every tile on its own and listed the links from its
// first rows to pixels of earlier tiles; a tile that overflowed the shared-memory labelling (tileovf) has all
// of its links made here instead, with the word-parallel global path.
// PASS 0: union.  PASS 1: L2 fold of the re-parented tile-local roots (separate launch: needs final roots).
// PASS 2: L4 merge of their bounding boxes (acc_all is then the box array).  One warp per tile.
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
            else if (PASS == 1) FoldAct{parent, acc, sum}(e.x, sq);
            else BboxFoldAct{parent, acc_all + (size_t)f * slots * 4}(e.x, sq);
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
        else if (PASS == 1) link_word(sp, FoldAct{parent, acc, sum}, w, bits, ny, nx, 0xffffffffu);
        else link_word(sp, BboxFoldAct{parent, acc_all + (size_t)f * slots * 4}, w, bits, ny, nx, 0xffffffffu);
    }
}

// rc_ccl_label: afterwards every parent entry is the plain root slot
__global__ void __launch_bounds__(256)
k_ccl_flatten(const uint32_t *__restrict__ maps, size_t MS, const uint16_t *__restrict__ wordpre_all, int NT,
              uint32_t *__restrict__ parent_all, int ny, int nx, uint32_t MW)
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
    while (todo) {
        const uint32_t k = __ffs(todo) - 1;
        todo &= todo - 1;
        const uint32_t s = sb + __popc(bits & ((1u << k) - 1u));
        const uint32_t root = uf_find_ro(parent, s);
        if (root != s) parent[s] = root;
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

// ---- L4: puddles that cross tile boundaries -----------------------------------------------------------
// After k_ccl_border<0> (links) and <2> (boxes) every open puddle has a final root g with the box of all its
// members.  It is reached through the cross-link lists (one warp per tile; all slots of an overflowed tile),
// claimed once through claim[g] and replayed in raster order over its box with the reference's arithmetic.
__global__ void __launch_bounds__(256)
k_l4_open(const uint32_t *__restrict__ maps, size_t MS, const uint16_t *__restrict__ wordpre_all, int NT,
          const uint32_t *__restrict__ tilecnt, const uint8_t *__restrict__ tileovf,
          const uint32_t *__restrict__ xcount, const uint2 *__restrict__ xlinks,
          const uint32_t *__restrict__ parent_all, uint32_t *__restrict__ claim_all,
          const uint32_t *__restrict__ bbox_all, const uint32_t *__restrict__ vp_all, int ny, int nx, int mode,
          uint32_t *__restrict__ map2_all, uint64_t *__restrict__ cent_all, uint32_t *__restrict__ rootcnt)
{
    const int tile = blockIdx.x * 8 + (threadIdx.x >> 5), f = blockIdx.y, lane = threadIdx.x & 31;
    if (tile >= NT) return;
    const size_t ti = (size_t)f * NT + tile;
    const size_t slots = (size_t)NT * TILE_PX;
    const uint32_t *map = maps + (size_t)f * MS;
    const uint16_t *wordpre = wordpre_all + (size_t)f * MS;
    const uint32_t *parent = parent_all + (size_t)f * slots;
    const uint32_t *vp = vp_all + (size_t)f * slots;
    uint32_t *claim = claim_all + (size_t)f * slots;
    const bool ovf = tileovf[ti] != 0;
    const uint32_t n = ovf ? tilecnt[ti] : xcount[ti];
    const uint2 *xl = xlinks + ti * CCL_XCAP;
    const uint32_t unx = (uint32_t)nx;
    for (uint32_t j = lane; j < n; j += 32) {
        const uint32_t a = ovf ? ((uint32_t)tile << TILE_LOG2) + j : xl[j].x;
        const uint32_t g = uf_find_ro(parent, a);
        if (atomicCAS(&claim[g], 0u, 1u) != 0u) continue;
        const uint4 bb = reinterpret_cast<const uint4 *>(bbox_all)[(size_t)f * slots + g];
        CentAcc ca;
        for (uint32_t r = bb.x; r <= bb.y; r++) {
            const uint32_t q0 = r * unx + bb.z, q1 = r * unx + bb.w;      // inclusive pixel range
            for (uint32_t ww = q0 >> 5; ww <= (q1 >> 5); ww++) {
                uint32_t mb = map[ww];
                if (ww == (q0 >> 5)) mb &= 0xffffffffu << (q0 & 31);
                if (ww == (q1 >> 5)) mb &= 0xffffffffu >> (31 - (q1 & 31));
                while (mb) {
                    const uint32_t kk = __ffs(mb) - 1;
                    mb &= mb - 1;
                    const uint32_t q = (ww << 5) + kk;
                    const uint32_t sq = slot_of(map, wordpre, q);
                    if (uf_find_ro(parent, sq) != g) continue;
                    ca.add(mode, r, q - r * unx, vp[sq] >> 16);
                }
            }
        }
        float fr, fc;
        ca.finish(mode, fr, fc);
        atomicAdd(&rootcnt[(size_t)f * NT + (g >> TILE_LOG2)], 1u);
        if (cent_all) cent_all[(size_t)f * slots + g] = (uint64_t)__float_as_uint(fr) | ((uint64_t)__float_as_uint(fc) << 32);
        uint32_t q;
        if (map2_all && centroid_pixel(fr, fc, ny, nx, q)) atomicOr(&map2_all[(size_t)f * MS + (q >> 5)], 1u << (q & 31));
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

// fold: 1 = L2 max, 2 = L2 sum, 3 = L4 (l4mode = centroiding method; closed puddles -> map2 / cent / rootcnt)
int launch_ccl_tiles(rc_ctx *ctx, const Geom &g, int fold, const uint32_t *maps, const uint16_t *wordpre,
                     const uint32_t *tilecnt, const uint32_t *vp, uint8_t *tileovf, uint32_t *xcount, void *xlinks,
                     uint32_t *parent, uint32_t *acc, int l4mode, uint32_t *bbox, uint32_t *map2, uint64_t *cent,
                     uint32_t *rootcnt, int F, cudaStream_t st)
{
    if (F <= 0) return 0;
    const int nt = F * g.NT;
    // CTAs per SM: RECODE_B200_CCL_CTAS, else few when several batches share the GPU (the rest of every SM is left to
    // the streaming kernel of the next batch), else what fits
    const int fit = fold == 3 ? 4 : 5;
    int per_sm = ctx->ccl_ctas_per_sm > 0 ? ctx->ccl_ctas_per_sm
                                          : ((ctx->use_priority == 0 || !ctx->pipelined) ? fit : 2);
    if (per_sm > fit) per_sm = fit;
    unsigned grid = (unsigned)nt;
    if ((unsigned)(per_sm * ctx->sm_count) < grid) grid = (unsigned)(per_sm * ctx->sm_count);
#define RC_CT(FO, TPB)                                                                                         \
    {                                                                                                          \
        constexpr size_t dyn = 2 * sizeof(CclStage<FO == 3>);                                                  \
        RC_CUDA(ctx, cudaFuncSetAttribute(k_ccl_tiles<FO, TPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn)); \
        k_ccl_tiles<FO, TPB><<<grid, TPB, dyn, st>>>(maps, g.MS, wordpre, g.NT, nt, tilecnt, vp, tileovf, xcount, \
                                                     (uint2 *)xlinks, parent, acc, g.ny, g.nx, l4mode, bbox, map2,  \
                                                     cent, rootcnt);                                           \
    }
    // threads per CTA: 128 halves the per-tile fixed work and wastes fewer lanes on a tile's few hundred pixels
    static const int tpb = getenv("RECODE_B200_CCL_TPB") ? atoi(getenv("RECODE_B200_CCL_TPB")) : 256;
    if (tpb == 128) {
        if (fold == 1) RC_CT(1, 128)
        else if (fold == 2) RC_CT(2, 128)
        else RC_CT(3, 128)
    } else {
        if (fold == 1) RC_CT(1, 256)
        else if (fold == 2) RC_CT(2, 256)
        else RC_CT(3, 256)
    }
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

// fold: 1 = + L2 max fold, 2 = + L2 sum fold (acc = statistics), 3 = + L4 box merge (acc = boxes)
int launch_ccl_border(rc_ctx *ctx, const Geom &g, int fold, const uint32_t *maps, const uint16_t *wordpre,
                      const uint8_t *tileovf, const uint32_t *xcount, const void *xlinks, uint32_t *parent,
                      uint32_t *acc, int F, cudaStream_t st)
{
    if (F <= 0) return 0;
    dim3 grid((unsigned)((g.NT + 7) / 8), F);
    k_ccl_border<0><<<grid, 256, 0, st>>>(maps, g.MS, wordpre, g.NT, tileovf, xcount, (const uint2 *)xlinks, parent, acc,
                                          g.ny, g.nx, (uint32_t)g.MW, 0);
    RC_LAUNCH_CHECK(ctx, "k_ccl_border<0>");
    if (fold == 3) {
        k_ccl_border<2><<<grid, 256, 0, st>>>(maps, g.MS, wordpre, g.NT, tileovf, xcount, (const uint2 *)xlinks, parent,
                                              acc, g.ny, g.nx, (uint32_t)g.MW, 0);
        RC_LAUNCH_CHECK(ctx, "k_ccl_border<2>");
    } else if (fold) {
        k_ccl_border<1><<<grid, 256, 0, st>>>(maps, g.MS, wordpre, g.NT, tileovf, xcount, (const uint2 *)xlinks, parent,
                                              acc, g.ny, g.nx, (uint32_t)g.MW, fold == 2);
        RC_LAUNCH_CHECK(ctx, "k_ccl_border<1>");
    }
    return 0;
}

int launch_l4_open(rc_ctx *ctx, const Geom &g, int mode, const uint32_t *maps, const uint16_t *wordpre,
                   const uint32_t *tilecnt, const uint8_t *tileovf, const uint32_t *xcount, const void *xlinks,
                   const uint32_t *parent, uint32_t *claim, const uint32_t *bbox, const uint32_t *vp, uint32_t *map2,
                   uint64_t *cent, uint32_t *rootcnt, int F, cudaStream_t st)
{
    if (F <= 0) return 0;
    dim3 grid((unsigned)((g.NT + 7) / 8), F);
    k_l4_open<<<grid, 256, 0, st>>>(maps, g.MS, wordpre, g.NT, tilecnt, tileovf, xcount, (const uint2 *)xlinks, parent,
                                    claim, bbox, vp, g.ny, g.nx, mode, map2, cent, rootcnt);
    RC_LAUNCH_CHECK(ctx, "k_l4_open");
    return 0;
}

int launch_ccl_flatten(rc_ctx *ctx, const Geom &g, const uint32_t *maps, const uint16_t *wordpre, uint32_t *parent,
                       int F, cudaStream_t st)
{
    if (F <= 0) return 0;
    dim3 grid((unsigned)((g.MW + 255) / 256), F);
    k_ccl_flatten<<<grid, 256, 0, st>>>(maps, g.MS, wordpre, g.NT, parent, g.ny, g.nx, (uint32_t)g.MW);
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
