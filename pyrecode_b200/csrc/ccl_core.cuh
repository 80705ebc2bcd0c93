// ccl_core.cuh -- union-find primitives and the per-word 8-connectivity linking shared by the tile-local
// labelling inside k_reduce_tiles (shared memory) and the global kernels of ccl.cu.
//
// Replaces scipy.ndimage.label(binary, structure=3x3 ones) (pyrecode/recode_writer.py:166,443).
//
// Union-find over foreground slots with "smaller index wins": the root of a puddle is its first pixel in
// raster order, which is scipy's label order.  A parent entry with UF_FLAG set is a pixel that was labelled
// inside its tile and points at its tile-local root; such entries are never modified again.
#pragma once
#include "common.cuh"

__device__ __forceinline__ uint32_t uf_find_ro(const uint32_t *parent, uint32_t x)
{
    // volatile: other threads lower parents concurrently; any value read is a valid ancestor
    uint32_t p = ((const volatile uint32_t *)parent)[x] & ~UF_FLAG;
    while (p != x) {
        x = p;
        p = ((const volatile uint32_t *)parent)[x] & ~UF_FLAG;
    }
    return x;
}

__device__ __forceinline__ void uf_union(uint32_t *parent, uint32_t a, uint32_t b)
{
    while (true) {
        a = uf_find_ro(parent, a);
        b = uf_find_ro(parent, b);
        if (a == b) return;
        if (a > b) { uint32_t t = a; a = b; b = t; }
        const uint32_t old = atomicMin(&parent[b], a);   // link the larger root under the smaller
        if (old == b) return;
        b = old;                                         // b was linked elsewhere meanwhile: merge with that
    }
}

// Neighbour masks of the 32 pixels of word w for nx % 32 == 0.  Bit k of `west` is set when pixel k's west
// neighbour is foreground, etc.  Rows above / below the frame and columns outside it contribute zeros.
struct Nbr {
    uint32_t west, east, n, nw, ne, s, sw, se;
};

// S provides: word(w) -> map word w of the frame (0 where S does not cover w); slot(q) -> slot of foreground
// pixel q; word_slot(w) -> slot of the first pixel of word w; parent -> union-find array indexed by those slots.
template <bool WITH_SOUTH, class S>
__device__ __forceinline__ Nbr neighbour_masks(const S &sp, uint32_t w, uint32_t bits, uint32_t wpr, uint32_t ny)
{
    Nbr m;
    const uint32_t row = (wpr & (wpr - 1u)) == 0 ? w >> (31 - __clz(wpr)) : w / wpr, wc = w - row * wpr;
    const bool has_l = wc > 0, has_r = wc + 1 < wpr;
    const uint32_t cl = has_l ? sp.word(w - 1) : 0, cr = has_r ? sp.word(w + 1) : 0;
    m.west = (bits << 1) | (cl >> 31);
    m.east = (bits >> 1) | (cr << 31);
    if (row > 0) {
        const uint32_t u = sp.word(w - wpr);
        const uint32_t ul = has_l ? sp.word(w - wpr - 1) : 0, ur = has_r ? sp.word(w - wpr + 1) : 0;
        m.n = u;
        m.nw = (u << 1) | (ul >> 31);
        m.ne = (u >> 1) | (ur << 31);
    } else {
        m.n = m.nw = m.ne = 0;
    }
    if (WITH_SOUTH && row + 1 < ny) {
        const uint32_t d = sp.word(w + wpr);
        const uint32_t dl = has_l ? sp.word(w + wpr - 1) : 0, dr = has_r ? sp.word(w + wpr + 1) : 0;
        m.s = d;
        m.sw = (d << 1) | (dl >> 31);
        m.se = (d >> 1) | (dr << 31);
    } else {
        m.s = m.sw = m.se = 0;
    }
    return m;
}

// Enumerates the links of every foreground pixel of word w (bits = sp.word(w) != 0) to its W / NW / N / NE
// neighbours q with q < q_hi and calls act(slot of the pixel, slot of the neighbour) for each.  Links that are
// implied by others (NW / NE when N is set) are skipped.  (A space that returns 0 for words it does not cover
// thereby also bounds q from below.)
template <class S, class A>
__device__ __forceinline__ void link_word(const S &sp, const A &act, uint32_t w, uint32_t bits, int ny, int nx,
                                          uint32_t q_hi)
{
    const uint32_t p0 = w << 5;
    if ((nx & 31) == 0) {
        const uint32_t wpr = (uint32_t)nx >> 5;
        const Nbr m = neighbour_masks<false>(sp, w, bits, wpr, (uint32_t)ny);
        uint32_t need = bits & (m.west | m.n | m.nw | m.ne);
        if (!need) return;
        const uint32_t sb = sp.word_slot(w);
        while (need) {
            const uint32_t k = __ffs(need) - 1;
            need &= need - 1;
            const uint32_t bk = 1u << k;
            const uint32_t s = sb + __popc(bits & (bk - 1u));
            const uint32_t p = p0 + k;
            if ((m.west & bk) && p - 1 < q_hi) act(s, k ? s - 1 : sp.slot(p - 1));
            if (m.n & bk) {
                // N is set: NW and NE are horizontally adjacent to N, their own W-links connect them
                if (p - nx < q_hi) act(s, sp.slot(p - nx));
            } else {
                if ((m.nw & bk) && p - nx - 1 < q_hi) act(s, sp.slot(p - nx - 1));
                if ((m.ne & bk) && p - nx + 1 < q_hi) act(s, sp.slot(p - nx + 1));
            }
        }
        return;
    }
    // generic geometry: per-pixel neighbour tests
    uint32_t s = sp.word_slot(w);
    uint32_t r = p0 / (uint32_t)nx, c = p0 - r * (uint32_t)nx;   // of bit 0; advanced incrementally
    uint32_t prev_k = 0, rest = bits;
    while (rest) {
        const uint32_t k = __ffs(rest) - 1;
        rest &= rest - 1;
        c += k - prev_k;
        prev_k = k;
        while (c >= (uint32_t)nx) { c -= nx; r++; }
        const uint32_t p = p0 + k;
        if (c > 0 && p - 1 < q_hi && sp.bit(p - 1)) act(s, sp.slot(p - 1));
        if (r > 0) {
            const uint32_t up = p - nx;
            if (sp.bit(up)) {
                if (up < q_hi) act(s, sp.slot(up));
            } else {
                if (c > 0 && up - 1 < q_hi && sp.bit(up - 1)) act(s, sp.slot(up - 1));
                if (c + 1 < (uint32_t)nx && up + 1 < q_hi && sp.bit(up + 1)) act(s, sp.slot(up + 1));
            }
        }
        s++;
    }
}

struct UnionAct {
    uint32_t *parent;
    __device__ __forceinline__ void operator()(uint32_t a, uint32_t b) const { uf_union(parent, a, b); }
};

// whole-frame space over the global arrays of one frame
struct GlobalSpace {
    const uint32_t *map;
    const uint16_t *wordpre;
    uint32_t *parent;
    __device__ __forceinline__ uint32_t word(uint32_t w) const { return map[w]; }
    __device__ __forceinline__ uint32_t bit(uint32_t q) const { return (map[q >> 5] >> (q & 31)) & 1u; }
    __device__ __forceinline__ uint32_t slot(uint32_t q) const { return slot_of(map, wordpre, q); }
    __device__ __forceinline__ uint32_t word_slot(uint32_t w) const { return word_slot_base(wordpre, w); }
};

// one tile in shared memory: words outside the tile read as 0, slots are tile-local ranks
struct TileSpace {
    const uint32_t *mask;      // [TILE_WORDS]
    const uint16_t *wpre;      // [TILE_WORDS]
    uint32_t *parent;          // [tile capacity]
    uint32_t w0;               // first word of the tile within the frame
    __device__ __forceinline__ uint32_t word(uint32_t w) const
    {
        const uint32_t i = w - w0;                       // wraps for w < w0
        return i < (uint32_t)TILE_WORDS ? mask[i] : 0u;
    }
    __device__ __forceinline__ uint32_t bit(uint32_t q) const { return (word(q >> 5) >> (q & 31)) & 1u; }
    __device__ __forceinline__ uint32_t slot(uint32_t q) const
    {
        const uint32_t i = (q >> 5) - w0;
        return wpre[i] + __popc(mask[i] & ((1u << (q & 31)) - 1u));
    }
    __device__ __forceinline__ uint32_t word_slot(uint32_t w) const { return wpre[w - w0]; }
};
