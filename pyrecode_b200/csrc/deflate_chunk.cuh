// deflate_chunk.cuh -- the GPU deflate encoder, written as per-thread phases.
//
// Replaces zlib.compress on the hot path (reference: pyrecode/recode_compressors.py:84-85, called from
// recode_writer.py:503-511,538-540).  Only the INFLATED payload has to match the reference, not the
// compressed bytes, so the encoder is designed for the GPU and for this data:
//
//   * input streams are cut into 16 KiB chunks; a chunk is encoded by one CTA of 256 threads, 64 input
//     bytes per thread.  A thread tokenizes its segment ONCE, writing the code bits into a private staging
//     area; a block scan of the bit counts then places the 256 private bit strings (funnel shift + atomicOr).
//   * LZ77 restricted to distance-1 matches (byte runs): binary maps are 85-99 % 0x00 bytes; a run costs one
//     compare per byte, or one compare per 4 bytes on the all-equal-word fast path.
//   * the dynamic-Huffman code is built ONCE, not per chunk: k_deflate_hist sums a token histogram,
//     k_deflate_tables builds the code (two-queue Huffman merge, zlib-style 15-bit length limiting, canonical
//     codes, RFC 1951 code-length header) and every chunk re-uses the prebuilt header bits, which removes all
//     serial work from the per-chunk kernel.  compression_level 1..5: one code per batch of streams from a
//     1-in-8 sample of their chunks, every producible symbol smoothed to a non-zero count (frames of one
//     acquisition share their statistics).  Levels 6..9: one code per stream from all of its chunks.
//   * every chunk is its own deflate block, starts byte aligned, ends with an empty stored block (the
//     Z_SYNC_FLUSH marker 00 00 FF FF) and never references bytes before its own start.  Chunks are therefore
//     independent: encoded by different CTAs, concatenated by byte copies, and found and inflated in parallel
//     again on the read side.  Stock zlib inflates the result (verified in tests).
//   * a chunk that would not shrink is emitted as a stored block.
//
// The phases are __host__ __device__ so that tests/csrc/codec_host_test.cpp can run the very same code on the
// CPU (threads simulated by a loop per phase) against stock zlib.  The CPU build is test infrastructure only;
// the product calls the CUDA kernels in deflate.cu.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define DF_HD __host__ __device__ __forceinline__
#else
#define DF_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define DF_ATOMIC_ADD(p, v) atomicAdd((p), (v))
#define DF_ATOMIC_OR(p, v) atomicOr((p), (v))
#else
#define DF_ATOMIC_ADD(p, v) (*(p) += (v))
#define DF_ATOMIC_OR(p, v) (*(p) |= (v))
#endif

constexpr int DF_THREADS = 256;
constexpr int DF_SEG = 64;                          // input bytes per thread
constexpr int DF_CHUNK = DF_THREADS * DF_SEG;       // 16384
constexpr int DF_SEG_WORDS = DF_SEG / 4;            // 16
constexpr int DF_SEG_STRIDE = DF_SEG_WORDS + 1;     // staging stride in words: 17 -> bank-conflict free per-thread access
constexpr int DF_STAGE_WORDS = DF_THREADS * DF_SEG_STRIDE;   // 4352 words = 17408 bytes
constexpr int DF_NSYM = 288;                        // literal/length alphabet (286 used)
constexpr int DF_LEN_SYMS = 286;                    // symbols a chunk can produce: 0..255, 256, 257..285 (length <= 258)
constexpr int DF_MAX_MATCH = 258;
constexpr int DF_ZRUN = DF_MAX_MATCH + 1;           // zero bytes one composite run token covers: literal 0 + match(258)
constexpr int DF_SLOT_BYTES = DF_CHUNK + 64;        // scratch slot per chunk (multiple of 16)
constexpr int DF_HDR_WORDS = 136;                   // 17 + 57 + 288 * 14 bits worst case
constexpr int DF_NTBL = DF_NSYM + DF_MAX_MATCH + 1; // token table: literals / EOB, then match lengths 0..258

// per-stream (or per-group) code, built once by k_deflate_tables and read by every chunk that uses it
struct DeflateTable {
    uint16_t code[DF_NSYM];           // bit-reversed canonical codes
    uint8_t len[DF_NSYM];             // code lengths
    uint32_t header[DF_HDR_WORDS];    // BFINAL=0/BTYPE=10 + code description, LSB first
    uint32_t header_bits;
    uint32_t pad[3];
};

// chunk staging: thread t's 64-byte segment occupies words [17 t, 17 t + 16) (one pad word per segment)
DF_HD int df_in_index(int t, int k) { return t * DF_SEG_STRIDE + k; }

DF_HD void df_store_word(uint32_t *in32, int byte_off, uint32_t w)
{
    const int t = byte_off / DF_SEG, k = (byte_off % DF_SEG) >> 2;
    in32[df_in_index(t, k)] = w;
}

// length -> (symbol, extra bits count, extra bits value); L in [3, 258]
DF_HD void df_len_code(int L, int &sym, int &ebits, int &eval)
{
    if (L == 258) { sym = 285; ebits = 0; eval = 0; return; }
    const int x = L - 3;
    int e = 0;
    if (x >= 8) {
#ifdef __CUDA_ARCH__
        e = 29 - __clz(x);                          // floor(log2 x) - 2
#else
        int hb = 0;
        for (int y = x; y > 1; y >>= 1) hb++;
        e = hb - 2;
#endif
    }
    sym = 257 + 4 * e + (x >> e);
    ebits = e;
    eval = x & ((1 << e) - 1);
}

// ---- sequential bit writer (single thread) ----------------------------------------------------------
struct DfBitWriter {
    uint32_t *buf;
    uint32_t pos;
    DF_HD void put(uint32_t bits, int n)
    {
        if (n == 0) return;
        const uint32_t w = pos >> 5, sh = pos & 31;
        buf[w] |= bits << sh;
        if (sh + n > 32) buf[w + 1] |= bits >> (32 - sh);
        pos += n;
    }
};

DF_HD uint32_t df_bitrev(uint32_t c, int n)
{
#ifdef __CUDA_ARCH__
    return n ? (__brev(c) >> (32 - n)) : 0;
#else
    uint32_t r = 0;
    for (int i = 0; i < n; i++) { r = (r << 1) | (c & 1); c >>= 1; }
    return r;
#endif
}

// ---- tokenizer ----------------------------------------------------------------------------------------
// E.lit(c) / E.match(L) are called in stream order for the nbytes of thread t's segment.
// A run never crosses a segment, and the first byte of a chunk is always a literal (chunk independence).
DF_HD uint32_t df_byte_at(const uint32_t *in32, int t, int pos)
{
    return reinterpret_cast<const uint8_t *>(in32)[t * (DF_SEG_STRIDE * 4) + pos];   // little-endian words
}

// bit i (i < 4) set when byte i of x differs from byte i of y
DF_HD uint32_t df_ne_nibble(uint32_t x, uint32_t y)
{
    const uint32_t e = x ^ y;
    uint32_t z = (e & 0x7f7f7f7fu) + 0x7f7f7f7fu;      // bit 7 of each byte: low 7 bits non-zero
    z = (z | e) & 0x80808080u;                         // bit 7 set <=> byte of e non-zero
    return ((z >> 7) * 0x00204081u >> 21) & 0xfu;     // gather bits 0, 8, 16, 24 -> bits 0..3
}

// Token masks of thread t's segment (bit i = byte i of the segment; [0] = bytes 0..31, [1] = bytes 32..63):
//   lit  bytes emitted as literals: bytes that differ from their predecessor, and members of runs shorter than 3
//   lng  members of runs (>= 3 bytes equal to the byte before the run): covered by a distance-1 match
//   ms   first byte of each such run = where the match token is emitted; its length is the run of lng bits
// All of it is branch-free SIMD-in-register work on the "same as predecessor" mask, so the 32 lanes of a warp
// stay converged; the only data-dependent loop left is one iteration per TOKEN.  Everything is kept in 32-bit
// halves: 64-bit shifts and find-first-set cost several instructions each on the GPU.
struct DfMasks {
    uint32_t lit[2], ms[2], lng[2];
};

DF_HD int df_ctz32(uint32_t x)                       // 32 for x == 0
{
#ifdef __CUDA_ARCH__
    return __clz(__brev(x));
#else
    return x ? __builtin_ctz(x) : 32;
#endif
}

DF_HD DfMasks df_token_masks(const uint32_t *in32, int t, int nbytes)
{
    uint32_t prev = 0x100;                           // "no previous byte": a chunk never looks behind its start
    if (t > 0) prev = in32[df_in_index(t - 1, DF_SEG_WORDS - 1)] >> 24;
    uint32_t blo = 0, bhi = 0;                       // break mask: byte differs from its predecessor
    uint32_t carry = prev & 0xffu;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int k = 0; k < DF_SEG_WORDS; k++) {
        const uint32_t x = in32[df_in_index(t, k)];
        const uint32_t sh = (x << 8) | carry;         // each byte's predecessor
        const uint32_t nz = df_ne_nibble(x, sh);
        if (k < 8) blo |= nz << (4 * k); else bhi |= nz << (4 * (k - 8));
        carry = x >> 24;
    }
    if (prev == 0x100) blo |= 1;                     // chunk start: byte 0 is always a literal
    const uint32_t vlo = nbytes >= 32 ? 0xffffffffu : ((1u << nbytes) - 1u);
    const uint32_t vhi = nbytes >= 64 ? 0xffffffffu : (nbytes > 32 ? ((1u << (nbytes - 32)) - 1u) : 0u);
    const uint32_t nlo = ~blo & vlo, nhi = ~bhi & vhi;            // same as predecessor
    // third or later member of a run: nb & (nb << 1) & (nb << 2) over the 64-bit mask
    const uint32_t alo = nlo & (nlo << 1) & (nlo << 2);
    const uint32_t ahi = nhi & ((nhi << 1) | (nlo >> 31)) & ((nhi << 2) | (nlo >> 30));
    DfMasks m;
    // every member of a run of >= 3: (a | a >> 1 | a >> 2) & nb
    m.lng[0] = (alo | (alo >> 1) | (ahi << 31) | (alo >> 2) | (ahi << 30)) & nlo;
    m.lng[1] = (ahi | (ahi >> 1) | (ahi >> 2)) & nhi;
    m.lit[0] = vlo & ~m.lng[0];
    m.lit[1] = vhi & ~m.lng[1];
    m.ms[0] = m.lng[0] & ~(m.lng[0] << 1);
    m.ms[1] = m.lng[1] & ~((m.lng[1] << 1) | (m.lng[0] >> 31));
    return m;
}

// Length of the run of lng bits that starts at bit pos of half h.  cont = number of lng bits at the bottom of
// the upper half (a run that reaches the end of the lower half continues there).
DF_HD uint32_t df_run_length(const DfMasks &m, int h, int pos, uint32_t cont)
{
    const uint32_t l0 = (uint32_t)df_ctz32(~(m.lng[h] >> pos));   // the NOT sets the pos vacated top bits
    return l0 + ((h == 0 && (uint32_t)pos + l0 == 32u) ? cont : 0u);
}

// E.lit(c) / E.match(L) in stream order, one loop iteration per token
template <typename E>
DF_HD void df_for_tokens(const uint32_t *in32, int t, const DfMasks &m, E &em)
{
    const uint32_t cont = (uint32_t)df_ctz32(~m.lng[1]);
    for (int h = 0; h < 2; h++) {
        uint32_t tk = m.lit[h] | m.ms[h];
        while (tk) {
            const int pos = df_ctz32(tk);
            tk &= tk - 1;
            if ((m.lit[h] >> pos) & 1) em.lit(df_byte_at(in32, t, 32 * h + pos));
            else em.match((int)df_run_length(m, h, pos, cont));
        }
    }
}

template <typename E>
DF_HD void df_tokenize(const uint32_t *in32, int t, int nbytes, E &em)
{
    if (nbytes <= 0) return;
    const DfMasks m = df_token_masks(in32, t, nbytes);
    df_for_tokens(in32, t, m, em);
}

DF_HD int df_seg_bytes(int t, int clen)
{
    int nbytes = clen - t * DF_SEG;
    return nbytes > DF_SEG ? DF_SEG : nbytes;
}

// ---- histogram (k_deflate_hist) -----------------------------------------------------------------------------
// hist: DF_NSYM counters (shared).  Order does not matter here, so literals and matches are counted in two
// loops whose bodies have no divergent branch.
DF_HD void df_phase_hist(const uint32_t *in32, uint32_t *hist, int t, int clen)
{
    const int nbytes = df_seg_bytes(t, clen);
    if (nbytes <= 0) return;
    const DfMasks m = df_token_masks(in32, t, nbytes);
    const uint32_t cont = (uint32_t)df_ctz32(~m.lng[1]);
    uint32_t n0 = 0;                                 // literal 0x00 is frequent on binary maps: count it privately
    for (int h = 0; h < 2; h++) {
        uint32_t lit = m.lit[h];
        while (lit) {
            const int pos = df_ctz32(lit);
            lit &= lit - 1;
            const uint32_t c = df_byte_at(in32, t, 32 * h + pos);
            if (c == 0) n0++;
            else DF_ATOMIC_ADD(&hist[c], 1u);
        }
    }
    if (n0) DF_ATOMIC_ADD(&hist[0], n0);
    for (int h = 0; h < 2; h++) {
        uint32_t ms = m.ms[h];
        while (ms) {
            const int pos = df_ctz32(ms);
            ms &= ms - 1;
            int sym, eb, ev;
            df_len_code((int)df_run_length(m, h, pos, cont), sym, eb, ev);
            DF_ATOMIC_ADD(&hist[sym], 1u);
        }
    }
}

// Adler-32 partials of thread t's segment: {sum b_i, sum (clen - i) * b_i}, i = position in the chunk
DF_HD void df_adler_partial(const uint32_t *in32, int t, int clen, uint32_t &a_out, uint32_t &b_out)
{
    const int nbytes = df_seg_bytes(t, clen);          // bytes past nbytes are staged as zeros
    uint32_t a = 0, b = 0;
    if (nbytes > 0) {
        const uint32_t w0 = (uint32_t)(clen - t * DF_SEG);          // weight of the segment's first byte
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
        for (int k = 0; k < DF_SEG_WORDS; k++) {
            const uint32_t x = in32[df_in_index(t, k)];
#ifdef __CUDA_ARCH__
            const uint32_t s = __dp4a(x, 0x01010101u, 0u);          // c0 + c1 + c2 + c3
            const uint32_t w = __dp4a(x, 0x03020100u, 0u);          // 0 c0 + 1 c1 + 2 c2 + 3 c3
#else
            const uint32_t c0 = x & 0xff, c1 = (x >> 8) & 0xff, c2 = (x >> 16) & 0xff, c3 = x >> 24;
            const uint32_t s = c0 + c1 + c2 + c3, w = c1 + 2 * c2 + 3 * c3;
#endif
            a += s;
            b += (w0 - 4u * (uint32_t)k) * s - w;                   // <= 64 * 16384 * 255 < 2^32
        }
    }
    a_out = a;
    b_out = b;
}

// ---- code construction (k_deflate_tables) ------------------------------------------------------------------
struct DfBuildShared {
    uint32_t keys[512];               // (count << 9 | symbol), sorted ascending; 0xffffffff = unused
    uint32_t node_w[DF_NSYM];         // Huffman internal node weights
    uint16_t node_parent[2 * DF_NSYM];  // [0,288) leaves (sorted order), [288, 576) internal nodes
    uint8_t node_depth[DF_NSYM];
    uint8_t seq[DF_NSYM + 2];         // code lengths to transmit
    uint8_t rl_sym[DF_NSYM + 2], rl_ext[DF_NSYM + 2];
    DeflateTable tab;
};

// sorted (weight, symbol) pairs -> code lengths limited to maxbits (zlib's gen_bitlen overflow repair, trees.c)
DF_HD void df_build_lengths(const uint32_t *keys, int n_used, int maxbits, uint8_t *len_out, int nsym,
                            uint32_t *node_w, uint16_t *node_parent, uint8_t *node_depth)
{
    for (int i = 0; i < nsym; i++) len_out[i] = 0;
    if (n_used == 0) return;
    if (n_used == 1) {                       // a lone symbol still needs one bit; add a dummy sibling
        const int s = keys[0] & 511;
        len_out[s] = 1;
        len_out[s == 0 ? 1 : 0] = 1;
        return;
    }
    // two-queue merge: leaves i (weight keys[i] >> 9), internal nodes j (weight node_w[j])
    int li = 0, ii = 0, ni = 0;
    for (int k = 0; k < n_used - 1; k++) {
        uint32_t w = 0;
        for (int pick = 0; pick < 2; pick++) {
            const bool leaf = li < n_used && (ii >= ni || (keys[li] >> 9) <= node_w[ii]);
            if (leaf) { w += keys[li] >> 9; node_parent[li] = (uint16_t)ni; li++; }
            else { w += node_w[ii]; node_parent[DF_NSYM + ii] = (uint16_t)ni; ii++; }
        }
        node_w[ni++] = w;
    }
    // depths top-down with clamping, counting overflow like zlib (every clamped node counts)
    uint32_t bl_count[16];
    for (int i = 0; i < 16; i++) bl_count[i] = 0;
    int overflow = 0;
    node_depth[ni - 1] = 0;
    for (int j = ni - 2; j >= 0; j--) {
        int d = node_depth[node_parent[DF_NSYM + j]] + 1;
        if (d > maxbits) { d = maxbits; overflow++; }
        node_depth[j] = (uint8_t)d;
    }
    for (int i = 0; i < n_used; i++) {
        int d = node_depth[node_parent[i]] + 1;
        if (d > maxbits) { d = maxbits; overflow++; }
        bl_count[d]++;
    }
    if (overflow > 0) {
        do {
            int bits = maxbits - 1;
            while (bl_count[bits] == 0) bits--;
            bl_count[bits]--;
            bl_count[bits + 1] += 2;
            bl_count[maxbits]--;
            overflow -= 2;
        } while (overflow > 0);
    }
    // least frequent leaves get the longest codes
    int i = 0;
    for (int bits = maxbits; bits >= 1; bits--)
        for (uint32_t c = 0; c < bl_count[bits]; c++) len_out[keys[i++] & 511] = (uint8_t)bits;
}

// canonical codes (RFC 1951 3.2.2), stored bit-reversed for LSB-first emission
DF_HD void df_assign_codes(const uint8_t *len, int nsym, uint16_t *code)
{
    uint32_t bl_count[16], next_code[16];
    for (int i = 0; i < 16; i++) bl_count[i] = 0;
    for (int s = 0; s < nsym; s++) bl_count[len[s]]++;
    bl_count[0] = 0;
    uint32_t c = 0;
    next_code[0] = 0;
    for (int bits = 1; bits <= 15; bits++) {
        c = (c + bl_count[bits - 1]) << 1;
        next_code[bits] = c;
    }
    for (int s = 0; s < nsym; s++) {
        const int l = len[s];
        code[s] = l ? (uint16_t)df_bitrev(next_code[l]++, l) : 0;
    }
}

// B.keys[0..n_used) sorted.  Fills B.tab (codes, lengths, header).  Single thread.
DF_HD void df_phase_build(DfBuildShared &B, int n_used)
{
    DeflateTable &T = B.tab;
    df_build_lengths(B.keys, n_used, 15, T.len, DF_NSYM, B.node_w, B.node_parent, B.node_depth);
    df_assign_codes(T.len, DF_NSYM, T.code);

    int nlit = 286;
    while (nlit > 257 && T.len[nlit - 1] == 0) nlit--;
    const int ndist = 2;                       // like zlib: always two distance codes of one bit each
    uint8_t *seq = B.seq;
    for (int i = 0; i < nlit; i++) seq[i] = T.len[i];
    seq[nlit] = 1; seq[nlit + 1] = 1;
    const int nseq = nlit + ndist;

    // run-length encode with symbols 16 / 17 / 18 (RFC 1951 3.2.7)
    uint8_t *rl_sym = B.rl_sym, *rl_ext = B.rl_ext;
    int nrl = 0;
    uint32_t cl_freq[19];
    for (int i = 0; i < 19; i++) cl_freq[i] = 0;
    for (int i = 0; i < nseq;) {
        const int v = seq[i];
        int run = 1;
        while (i + run < nseq && seq[i + run] == v) run++;
        if (v == 0) {
            int left = run;
            while (left >= 11) { const int r = left > 138 ? 138 : left; rl_sym[nrl] = 18; rl_ext[nrl++] = (uint8_t)(r - 11); left -= r; }
            if (left >= 3) { rl_sym[nrl] = 17; rl_ext[nrl++] = (uint8_t)(left - 3); left = 0; }
            while (left-- > 0) { rl_sym[nrl] = 0; rl_ext[nrl++] = 0; }
        } else {
            rl_sym[nrl] = (uint8_t)v; rl_ext[nrl++] = 0;
            int left = run - 1;
            while (left >= 3) { const int r = left > 6 ? 6 : left; rl_sym[nrl] = 16; rl_ext[nrl++] = (uint8_t)(r - 3); left -= r; }
            while (left-- > 0) { rl_sym[nrl] = (uint8_t)v; rl_ext[nrl++] = 0; }
        }
        i += run;
    }
    for (int i = 0; i < nrl; i++) cl_freq[rl_sym[i]]++;

    // code-length code: 19 symbols, 7 bits max; tiny insertion sort
    uint32_t ckeys[19];
    int cused = 0;
    for (int s = 0; s < 19; s++) if (cl_freq[s]) {
        const uint32_t key = (cl_freq[s] << 9) | (uint32_t)s;
        int j = cused++;
        while (j > 0 && ckeys[j - 1] > key) { ckeys[j] = ckeys[j - 1]; j--; }
        ckeys[j] = key;
    }
    uint8_t cl_len[19];
    uint16_t cl_code[19];
    df_build_lengths(ckeys, cused, 7, cl_len, 19, B.node_w, B.node_parent, B.node_depth);
    df_assign_codes(cl_len, 19, cl_code);

    const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    int ncl = 19;
    while (ncl > 4 && cl_len[order[ncl - 1]] == 0) ncl--;

    for (int i = 0; i < DF_HDR_WORDS; i++) T.header[i] = 0;
    DfBitWriter bw{T.header, 0};
    bw.put(0, 1);                 // BFINAL = 0 (the stream is closed by the assembler)
    bw.put(2, 2);                 // BTYPE = 10 dynamic
    bw.put((uint32_t)(nlit - 257), 5);
    bw.put((uint32_t)(ndist - 1), 5);
    bw.put((uint32_t)(ncl - 4), 4);
    for (int i = 0; i < ncl; i++) bw.put(cl_len[order[i]], 3);
    for (int i = 0; i < nrl; i++) {
        const int s = rl_sym[i];
        bw.put(cl_code[s], cl_len[s]);
        if (s == 16) bw.put(rl_ext[i], 2);
        else if (s == 17) bw.put(rl_ext[i], 3);
        else if (s == 18) bw.put(rl_ext[i], 7);
    }
    T.header_bits = bw.pos;
}

// ---- per-chunk emission (k_deflate_chunks) --------------------------------------------------------------------
struct DfEmitShared {
    uint32_t io[DF_STAGE_WORDS];      // staged input; after the encode pass it is reused as the output bit stream
    uint32_t priv[DF_STAGE_WORDS];    // private code bits of thread t at [17 t, 17 t + 17): 544 bits
    uint32_t tbl[DF_NTBL];            // (bits << 24) | code: literals / EOB at [sym], matches of length L at [288 + L]
    uint32_t tbits[DF_THREADS];       // per-thread bit counts -> exclusive bit offsets
    uint32_t ecode[DF_ZRUN + 1];      // sparse path: code bits of a run of g zero bytes, g = 0..259 (see df_load_table)
    uint8_t elen[DF_ZRUN + 1];        // ... and their number
    uint32_t header_bits;
    uint32_t out_bytes;
    uint32_t overflow;                // a thread's code bits did not fit its private area: store the chunk
    uint32_t e_bad;                   // a zero-run composite does not fit 32 bits / lacks a code: no sparse path
};
constexpr int DF_PRIV_BITS = DF_SEG_STRIDE * 32;

// fills S.tbl from the code table; thread i handles entries i, i + nthreads, ...
DF_HD void df_load_table(DfEmitShared &S, const DeflateTable &T, int i, int nthreads)
{
    for (int s = i; s < DF_NSYM; s += nthreads) S.tbl[s] = ((uint32_t)T.len[s] << 24) | T.code[s];
    for (int L = i; L <= DF_MAX_MATCH; L += nthreads) {
        uint32_t e = 0;
        if (L >= 3) {
            int sym, eb, ev;
            df_len_code(L, sym, eb, ev);
            const uint32_t l = T.len[sym];
            // length code, extra bits, then distance symbol 0 = code '0' (1 bit, distance 1 has no extra bits)
            e = ((l + eb + 1) << 24) | (uint32_t)T.code[sym] | ((uint32_t)ev << l);
        }
        S.tbl[DF_NSYM + L] = e;
    }
    // sparse path: a run of g zero bytes that follows a non-zero byte (or starts the chunk) is
    //   g = 1..3: g literals 0x00;   g = 4..259: literal 0x00 + a distance-1 match of g - 1
    // (what the tokenizer above produces for such a run), pre-merged into one code of at most 32 bits
    const uint32_t l0 = T.len[0], c0 = T.code[0];
    for (int g = i; g <= DF_ZRUN; g += nthreads) {
        uint32_t code = 0, len = 0;
        bool bad = l0 == 0;
        if (g >= 1 && g <= 3) {
            len = (uint32_t)g * l0;
            bad = bad || len > 32;
            if (!bad) for (int k = 0; k < g; k++) code |= c0 << (k * l0);
        } else if (g >= 4) {
            int sym, eb, ev;
            df_len_code(g - 1, sym, eb, ev);
            const uint32_t l = T.len[sym];
            len = l0 + l + (uint32_t)eb + 1u;
            bad = bad || l == 0 || len > 32;
            if (!bad) code = c0 | ((uint32_t)T.code[sym] << l0) | ((uint32_t)ev << (l0 + l));
        }
        S.ecode[g] = code;
        S.elen[g] = (uint8_t)len;
        if (bad && g > 0) S.e_bad = 1;
    }
}

// Tokenizes thread t's segment once: code bits go to the thread's private area, the bit count is returned.
// The loop body is branch-free apart from the 32-bit flush, so lanes only diverge in the trip count.
DF_HD uint32_t df_encode_segment(DfEmitShared &S, int t, int clen)
{
    const int nbytes = df_seg_bytes(t, clen);
    if (nbytes <= 0) return 0;
    const DfMasks m = df_token_masks(S.io, t, nbytes);
    const uint8_t *seg = reinterpret_cast<const uint8_t *>(S.io) + t * (DF_SEG_STRIDE * 4);
    uint32_t *priv = S.priv + t * DF_SEG_STRIDE;
    // Every byte of the segment belongs to exactly one token (a literal, or a match that covers a run up to the
    // literal that ends it), so a token's length is the distance to the next token start -- which the loop needs
    // anyway -- or to the end of the segment: no run-length scan per token.
    const uint32_t tk0 = m.lit[0] | m.ms[0], tk1 = m.lit[1] | m.ms[1];
    const uint32_t end = (uint32_t)nbytes;
    const uint32_t first1 = tk1 ? 32u + (uint32_t)df_ctz32(tk1) : end;     // first token start of the upper half
    uint32_t lo = 0, hi = 0;                         // bit accumulator: lo is flushed when nb reaches 32
    uint32_t nb = 0, w = 0, total = 0;
    for (int h = 0; h < 2; h++) {
        uint32_t tk = h ? tk1 : tk0;
        if (!tk) continue;
        const uint32_t lit = m.lit[h];
        const uint32_t lim = h ? end - 32u : first1; // where the half's last token ends (relative to the half)
        uint32_t pos = (uint32_t)df_ctz32(tk);
        while (true) {
            tk &= tk - 1;
            const uint32_t nxt = tk ? (uint32_t)df_ctz32(tk) : lim;
            const uint32_t idx = ((lit >> pos) & 1) ? (uint32_t)seg[32 * h + pos] : (uint32_t)DF_NSYM + (nxt - pos);
            const uint32_t e = S.tbl[idx];
            const uint32_t code = e & 0xffffffu;
            lo |= code << nb;
#ifdef __CUDA_ARCH__
            hi = __funnelshift_l(code, 0u, nb);      // bits of (code << nb) above bit 31; 0 for nb == 0
#else
            hi = nb ? code >> (32 - nb) : 0u;
#endif
            nb += e >> 24;
            total += e >> 24;
            if (nb >= 32) {
                if (w < (uint32_t)DF_SEG_STRIDE) priv[w] = lo;
                w++;
                lo = hi;
                nb -= 32;
            }
            if (!tk) break;
            pos = nxt;
        }
    }
    if (nb) {
        if (w < (uint32_t)DF_SEG_STRIDE) priv[w] = lo;
        w++;
    }
    if (w > (uint32_t)DF_SEG_STRIDE) S.overflow = 1;
    return total;
}

// ---- sparse chunks (binary maps: 85-99 % zero bytes) ---------------------------------------------------------
// A chunk with few non-zero bytes is encoded from the LIST of its non-zero bytes instead of byte by byte: every list
// entry is one literal preceded by the pre-merged code of the zero run in front of it (S.ecode), so the work is
// proportional to the non-zero bytes, a thread's share is a contiguous range of the list (every lane equally busy),
// and runs are not cut at 64-byte segments.  The tokens are a legal deflate stream for ANY input (a non-zero byte is
// always sent as a literal); the kernel takes this path only when the list fits (DF_STAGE_WORDS entries, i.e. up to a
// quarter of the bytes non-zero) and the code is the smoothed shared one (every length symbol has a code).
//
// bit i of nz[h] = byte 32 h + i of thread t's segment is non-zero
DF_HD uint32_t df_nz_masks(const uint32_t *in32, int t, int nbytes, uint32_t (&nz)[2])
{
    nz[0] = nz[1] = 0;
    if (nbytes <= 0) return 0;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int k = 0; k < DF_SEG_WORDS; k++) {
        const uint32_t x = in32[df_in_index(t, k)];                 // bytes past nbytes are staged as zeros
        uint32_t z = (x & 0x7f7f7f7fu) + 0x7f7f7f7fu;
        z = (z | x) & 0x80808080u;                                   // bit 7 of each non-zero byte
        const uint32_t nib = ((z >> 7) * 0x00204081u >> 21) & 0xfu;
        if (k < 8) nz[0] |= nib << (4 * k); else nz[1] |= nib << (4 * (k - 8));
    }
#ifdef __CUDA_ARCH__
    return (uint32_t)(__popc(nz[0]) + __popc(nz[1]));
#else
    return (uint32_t)(__builtin_popcount(nz[0]) + __builtin_popcount(nz[1]));
#endif
}

// list[base + k] = (position in the chunk << 8) | byte, in stream order
DF_HD void df_nz_scatter(const uint32_t *in32, uint32_t *list, int t, uint32_t base, const uint32_t (&nz)[2])
{
    for (int h = 0; h < 2; h++) {
        uint32_t m = nz[h];
        while (m) {
            const int pos = df_ctz32(m);
            m &= m - 1;
            list[base++] = ((uint32_t)(t * DF_SEG + 32 * h + pos) << 8) | df_byte_at(in32, t, 32 * h + pos);
        }
    }
}

// thread t's share of the n list entries: [lo, hi)
DF_HD void df_sparse_range(int t, uint32_t n, uint32_t &lo, uint32_t &hi)
{
    lo = (uint32_t)(((uint64_t)n * (uint32_t)t) / DF_THREADS);
    hi = (uint32_t)(((uint64_t)n * (uint32_t)(t + 1)) / DF_THREADS);
}

// code bits of a run of g zero bytes: whole composites of 259, then the composite of the rest
DF_HD uint32_t df_zrun_bits(const DfEmitShared &S, uint32_t g)
{
    if (g <= (uint32_t)DF_ZRUN) return S.elen[g];                  // the common case: no division
    const uint32_t rep = g / (uint32_t)DF_ZRUN;
    return rep * S.elen[DF_ZRUN] + S.elen[g - rep * (uint32_t)DF_ZRUN];
}

// code bits of thread t's entries (the zero bytes after the last entry of the chunk belong to the last thread) and
// the Adler-32 partials of its entries {sum b, sum (clen - i) b}: the zero bytes contribute nothing to either sum
DF_HD uint32_t df_sparse_bits(const DfEmitShared &S, const uint32_t *list, int t, uint32_t n, int clen,
                              uint32_t &adler_a, uint32_t &adler_b)
{
    uint32_t lo, hi;
    df_sparse_range(t, n, lo, hi);
    uint32_t bits = 0, a = 0, b = 0;
    uint32_t prev = lo ? (list[lo - 1] >> 8) + 1u : 0u;            // first byte not yet covered
    for (uint32_t k = lo; k < hi; k++) {
        const uint32_t e = list[k];
        const uint32_t pos = e >> 8, val = e & 0xffu;
        bits += df_zrun_bits(S, pos - prev) + (S.tbl[val] >> 24);
        prev = pos + 1u;
        a += val;
        b += ((uint32_t)clen - pos) * val;                          // <= 17 entries * 16384 * 255 < 2^32
    }
    if (t == DF_THREADS - 1) bits += df_zrun_bits(S, (uint32_t)clen - prev);
    adler_a = a;
    adler_b = b;
    return bits;
}

// ORs `n` (<= 47) code bits into the output stream at bit `pos`: at most three words
DF_HD void df_or_bits64(uint32_t *out, uint32_t &pos, uint32_t lo, uint32_t hi, uint32_t n)
{
    const uint32_t w = pos >> 5, sh = pos & 31;
    pos += n;
    uint32_t x0 = lo << sh, x1, x2 = 0;
    if (sh) {
        x1 = (lo >> (32 - sh)) | (hi << sh);
        x2 = hi >> (32 - sh);
    } else {
        x1 = hi;
    }
    if (x0) DF_ATOMIC_OR(&out[w], x0);
    if (x1) DF_ATOMIC_OR(&out[w + 1], x1);
    if (x2) DF_ATOMIC_OR(&out[w + 2], x2);
}

// ORs the codes of thread t's entries into the (zero-initialised) output bit stream from bit `pos`
DF_HD void df_sparse_emit(DfEmitShared &S, const uint32_t *list, int t, uint32_t n, int clen, uint32_t pos)
{
    uint32_t lo, hi;
    df_sparse_range(t, n, lo, hi);
    uint32_t prev = lo ? (list[lo - 1] >> 8) + 1u : 0u;
    for (uint32_t k = lo; k <= hi; k++) {
        uint32_t g, lit = 0;
        if (k < hi) {
            const uint32_t e = list[k];
            g = (e >> 8) - prev;
            prev = (e >> 8) + 1u;
            lit = S.tbl[e & 0xffu];
        } else {
            if (t != DF_THREADS - 1) break;
            g = (uint32_t)clen - prev;                                // the zero bytes that end the chunk
        }
        for (; g > (uint32_t)DF_ZRUN; g -= DF_ZRUN) df_or_bits64(S.io, pos, S.ecode[DF_ZRUN], 0u, S.elen[DF_ZRUN]);
        // the run's composite (<= 32 bits) and the literal (<= 15 bits) as one bit string
        const uint32_t el = S.elen[g], ec = S.ecode[g], lc = lit & 0xffffffu, ll = lit >> 24;
        uint32_t c_lo = ec, c_hi = 0;
        if (el < 32) { c_lo |= lc << el; c_hi = el ? lc >> (32 - el) : 0u; }
        else c_hi = lc;
        df_or_bits64(S.io, pos, c_lo, c_hi, el + ll);
    }
}

// ORs thread t's private bit string into the output stream at bit offset S.header_bits + S.tbits[t] (exclusive scan)
DF_HD void df_place_segment(DfEmitShared &S, int t, uint32_t nbits)
{
    if (!nbits) return;
    const uint32_t o = S.header_bits + S.tbits[t];
    const uint32_t sh = o & 31;
    uint32_t wpos = o >> 5;
    const uint32_t *priv = S.priv + t * DF_SEG_STRIDE;
    const uint32_t nw = (nbits + 31) >> 5;
    for (uint32_t j = 0; j < nw; j++, wpos++) {
        const uint32_t x = priv[j];
        DF_ATOMIC_OR(&S.io[wpos], x << sh);
        if (sh && (x >> (32 - sh))) DF_ATOMIC_OR(&S.io[wpos + 1], x >> (32 - sh));
    }
}

// body_bits = header + all tokens.  Appends EOB and the sync-flush marker; sets out_bytes.  Single thread.
DF_HD void df_phase_finish(DfEmitShared &S, uint32_t body_bits)
{
    DfBitWriter bw{S.io, body_bits};
    bw.put(S.tbl[256] & 0xffffffu, S.tbl[256] >> 24);  // end of block
    bw.put(0, 3);                         // BFINAL=0, BTYPE=00: empty stored block
    bw.pos = (bw.pos + 7) & ~7u;          // pad to a byte boundary (zero bits)
    bw.put(0x0000, 16);                   // LEN = 0
    bw.put(0xffff, 16);                   // NLEN
    S.out_bytes = bw.pos >> 3;
}

// size in bytes of the dynamic form given the body bits (header + tokens)
DF_HD uint32_t df_dynamic_bytes(uint32_t body_bits, int eob_len)
{
    return ((body_bits + eob_len + 3 + 7) >> 3) + 4;
}
