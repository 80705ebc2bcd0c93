// deflate_chunk.cuh -- the GPU deflate encoder, written as per-thread phases.
//
// Replaces zlib.compress on the hot path (reference: pyrecode/recode_compressors.py:84-85, called from
// recode_writer.py:503-511,538-540).  Only the INFLATED payload has to match the reference, not the
// compressed bytes, so the encoder is designed for the GPU and for this data:
//
//   * input streams are cut into 16 KiB chunks; a chunk is encoded by one CTA of 256 threads, 64 input
//     bytes per thread, all threads in lock step.
//   * LZ77 restricted to distance-1 matches (byte runs): binary maps are 85-99 % 0x00 bytes; a run costs one
//     compare per byte, or one compare per 4 bytes on the all-equal-word fast path.
//   * ONE dynamic-Huffman code per stream: k_deflate_hist sums the token histogram of all chunks of a
//     stream, k_deflate_tables builds the code once (two-queue Huffman merge, zlib-style 15-bit length
//     limiting, canonical codes, RFC 1951 code-length header) and every chunk of the stream re-uses the
//     prebuilt header bits.  That removes all serial work from the per-chunk kernel.
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
constexpr int DF_NSYM = 288;                        // literal/length alphabet (286 used)
constexpr int DF_OUT_WORDS = DF_CHUNK / 4 + 64;     // a compressed chunk never exceeds the stored form
constexpr int DF_SLOT_BYTES = DF_CHUNK + 64;        // scratch slot per chunk (multiple of 16)
constexpr int DF_HDR_WORDS = 136;                   // 17 + 57 + 288 * 14 bits worst case

// per-stream code, built once by k_deflate_tables and read by every chunk of the stream
struct DeflateTable {
    uint16_t code[DF_NSYM];           // bit-reversed canonical codes
    uint8_t len[DF_NSYM];             // code lengths
    uint32_t header[DF_HDR_WORDS];    // BFINAL=0/BTYPE=10 + code description, LSB first
    uint32_t header_bits;
    uint32_t pad[3];
};

// chunk staging: word k of thread t at [k*256 + (t ^ ((k>>2)<<3))] (transposed + swizzled: conflict-free both
// for the coalesced 128-bit fill and for the per-thread word reads)
DF_HD int df_in_index(int t, int k) { return k * DF_THREADS + (t ^ ((k >> 2) << 3)); }

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
    return (in32[df_in_index(t, pos >> 2)] >> (8 * (pos & 3))) & 0xffu;
}

// bit i (i < 4) set when byte i of x differs from byte i of y
DF_HD uint32_t df_ne_nibble(uint32_t x, uint32_t y)
{
    const uint32_t e = x ^ y;
    uint32_t z = (e & 0x7f7f7f7fu) + 0x7f7f7f7fu;      // bit 7 of each byte: low 7 bits non-zero
    z = (z | e) & 0x80808080u;                         // bit 7 set <=> byte of e non-zero
    return ((z >> 7) * 0x00204081u >> 21) & 0xfu;     // gather bits 0, 8, 16, 24 -> bits 0..3
}

// Token masks of thread t's segment (bit i = byte i of the segment):
//   lit  bytes emitted as literals: bytes that differ from their predecessor, and members of runs shorter than 3
//   lng  members of runs (>= 3 bytes equal to the byte before the run): covered by a distance-1 match
//   ms   first byte of each such run = where the match token is emitted; its length is the run of lng bits
// All of it is branch-free SIMD-in-register work on the 64-bit "same as predecessor" mask, so the 32 lanes of
// a warp stay converged; the only data-dependent loop left is one iteration per TOKEN (df_for_tokens).
struct DfMasks {
    uint64_t lit, ms, lng;
};

DF_HD int df_ctz64(uint64_t x)
{
#ifdef __CUDA_ARCH__
    return __ffsll((long long)x) - 1;
#else
    return __builtin_ctzll(x);
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
    uint64_t brk = ((uint64_t)bhi << 32) | blo;
    if (prev == 0x100) brk |= 1;                     // chunk start: byte 0 is always a literal
    const uint64_t valid = nbytes >= DF_SEG ? ~0ull : ((1ull << nbytes) - 1);
    const uint64_t nb = ~brk & valid;                // same as predecessor
    const uint64_t a = nb & (nb << 1) & (nb << 2);   // third or later member of a run
    DfMasks m;
    m.lng = (a | (a >> 1) | (a >> 2)) & nb;          // every member of a run of >= 3
    m.lit = valid & ~m.lng;
    m.ms = m.lng & ~(m.lng << 1);
    return m;
}

// E.lit(c) / E.match(L) in stream order, one loop iteration per token
template <typename E>
DF_HD void df_for_tokens(const uint32_t *in32, int t, const DfMasks &m, E &em)
{
    uint64_t tk = m.lit | m.ms;
    while (tk) {
        const int pos = df_ctz64(tk);
        tk &= tk - 1;
        if ((m.lit >> pos) & 1) {
            em.lit(df_byte_at(in32, t, pos));
        } else {
            const uint64_t r = ~(m.lng >> pos);
            em.match(r ? df_ctz64(r) : DF_SEG);
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

// ---- histogram + Adler-32 partials (k_deflate_hist) -----------------------------------------------------
struct DfHistEmit {
    uint32_t *hist;
    uint32_t n0;                      // literal 0x00 is a third of all tokens on binary maps: count it privately
    DF_HD void lit(uint32_t c)
    {
        if (c == 0) n0++;
        else DF_ATOMIC_ADD(&hist[c], 1u);
    }
    DF_HD void match(int L)
    {
        int sym, eb, ev;
        df_len_code(L, sym, eb, ev);
        DF_ATOMIC_ADD(&hist[sym], 1u);
    }
};

// hist: DF_NSYM counters (shared);  adler: {sum b_i, sum (clen - i) * b_i} mod 65521 (shared)
DF_HD void df_phase_hist(const uint32_t *in32, uint32_t *hist, uint32_t *adler, int t, int clen, bool tokens)
{
    const int nbytes = df_seg_bytes(t, clen);
    if (nbytes <= 0) return;
    if (tokens) {
        DfHistEmit em{hist, 0};
        df_tokenize(in32, t, nbytes, em);
        if (em.n0) DF_ATOMIC_ADD(&hist[0], em.n0);
    }
    uint32_t a = 0, b = 0;
    const int nw = (nbytes + 3) >> 2;
    for (int k = 0; k < nw; k++) {
        const uint32_t x = in32[df_in_index(t, k)];
        if (x == 0) continue;
        for (int bi = 0; bi < 4; bi++) {
            const int i = k * 4 + bi;
            if (i < nbytes) {
                const uint32_t c = (x >> (8 * bi)) & 0xffu;
                a += c;
                b += (uint32_t)(clen - (t * DF_SEG + i)) * c;   // <= 64 * 16384 * 255 < 2^32
            }
        }
    }
    if (a) {
        DF_ATOMIC_ADD(&adler[0], a % 65521u);
        DF_ATOMIC_ADD(&adler[1], b % 65521u);
    }
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
    uint32_t in32[DF_CHUNK / 4];
    uint32_t out[DF_OUT_WORDS];       // bit stream, zero initialised
    uint32_t cl[DF_NSYM];             // (length << 16) | bit-reversed code, literals and end-of-block
    uint32_t mt[DF_SEG + 1];          // per match length 3..64: (total bits << 24) | length code + extra + distance bit
    uint32_t tbits[DF_THREADS];       // per-thread bit counts -> exclusive bit offsets
    uint32_t header_bits;
    uint32_t out_bytes;
};

// fills S.cl / S.mt from the stream's table; thread i handles entries i, i + nthreads, ...
DF_HD void df_load_table(DfEmitShared &S, const DeflateTable &T, int i, int nthreads)
{
    for (int s = i; s < DF_NSYM; s += nthreads) S.cl[s] = ((uint32_t)T.len[s] << 16) | T.code[s];
    for (int L = 3 + i; L <= DF_SEG; L += nthreads) {
        int sym, eb, ev;
        df_len_code(L, sym, eb, ev);
        const uint32_t l = T.len[sym];
        // length code, extra bits, then distance symbol 0 = code '0' (1 bit, distance 1 has no extra bits)
        S.mt[L] = ((l + eb + 1) << 24) | (uint32_t)T.code[sym] | ((uint32_t)ev << l);
    }
}

struct DfSizeEmit {
    const uint32_t *cl, *mt;
    uint32_t bits;
    DF_HD void lit(uint32_t c) { bits += cl[c] >> 16; }
    DF_HD void match(int L) { bits += mt[L] >> 24; }
};

// masks computed once per thread (k_deflate_chunks keeps them in registers for the size and the emit pass)
DF_HD DfMasks df_phase_masks(const DfEmitShared &S, int t, int clen)
{
    const int nbytes = df_seg_bytes(t, clen);
    DfMasks m{0, 0, 0};
    if (nbytes > 0) m = df_token_masks(S.in32, t, nbytes);
    return m;
}

DF_HD void df_phase_size_m(DfEmitShared &S, int t, const DfMasks &m)
{
    DfSizeEmit em{S.cl, S.mt, 0};
    df_for_tokens(S.in32, t, m, em);
    S.tbits[t] = em.bits;
}

DF_HD void df_phase_size(DfEmitShared &S, int t, int clen)
{
    const DfMasks m = df_phase_masks(S, t, clen);
    df_phase_size_m(S, t, m);
}

struct DfBitEmit {
    uint32_t *out;
    const uint32_t *cl, *mt;
    uint64_t acc;
    uint32_t nb, wpos;
    DF_HD void add(uint32_t bits, int n)
    {
        acc |= (uint64_t)bits << nb;
        nb += n;
        if (nb >= 32) {
            DF_ATOMIC_OR(&out[wpos], (uint32_t)acc);
            acc >>= 32;
            nb -= 32;
            wpos++;
        }
    }
    DF_HD void lit(uint32_t c) { const uint32_t e = cl[c]; add(e & 0xffffu, e >> 16); }
    DF_HD void match(int L) { const uint32_t e = mt[L]; add(e & 0xffffffu, e >> 24); }
    DF_HD void flush() { if (nb) DF_ATOMIC_OR(&out[wpos], (uint32_t)acc); }
};

// S.tbits[t] must hold the exclusive prefix (bit offset relative to the header end)
DF_HD void df_phase_emit_m(DfEmitShared &S, int t, const DfMasks &m)
{
    if (!(m.lit | m.ms)) return;
    const uint32_t o = S.header_bits + S.tbits[t];
    DfBitEmit em{S.out, S.cl, S.mt, 0, o & 31, o >> 5};
    df_for_tokens(S.in32, t, m, em);
    em.flush();
}

DF_HD void df_phase_emit(DfEmitShared &S, int t, int clen)
{
    const DfMasks m = df_phase_masks(S, t, clen);
    df_phase_emit_m(S, t, m);
}

// body_bits = header + all tokens.  Appends EOB and the sync-flush marker; sets out_bytes.  Single thread.
DF_HD void df_phase_finish(DfEmitShared &S, uint32_t body_bits)
{
    DfBitWriter bw{S.out, body_bits};
    bw.put(S.cl[256] & 0xffffu, S.cl[256] >> 16);      // end of block
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
