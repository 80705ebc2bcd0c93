// deflate_chunk.cuh -- one deflate chunk (<= 16 KiB of input) encoded by one CTA of 256 threads.
//
// Replaces zlib.compress on the hot path (reference: pyrecode/recode_compressors.py:84-85, called from
// recode_writer.py:503-511,538-540).  Only the INFLATED payload has to match the reference, not the
// compressed bytes, so the encoder is designed for the GPU and for this data:
//
//   * LZ77 restricted to distance-1 matches (byte runs): binary maps are 85-99 % 0x00 bytes, and a run is
//     found with one compare per byte, in lock step across the 256 threads (64 input bytes per thread).
//   * one dynamic-Huffman block per chunk: 286-symbol histogram in shared memory, CTA-wide bitonic sort,
//     two-queue Huffman merge + zlib-style 15-bit length limiting + canonical codes by one thread.
//   * every chunk starts byte aligned and ends with an empty stored block (the Z_SYNC_FLUSH marker
//     00 00 FF FF), and never references bytes before its own start.  Chunks of a stream are therefore
//     independent: they are encoded by different CTAs, concatenated by byte copies, and can be found and
//     inflated in parallel again on the read side.  Stock zlib inflates the result (verified in tests).
//   * a chunk that would not shrink is emitted as a stored block.
//
// The per-thread phases are __host__ __device__ so that tests/csrc/deflate_host_test.cpp can run the very
// same code on the CPU (threads simulated by a loop per phase) against stock zlib.  The CPU build is test
// infrastructure only; the product calls the CUDA kernel in deflate.cu.
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
constexpr int DF_OUT_WORDS = DF_CHUNK / 4 + 64;     // compressed chunk never exceeds the stored form
constexpr int DF_SLOT_BYTES = DF_CHUNK + 64;        // scratch slot per chunk (multiple of 16)
constexpr int DF_MAX_HEADER_BITS = 17 + 19 * 3 + (286 + 2) * (7 + 7);   // loose bound

struct DeflateShared {
    uint32_t in32[DF_CHUNK / 4];      // transposed + swizzled: word k of thread t at [k*256 + (t ^ ((k>>2)<<3))]
    uint32_t out[DF_OUT_WORDS];       // bit stream, zero initialised
    uint32_t hist[DF_NSYM];           // symbol counts; reused as sort keys (count << 9 | symbol)
    uint32_t keys[512];               // sort buffer
    uint16_t code[DF_NSYM];           // bit-reversed canonical codes
    uint8_t len[DF_NSYM];             // code lengths
    uint32_t tbits[DF_THREADS];       // per-thread bit counts -> exclusive bit offsets
    uint32_t node_w[DF_NSYM];         // Huffman internal node weights
    uint16_t node_parent[2 * DF_NSYM];  // [0,288) leaves (sorted order), [288, 576) internal nodes
    uint8_t node_depth[DF_NSYM];
    uint32_t header_bits;             // bits of block header + tables
    uint32_t total_bits;              // bits of the whole dynamic block incl. EOB
    uint32_t out_bytes;               // final size of the chunk piece
    uint32_t adler_a, adler_b;        // sum b_i mod 65521, sum (clen - i) * b_i mod 65521
    uint32_t n_match;
    int stored;
};

DF_HD int df_in_index(int t, int k) { return k * DF_THREADS + (t ^ ((k >> 2) << 3)); }

// length -> (symbol, extra bits count, extra bits value); L in [3, 258]
DF_HD void df_len_code(int L, int &sym, int &ebits, int &eval)
{
    if (L == 258) { sym = 285; ebits = 0; eval = 0; return; }
    const int x = L - 3;
    int e = 0;
    if (x >= 8) {
        int hb = 0;
        for (int y = x; y > 1; y >>= 1) hb++;        // floor(log2 x); x < 256 -> at most 7 steps
        e = hb - 2;
    }
    sym = 257 + 4 * e + (x >> e);
    ebits = e;
    eval = x & ((1 << e) - 1);
}

// ---- sequential bit writer (single thread; header and trailer) ----------------------------------
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
    uint32_t r = 0;
    for (int i = 0; i < n; i++) { r = (r << 1) | (c & 1); c >>= 1; }
    return r;
}

// ---- tokenizer --------------------------------------------------------------------------------
// Walks the nbytes of thread t's segment; E.lit(c) / E.match(L) are called in stream order.
// prev = 0x100 means "no previous byte" (chunk start).
template <typename E>
DF_HD void df_tokenize(const DeflateShared &S, int t, int nbytes, E &em)
{
    if (nbytes <= 0) return;
    uint32_t prev = 0x100;
    if (t > 0) prev = S.in32[df_in_index(t - 1, DF_SEG_WORDS - 1)] >> 24;
    int run = 0;
    const int nw = (nbytes + 3) >> 2;
    for (int k = 0; k < nw; k++) {
        const uint32_t x = S.in32[df_in_index(t, k)];
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
        for (int bi = 0; bi < 4; bi++) {
            if (k * 4 + bi < nbytes) {
                const uint32_t c = (x >> (8 * bi)) & 0xffu;
                if (c == prev) {
                    run++;
                } else {
                    if (run >= 3) em.match(run);
                    else for (int i = 0; i < run; i++) em.lit(prev);
                    em.lit(c);
                    prev = c;
                    run = 0;
                }
            }
        }
    }
    if (run >= 3) em.match(run);
    else for (int i = 0; i < run; i++) em.lit(prev);
}

// ---- phase 0: load (device version lives in deflate.cu; this is the scalar form) -----------------
DF_HD void df_store_word(DeflateShared &S, int byte_off, uint32_t w)
{
    const int t = byte_off / DF_SEG, k = (byte_off % DF_SEG) >> 2;
    S.in32[df_in_index(t, k)] = w;
}

// ---- phase 1: histogram + adler partial sums -----------------------------------------------------
struct DfHistEmit {
    uint32_t *hist;
    uint32_t nmatch;
    DF_HD void lit(uint32_t c) { DF_ATOMIC_ADD(&hist[c], 1u); }
    DF_HD void match(int L)
    {
        int sym, eb, ev;
        df_len_code(L, sym, eb, ev);
        DF_ATOMIC_ADD(&hist[sym], 1u);
        nmatch++;
    }
};

DF_HD void df_phase_hist(DeflateShared &S, int t, int clen)
{
    int nbytes = clen - t * DF_SEG;
    if (nbytes > DF_SEG) nbytes = DF_SEG;
    if (nbytes <= 0) return;
    DfHistEmit em{S.hist, 0};
    df_tokenize(S, t, nbytes, em);
    if (em.nmatch) DF_ATOMIC_ADD(&S.n_match, em.nmatch);
    // adler partials: A = sum b, B = sum (clen - i) * b_i  (i = absolute index in chunk)
    uint32_t a = 0, b = 0;
    const int nw = (nbytes + 3) >> 2;
    for (int k = 0; k < nw; k++) {
        const uint32_t x = S.in32[df_in_index(t, k)];
        for (int bi = 0; bi < 4; bi++) {
            const int i = k * 4 + bi;
            if (i < nbytes) {
                const uint32_t c = (x >> (8 * bi)) & 0xffu;
                a += c;
                b += (uint32_t)(clen - (t * DF_SEG + i)) * c;   // <= 64 * 16384 * 255 < 2^32
            }
        }
    }
    DF_ATOMIC_ADD(&S.adler_a, a % 65521u);
    DF_ATOMIC_ADD(&S.adler_b, b % 65521u);
}

// ---- phase 2: Huffman construction (one thread) --------------------------------------------------
// keys[0..n_used) = (count << 9 | symbol) sorted ascending.  Produces S.len / S.code for the literal/length
// alphabet, writes the dynamic block header into S.out and sets S.header_bits.
//
// generic builder: sorted (weight, symbol) pairs -> code lengths limited to maxbits (zlib's gen_bitlen
// overflow repair, trees.c), using caller scratch.
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

DF_HD void df_phase_build(DeflateShared &S, int n_used)
{
    df_build_lengths(S.keys, n_used, 15, S.len, DF_NSYM, S.node_w, S.node_parent, S.node_depth);
    df_assign_codes(S.len, DF_NSYM, S.code);

    // ---- header ----
    int nlit = 286;
    while (nlit > 257 && S.len[nlit - 1] == 0) nlit--;
    const int ndist = 2;                       // like zlib: always two distance codes of one bit each
    uint8_t seq[286 + 2];
    for (int i = 0; i < nlit; i++) seq[i] = S.len[i];
    seq[nlit] = 1; seq[nlit + 1] = 1;
    const int nseq = nlit + ndist;

    // run-length encode with symbols 16 / 17 / 18 (RFC 1951 3.2.7); rl_sym/rl_ext hold the result
    uint8_t rl_sym[286 + 2], rl_ext[286 + 2];
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
    df_build_lengths(ckeys, cused, 7, cl_len, 19, S.node_w, S.node_parent, S.node_depth);
    df_assign_codes(cl_len, 19, cl_code);

    const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    int ncl = 19;
    while (ncl > 4 && cl_len[order[ncl - 1]] == 0) ncl--;

    DfBitWriter bw{S.out, 0};
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
    S.header_bits = bw.pos;
}

// ---- phase 3: per-thread bit counts ---------------------------------------------------------------
struct DfSizeEmit {
    const uint8_t *len;
    uint32_t bits;
    DF_HD void lit(uint32_t c) { bits += len[c]; }
    DF_HD void match(int L)
    {
        int sym, eb, ev;
        df_len_code(L, sym, eb, ev);
        bits += len[sym] + eb + 1;        // + one-bit distance code (distance 1 = symbol 0, no extra bits)
    }
};

DF_HD void df_phase_size(DeflateShared &S, int t, int clen)
{
    int nbytes = clen - t * DF_SEG;
    if (nbytes > DF_SEG) nbytes = DF_SEG;
    DfSizeEmit em{S.len, 0};
    if (nbytes > 0) df_tokenize(S, t, nbytes, em);
    S.tbits[t] = em.bits;
}

// ---- phase 4: emit ---------------------------------------------------------------------------------
struct DfBitEmit {
    uint32_t *out;
    const uint16_t *code;
    const uint8_t *len;
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
    DF_HD void lit(uint32_t c) { add(code[c], len[c]); }
    DF_HD void match(int L)
    {
        int sym, eb, ev;
        df_len_code(L, sym, eb, ev);
        // length code, extra bits, then distance symbol 0 = code '0' (1 bit)
        add((uint32_t)code[sym] | ((uint32_t)ev << len[sym]), len[sym] + eb + 1);
    }
    DF_HD void flush() { if (nb) DF_ATOMIC_OR(&out[wpos], (uint32_t)acc); }
};

// S.tbits[t] must hold the exclusive prefix (bit offset relative to header end)
DF_HD void df_phase_emit(DeflateShared &S, int t, int clen)
{
    int nbytes = clen - t * DF_SEG;
    if (nbytes > DF_SEG) nbytes = DF_SEG;
    if (nbytes <= 0) return;
    const uint32_t o = S.header_bits + S.tbits[t];
    DfBitEmit em{S.out, S.code, S.len, 0, o & 31, o >> 5};
    df_tokenize(S, t, nbytes, em);
    em.flush();
}

// ---- phase 5: trailer (one thread) -------------------------------------------------------------------
// body_bits = header + all tokens.  Appends EOB and the sync-flush marker; sets out_bytes.
DF_HD void df_phase_finish(DeflateShared &S, uint32_t body_bits)
{
    DfBitWriter bw{S.out, body_bits};
    bw.put(S.code[256], S.len[256]);      // end of block
    bw.put(0, 3);                         // BFINAL=0, BTYPE=00: empty stored block
    bw.pos = (bw.pos + 7) & ~7u;          // pad to a byte boundary (zero bits)
    bw.put(0x0000, 16);                   // LEN = 0
    bw.put(0xffff, 16);                   // NLEN
    S.out_bytes = bw.pos >> 3;
}

// size in bytes of the dynamic form given the body bits (header + tokens)
DF_HD uint32_t df_dynamic_bytes(const DeflateShared &S, uint32_t body_bits)
{
    return ((body_bits + S.len[256] + 3 + 7) >> 3) + 4;
}
