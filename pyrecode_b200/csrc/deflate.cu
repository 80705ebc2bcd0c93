// deflate.cu -- batched zlib-format deflate on the GPU and assembly of the ReCoDe frame records.
//
// Replaces zlib.compress (pyrecode/recode_compressors.py:84-85) and the record assembly of
// ReCoDeWriter._reduce_compress / _write_to_frame_buffer (pyrecode/recode_writer.py:482-574).
//
// Pipeline over S input streams (for the writer: 2 per frame, map and packed values):
//   k_deflate_plan      1 CTA    chunks per stream (16 KiB each), exclusive scan -> chunk_base, total
//   k_deflate_hist      token histogram: a 1-in-8 sample of the chunks into one shared histogram (levels 1..5)
//                       or every chunk into per-stream histograms (levels 6..9)
//   k_deflate_tables    Huffman code + block header, once per group (or per stream)
//   k_deflate_chunks    persistent CTAs pull chunk tickets; each chunk is an independent, byte-aligned
//                       piece (deflate_chunk.cuh) written to its scratch slot, plus its Adler-32 partials
//   k_stream_finalize   1 CTA per stream: prefix of piece sizes, Adler-32 combine, total stream size
//   k_layout_*          destination offset of every stream (fixed stride, or ReCoDe records incl. the
//                       [frame_id][sizes...] header and the exclusive scan of record sizes)
//   k_copy_pieces       one warp per piece and ticket copies it to its final byte offset (funnel-shifted 32-bit
//                       words) and write the zlib header / final block / Adler-32 trailer
// Only compact records cross PCIe afterwards.
#include "common.cuh"
#include "kernels.cuh"
#include "deflate_chunk.cuh"
#include <stdlib.h>

__global__ void __launch_bounds__(256)
k_deflate_plan(const uint32_t *__restrict__ in_bytes, int n_streams, uint32_t *__restrict__ chunk_base,
               uint32_t *__restrict__ counters, uint32_t *__restrict__ ghist)
{
    __shared__ uint32_t s_warp[9];
    uint32_t carry = 0;
    for (int s0 = 0; s0 < n_streams; s0 += 256) {
        const int s = s0 + threadIdx.x;
        const uint32_t nc = s < n_streams ? (in_bytes[s] + DF_CHUNK - 1) / DF_CHUNK : 0;
        uint32_t total;
        const uint32_t e = block_excl_scan<8>(nc, s_warp, &total);
        if (s < n_streams) chunk_base[s] = carry + e;
        carry += total;
        __syncthreads();
    }
    if (ghist)
        for (int i = threadIdx.x; i < n_streams * DF_NSYM; i += 256) ghist[i] = 0;
    if (threadIdx.x == 0) {
        chunk_base[n_streams] = carry;
        counters[0] = 0;      // histogram ticket
        counters[1] = 0;      // copy ticket
        counters[2] = 0;      // emit ticket
        counters[3] = 0;      // bytes tokenized by k_deflate_hist
    }
}

__device__ __forceinline__ int find_stream(const uint32_t *__restrict__ chunk_base, int n_streams, uint32_t gci)
{
    int lo = 0, hi = n_streams;         // chunk_base[lo] <= gci < chunk_base[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (chunk_base[mid] <= gci) lo = mid; else hi = mid;
    }
    return lo;
}

// coalesced 128-bit loads of one chunk -> per-thread segments in shared staging (zero padded past clen)
__device__ __forceinline__ void stage_chunk(uint32_t *in32, const uint8_t *__restrict__ src, int clen, int t)
{
    if ((((uintptr_t)src & 15) == 0) && clen == DF_CHUNK) {          // uniform branch: the common case
#pragma unroll
        for (int u = t; u < DF_CHUNK / 16; u += DF_THREADS) {
            const uint4 v = *reinterpret_cast<const uint4 *>(src + u * 16);
            uint32_t *d = in32 + df_in_index(u >> 2, (u & 3) * 4);
            d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
        }
        return;
    }
    const bool aligned = ((uintptr_t)src & 15) == 0;
    for (int u = t; u < DF_CHUNK / 16; u += DF_THREADS) {
        const int o = u * 16;
        uint32_t w[4] = {0, 0, 0, 0};
        if (aligned && o + 16 <= clen) {
            const uint4 v = *reinterpret_cast<const uint4 *>(src + o);
            w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
        } else if (o < clen) {
            for (int b = 0; b < 16 && o + b < clen; b++) w[b >> 2] |= (uint32_t)src[o + b] << (8 * (b & 3));
        }
        const int tt = u >> 2, k0 = (u & 3) * 4;
#pragma unroll
        for (int i = 0; i < 4; i++) in32[df_in_index(tt, k0 + i)] = w[i];
    }
}

// Token histogram.  shared_table = 1 (levels 1..5): every DF_SAMPLE-th chunk of the group is tokenized and all
// streams share one histogram (ghist[0]); shared_table = 0 (levels 6..9): every chunk, one histogram per stream.
// counters[3] accumulates the number of input bytes that were tokenized (for the entropy estimate).
constexpr uint32_t DF_SAMPLE = 8;

__global__ void __launch_bounds__(DF_THREADS)
k_deflate_hist(const uint8_t *__restrict__ in, const uint64_t *__restrict__ in_off,
               const uint32_t *__restrict__ in_bytes, int n_streams, const uint32_t *__restrict__ chunk_base,
               uint32_t *__restrict__ counters, int shared_table, uint32_t *__restrict__ ghist)
{
    __shared__ uint32_t s_in[DF_STAGE_WORDS];
    __shared__ uint32_t s_hist[DF_NSYM];
    __shared__ uint32_t s_ticket;
    const int t = threadIdx.x;
    const uint32_t total_chunks = chunk_base[n_streams];
    const uint32_t step = shared_table ? DF_SAMPLE : 1u;
    const uint32_t n_tasks = (total_chunks + step - 1) / step;
    while (true) {
        if (t == 0) s_ticket = atomicAdd(&counters[0], 1u);
        __syncthreads();
        const uint32_t task = s_ticket;
        if (task >= n_tasks) break;
        const uint32_t gci = task * step;
        const int s = find_stream(chunk_base, n_streams, gci);
        const uint32_t ci = gci - chunk_base[s];
        const int clen = (int)min((uint32_t)DF_CHUNK, in_bytes[s] - ci * DF_CHUNK);
        for (int i = t; i < DF_NSYM; i += DF_THREADS) s_hist[i] = 0;
        stage_chunk(s_in, in + in_off[s] + (size_t)ci * DF_CHUNK, clen, t);
        __syncthreads();
        df_phase_hist(s_in, s_hist, t, clen);
        __syncthreads();
        uint32_t *gh = ghist + (shared_table ? 0 : (size_t)s * DF_NSYM);
        for (int i = t; i < DF_NSYM; i += DF_THREADS) {
            const uint32_t h = s_hist[i];
            if (h) atomicAdd(&gh[i], h);
        }
        if (t == 0) atomicAdd(&counters[3], (uint32_t)clen);
        __syncthreads();
    }
}

// one CTA per code: sort the symbols by count (bitonic, 512 keys), build the code and its header.
// shared_table = 1: a single CTA builds tables[0] from ghist[0], with every producible symbol smoothed to a
// non-zero count so that chunks that were not sampled can always be encoded.
__global__ void __launch_bounds__(DF_THREADS)
k_deflate_tables(const uint32_t *__restrict__ ghist, const uint32_t *__restrict__ chunk_base,
                 const uint32_t *__restrict__ in_bytes, const uint32_t *__restrict__ counters, int shared_table,
                 int n_streams, DeflateTable *__restrict__ tables)
{
    __shared__ DfBuildShared B;
    __shared__ float s_bits[DF_THREADS];
    __shared__ uint32_t s_tok[DF_THREADS];
    __shared__ int s_skip;
    const int s = blockIdx.x, t = threadIdx.x;
    uint32_t sampled_bytes;
    if (shared_table) {
        if (chunk_base[n_streams] == 0) return;
        sampled_bytes = counters[3];
    } else {
        if (chunk_base[s + 1] == chunk_base[s]) return;      // empty stream: no chunk will ask for a table
        sampled_bytes = in_bytes[s];
    }
    // Entropy estimate of the token stream: data that would hardly shrink (bit-packed intensities are close to
    // random bytes) is emitted as stored blocks, which skips the code construction and the tokenizer for all
    // of its chunks -- and lets the reader copy them instead of decoding them.
    {
        uint32_t tok = 0;
        for (int i = t; i < DF_NSYM; i += DF_THREADS) tok += ghist[(size_t)s * DF_NSYM + i];
        s_tok[t] = tok;
        __syncthreads();
        if (t == 0) { uint32_t n = 0; for (int i = 0; i < DF_THREADS; i++) n += s_tok[i]; s_tok[0] = n; }
        __syncthreads();
        const float n = (float)s_tok[0];
        float bits = 0.f;
        for (int i = t; i < DF_NSYM; i += DF_THREADS) {
            const uint32_t h = ghist[(size_t)s * DF_NSYM + i];
            if (h) bits += (float)h * (log2f(n / (float)h) + (i > 264 ? 2.f : (i > 256 ? 1.f : 0.f)));
        }
        s_bits[t] = bits;
        __syncthreads();
        if (t == 0) {
            float b = 0.f;
            for (int i = 0; i < DF_THREADS; i++) b += s_bits[i];
            // levels 1..5 ("fast"): a group that would shrink by less than 10 % is stored -- bit-packed 12-bit
            // values gain about 5 % but cost the reader a Huffman decode of every byte; levels 6..9: 3 %
            s_skip = b * 0.125f + 128.f >= (shared_table ? 0.90f : 0.97f) * (float)sampled_bytes;
        }
        __syncthreads();
        if (s_skip) {
            if (t == 0) tables[s].header_bits = 0xffffffffu;    // sentinel: store every chunk that uses this code
            return;
        }
    }
    // ---- the code, built by the whole CTA; only the Huffman merge itself is serial (thread 0) ------------------
    // (same result as the serial df_phase_build of deflate_chunk.cuh, which the CPU harness runs)
    __shared__ uint32_t s_raw[DF_NSYM];         // unsorted keys
    __shared__ uint32_t s_cnt[16], s_start[16], s_next[16];
    __shared__ uint32_t s_misc[8];              // 0 n_used, 1 leaf overflow, 2 nlit, 3 nrl, 4 ncl, 5 header payload bits
    __shared__ uint16_t s_pos[DF_NSYM + 4];     // per head: output position of its run-length symbols; then bit offsets
    __shared__ uint8_t s_cllen[20];
    __shared__ uint16_t s_clcode[20];
    __shared__ uint32_t s_clfreq[20];
    __shared__ uint32_t s_warp[9];
    DeflateTable &T = B.tab;
    // counts are scaled to < 2^21 so that (count << 9 | symbol) keys and all weight sums fit 32 bits
    const uint32_t ntok = s_tok[0];
    const int cshift = ntok >> 21 ? 32 - __clz(ntok >> 21) : 0;
    for (int i = t; i < DF_NSYM; i += DF_THREADS) {
        uint32_t c = ghist[(size_t)s * DF_NSYM + i];
        if (c) c = max(c >> cshift, 1u);
        if (shared_table && i < DF_LEN_SYMS) c = c * 2u + 1u; // smoothing: every producible symbol gets a code
        if (i == 256 && c == 0) c = 1;                        // end-of-block is used once per chunk; any count > 0 works
        s_raw[i] = c ? ((c << 9) | (uint32_t)i) : 0xffffffffu;
        T.len[i] = 0;
        T.code[i] = 0;
    }
    for (int i = t; i < DF_HDR_WORDS; i += DF_THREADS) T.header[i] = 0;
    if (t < 16) s_cnt[t] = 0;
    if (t < 20) s_clfreq[t] = 0;
    if (t < 8) s_misc[t] = 0;
    __syncthreads();
    // rank sort (keys are unique: the symbol is in the low bits)
    for (int i = t; i < DF_NSYM; i += DF_THREADS) {
        const uint32_t k = s_raw[i];
        if (k != 0xffffffffu) {
            uint32_t r = 0;
            for (int j = 0; j < DF_NSYM; j++) r += s_raw[j] < k;
            B.keys[r] = k;
            atomicAdd(&s_misc[0], 1u);
        }
    }
    __syncthreads();
    const int n_used = (int)s_misc[0];
    if (n_used == 1) {
        // a lone symbol still needs one bit; add a dummy sibling (as df_build_lengths does)
        if (t == 0) {
            const int sy = B.keys[0] & 511;
            T.len[sy] = 1;
            T.len[sy == 0 ? 1 : 0] = 1;
        }
    } else {
        if (t == 0) {
            // two-queue merge: leaves i (weight keys[i] >> 9), internal nodes j (weight node_w[j])
            int li = 0, ii = 0, ni = 0;
            uint32_t lw = B.keys[0] >> 9, iw = 0xffffffffu;
            for (int k = 0; k < n_used - 1; k++) {
                uint32_t w = 0;
#pragma unroll
                for (int pick = 0; pick < 2; pick++) {
                    if (li < n_used && (ii >= ni || lw <= iw)) {
                        w += lw; B.node_parent[li] = (uint16_t)ni; li++;
                        lw = li < n_used ? B.keys[li] >> 9 : 0xffffffffu;
                    } else {
                        w += iw; B.node_parent[DF_NSYM + ii] = (uint16_t)ni; ii++;
                        iw = ii < ni ? B.node_w[ii] : 0xffffffffu;
                    }
                }
                B.node_w[ni] = w;
                if (ii == ni) iw = w;          // the new node is the head of the internal queue
                ni++;
            }
            // depths of the internal nodes top-down with clamping; every clamped node counts (zlib gen_bitlen)
            uint32_t overflow = 0;
            B.node_depth[ni - 1] = 0;
            for (int j = ni - 2; j >= 0; j--) {
                int d = B.node_depth[B.node_parent[DF_NSYM + j]] + 1;
                if (d > 15) { d = 15; overflow++; }
                B.node_depth[j] = (uint8_t)d;
            }
            s_misc[1] = overflow;
        }
        __syncthreads();
        for (int i = t; i < n_used; i += DF_THREADS) {
            int d = B.node_depth[B.node_parent[i]] + 1;
            if (d > 15) { d = 15; atomicAdd(&s_misc[1], 1u); }
            atomicAdd(&s_cnt[d], 1u);
        }
        __syncthreads();
        if (t == 0) {
            int overflow = (int)s_misc[1];
            while (overflow > 0) {
                int bits = 14;
                while (s_cnt[bits] == 0) bits--;
                s_cnt[bits]--;
                s_cnt[bits + 1] += 2;
                s_cnt[15]--;
                overflow -= 2;
            }
            uint32_t i = 0;                  // least frequent leaves (front of the sorted keys) get the longest codes
            for (int bits = 15; bits >= 1; bits--) { s_start[bits] = i; i += s_cnt[bits]; }
        }
        __syncthreads();
        for (int i = t; i < n_used; i += DF_THREADS) {
            int bits = 15;
            while (bits > 1 && !((uint32_t)i >= s_start[bits] && (uint32_t)i < s_start[bits] + s_cnt[bits])) bits--;
            T.len[B.keys[i] & 511] = (uint8_t)bits;
        }
    }
    __syncthreads();
    // canonical codes (RFC 1951 3.2.2): next_code per length, then rank among equal lengths by symbol order
    if (t < 16) s_cnt[t] = 0;
    __syncthreads();
    for (int i = t; i < DF_NSYM; i += DF_THREADS) {
        const int l = T.len[i];
        if (l) atomicAdd(&s_cnt[l], 1u);
        if (l && i < 286) atomicMax(&s_misc[2], (uint32_t)i + 1u);
    }
    __syncthreads();
    if (t == 0) {
        uint32_t c = 0;
        s_next[0] = 0;
        uint32_t prev = 0;
        for (int bits = 1; bits <= 15; bits++) { c = (c + prev) << 1; s_next[bits] = c; prev = s_cnt[bits]; }
        if (s_misc[2] < 257) s_misc[2] = 257;
    }
    __syncthreads();
    for (int i = t; i < DF_NSYM; i += DF_THREADS) {
        const int l = T.len[i];
        if (l) {
            uint32_t r = 0;
            for (int j = 0; j < i; j++) r += T.len[j] == l;
            T.code[i] = (uint16_t)df_bitrev(s_next[l] + r, l);
        }
    }
    // ---- header: HLIT/HDIST/HCLEN, code-length code, run-length coded lengths (RFC 1951 3.2.7) ---------------
    const int nlit = (int)s_misc[2];
    const int nseq = nlit + 2;                  // like zlib: always two distance codes of one bit each
    for (int i = t; i < nseq; i += DF_THREADS) B.seq[i] = i < nlit ? T.len[i] : 1;
    __syncthreads();
    // run heads -> number of run-length symbols each run emits
    for (int i0 = 0; i0 < nseq; i0 += DF_THREADS) {
        const int i = i0 + t;
        uint32_t n = 0;
        if (i < nseq && (i == 0 || B.seq[i] != B.seq[i - 1])) {
            const int v = B.seq[i];
            int R = 1;
            while (i + R < nseq && B.seq[i + R] == v) R++;
            if (v == 0) {
                const int rem = R % 138;
                n = R / 138 + (rem >= 3 ? 1 : rem);
            } else {
                const int left = R - 1, rem = left % 6;
                n = 1 + left / 6 + (rem >= 3 ? 1 : rem);
            }
        }
        uint32_t total;
        const uint32_t e = block_excl_scan<8>(n, s_warp, &total);
        if (i < nseq) s_pos[i] = (uint16_t)(s_misc[3] + e);
        __syncthreads();
        if (t == 0) s_misc[3] += total;
        __syncthreads();
    }
    const int nrl = (int)s_misc[3];
    for (int i = t; i < nseq; i += DF_THREADS) {
        if (i == 0 || B.seq[i] != B.seq[i - 1]) {
            const int v = B.seq[i];
            int R = 1;
            while (i + R < nseq && B.seq[i + R] == v) R++;
            int o = s_pos[i];
            if (v == 0) {
                int left = R;
                while (left >= 11) { const int r = left > 138 ? 138 : left; B.rl_sym[o] = 18; B.rl_ext[o++] = (uint8_t)(r - 11); left -= r; }
                if (left >= 3) { B.rl_sym[o] = 17; B.rl_ext[o++] = (uint8_t)(left - 3); left = 0; }
                while (left-- > 0) { B.rl_sym[o] = 0; B.rl_ext[o++] = 0; }
            } else {
                B.rl_sym[o] = (uint8_t)v; B.rl_ext[o++] = 0;
                int left = R - 1;
                while (left >= 3) { const int r = left > 6 ? 6 : left; B.rl_sym[o] = 16; B.rl_ext[o++] = (uint8_t)(r - 3); left -= r; }
                while (left-- > 0) { B.rl_sym[o] = (uint8_t)v; B.rl_ext[o++] = 0; }
            }
        }
    }
    __syncthreads();
    for (int i = t; i < nrl; i += DF_THREADS) atomicAdd(&s_clfreq[B.rl_sym[i]], 1u);
    __syncthreads();
    const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    if (t == 0) {
        // code-length code: 19 symbols, 7 bits max; tiny insertion sort + the serial builder
        uint32_t ckeys[19];
        int cused = 0;
        for (int sy = 0; sy < 19; sy++) if (s_clfreq[sy]) {
            const uint32_t key = (s_clfreq[sy] << 9) | (uint32_t)sy;
            int j = cused++;
            while (j > 0 && ckeys[j - 1] > key) { ckeys[j] = ckeys[j - 1]; j--; }
            ckeys[j] = key;
        }
        df_build_lengths(ckeys, cused, 7, s_cllen, 19, B.node_w, B.node_parent, B.node_depth);
        df_assign_codes(s_cllen, 19, s_clcode);
        int ncl = 19;
        while (ncl > 4 && s_cllen[order[ncl - 1]] == 0) ncl--;
        s_misc[4] = (uint32_t)ncl;
        DfBitWriter bw{T.header, 0};
        bw.put(0, 1);                 // BFINAL = 0 (the stream is closed by the assembler)
        bw.put(2, 2);                 // BTYPE = 10 dynamic
        bw.put((uint32_t)(nlit - 257), 5);
        bw.put(1, 5);                 // HDIST = 2 - 1
        bw.put((uint32_t)(ncl - 4), 4);
    }
    __syncthreads();
    const int ncl = (int)s_misc[4];
    if (t < ncl) {
        const uint32_t pos = 17u + 3u * (uint32_t)t, v = s_cllen[order[t]];
        atomicOr(&T.header[pos >> 5], v << (pos & 31));
        if ((pos & 31) > 29) atomicOr(&T.header[(pos >> 5) + 1], v >> (32 - (pos & 31)));
    }
    // bit offset of every run-length symbol, then parallel emission
    const uint32_t hbase = 17u + 3u * (uint32_t)ncl;
    for (int i0 = 0; i0 < nrl; i0 += DF_THREADS) {
        const int i = i0 + t;
        uint32_t nb = 0, bits = 0;
        if (i < nrl) {
            const int sy = B.rl_sym[i];
            const uint32_t l = s_cllen[sy];
            const uint32_t eb = sy == 16 ? 2u : (sy == 17 ? 3u : (sy == 18 ? 7u : 0u));
            bits = (uint32_t)s_clcode[sy] | ((uint32_t)B.rl_ext[i] << l);
            nb = l + eb;
        }
        uint32_t total;
        const uint32_t e = block_excl_scan<8>(nb, s_warp, &total);
        if (nb) {
            const uint32_t pos = hbase + s_misc[5] + e;
            atomicOr(&T.header[pos >> 5], bits << (pos & 31));
            if ((pos & 31) + nb > 32) atomicOr(&T.header[(pos >> 5) + 1], bits >> (32 - (pos & 31)));
        }
        __syncthreads();
        if (t == 0) s_misc[5] += total;
        __syncthreads();
    }
    if (t == 0) T.header_bits = hbase + s_misc[5];
    __syncthreads();
    const uint32_t *src = reinterpret_cast<const uint32_t *>(&B.tab);
    uint32_t *dst = reinterpret_cast<uint32_t *>(&tables[s]);
    for (int i = t; i < (int)(sizeof(DeflateTable) / 4); i += DF_THREADS) dst[i] = src[i];
}

// every chunk: Adler-32 partials, then encode with its code (or store)
__global__ void __launch_bounds__(DF_THREADS)
k_deflate_chunks(const uint8_t *__restrict__ in, const uint64_t *__restrict__ in_off,
                 const uint32_t *__restrict__ in_bytes, int n_streams, const uint32_t *__restrict__ chunk_base,
                 uint32_t *__restrict__ counters, int level, int shared_table,
                 const DeflateTable *__restrict__ tables, uint8_t *__restrict__ scratch,
                 uint32_t *__restrict__ chunk_bytes, uint2 *__restrict__ chunk_adler)
{
    __shared__ DfEmitShared S;
    __shared__ uint32_t s_ticket;
    __shared__ uint32_t s_warp[9];
    __shared__ uint32_t s_adler[2];
    const int t = threadIdx.x;
    const uint32_t total_chunks = chunk_base[n_streams];
    // one code for all chunks (levels 1..5): its token tables are set up once per CTA, not once per chunk
    bool table_loaded = false;
    if (t == 0) S.e_bad = 0;
    __syncthreads();
    if (shared_table && level > 0 && tables[0].header_bits != 0xffffffffu) {
        df_load_table(S, tables[0], t, DF_THREADS);
        table_loaded = true;
    }

    while (true) {
        if (t == 0) { s_ticket = atomicAdd(&counters[2], 1u); s_adler[0] = 0; s_adler[1] = 0; S.overflow = 0; }
        __syncthreads();
        const uint32_t gci = s_ticket;
        if (gci >= total_chunks) break;
        const int s = find_stream(chunk_base, n_streams, gci);
        const uint32_t ci = gci - chunk_base[s];
        const int clen = (int)min((uint32_t)DF_CHUNK, in_bytes[s] - ci * DF_CHUNK);
        const uint8_t *src = in + in_off[s] + (size_t)ci * DF_CHUNK;

        stage_chunk(S.io, src, clen, t);
        bool stored = level == 0;
        uint32_t hb = 0;
        const DeflateTable &T = tables[shared_table ? 0 : s];
        if (!stored) {
            hb = T.header_bits;
            stored = hb == 0xffffffffu;
        }
        if (!stored && !table_loaded) df_load_table(S, T, t, DF_THREADS);     // per-stream codes (levels 6..9)
        __syncthreads();

        // A chunk with few non-zero bytes (a binary map) is encoded from the list of those bytes -- work per non-zero
        // byte, equal shares -- and its Adler-32 partials come from the same list; everything else is tokenized
        // byte by byte.
        uint32_t body_bits = 0, my_bits = 0, my_off = 0, n_nz = 0;
        uint32_t ad_a = 0, ad_b = 0;
        bool sparse = false;
        if (!stored) {
            if (t == 0) S.header_bits = hb;
            if (shared_table && !S.e_bad) {
                uint32_t nz[2];
                const uint32_t cnt = df_nz_masks(S.io, t, df_seg_bytes(t, clen), nz);
                const uint32_t base = block_excl_scan<8>(cnt, s_warp, &n_nz);
                sparse = n_nz <= (uint32_t)DF_STAGE_WORDS;
                if (sparse) df_nz_scatter(S.io, S.priv, t, base, nz);
                __syncthreads();
                if (sparse) my_bits = df_sparse_bits(S, S.priv, t, n_nz, clen, ad_a, ad_b);
            }
        }
        if (!sparse) df_adler_partial(S.io, t, clen, ad_a, ad_b);
        {
            // Adler-32 partials of the chunk (warp shuffle reduction, one shared atomic per warp)
            uint32_t a = ad_a % 65521u, b = ad_b % 65521u;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                a += __shfl_down_sync(0xffffffffu, a, d);
                b += __shfl_down_sync(0xffffffffu, b, d);
            }
            if ((t & 31) == 0 && (a | b)) { atomicAdd(&s_adler[0], a % 65521u); atomicAdd(&s_adler[1], b % 65521u); }
        }
        if (!stored) {
            if (!sparse) my_bits = df_encode_segment(S, t, clen);
            uint32_t tok_bits;
            const uint32_t e = block_excl_scan<8>(my_bits, s_warp, &tok_bits);     // two barriers: encode pass is over
            S.tbits[t] = e;
            my_off = hb + e;
            body_bits = hb + tok_bits;
            stored = S.overflow || df_dynamic_bytes(body_bits, (int)(S.tbl[256] >> 24)) >= (uint32_t)clen + 10u;
        } else {
            __syncthreads();
        }

        uint8_t *slot = scratch + (size_t)gci * DF_SLOT_BYTES;
        if (!stored) {
            // the staged input is no longer needed: its area becomes the zero-initialised output bit stream
            const int hw = (int)((hb + 31) >> 5);
            const int ow = (int)((df_dynamic_bytes(body_bits, (int)(S.tbl[256] >> 24)) + 3 + 4) >> 2);
            for (int i = t; i < ow && i < DF_STAGE_WORDS; i += DF_THREADS) S.io[i] = i < hw ? T.header[i] : 0;
            __syncthreads();
            if (sparse) df_sparse_emit(S, S.priv, t, n_nz, clen, my_off);
            else df_place_segment(S, t, my_bits);
            __syncthreads();
            if (t == 0) df_phase_finish(S, body_bits);
            __syncthreads();
            const uint32_t nb = S.out_bytes;
            uint4 *dst = reinterpret_cast<uint4 *>(slot);
            const uint4 *so = reinterpret_cast<const uint4 *>(S.io);
            for (uint32_t i = t; i < (nb + 15) / 16; i += DF_THREADS) dst[i] = so[i];
            if (t == 0) chunk_bytes[gci] = nb;
        } else {
            // stored block: 00 | LEN | ~LEN | data | sync marker (00 0000 FFFF), straight from the staged input
            if (t == 0) {
                slot[0] = 0;
                slot[1] = (uint8_t)(clen & 0xff); slot[2] = (uint8_t)(clen >> 8);
                slot[3] = (uint8_t)(~clen & 0xff); slot[4] = (uint8_t)((~clen >> 8) & 0xff);
                slot[5 + clen] = 0; slot[6 + clen] = 0; slot[7 + clen] = 0; slot[8 + clen] = 0xff; slot[9 + clen] = 0xff;
                chunk_bytes[gci] = (uint32_t)clen + 10u;
            }
            const uint8_t *sb = reinterpret_cast<const uint8_t *>(S.io);
            for (int i = t; i < clen; i += DF_THREADS)
                slot[5 + i] = sb[(i >> 6) * (DF_SEG_STRIDE * 4) + (i & 63)];
        }
        __syncthreads();
        if (t == 0) chunk_adler[gci] = make_uint2(s_adler[0] % 65521u, s_adler[1] % 65521u);
        __syncthreads();
    }
}

// ---- per-stream totals ---------------------------------------------------------------------------
// wrap = 1: zlib wrapper (2-byte header, final empty fixed block 03 00, Adler-32): stream = 2 + pieces + 6
// wrap = 0: raw concatenation (mode 0: chunk_bytes must already hold the raw chunk lengths)
__global__ void __launch_bounds__(256)
k_stream_finalize(const uint32_t *__restrict__ in_bytes, const uint32_t *__restrict__ chunk_base,
                  const uint32_t *__restrict__ chunk_bytes, const uint2 *__restrict__ chunk_adler, int wrap,
                  uint32_t *__restrict__ chunk_rel, uint32_t *__restrict__ stream_bytes,
                  uint32_t *__restrict__ stream_adler)
{
    __shared__ uint32_t s_warp[9];
    const int s = blockIdx.x, t = threadIdx.x;
    const uint32_t c0 = chunk_base[s], c1 = chunk_base[s + 1];
    const uint32_t slen = in_bytes[s];
    uint32_t carry = 0;
    // Adler-32 over the concatenated chunks from their (A_c, B_c, len_c) partials, all chunks in parallel:
    //   s1 before chunk c = 1 + sum of A_k (k < c);   s2 = sum over c of (len_c * s1 before c + B_c)      (mod 65521)
    uint32_t s1 = 1, s2 = 0;
    for (uint32_t i0 = c0; i0 < c1; i0 += 256) {
        const uint32_t i = i0 + t;
        const bool in = i < c1;
        const uint32_t v = in ? chunk_bytes[i] : 0;
        const uint2 ab = (in && wrap) ? chunk_adler[i] : make_uint2(0, 0);
        uint32_t total;
        const uint32_t e = block_excl_scan<8>(v, s_warp, &total);
        if (in) chunk_rel[i] = carry + e;
        carry += total;
        __syncthreads();
        if (wrap) {
            uint32_t a_tot;
            const uint32_t a_ex = block_excl_scan<8>(ab.x, s_warp, &a_tot);            // A_k < 65521: no overflow
            __syncthreads();
            const uint32_t s1b = (s1 + a_ex) % 65521u;
            const uint32_t clen = in ? min((uint32_t)DF_CHUNK, slen - (i - c0) * DF_CHUNK) : 0u;
            const uint32_t term = in ? (uint32_t)(((uint64_t)clen * s1b + ab.y) % 65521u) : 0u;
            uint32_t t_tot;
            block_excl_scan<8>(term, s_warp, &t_tot);
            __syncthreads();
            s2 = (s2 + t_tot % 65521u) % 65521u;
            s1 = (s1 + a_tot % 65521u) % 65521u;
        }
    }
    if (t == 0) {
        stream_bytes[s] = wrap ? carry + 8u : carry;
        if (wrap) stream_adler[s] = (s2 << 16) | s1;
    }
}

// ---- layouts ---------------------------------------------------------------------------------------
__global__ void k_layout_strided(const uint32_t *__restrict__ stream_bytes, int n_streams, size_t stride,
                                 uint64_t *__restrict__ stream_dst, uint32_t *__restrict__ out_bytes)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_streams) return;
    stream_dst[s] = (uint64_t)s * stride;
    out_bytes[s] = stream_bytes[s];
}

__device__ __forceinline__ void put_u32le(uint8_t *p, uint32_t v)
{
    p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24);
}

// ReCoDe records.  streams_per_frame = 2 (L1/L2: map, vals) or 1 (L3/L4: map).  The map streams and the value
// streams are two separate deflate groups (they are encoded on different CUDA streams): stream f of each.
// mode 1 header: [frame_id][n_comp_map]([n_comp_vals][n_packed]); mode 0 header: [frame_id]([n_packed]).
__global__ void __launch_bounds__(256)
k_layout_records(const uint32_t *__restrict__ map_bytes, const uint32_t *__restrict__ val_bytes,
                 const uint32_t *__restrict__ packed_bytes, int n_frames,
                 int spf, int mode, uint32_t first_frame_id, size_t capacity, uint8_t *__restrict__ records,
                 uint64_t *__restrict__ record_off, uint64_t *__restrict__ map_dst, uint64_t *__restrict__ val_dst,
                 uint32_t *__restrict__ status)
{
    __shared__ uint32_t s_warp[9];
    __shared__ uint64_t s_carry;
    const int t = threadIdx.x;
    if (t == 0) s_carry = 0;
    __syncthreads();
    const uint32_t hdr = mode == 1 ? (spf == 2 ? 16u : 8u) : (spf == 2 ? 8u : 4u);
    for (int f0 = 0; f0 < n_frames; f0 += 256) {
        const int f = f0 + t;
        uint32_t len = 0, l0 = 0, l1 = 0;
        if (f < n_frames) {
            l0 = map_bytes[f];
            l1 = spf == 2 ? val_bytes[f] : 0;
            len = hdr + l0 + l1;        // one record: < 2^32 (two streams of < 2^31 bytes each)
        }
        // the sum over a group of 256 records can pass 2^32: scan 16 MiB units and remainders separately
        uint32_t total_hi, total_lo;
        const uint32_t e_hi = block_excl_scan<8>(len >> 24, s_warp, &total_hi);
        __syncthreads();
        const uint32_t e_lo = block_excl_scan<8>(len & 0xffffffu, s_warp, &total_lo);
        const uint64_t off = s_carry + ((uint64_t)e_hi << 24) + e_lo;
        const uint64_t total = ((uint64_t)total_hi << 24) + total_lo;
        if (f < n_frames) {
            record_off[f] = off;
            map_dst[f] = off + hdr;
            if (spf == 2) val_dst[f] = off + hdr + l0;
            if (off + len <= capacity) {
                uint8_t *r = records + off;
                put_u32le(r, first_frame_id + (uint32_t)f);
                if (mode == 1) {
                    put_u32le(r + 4, l0);
                    if (spf == 2) { put_u32le(r + 8, l1); put_u32le(r + 12, packed_bytes[f]); }
                } else if (spf == 2) {
                    put_u32le(r + 4, packed_bytes[f]);
                }
            } else {
                atomicOr(status, RC_STATUS_RECORDS_OVERFLOW);
            }
        }
        __syncthreads();
        if (t == 0) s_carry += total;
        __syncthreads();
    }
    if (t == 0) record_off[n_frames] = s_carry;
}

// ---- piece copy ------------------------------------------------------------------------------------
// src is 4-byte aligned (scratch slot or a 16-byte aligned raw stream + multiple of 16 KiB); dst arbitrary.
__device__ __forceinline__ void copy_bytes_shifted(uint8_t *__restrict__ dst, const uint8_t *__restrict__ src,
                                                   uint32_t n, int t, int nt)
{
    // head bytes up to the first 4-byte aligned destination address
    uint32_t head = (uint32_t)((4 - ((uintptr_t)dst & 3)) & 3);
    if (head > n) head = n;
    if ((uint32_t)t < head) dst[t] = src[t];
    const uint32_t nw = (n - head) >> 2;
    uint32_t *d32 = reinterpret_cast<uint32_t *>(dst + head);
    const uint32_t *s32 = reinterpret_cast<const uint32_t *>(src);
    const uint32_t sh = head * 8;     // source byte offset of the first aligned destination word is `head`
    if (sh == 0) {
#pragma unroll 4
        for (uint32_t i = t; i < nw; i += nt) d32[i] = s32[i];
    } else {
#pragma unroll 4
        for (uint32_t i = t; i < nw; i += nt) d32[i] = __funnelshift_r(s32[i], s32[i + 1], sh);
    }
    const uint32_t done = head + nw * 4;
    if (done + (uint32_t)t < n) dst[done + t] = src[done + t];      // tail: at most 3 bytes
}

__global__ void __launch_bounds__(256)
k_copy_pieces(const uint8_t *__restrict__ raw_in, const uint64_t *__restrict__ in_off,
              const uint32_t *__restrict__ in_bytes, int n_streams, const uint32_t *__restrict__ chunk_base,
              const uint8_t *__restrict__ scratch, const uint32_t *__restrict__ chunk_bytes,
              const uint32_t *__restrict__ chunk_rel, const uint32_t *__restrict__ stream_bytes,
              const uint32_t *__restrict__ stream_adler, const uint64_t *__restrict__ stream_dst, int wrap,
              uint8_t *__restrict__ out, size_t capacity, uint32_t *__restrict__ counters,
              uint32_t *__restrict__ status)
{
    // one piece per WARP and ticket: a piece is a few KiB, so a whole CTA per piece spends most of its instructions on
    // the ticket, the stream search and the head / tail bytes rather than on the copy
    const int t = threadIdx.x, lane = t & 31;
    const uint32_t total_chunks = chunk_base[n_streams];
    // empty streams have no chunk: their wrapper is written by the first CTA
    if (wrap && blockIdx.x == 0) {
        for (int s = t; s < n_streams; s += 256) {
            if (chunk_base[s + 1] == chunk_base[s] && stream_dst[s] + 8 <= capacity) {
                uint8_t *d = out + stream_dst[s];
                d[0] = 0x78; d[1] = 0x01; d[2] = 0x03; d[3] = 0x00; d[4] = 0; d[5] = 0; d[6] = 0; d[7] = 1;
            }
        }
    }
    while (true) {
        uint32_t gci = 0;
        if (lane == 0) gci = atomicAdd(&counters[1], 1u);
        gci = __shfl_sync(0xffffffffu, gci, 0);
        if (gci >= total_chunks) break;
        const int s = find_stream(chunk_base, n_streams, gci);
        const uint32_t ci = gci - chunk_base[s];
        const uint64_t sd = stream_dst[s];
        const uint32_t sb = stream_bytes[s];
        if (sd + sb > capacity) {
            if (lane == 0) atomicOr(status, RC_STATUS_RECORDS_OVERFLOW);
            continue;
        }
        const uint32_t nb = chunk_bytes[gci];
        uint8_t *dst = out + sd + (wrap ? 2 : 0) + chunk_rel[gci];
        const uint8_t *src = wrap ? scratch + (size_t)gci * DF_SLOT_BYTES
                                  : raw_in + in_off[s] + (size_t)ci * DF_CHUNK;
        copy_bytes_shifted(dst, src, nb, lane, 32);
        if (wrap && lane == 0) {
            if (ci == 0) { out[sd] = 0x78; out[sd + 1] = 0x01; }
            if (gci + 1 == chunk_base[s + 1]) {
                uint8_t *tr = out + sd + sb - 6;
                const uint32_t ad = stream_adler[s];
                tr[0] = 0x03; tr[1] = 0x00;          // BFINAL=1, BTYPE=01, end-of-block: the final empty block
                tr[2] = (uint8_t)(ad >> 24); tr[3] = (uint8_t)(ad >> 16); tr[4] = (uint8_t)(ad >> 8); tr[5] = (uint8_t)ad;
            }
        }
    }
}

// mode 0: chunk_bytes = raw chunk lengths
__global__ void k_raw_chunk_bytes(const uint32_t *__restrict__ in_bytes, const uint32_t *__restrict__ chunk_base,
                                  int n_streams, uint32_t *__restrict__ chunk_bytes)
{
    const int s = blockIdx.x;
    const uint32_t c0 = chunk_base[s], c1 = chunk_base[s + 1], slen = in_bytes[s];
    for (uint32_t i = c0 + threadIdx.x; i < c1; i += blockDim.x)
        chunk_bytes[i] = min((uint32_t)DF_CHUNK, slen - (i - c0) * DF_CHUNK);
}

// ---- host side -------------------------------------------------------------------------------------
size_t deflate_table_bytes() { return sizeof(DeflateTable); }

size_t deflate_max_chunks(int n_streams, size_t max_in_bytes)
{
    return (size_t)n_streams * ((max_in_bytes + DF_CHUNK - 1) / DF_CHUNK);
}

DeflateWs carve_deflate_ws(Carver &c, int n_streams, size_t max_chunks, bool need_scratch)
{
    DeflateWs w;
    w.chunk_base = c.take<uint32_t>((size_t)n_streams + 1);
    w.counters = c.take<uint32_t>(8);
    w.chunk_bytes = c.take<uint32_t>(max_chunks + 1);
    w.chunk_rel = c.take<uint32_t>(max_chunks + 1);
    w.chunk_adler = c.take<uint2>(max_chunks + 1);
    w.stream_bytes = c.take<uint32_t>((size_t)n_streams + 1);
    w.stream_adler = c.take<uint32_t>((size_t)n_streams + 1);
    w.stream_dst = c.take<uint64_t>((size_t)n_streams + 1);
    w.ghist = need_scratch ? c.take<uint32_t>((size_t)n_streams * DF_NSYM) : nullptr;
    w.tables = need_scratch ? c.take<DeflateTable>((size_t)n_streams) : nullptr;
    w.scratch = need_scratch ? c.take<uint8_t>(max_chunks * DF_SLOT_BYTES + 64) : nullptr;
    w.max_chunks = max_chunks;
    return w;
}

// encodes (wrap = 1) or sizes (wrap = 0, reduce-only mode) the chunks and finalizes per-stream totals
// shared_table: the streams are statistically alike (the frames of one batch): one sampled code for all of them.
// kept_table (shared_table only): a code in context-owned memory that survives the call; it is rebuilt from this
// batch when build_table is set and re-used as it is otherwise (no histogram / table kernels at all).
int launch_deflate_streams(rc_ctx *ctx, int level, int wrap, int shared_table, void *kept_table, int build_table,
                           const uint8_t *in, const uint64_t *in_off, const uint32_t *in_bytes, int n_streams,
                           const DeflateWs &w, cudaStream_t st)
{
    if (n_streams <= 0) return 0;
    DeflateTable *tables = (shared_table && kept_table) ? (DeflateTable *)kept_table : (DeflateTable *)w.tables;
    if (!(shared_table && kept_table)) build_table = 1;
    k_deflate_plan<<<1, 256, 0, st>>>(in_bytes, n_streams, w.chunk_base, w.counters, wrap ? w.ghist : nullptr);
    RC_LAUNCH_CHECK(ctx, "k_deflate_plan");
    if (wrap) {
        size_t want;
        if (level > 0 && build_table) {
            const size_t tasks = shared_table ? (w.max_chunks + DF_SAMPLE - 1) / DF_SAMPLE : w.max_chunks;
            want = tasks < (size_t)ctx->sm_count * 8 ? tasks : (size_t)ctx->sm_count * 8;
            if (want < 1) want = 1;
            k_deflate_hist<<<(unsigned)want, DF_THREADS, 0, st>>>(in, in_off, in_bytes, n_streams, w.chunk_base,
                                                                   w.counters, shared_table, w.ghist);
            RC_LAUNCH_CHECK(ctx, "k_deflate_hist");
            k_deflate_tables<<<shared_table ? 1 : n_streams, DF_THREADS, 0, st>>>(
                w.ghist, w.chunk_base, in_bytes, w.counters, shared_table, n_streams, tables);
            RC_LAUNCH_CHECK(ctx, "k_deflate_tables");
        }
        // Persistent CTAs per SM (40 KiB of shared memory each).  A batch alone on the GPU takes all that fit; with
        // several batches in flight (rc_set_pipelined) the encoder runs thin and long -- ONE CTA per SM -- beside the
        // DRAM-bound streaming kernel of the next batch instead of displacing its CTAs: 81.1 k vs 76.6 k L2 frames/s
        // (profiles/r02_sweep_residency.txt).  RECODE_B200_DEFLATE_CTAS overrides.
        static const int def_ctas = getenv("RECODE_B200_DEFLATE_CTAS") ? atoi(getenv("RECODE_B200_DEFLATE_CTAS")) : 0;
        const size_t per_sm = (size_t)(def_ctas > 0 ? def_ctas : (ctx->pipelined ? 1 : 6));
        want = w.max_chunks < (size_t)ctx->sm_count * per_sm ? w.max_chunks : (size_t)ctx->sm_count * per_sm;
        if (want < 1) want = 1;
        // RC_ABLATE bit 1 (timing experiments only): no chunk encoder, every piece is empty
        static const int ablate = getenv("RC_ABLATE") ? atoi(getenv("RC_ABLATE")) : 0;
        if (ablate & 2) {
            cudaMemsetAsync(w.chunk_bytes, 0, (w.max_chunks + 1) * sizeof(uint32_t), st);
            cudaMemsetAsync(w.chunk_adler, 0, (w.max_chunks + 1) * sizeof(uint2), st);
        } else
        k_deflate_chunks<<<(unsigned)want, DF_THREADS, 0, st>>>(in, in_off, in_bytes, n_streams, w.chunk_base,
                                                                 w.counters, level, shared_table,
                                                                 tables, w.scratch,
                                                                 w.chunk_bytes, w.chunk_adler);
        RC_LAUNCH_CHECK(ctx, "k_deflate_chunks");
    } else {
        k_raw_chunk_bytes<<<n_streams, 128, 0, st>>>(in_bytes, w.chunk_base, n_streams, w.chunk_bytes);
        RC_LAUNCH_CHECK(ctx, "k_raw_chunk_bytes");
    }
    k_stream_finalize<<<n_streams, 256, 0, st>>>(in_bytes, w.chunk_base, w.chunk_bytes, w.chunk_adler, wrap,
                                                 w.chunk_rel, w.stream_bytes, w.stream_adler);
    RC_LAUNCH_CHECK(ctx, "k_stream_finalize");
    return 0;
}

int launch_layout_strided(rc_ctx *ctx, const DeflateWs &w, int n_streams, size_t stride, uint32_t *out_bytes,
                          cudaStream_t st)
{
    k_layout_strided<<<(n_streams + 255) / 256, 256, 0, st>>>(w.stream_bytes, n_streams, stride, w.stream_dst, out_bytes);
    RC_LAUNCH_CHECK(ctx, "k_layout_strided");
    return 0;
}

int launch_layout_records(rc_ctx *ctx, const DeflateWs &wm, const DeflateWs &wv, const uint32_t *packed_bytes,
                          int n_frames, int spf, int mode, uint32_t first_frame_id, uint8_t *records, size_t capacity,
                          uint64_t *record_off, uint32_t *status, cudaStream_t st)
{
    k_layout_records<<<1, 256, 0, st>>>(wm.stream_bytes, wv.stream_bytes, packed_bytes, n_frames, spf, mode,
                                        first_frame_id, capacity, records, record_off, wm.stream_dst, wv.stream_dst,
                                        status);
    RC_LAUNCH_CHECK(ctx, "k_layout_records");
    return 0;
}

int launch_copy_pieces(rc_ctx *ctx, const DeflateWs &w, int wrap, const uint8_t *raw_in, const uint64_t *in_off,
                       const uint32_t *in_bytes, int n_streams, uint8_t *out, size_t capacity, uint32_t *status,
                       cudaStream_t st)
{
    size_t want = (w.max_chunks + 7) / 8;                    // 8 warps per CTA, one piece per warp and ticket
    if (want > (size_t)ctx->sm_count * 4) want = (size_t)ctx->sm_count * 4;
    if (want < 1) want = 1;
    k_copy_pieces<<<(unsigned)want, 256, 0, st>>>(raw_in, in_off, in_bytes, n_streams, w.chunk_base, w.scratch,
                                                  w.chunk_bytes, w.chunk_rel, w.stream_bytes, w.stream_adler,
                                                  w.stream_dst, wrap, out, capacity, w.counters, status);
    RC_LAUNCH_CHECK(ctx, "k_copy_pieces");
    return 0;
}
