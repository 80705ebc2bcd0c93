// inflate.cu -- batched zlib-format inflate on the GPU (read path).
//
// Replaces zlib.decompress (pyrecode/recode_compressors.py:42-43).  A deflate stream is inherently serial, so
// parallelism comes from (a) many streams per batch (2 per frame) and (b), for streams written by deflate.cu,
// from their independent 16 KiB chunks: every chunk ends with the sync-flush marker 00 00 FF FF and starts byte
// aligned, so chunk starts can be found by a byte scan and decoded speculatively in parallel:
//
//   k_inflate_scan      1 CTA per stream: checks the zlib header, lists candidate chunk starts (offset 2 and
//                       the byte after every 00 00 FF FF) in order
//   k_scan_u32          exclusive scan of candidates per stream -> task table
//   k_inflate_chunks    one warp per candidate (lane 0 decodes with shared-memory tables): candidate j writes
//                       its output at j * 16 KiB, records end offset / length / Adler-32 partials
//   k_inflate_validate  1 thread per stream: the candidates must chain exactly (each starts where the previous
//                       one ended, every chunk but the last inflates to 16 KiB, the last one ends in the final
//                       block, Adler-32 matches).  Anything else -- a foreign stream with other chunking, a
//                       marker pattern inside compressed data -- flags the stream for
//   k_inflate_serial    one warp per flagged stream: plain sequential inflate of the whole stream.
// Reference-written files (one zlib stream, no markers) have a single candidate that decodes the entire
// stream, which k_inflate_validate accepts directly.
#include "common.cuh"
#include "kernels.cuh"
#include "inflate_core.cuh"

constexpr int INF_CHUNK = 16384;
constexpr int IF_RETRY = 100;           // task code: left to k_inflate_chunks by k_inflate_lanes
constexpr int INF_WARPS = 2;             // decoding warps per CTA (about 21 KiB of shared memory each)

__global__ void __launch_bounds__(256)
k_inflate_scan(const uint8_t *__restrict__ in, const uint64_t *__restrict__ in_off,
               const uint32_t *__restrict__ in_bytes, uint32_t cmax, uint32_t *__restrict__ cand,
               uint32_t *__restrict__ ncand, uint32_t *__restrict__ status)
{
    __shared__ uint32_t s_warp[9];
    const int s = blockIdx.x, t = threadIdx.x;
    const uint8_t *p = in + in_off[s];
    const uint32_t n = in_bytes[s];
    uint32_t *c = cand + (size_t)s * cmax;
    bool ok = n >= 8;
    if (ok) {
        const uint32_t cmf = p[0], flg = p[1];
        ok = (cmf & 0x0f) == 8 && (cmf >> 4) <= 7 && ((cmf << 8) | flg) % 31 == 0 && !(flg & 0x20);
    }
    if (!ok) {
        if (t == 0) { ncand[s] = 0; status[s] = RC_STATUS_BAD_STREAM; }
        return;
    }
    if (t == 0) { status[s] = RC_STATUS_OK; c[0] = 2; }
    uint32_t carry = 1;
    // a marker at byte q means a candidate block start at q + 4; the 6 trailer bytes can never hold a start.
    // Every thread examines 16 consecutive positions per round (19 bytes, read as words when aligned), so a
    // round covers 4096 bytes with one block scan.
    const uint32_t last = n - 6;
    for (uint32_t q0 = 2; q0 + 4 <= last; q0 += 4096) {
        const uint32_t q = q0 + (uint32_t)t * 16;
        uint32_t hits = 0;                           // bit i: marker at q + i
        if (q + 4 <= last) {
            uint8_t b[19];
#pragma unroll
            for (int i = 0; i < 19; i++) b[i] = q + i < n ? p[q + i] : 0x55;
#pragma unroll
            for (int i = 0; i < 16; i++)
                if (q + i + 4 <= last && b[i] == 0 && b[i + 1] == 0 && b[i + 2] == 0xff && b[i + 3] == 0xff) hits |= 1u << i;
        }
        uint32_t total;
        uint32_t e = carry + block_excl_scan<8>(__popc(hits), s_warp, &total);
        while (hits) {
            const uint32_t i = __ffs(hits) - 1;
            hits &= hits - 1;
            if (e < cmax) c[e] = q + i + 4;
            e++;
        }
        carry += total;
        __syncthreads();
    }
    if (t == 0) ncand[s] = carry;    // may exceed cmax: the validator then falls back to serial decoding
}

__global__ void __launch_bounds__(256)
k_scan_u32(const uint32_t *__restrict__ v, int n, uint32_t clamp, uint32_t *__restrict__ out,
           uint32_t *__restrict__ counters)
{
    __shared__ uint32_t s_warp[9];
    uint32_t carry = 0;
    for (int i0 = 0; i0 < n; i0 += 256) {
        const int i = i0 + threadIdx.x;
        uint32_t x = i < n ? v[i] : 0;
        if (x > clamp) x = 0;                 // too many candidates: nothing to decode in parallel
        uint32_t total;
        const uint32_t e = block_excl_scan<8>(x, s_warp, &total);
        if (i < n) out[i] = carry + e;
        carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) { out[n] = carry; counters[0] = 0; counters[1] = 0; }
}

struct InfTask {
    uint32_t end;        // byte offset after the consumed data
    uint32_t out_len;
    uint32_t s1, s2;     // Adler-32 partials from (0, 0)
    int32_t code;        // IF_END_SYNC / IF_END_FINAL / error
};

// Per warp: a 16 KiB output buffer (one whole chunk of our encoder), a 2 KiB sliding window over the compressed
// bytes and the decoding tables, all in shared memory.  Lane 0 decodes; the 32 lanes zero the buffer before and
// copy it out -- and sum its Adler-32 partials -- afterwards, so global memory sees only coalesced 128-bit
// accesses.  Output beyond 16 KiB (a foreign stream decoded from its only candidate) goes to global memory byte
// by byte as before.
struct InfWarpShared {
    __align__(16) uint8_t out[INF_CHUNK];
    __align__(16) uint32_t win[IF_WIN_BYTES / 4];
    IfTables tab;
};

__global__ void __launch_bounds__(INF_WARPS * 32)
k_inflate_chunks(const uint8_t *__restrict__ in, const uint64_t *__restrict__ in_off,
                 const uint32_t *__restrict__ in_bytes, int n_streams, uint32_t cmax,
                 const uint32_t *__restrict__ cand, const uint32_t *__restrict__ task_base,
                 uint32_t *__restrict__ counters, uint8_t *__restrict__ out, size_t out_stride,
                 InfTask *__restrict__ tasks, int retry_only)
{
    __shared__ InfWarpShared s_w[INF_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    InfWarpShared &W = s_w[warp];
    const uint32_t total = task_base[n_streams];
    while (true) {
        uint32_t ti = 0;
        if (lane == 0) ti = atomicAdd(&counters[0], 1u);
        ti = __shfl_sync(0xffffffffu, ti, 0);
        if (ti >= total) break;
        if (retry_only && tasks[ti].code != IF_RETRY) continue;      // finished by k_inflate_lanes
        // stream / candidate of this task (all lanes)
        int s;
        {
            int lo = 0, hi = n_streams;
            while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (task_base[mid] <= ti) lo = mid; else hi = mid; }
            s = lo;
        }
        const uint32_t j = ti - task_base[s];
        const uint64_t ooff = (uint64_t)j * INF_CHUNK;
        const uint64_t cap = ooff < out_stride ? out_stride - ooff : 0;
        uint8_t *dst = out + (size_t)s * out_stride + ooff;
        // Fast path for the pieces our encoder stores: [00][LEN][~LEN][LEN bytes][00 00 00 FF FF], byte aligned.
        // The 32 lanes copy the payload and sum its Adler-32 partials; anything else takes the serial decoder.
        {
            const uint8_t *p = in + in_off[s];
            const uint32_t nb = in_bytes[s], st0 = cand[(size_t)s * cmax + j];
            bool fast = false;
            uint32_t len = 0;
            if ((uint64_t)st0 + 10 <= nb && (p[st0] & 7) == 0) {
                len = (uint32_t)p[st0 + 1] | ((uint32_t)p[st0 + 2] << 8);
                const uint32_t nlen = (uint32_t)p[st0 + 3] | ((uint32_t)p[st0 + 4] << 8);
                const uint64_t m = (uint64_t)st0 + 5 + len;
                fast = len > 0 && (len ^ 0xffffu) == nlen && m + 5 <= nb && len <= cap && p[m] == 0 && p[m + 1] == 0 &&
                       p[m + 2] == 0 && p[m + 3] == 0xff && p[m + 4] == 0xff;
            }
            if (fast) {
                const uint8_t *src = p + st0 + 5;
                uint32_t a = 0, b = 0;                          // sum c_i, sum (len - i) c_i
                for (uint32_t i = lane; i < len; i += 32) {
                    const uint32_t cc = src[i];
                    dst[i] = (uint8_t)cc;
                    a += cc;
                    b = (b + (len - i) * cc) % 65521u;
                }
                a %= 65521u;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) {
                    a += __shfl_down_sync(0xffffffffu, a, d);
                    b += __shfl_down_sync(0xffffffffu, b, d);
                }
                if (lane == 0) {
                    InfTask r;
                    r.end = st0 + 5 + len + 5; r.out_len = len; r.s1 = a % 65521u; r.s2 = b % 65521u; r.code = IF_END_SYNC;
                    tasks[ti] = r;
                }
                __syncwarp();
                continue;
            }
        }
        // zero the output buffer (runs of 0x00 are then not stored at all)
#pragma unroll 4
        for (int i = lane; i < INF_CHUNK / 16; i += 32) reinterpret_cast<uint4 *>(W.out)[i] = make_uint4(0, 0, 0, 0);
        __syncwarp();
        uint32_t o_n = 0, o_s1 = 0, o_s2 = 0, o_end = 0;
        int code = 0;
        if (lane == 0) {
            IfOut O;
            // a piece of our own encoder never exceeds one chunk; anything longer (a foreign stream: one block sequence
            // of megabytes) stops here with IF_ERR_OUT, fails validation and is decoded by k_inflate_serial through its
            // shared-memory history ring instead of byte by byte through global memory
            O.init(dst, cap < (uint64_t)INF_CHUNK ? cap : (uint64_t)INF_CHUNK, W.out, (uint32_t)INF_CHUNK);
            uint64_t end = 0;
            code = if_inflate(in + in_off[s], in_bytes[s], cand[(size_t)s * cmax + j], O, W.tab, true, &end, W.win);
            o_n = (uint32_t)O.n; o_s1 = O.s1 % 65521u; o_s2 = O.s2 % 65521u; o_end = (uint32_t)end;
#ifdef RC_DEBUG
            printf("[chunks] ti=%u s=%d j=%u start=%u code=%d end=%u out=%u\n", ti, s, j, cand[(size_t)s * cmax + j], code, o_end, o_n);
#endif
        }
        __syncwarp();
        o_n = __shfl_sync(0xffffffffu, o_n, 0);
        code = __shfl_sync(0xffffffffu, code, 0);
        if (code == IF_END_SYNC || code == IF_END_FINAL) {
            // copy the buffered part out and sum its Adler-32 partials {sum c_i, sum (nsh - i) c_i}; the bytes of the
            // buffer past nsh are still zero
            const uint32_t nsh = o_n < (uint32_t)INF_CHUNK ? o_n : (uint32_t)INF_CHUNK;
            const uint32_t nw = (nsh + 3) >> 2;
            uint32_t a = 0, b = 0;
            for (uint32_t k = lane; k < nw; k += 32) {
                const uint32_t x = reinterpret_cast<const uint32_t *>(W.out)[k];
                const uint32_t sm = __dp4a(x, 0x01010101u, 0u), wt = __dp4a(x, 0x03020100u, 0u);
                a += sm;
                b += (nsh - 4u * k) * sm - wt;          // per lane < 512 * 255 * 16384 < 2^32
            }
            a %= 65521u; b %= 65521u;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                a += __shfl_down_sync(0xffffffffu, a, d);
                b += __shfl_down_sync(0xffffffffu, b, d);
            }
            if (((uintptr_t)dst & 15) == 0) {
                const uint32_t nv = nsh >> 4;
                for (uint32_t k = lane; k < nv; k += 32) reinterpret_cast<uint4 *>(dst)[k] = reinterpret_cast<const uint4 *>(W.out)[k];
                for (uint32_t k = (nv << 4) + lane; k < nsh; k += 32) dst[k] = W.out[k];
            } else {
                for (uint32_t k = lane; k < nsh; k += 32) dst[k] = W.out[k];
            }
            if (lane == 0) {
                // buffered part A (nsh bytes) followed by the directly written part B (o_n - nsh bytes)
                a %= 65521u; b %= 65521u;
                const uint32_t nB = o_n - nsh;
                InfTask r;
                r.end = o_end; r.out_len = o_n; r.code = code;
                r.s1 = (a + o_s1) % 65521u;
                r.s2 = (uint32_t)(((uint64_t)b + (uint64_t)nB % 65521u * a + o_s2) % 65521u);
                tasks[ti] = r;
            }
        } else if (lane == 0) {
            InfTask r;
            r.end = o_end; r.out_len = o_n; r.s1 = 0; r.s2 = 0; r.code = code;
            tasks[ti] = r;
        }
        __syncwarp();
    }
}

// ---- lane-per-chunk decoding of our own streams -------------------------------------------------------------
// Every chunk written by deflate.cu is one dynamic block, and all chunks of a stream carry the SAME code (one
// code per group of streams at levels 1..5, one per stream at levels 6..9), i.e. bit-identical block headers.
// k_inflate_tables parses the header of a stream's first chunk once and leaves the decoding tables in global
// memory; k_inflate_lanes then gives every chunk whose header bits equal the first chunk's to ONE LANE: 32 chunks
// decode per warp instead of one, which is what the latency-bound serial decoding needs to fill the machine.
// A lane reads its compressed bytes through a preloaded 32-bit word (the load for the next refill is always in
// flight), keeps the current output word in a register and stores only non-zero words into the zero-filled
// output; distance-1 runs (the only matches our encoder emits) cost no memory access at all.  Anything a lane does
// not understand -- another header, another distance, output beyond 16 KiB, a stored piece -- is left to
// k_inflate_chunks (task code IF_RETRY).
constexpr int INF_LANES = 64;            // chunks per CTA of k_inflate_lanes

constexpr int INF_LUT_BITS = 15;         // k_inflate_lanes: every code of the stream in one table lookup (64 KiB)
constexpr int INF_LUT_SIZE = 1 << INF_LUT_BITS;

struct __align__(16) InfStreamTable {
    uint16_t lut[INF_LUT_SIZE];   // literal / length code: (len << 12) | symbol, 0 = unused bit pattern
    uint16_t dlut[IF_LUT_SIZE];   // distance code
    uint32_t hdr_bits;       // bits from the chunk start to its first token; 0 = no lane decoding for this stream
    uint32_t pad[3];
};

__global__ void __launch_bounds__(32)
k_inflate_tables(const uint8_t *__restrict__ in, const uint64_t *__restrict__ in_off,
                 const uint32_t *__restrict__ in_bytes, const uint32_t *__restrict__ ncand, uint32_t cmax,
                 InfStreamTable *__restrict__ tabs)
{
    __shared__ IfTables s_tab;
    __shared__ uint32_t s_bits;
    const int s = blockIdx.x, lane = threadIdx.x;
    const uint32_t nc = ncand[s];
    if (lane == 0) {
        s_bits = 0;
        if (nc >= 2 && nc <= cmax) {
            IfBits B;
            B.init(in + in_off[s], in_bytes[s], 2);
            const uint32_t bfinal = B.get(1), btype = B.get(2);
            if (!bfinal && btype == 2) {
                uint8_t lens[320];
                int nl = 0, nd = 0;
                if (if_dynamic_lengths(B, s_tab, lens, nl, nd) == IF_OK && !B.overrun() &&
                    if_build(s_tab.ll, lens, nl) == 0 && if_build(s_tab.d, lens + 288, nd) == 0)
                    s_bits = (uint32_t)(B.bitpos() - 16);
            }
        }
    }
    __syncwarp();
    InfStreamTable &T = tabs[s];
    if (lane == 0) T.hdr_bits = s_bits;
    if (!s_bits) return;
    // the wide LUT, all lanes: canonical code of the idx-th symbol in (length, value) order
    for (int i = lane; i < INF_LUT_SIZE; i += 32) T.lut[i] = 0;
    for (int i = lane; i < IF_LUT_SIZE; i += 32) T.dlut[i] = s_tab.d.lut[i];
    __syncwarp();
    uint32_t offs[17], first[17];
    offs[1] = 0; first[1] = 0;
    for (int l = 1; l <= 15; l++) {
        offs[l + 1] = offs[l] + s_tab.ll.count[l];
        first[l + 1] = (first[l] + s_tab.ll.count[l]) << 1;
    }
    for (uint32_t idx = lane; idx < offs[16]; idx += 32) {
        int l = 1;
        while (idx >= offs[l + 1]) l++;
        if (l > INF_LUT_BITS) continue;
        const uint32_t code = first[l] + (idx - offs[l]);
        const uint32_t r = __brev(code) >> (32 - l);
        const uint16_t e = (uint16_t)((l << 12) | s_tab.ll.symbol[idx]);
        for (uint32_t x = r; x < (uint32_t)INF_LUT_SIZE; x += 1u << l) T.lut[x] = e;
    }
}

__global__ void __launch_bounds__(INF_LANES)
k_inflate_lanes(const uint8_t *__restrict__ in, const uint64_t *__restrict__ in_off,
                const uint32_t *__restrict__ in_bytes, uint32_t cmax, const uint32_t *__restrict__ cand,
                const uint32_t *__restrict__ ncand, const uint32_t *__restrict__ task_base,
                const InfStreamTable *__restrict__ tabs, uint8_t *__restrict__ out, size_t out_stride,
                InfTask *__restrict__ tasks)
{
    extern __shared__ __align__(16) uint16_t s_lanes[];
    uint16_t *s_lut = s_lanes;                       // [INF_LUT_SIZE]
    uint16_t *s_dlut = s_lanes + INF_LUT_SIZE;       // [IF_LUT_SIZE]
    const int s = blockIdx.y, t = threadIdx.x;
    const uint32_t nc = ncand[s];
    const uint32_t j = blockIdx.x * INF_LANES + t;
    if (nc > cmax || blockIdx.x * INF_LANES >= nc) return;
    const InfStreamTable &T = tabs[s];
    const uint32_t hdr_bits = T.hdr_bits;
    const uint32_t tb = task_base[s];
    if (hdr_bits == 0) {
        if (j < nc) tasks[tb + j].code = IF_RETRY;
        return;
    }
    for (int i = t; i < INF_LUT_SIZE / 8; i += INF_LANES)
        reinterpret_cast<uint4 *>(s_lut)[i] = reinterpret_cast<const uint4 *>(T.lut)[i];
    for (int i = t; i < IF_LUT_SIZE / 8; i += INF_LANES)
        reinterpret_cast<uint4 *>(s_dlut)[i] = reinterpret_cast<const uint4 *>(T.dlut)[i];
    __syncthreads();
    if (j >= nc) return;
    const uint8_t *p = in + in_off[s];
    const uint32_t nb = in_bytes[s];
    const uint32_t st0 = cand[(size_t)s * cmax + j];
    InfTask r;
    r.end = 0; r.out_len = 0; r.s1 = 0; r.s2 = 0; r.code = IF_RETRY;
    // the stream's closing piece: an empty final fixed block (03 00) in front of the Adler-32 trailer
    if ((uint64_t)st0 + 6 == nb && p[st0] == 0x03 && p[st0 + 1] == 0x00) {
        r.end = st0 + 2; r.code = IF_END_FINAL;
        tasks[tb + j] = r;
        return;
    }
    const uint64_t ooff = (uint64_t)j * INF_CHUNK;
    const uint64_t room = ooff < out_stride ? out_stride - ooff : 0;
    const uint32_t cap = room < (uint64_t)INF_CHUNK ? (uint32_t)room : (uint32_t)INF_CHUNK;
    const uint32_t hbytes = (hdr_bits + 7) >> 3;
    bool ok = (uint64_t)st0 + hbytes + 8 <= nb;
    if (ok && j > 0) {
        // same header bits as the first chunk (whose tables these are)?
        const uint8_t *a = p + 2, *b = p + st0;
        uint32_t diff = 0;
        for (uint32_t i = 0; i + 1 < hbytes; i++) diff |= (uint32_t)(a[i] ^ b[i]);
        const uint32_t lastmask = (hdr_bits & 7) ? ((1u << (hdr_bits & 7)) - 1u) : 0xffu;
        diff |= (uint32_t)(a[hbytes - 1] ^ b[hbytes - 1]) & lastmask;
        ok = diff == 0;
    }
    if (!ok) { tasks[tb + j] = r; return; }

    // ---- bit reader: 64-bit buffer fed by aligned 32-bit words, the next word always preloaded
    const uint8_t *first = p + st0 + (hdr_bits >> 3);
    const uint32_t *wp = reinterpret_cast<const uint32_t *>((uintptr_t)first & ~(uintptr_t)3);
    const uint32_t *wend = reinterpret_cast<const uint32_t *>(((uintptr_t)(p + nb) + 3) & ~(uintptr_t)3);
    const uint32_t *wfirst = wp;
    const uint32_t skip = (uint32_t)((uintptr_t)first & 3) * 8u + (hdr_bits & 7u);
    uint64_t buf = (uint64_t)(wp < wend ? __ldg(wp) : 0u);
    wp++;
    buf |= (uint64_t)(wp < wend ? __ldg(wp) : 0u) << 32;
    wp++;
    buf >>= skip;
    int cnt = 64 - (int)skip;
    uint32_t nxt = wp < wend ? __ldg(wp) : 0u;
    wp++;
#define LN_REFILL()                                                      \
    if (cnt <= 32) {                                                     \
        buf |= (uint64_t)nxt << cnt;                                     \
        cnt += 32;                                                       \
        nxt = wp < wend ? __ldg(wp) : 0u;                                \
        wp++;                                                            \
    }
    // bits consumed since `first`'s word: 32 * (words moved into buf) - cnt
#define LN_BITPOS() ((uint64_t)((wp - wfirst) - 1) * 32u - (uint64_t)cnt)

    uint32_t *out32 = reinterpret_cast<uint32_t *>(out + (size_t)s * out_stride + ooff);
    uint32_t n = 0, w = 0, last = 0, s1 = 0;
    uint64_t s2 = 0;
    int code = IF_OK;
    while (true) {
        LN_REFILL();
        // literal / length symbol
        const uint32_t e = s_lut[(uint32_t)buf & (INF_LUT_SIZE - 1)];
        if (e == 0) { code = IF_RETRY; break; }      // not a code of this stream
        buf >>= (e >> 12); cnt -= (int)(e >> 12);
        const uint32_t sym = e & 0xfffu;
        if (sym == 256) break;
        // one straight-line path for literals and matches: a literal is a "run" of one byte; a match reads its
        // length bits and distance code, a literal reads nothing (zero bits)
        const bool is_match = sym > 256;
        const uint32_t li = is_match ? sym - 257 : 0u;
        if (li >= 29) { code = IF_RETRY; break; }
        const uint32_t eb = (li < 8 || li == 28) ? 0u : (li - 4) >> 2;
        const uint32_t lb = li < 8 ? li + 3 : (li == 28 ? 258u : ((4u + (li & 3u)) << eb) + 3u);
        const uint32_t L = is_match ? lb + ((uint32_t)buf & ((1u << eb) - 1u)) : 1u;
        buf >>= eb; cnt -= (int)eb;
        // (at least 33 bits were there: 15 code + 5 length bits + 9 distance code bits never run dry)
        const uint32_t d = s_dlut[(uint32_t)buf & (IF_LUT_SIZE - 1)];
        // only distance 1 (symbol 0, a run of the previous byte) is decoded here
        if (is_match && (d == 0 || (d & 0xfffu) != 0 || n == 0)) { code = IF_RETRY; break; }
        const uint32_t dl = is_match ? d >> 12 : 0u;
        buf >>= dl; cnt -= (int)dl;
        const uint32_t c = is_match ? last : sym;
        if (n + L > cap) { code = IF_RETRY; break; }
        // Adler-32 partials of L copies of c in closed form
        s2 += (uint64_t)L * s1 + (uint64_t)(c * ((L * (L + 1)) >> 1));
        s1 += c * L;
        if (c != 0 && L > 1) {
            // run of a non-zero byte: rare in binary maps
            for (uint32_t i = 0; i < L; i++) {
                w |= c << ((n & 3u) * 8u);
                n++;
                if ((n & 3u) == 0) { out32[(n >> 2) - 1] = w; w = 0; }
            }
        } else {
            // a zero run of any length or ONE non-zero byte: the same few instructions for every lane of the warp
            const uint32_t nn = n + L;
            w |= c << ((n & 3u) * 8u);
            if ((n ^ nn) >> 2) {
                if (w) out32[n >> 2] = w;
                w = 0;
            }
            n = nn;
        }
        last = c;
    }
    if (code == IF_OK) {
        // the chunk's block is followed by the sync marker: 000 (stored, not final), pad to the byte, 00 00 FF FF
        LN_REFILL();
        const uint32_t hb = (uint32_t)buf & 7u;
        buf >>= 3; cnt -= 3;
        const uint64_t bp = LN_BITPOS();
        const uint32_t pad = (uint32_t)((8u - (bp & 7u)) & 7u);
        buf >>= pad; cnt -= (int)pad;
        LN_REFILL();
        if (hb == 0 && (uint32_t)buf == 0xffff0000u) {
            const uint64_t endbit = LN_BITPOS() + 32u;
            const uint64_t endbyte = (uint64_t)((const uint8_t *)wfirst - p) + (endbit >> 3);
            if (endbyte <= nb) {
                if (w) {
                    // trailing partial word: byte stores, the bytes after it belong to someone else
                    uint8_t *o8 = reinterpret_cast<uint8_t *>(out32);
                    for (uint32_t k = n & ~3u; k < n; k++) o8[k] = (uint8_t)(w >> ((k & 3u) * 8u));
                }
                r.end = (uint32_t)endbyte; r.out_len = n; r.s1 = s1 % 65521u; r.s2 = (uint32_t)(s2 % 65521u);
                r.code = IF_END_SYNC;
            }
        }
    }
#undef LN_REFILL
#undef LN_BITPOS
    tasks[tb + j] = r;
}

__global__ void k_inflate_validate(const uint8_t *__restrict__ in, const uint64_t *__restrict__ in_off,
                                   const uint32_t *__restrict__ in_bytes, int n_streams, uint32_t cmax,
                                   const uint32_t *__restrict__ cand, const uint32_t *__restrict__ ncand,
                                   const uint32_t *__restrict__ task_base, const InfTask *__restrict__ tasks,
                                   uint32_t *__restrict__ out_bytes, uint32_t *__restrict__ status,
                                   uint32_t *__restrict__ need_serial)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_streams) return;
    need_serial[s] = 0;
    if (status[s] != RC_STATUS_OK) { out_bytes[s] = 0; return; }
    const uint32_t nc = ncand[s];
    bool ok = nc >= 1 && nc <= cmax;
    uint32_t pos = 2, total = 0, s1 = 1, s2 = 0;
    bool finished = false;
    for (uint32_t j = 0; ok && j < nc; j++) {
        const InfTask r = tasks[task_base[s] + j];
        if (cand[(size_t)s * cmax + j] != pos) { ok = false; break; }
        if (r.code != IF_END_SYNC && r.code != IF_END_FINAL) { ok = false; break; }
        if (r.out_len && total != j * (uint32_t)INF_CHUNK) { ok = false; break; }
        s2 = (uint32_t)(((uint64_t)s2 + (uint64_t)r.out_len * s1 + r.s2) % 65521u);
        s1 = (s1 + r.s1) % 65521u;
        total += r.out_len;
        pos = r.end;
        if (r.code == IF_END_FINAL) { finished = j + 1 == nc; ok = finished; break; }
    }
    if (ok && finished && (uint64_t)pos + 4 <= in_bytes[s]) {
        const uint8_t *tr = in + in_off[s] + pos;
        const uint32_t want = ((uint32_t)tr[0] << 24) | ((uint32_t)tr[1] << 16) | ((uint32_t)tr[2] << 8) | tr[3];
        if (want == ((s2 << 16) | s1)) { out_bytes[s] = total; return; }
    }
#ifdef RC_DEBUG
    printf("[validate] s=%d nc=%u ok=%d finished=%d pos=%u total=%u in_bytes=%u\n", s, nc, (int)ok, (int)finished, pos, total, in_bytes[s]);
#endif
    need_serial[s] = 1;
}

__global__ void __launch_bounds__(32)
k_inflate_serial(const uint8_t *__restrict__ in, const uint64_t *__restrict__ in_off,
                 const uint32_t *__restrict__ in_bytes, const uint32_t *__restrict__ need_serial,
                 uint8_t *__restrict__ out, size_t out_stride, uint32_t *__restrict__ out_bytes,
                 uint32_t *__restrict__ status)
{
    // A stream that is not a chain of our own chunks (a file written by the reference: one block sequence per stream,
    // matches at any distance) is decoded by one lane through a 32 KiB shared-memory history ring and a 2 KiB input
    // window: match copies and input bytes never wait for global memory.
    __shared__ IfTables s_tab;
    __shared__ __align__(16) uint8_t s_ring[IF_RING_BYTES];
    __shared__ __align__(16) uint32_t s_win[IF_WIN_BYTES / 4];
    const int s = blockIdx.x;
    if (!need_serial[s] || threadIdx.x != 0) return;
    IfOut O;
    O.init_ring(out + (size_t)s * out_stride, out_stride, s_ring);
    uint64_t end = 0;
    const int code = if_inflate(in + in_off[s], in_bytes[s], 2, O, s_tab, false, &end, s_win);
    O.flush(O.n);
    uint32_t st = RC_STATUS_OK;
    if (code == IF_ERR_OUT) st = RC_STATUS_OUT_OVERFLOW;
    else if (code != IF_END_FINAL || end + 4 > in_bytes[s]) st = RC_STATUS_BAD_STREAM;
    else {
        // Adler-32 from (0,0) partials: s1 = 1 + A, s2 = n + B
        const uint32_t a = (1u + O.s1 % 65521u) % 65521u;
        const uint32_t b = (uint32_t)(((uint64_t)O.n + O.s2) % 65521u);
        const uint8_t *tr = in + in_off[s] + end;
        const uint32_t want = ((uint32_t)tr[0] << 24) | ((uint32_t)tr[1] << 16) | ((uint32_t)tr[2] << 8) | tr[3];
        if (want != ((b << 16) | a)) st = RC_STATUS_BAD_STREAM;
    }
#ifdef RC_DEBUG
    printf("[serial] s=%d code=%d end=%llu n=%llu st=%u\n", s, code, (unsigned long long)end, (unsigned long long)O.n, st);
#endif
    out_bytes[s] = (uint32_t)O.n;
    status[s] = st;
}

size_t inflate_cmax(size_t out_stride) { return out_stride / INF_CHUNK + 3; }

size_t inflate_workspace_bytes(int n_streams, size_t out_stride)
{
    Carver c(nullptr);
    const size_t cmax = inflate_cmax(out_stride);
    c.take<uint32_t>((size_t)n_streams * cmax);
    c.take<uint32_t>((size_t)n_streams + 1);
    c.take<uint32_t>((size_t)n_streams + 1);
    c.take<uint32_t>(8);
    c.take<InfTask>((size_t)n_streams * cmax + 1);
    c.take<uint32_t>((size_t)n_streams + 1);
    c.take<InfStreamTable>((size_t)n_streams);
    return c.used();
}

int launch_inflate(rc_ctx *ctx, const uint8_t *in, const uint64_t *in_off, const uint32_t *in_bytes, int n_streams,
                   void *ws, uint8_t *out, size_t out_stride, uint32_t *out_bytes, uint32_t *status, cudaStream_t st)
{
    if (n_streams <= 0) return 0;
    Carver c(ws);
    const uint32_t cmax = (uint32_t)inflate_cmax(out_stride);
    uint32_t *cand = c.take<uint32_t>((size_t)n_streams * cmax);
    uint32_t *ncand = c.take<uint32_t>((size_t)n_streams + 1);
    uint32_t *task_base = c.take<uint32_t>((size_t)n_streams + 1);
    uint32_t *counters = c.take<uint32_t>(8);
    InfTask *tasks = c.take<InfTask>((size_t)n_streams * cmax + 1);
    uint32_t *need_serial = c.take<uint32_t>((size_t)n_streams + 1);
    InfStreamTable *tabs = c.take<InfStreamTable>((size_t)n_streams);

    ctx->n_dmarks = 0;                               // profile level 2: scan + tables | lanes | the rest
    rc_dmark(ctx, 0, st);
    k_inflate_scan<<<n_streams, 256, 0, st>>>(in, in_off, in_bytes, cmax, cand, ncand, status);
    RC_LAUNCH_CHECK(ctx, "k_inflate_scan");
    k_scan_u32<<<1, 256, 0, st>>>(ncand, n_streams, cmax, task_base, counters);
    RC_LAUNCH_CHECK(ctx, "k_scan_u32");
    // lane-per-chunk pass over the streams of our own encoder (word stores into a zero-filled output)
    const int lanes = (((uintptr_t)out | (uintptr_t)out_stride) & 3) == 0;
    if (lanes) {
        RC_CUDA(ctx, cudaMemsetAsync(out, 0, (size_t)n_streams * out_stride, st));
        k_inflate_tables<<<n_streams, 32, 0, st>>>(in, in_off, in_bytes, ncand, cmax, tabs);
        RC_LAUNCH_CHECK(ctx, "k_inflate_tables");
        dim3 grid((cmax + INF_LANES - 1) / INF_LANES, (unsigned)n_streams);
        constexpr int lanes_smem = (INF_LUT_SIZE + IF_LUT_SIZE) * (int)sizeof(uint16_t);
        if (!ctx->inflate_attr_set) {
            RC_CUDA(ctx, cudaFuncSetAttribute(k_inflate_lanes, cudaFuncAttributeMaxDynamicSharedMemorySize, lanes_smem));
            ctx->inflate_attr_set = true;
        }
        rc_dmark(ctx, 1, st);
        k_inflate_lanes<<<grid, INF_LANES, lanes_smem, st>>>(in, in_off, in_bytes, cmax, cand, ncand, task_base, tabs, out,
                                                   out_stride, tasks);
        RC_LAUNCH_CHECK(ctx, "k_inflate_lanes");
        rc_dmark(ctx, 2, st);
    }
    size_t blocks = ((size_t)n_streams * cmax + INF_WARPS - 1) / INF_WARPS;
    const size_t cap = (size_t)ctx->sm_count * 5;            // 5 CTAs of 43 KiB fit one SM
    if (blocks > cap) blocks = cap;
    k_inflate_chunks<<<(unsigned)blocks, INF_WARPS * 32, 0, st>>>(in, in_off, in_bytes, n_streams, cmax, cand, task_base,
                                                                 counters, out, out_stride, tasks, lanes);
    RC_LAUNCH_CHECK(ctx, "k_inflate_chunks");
    k_inflate_validate<<<(n_streams + 127) / 128, 128, 0, st>>>(in, in_off, in_bytes, n_streams, cmax, cand, ncand,
                                                               task_base, tasks, out_bytes, status, need_serial);
    RC_LAUNCH_CHECK(ctx, "k_inflate_validate");
    k_inflate_serial<<<n_streams, 32, 0, st>>>(in, in_off, in_bytes, need_serial, out, out_stride, out_bytes, status);
    RC_LAUNCH_CHECK(ctx, "k_inflate_serial");
    rc_dmark(ctx, 3, st);
    return 0;
}
