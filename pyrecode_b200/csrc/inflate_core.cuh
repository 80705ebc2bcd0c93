// inflate_core.cuh -- sequential deflate decoder used by one lane of a warp.
//
// Replaces zlib.decompress on the read path (reference: pyrecode/recode_compressors.py:42-43, called from
// recode_reader.py:397-400,428-431,454).  Accepts every deflate block type (stored, fixed, dynamic) because
// reference-written files are ordinary multi-block dynamic streams.
//
// The decoder works on a [start, ...) bit range of a compressed buffer and can stop at the first empty
// stored block (the sync-flush marker our encoder puts after every chunk), which is what makes streams
// written by deflate.cu decodable chunk-parallel.  __host__ __device__ so tests/csrc can run the same code
// against stock zlib output on the CPU (test infrastructure only).
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define IF_HD __host__ __device__ __forceinline__
#else
#define IF_HD inline
#endif

constexpr int IF_LUT_BITS = 9;
constexpr int IF_LUT_SIZE = 1 << IF_LUT_BITS;

// decoding tables of one Huffman code: fast LUT + canonical (count/symbol) fallback for longer codes
struct IfTable {
    uint16_t lut[IF_LUT_SIZE];     // (len << 12) | symbol, len = 0 -> not in LUT
    uint16_t count[16];            // codes per length
    uint16_t symbol[288];          // symbols ordered by (length, value)
};

struct IfTables {
    IfTable ll;    // literal / length
    IfTable d;     // distance (only the first 32 symbol slots used)
};

enum { IF_OK = 0, IF_END_SYNC = 1, IF_END_FINAL = 2, IF_ERR_DATA = -1, IF_ERR_OUT = -2, IF_ERR_IN = -3 };

constexpr int IF_WIN_BYTES = 2048;        // shared-memory input window of a decoding lane (device)

struct IfBits {
    const uint8_t *in;
    uint64_t nbytes;     // bytes available
    uint64_t pos;        // next byte to load
    uint64_t buf;
    int cnt;             // valid bits in buf
    // Device: a shared-memory window over the compressed bytes.  Window byte k mirrors the 16-byte aligned global
    // address `in + win_lo + k`; the decoding lane slides it forward itself with 128-bit loads (many in flight)
    // instead of paying one global-memory latency per input byte.
    uint32_t *win;       // IF_WIN_BYTES / 4 words, or nullptr (host, serial fallback)
    int64_t win_lo, win_hi;   // stream offsets covered by the window: [win_lo, win_hi)
    IF_HD void init(const uint8_t *p, uint64_t n, uint64_t start_byte, uint32_t *window = nullptr)
    {
        in = p; nbytes = n; pos = start_byte; buf = 0; cnt = 0;
        win = window; win_lo = 0; win_hi = 0;
    }
#ifdef __CUDA_ARCH__
    __device__ __forceinline__ bool slide()
    {
        // whole 16-byte units that lie inside the stream's buffer, starting at the unit that holds `pos`
        const uintptr_t a0 = (uintptr_t)(in + pos) & ~(uintptr_t)15;
        const uintptr_t a1 = (uintptr_t)(in + nbytes) & ~(uintptr_t)15;
        if (a0 + 16 > a1) return false;
        uint32_t units = (uint32_t)((a1 - a0) >> 4);
        if (units > IF_WIN_BYTES / 16) units = IF_WIN_BYTES / 16;
        const uint4 *src = reinterpret_cast<const uint4 *>(a0);
        uint4 *dst = reinterpret_cast<uint4 *>(win);
        uint32_t u = 0;
        for (; u + 8 <= units; u += 8) {
            uint4 v[8];
#pragma unroll
            for (int k = 0; k < 8; k++) v[k] = src[u + k];
#pragma unroll
            for (int k = 0; k < 8; k++) dst[u + k] = v[k];
        }
        for (; u < units; u++) dst[u] = src[u];
        win_lo = (int64_t)pos - (int64_t)((uintptr_t)(in + pos) - a0);
        win_hi = win_lo + (int64_t)units * 16;
        return true;
    }
#endif
    IF_HD void refill()
    {
#ifdef __CUDA_ARCH__
        if (win) {
            if (!((int64_t)pos >= win_lo && (int64_t)pos + 12 <= win_hi)) {
                if (pos + 12 <= nbytes) slide();
            }
            if ((int64_t)pos >= win_lo && (int64_t)pos + 12 <= win_hi) {
                // 8 stream bytes at `pos` from three aligned words; the bytes above `cnt` that are ORed in early
                // are the same bytes a later refill puts there again
                const uint32_t o = (uint32_t)((int64_t)pos - win_lo);
                const uint32_t *w = win + (o >> 2);
                const uint32_t sh = (o & 3u) * 8u;
                const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
                const uint32_t lo = __funnelshift_r(w0, w1, sh), hi = __funnelshift_r(w1, w2, sh);
                buf |= (((uint64_t)hi << 32) | lo) << cnt;
                const int nb = (63 - cnt) >> 3;
                pos += (uint64_t)nb;
                cnt += nb * 8;
                return;
            }
        }
#endif
        while (cnt <= 56) {
            const uint64_t b = pos < nbytes ? in[pos] : 0;    // zeros past the end; overrun is checked by callers
            pos++;
            buf |= b << cnt;
            cnt += 8;
        }
    }
    IF_HD uint32_t peek(int n) const { return (uint32_t)(buf & ((1ull << n) - 1)); }
    IF_HD void drop(int n) { buf >>= n; cnt -= n; }
    IF_HD uint32_t get(int n)
    {
        if (cnt < n) refill();
        const uint32_t v = peek(n);
        drop(n);
        return v;
    }
    // position (in bits) of the next unread bit
    IF_HD uint64_t bitpos() const { return pos * 8 - (uint64_t)cnt; }
    IF_HD bool overrun() const { return bitpos() > nbytes * 8; }
    IF_HD void align_byte() { drop(cnt & 7); }
};

IF_HD uint32_t if_bitrev(uint32_t c, int n)
{
    uint32_t r = 0;
    for (int i = 0; i < n; i++) { r = (r << 1) | (c & 1); c >>= 1; }
    return r;
}

// builds a table from code lengths; returns 0 ok, <0 over-subscribed / incomplete (incomplete is accepted only
// for a single one-bit code, like zlib's inflate_table)
IF_HD int if_build(IfTable &T, const uint8_t *len, int n)
{
    for (int i = 0; i < 16; i++) T.count[i] = 0;
    for (int s = 0; s < n; s++) T.count[len[s]]++;
    for (int i = 0; i < IF_LUT_SIZE; i++) T.lut[i] = 0;
    if (T.count[0] == n) return 0;                 // no codes at all (legal for distances)
    int left = 1;
    for (int l = 1; l <= 15; l++) {
        left <<= 1;
        left -= T.count[l];
        if (left < 0) return -1;                   // over-subscribed
    }
    uint16_t offs[16];
    offs[1] = 0;
    for (int l = 1; l < 15; l++) offs[l + 1] = offs[l] + T.count[l];
    for (int s = 0; s < n; s++) if (len[s]) T.symbol[offs[len[s]]++] = (uint16_t)s;
    // fast LUT for codes up to IF_LUT_BITS
    uint32_t code = 0;
    int idx = 0;
    for (int l = 1; l <= 15; l++) {
        for (int c = 0; c < T.count[l]; c++, idx++, code++) {
            if (l <= IF_LUT_BITS) {
                const uint32_t r = if_bitrev(code, l);
                const uint16_t e = (uint16_t)((l << 12) | T.symbol[idx]);
                for (uint32_t x = r; x < (uint32_t)IF_LUT_SIZE; x += 1u << l) T.lut[x] = e;
            }
        }
        code <<= 1;
    }
    if (left > 0) {
        // incomplete: only tolerated when there is exactly one code, of length 1
        int total = 0;
        for (int l = 1; l <= 15; l++) total += T.count[l];
        if (!(total == 1 && T.count[1] == 1)) return -2;
    }
    return 0;
}

// decode one symbol; returns -1 on invalid code
IF_HD int if_decode(IfBits &B, const IfTable &T)
{
    if (B.cnt < 15) B.refill();
    const uint16_t e = T.lut[B.peek(IF_LUT_BITS)];
    if (e) {
        B.drop(e >> 12);
        return e & 0xfff;
    }
    // canonical walk (puff-style), MSB-first code
    int code = 0, first = 0, index = 0;
    uint64_t bits = B.buf;
    for (int l = 1; l <= 15; l++) {
        code |= (int)(bits & 1);
        bits >>= 1;
        const int count = T.count[l];
        if (code - count < first) {
            B.drop(l);
            return T.symbol[index + (code - first)];
        }
        index += count;
        first += count;
        first <<= 1;
        code <<= 1;
    }
    return -1;
}

constexpr uint32_t IF_RING_BYTES = 32768;   // ring mode: deflate's whole history window (a power of two)
constexpr uint32_t IF_FLUSH_BYTES = 4096;   // ... written back to global memory in segments of this size

struct IfOut {
    uint8_t *out;
    uint64_t cap;
    uint64_t n;          // bytes produced
    uint32_t s1, s2;     // Adler-32 partials, started from (0, 0), of the bytes produced at index >= sh_cap
                         // (ring mode: of the bytes flushed so far, reduced mod 65521 -- all of them after flush(n))
    // Device, own chunks: the first sh_cap bytes (a whole chunk of our own encoder) are produced in a shared-memory
    // buffer, zero-filled beforehand, that the warp copies out -- and checksums -- together afterwards.  Runs of 0x00,
    // most of a binary map, then need no stores at all, and the byte before a run is read back from shared memory.
    uint8_t *sh;
    uint32_t sh_cap;
    // Ring mode (foreign streams: one block sequence of megabytes with matches at any distance up to 32 KiB): `sh` is
    // a circular window of IF_RING_BYTES over ALL the output, so a match copies shared memory to shared memory
    // instead of paying a global store -> load round trip per byte; finished IF_FLUSH_BYTES segments go to `out`
    // with 128-bit stores.
    bool ring;
    uint64_t flushed;
    IF_HD void init(uint8_t *o, uint64_t capacity, uint8_t *shared = nullptr, uint32_t shared_cap = 0)
    {
        out = o; cap = capacity; n = 0; s1 = 0; s2 = 0; sh = shared; sh_cap = shared_cap; ring = false; flushed = 0;
    }
    IF_HD void init_ring(uint8_t *o, uint64_t capacity, uint8_t *ring_buf)
    {
        out = o; cap = capacity; n = 0; s1 = 0; s2 = 0; sh = ring_buf; sh_cap = 0; ring = true; flushed = 0;
    }
    IF_HD uint8_t at(uint64_t i) const
    {
        if (ring) return sh[(uint32_t)i & (IF_RING_BYTES - 1)];
        return i < sh_cap ? sh[i] : out[i];
    }
    // ring mode: copy [flushed, upto) to global memory and fold those bytes into the Adler-32 partials (nothing is
    // summed per byte while decoding); upto = n at the end of the stream, else a segment boundary.  Pieces of at most
    // IF_FLUSH_BYTES: with L bytes c_0.. after state (s1, s2), s1' = s1 + sum c_i, s2' = s2 + L s1 + sum (L - i) c_i,
    // and the last sum stays below 255 * 4096 * 4097 / 2 < 2^32.
    IF_HD void flush(uint64_t upto)
    {
        while (flushed < upto) {
            const uint64_t lim = upto - flushed > IF_FLUSH_BYTES ? flushed + IF_FLUSH_BYTES : upto;
            const uint32_t L = (uint32_t)(lim - flushed);
            uint32_t A = 0, W = 0;
            uint64_t i = flushed;
#ifdef __CUDA_ARCH__
            if (((((uintptr_t)out) | (uintptr_t)flushed) & 15) == 0)
                for (; i + 16 <= lim; i += 16) {
                    const uint4 v = *reinterpret_cast<const uint4 *>(sh + ((uint32_t)i & (IF_RING_BYTES - 1)));
                    *reinterpret_cast<uint4 *>(out + i) = v;
                    const uint32_t a = __dp4a(v.x, 0x01010101u, __dp4a(v.y, 0x01010101u,
                                       __dp4a(v.z, 0x01010101u, __dp4a(v.w, 0x01010101u, 0u))));
                    const uint32_t w = __dp4a(v.x, 0x03020100u, __dp4a(v.y, 0x07060504u,
                                       __dp4a(v.z, 0x0b0a0908u, __dp4a(v.w, 0x0f0e0d0cu, 0u))));
                    A += a;
                    W += (uint32_t)(lim - i) * a - w;
                }
#endif
            for (; i < lim; i++) {
                const uint32_t c = sh[(uint32_t)i & (IF_RING_BYTES - 1)];
                out[i] = (uint8_t)c;
                A += c;
                W += (uint32_t)(lim - i) * c;
            }
            s2 = (uint32_t)(((uint64_t)s2 + (uint64_t)L * s1 + W) % 65521u);
            s1 = (s1 + A) % 65521u;
            flushed = lim;
        }
    }
    IF_HD void put(uint8_t c)
    {
        if (ring) {
            sh[(uint32_t)n & (IF_RING_BYTES - 1)] = c;
            n++;
            if (((uint32_t)n & (IF_FLUSH_BYTES - 1)) == 0) flush(n);
            return;
        }
        if (n < sh_cap) { sh[n++] = c; return; }
        out[n++] = c;
        s1 += c; s2 += s1;
        if ((n & 2047) == 0) { s1 %= 65521u; s2 %= 65521u; }
    }
    // len copies of c (a distance-1 match): Adler-32 in closed form, s1' = s1 + len c,
    // s2' = s2 + len s1 + c len (len + 1) / 2
    IF_HD void put_run(uint8_t c, uint32_t len)
    {
        if (ring) {
            const uint32_t at0 = (uint32_t)n;
            for (uint32_t i = 0; i < len; i++) sh[(at0 + i) & (IF_RING_BYTES - 1)] = c;
            n += len;
            if ((at0 ^ (uint32_t)n) & ~(IF_FLUSH_BYTES - 1)) flush(n & ~(uint64_t)(IF_FLUSH_BYTES - 1));
            return;
        }
        if (n < sh_cap) {
            const uint32_t k = (uint64_t)len < sh_cap - n ? len : (uint32_t)(sh_cap - n);
            if (c != 0)
                for (uint32_t i = 0; i < k; i++) sh[n + i] = c;
            n += k;
            len -= k;
            if (len == 0) return;
        }
        for (uint32_t i = 0; i < len; i++) out[n + i] = c;
        n += len;
        s1 %= 65521u;
        s2 = (uint32_t)(((uint64_t)s2 + (uint64_t)len * s1 + (uint64_t)c * (len * (len + 1) / 2)) % 65521u);
        s1 = (s1 + len * (uint32_t)c) % 65521u;
    }
    // a match at distance >= 2 (1 <= dist <= n, len <= 258).  Ring mode copies min(dist, 8) bytes per round: the
    // source bytes of a round lie entirely before its destination, so the loads are independent of each other and of
    // the round's stores -- one shared-memory round trip per 8 bytes instead of one per byte (stock zlib's matches on
    // a binary map are short, 8.6 bytes on average, and almost never at distance 1).
    IF_HD void copy_match(uint32_t dist, uint32_t len)
    {
        if (ring) {
            const uint32_t M = IF_RING_BYTES - 1, d0 = (uint32_t)n;
            uint32_t d = d0, s = d0 - dist;
            n += len;
            const uint32_t step = dist < 8u ? dist : 8u;
            while (len) {
                const uint32_t k = len < step ? len : step;
                uint8_t b[8];
                for (uint32_t j = 0; j < 8; j++)
                    if (j < k) b[j] = sh[(s + j) & M];
                for (uint32_t j = 0; j < 8; j++)
                    if (j < k) sh[(d + j) & M] = b[j];
                s += k; d += k; len -= k;
            }
            if ((d0 ^ (uint32_t)n) & ~(IF_FLUSH_BYTES - 1)) flush(n & ~(uint64_t)(IF_FLUSH_BYTES - 1));
            return;
        }
        for (uint32_t i = 0; i < len; i++) put(at(n - dist));
    }
};

// RFC 1951 3.2.5 in closed form (no table in local memory on the device): length symbol 257 + li -> base length and
// extra bits; distance symbol ds -> base distance and extra bits
IF_HD uint32_t if_len_base(int li, int &ext)
{
    if (li < 8) { ext = 0; return 3u + (uint32_t)li; }
    if (li == 28) { ext = 0; return 258u; }
    ext = (li >> 2) - 1;
    return 3u + ((4u + ((uint32_t)li & 3u)) << ext);
}
IF_HD uint32_t if_dist_base(int ds, int &ext)
{
    if (ds < 4) { ext = 0; return 1u + (uint32_t)ds; }
    ext = (ds >> 1) - 1;
    return 1u + ((2u + ((uint32_t)ds & 1u)) << ext);
}

// Reads the code description of a dynamic block (RFC 1951 3.2.7) that follows the 3 block-header bits:
// lens[0..nlen_codes) literal/length code lengths, lens[288..288 + ndist_codes) distance code lengths.
// T.d is used as scratch for the code-length code.
IF_HD int if_dynamic_lengths(IfBits &B, IfTables &T, uint8_t *lens, int &nlen_codes, int &ndist_codes)
{
    const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    nlen_codes = (int)B.get(5) + 257;
    ndist_codes = (int)B.get(5) + 1;
    const int ncl = (int)B.get(4) + 4;
    if (nlen_codes > 286 || ndist_codes > 30) return IF_ERR_DATA;
    uint8_t cl[19];
    for (int i = 0; i < 19; i++) cl[i] = 0;
    for (int i = 0; i < ncl; i++) cl[order[i]] = (uint8_t)B.get(3);
    if (if_build(T.d, cl, 19) != 0) return IF_ERR_DATA;     // T.d reused as the code-length table
    int idx = 0;
    while (idx < nlen_codes + ndist_codes) {
        const int sym = if_decode(B, T.d);
        if (sym < 0) return IF_ERR_DATA;
        if (sym < 16) lens[idx++] = (uint8_t)sym;
        else {
            int rep, val = 0;
            if (sym == 16) {
                if (idx == 0) return IF_ERR_DATA;
                val = lens[idx - 1];
                rep = 3 + (int)B.get(2);
            } else if (sym == 17) rep = 3 + (int)B.get(3);
            else rep = 11 + (int)B.get(7);
            if (idx + rep > nlen_codes + ndist_codes) return IF_ERR_DATA;
            while (rep--) lens[idx++] = (uint8_t)val;
        }
    }
    if (lens[256] == 0) return IF_ERR_DATA;
    // move distance lengths to a fixed place
    uint8_t dl[30];
    for (int i = 0; i < ndist_codes; i++) dl[i] = lens[nlen_codes + i];
    for (int i = 0; i < ndist_codes; i++) lens[288 + i] = dl[i];
    return IF_OK;
}

// Decodes blocks starting at byte `start` of in[0..nbytes).  Stops after the final block (IF_END_FINAL) or,
// if stop_at_sync, right after an empty stored block (IF_END_SYNC).  *end_byte receives the byte offset
// following the consumed data (only meaningful for the two END codes: both end byte aligned... the final
// block does not: it is rounded up).  Window = everything written to O.out so far (O.n); a distance reaching
// before O.out[0] is an error, i.e. independent chunks only.
IF_HD int if_inflate(const uint8_t *in, uint64_t nbytes, uint64_t start, IfOut &O, IfTables &T, bool stop_at_sync,
                     uint64_t *end_byte, uint32_t *window = nullptr)
{
    IfBits B;
    B.init(in, nbytes, start, window);
    while (true) {
        const uint32_t bfinal = B.get(1);
        const uint32_t btype = B.get(2);
        if (btype == 0) {
            B.align_byte();
            const uint32_t len = B.get(16), nlen = B.get(16);
            if ((len ^ 0xffffu) != nlen) return IF_ERR_DATA;
            if (B.overrun()) return IF_ERR_IN;
            if (O.n + len > O.cap) return IF_ERR_OUT;
            // the bit buffer holds whole bytes here
            for (uint32_t i = 0; i < len; i++) O.put((uint8_t)B.get(8));
            if (B.overrun()) return IF_ERR_IN;
            if (len == 0 && !bfinal && stop_at_sync) { *end_byte = B.bitpos() >> 3; return IF_END_SYNC; }
        } else if (btype == 1 || btype == 2) {
            uint8_t lens[320];
            int nlen_codes, ndist_codes;
            if (btype == 1) {
                for (int i = 0; i < 144; i++) lens[i] = 8;
                for (int i = 144; i < 256; i++) lens[i] = 9;
                for (int i = 256; i < 280; i++) lens[i] = 7;
                for (int i = 280; i < 288; i++) lens[i] = 8;
                for (int i = 0; i < 32; i++) lens[288 + i] = 5;     // 30 and 31 never occur in valid data
                nlen_codes = 288; ndist_codes = 32;
            } else {
                const int hrc = if_dynamic_lengths(B, T, lens, nlen_codes, ndist_codes);
                if (hrc != IF_OK) return hrc;
            }
            if (B.overrun()) return IF_ERR_IN;
            if (if_build(T.ll, lens, nlen_codes) != 0) return IF_ERR_DATA;
            if (if_build(T.d, lens + 288, ndist_codes) != 0) return IF_ERR_DATA;
            while (true) {
                const int sym = if_decode(B, T.ll);
                if (sym < 0) return IF_ERR_DATA;
                if (sym < 256) {
                    if (O.n >= O.cap) return IF_ERR_OUT;
                    O.put((uint8_t)sym);
                } else if (sym == 256) {
                    break;
                } else {
                    const int li = sym - 257;
                    if (li >= 29) return IF_ERR_DATA;
                    int ext;
                    uint32_t len = if_len_base(li, ext);
                    if (ext) len += B.get(ext);
                    const int ds = if_decode(B, T.d);
                    if (ds < 0 || ds >= 30) return IF_ERR_DATA;
                    uint32_t dist = if_dist_base(ds, ext);
                    if (ext) dist += B.get(ext);
                    if (dist > O.n) return IF_ERR_DATA;          // reaches before this range's own output
                    if (O.n + len > O.cap) return IF_ERR_OUT;
                    if (dist == 1) {
                        // byte run (the only match our own encoder emits): one read of the previous byte, then
                        // stores only -- no store -> load round trip through memory per byte
                        O.put_run(O.at(O.n - 1), len);
                    } else {
                        O.copy_match(dist, len);
                    }
                }
                if (B.pos > B.nbytes && B.overrun()) return IF_ERR_IN;    // overrun implies pos > nbytes
            }
        } else {
            return IF_ERR_DATA;
        }
        if (bfinal) {
            *end_byte = (B.bitpos() + 7) >> 3;
            return B.overrun() ? IF_ERR_IN : IF_END_FINAL;
        }
    }
}
