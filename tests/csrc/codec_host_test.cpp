// codec_host_test.cpp -- TEST INFRASTRUCTURE: runs the __host__ __device__ phases of the GPU deflate encoder
// (pyrecode_b200/csrc/deflate_chunk.cuh) and the inflate core (inflate_core.cuh) on the CPU, one simulated
// thread at a time, in exactly the order the kernels in deflate.cu / inflate.cu run them.  Lets the non-GPU
// test suite check the codec logic against stock zlib.  Never linked into the product library.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>
#include "../../pyrecode_b200/csrc/deflate_chunk.cuh"
#include "../../pyrecode_b200/csrc/inflate_core.cuh"

// mirrors k_deflate_hist / k_deflate_tables / k_deflate_chunks for one stream (a group of one stream)
extern "C" long host_deflate_stream(const uint8_t *in, uint32_t n, int level, uint8_t *out, uint32_t cap)
{
    static DfBuildShared B;
    static DfEmitShared S;
    static uint32_t in32[DF_STAGE_WORDS];
    const uint32_t nchunks = (n + DF_CHUNK - 1) / DF_CHUNK;
    const bool shared_table = level < 6;
    const uint32_t step = shared_table ? 8 : 1;
    std::vector<uint32_t> ghist(DF_NSYM, 0);
    auto stage = [&](uint32_t *dst, uint32_t off, int clen) {
        for (int i = 0; i < DF_STAGE_WORDS; i++) dst[i] = 0;
        for (int o = 0; o < DF_CHUNK; o += 4) {
            uint32_t w = 0;
            for (int b = 0; b < 4; b++) if (o + b < clen) w |= (uint32_t)in[off + o + b] << (8 * b);
            df_store_word(dst, o, w);
        }
    };
    // histogram (sampled chunks when the table is shared)
    uint32_t sampled = 0;
    if (level > 0)
        for (uint32_t c = 0; c < nchunks; c += step) {
            const int clen = (int)std::min<uint32_t>(DF_CHUNK, n - c * DF_CHUNK);
            uint32_t hist[DF_NSYM] = {0};
            stage(in32, c * DF_CHUNK, clen);
            for (int t = 0; t < DF_THREADS; t++) df_phase_hist(in32, hist, t, clen);
            for (int i = 0; i < DF_NSYM; i++) ghist[i] += hist[i];
            sampled += clen;
        }
    // tables
    bool store_all = level == 0;
    if (level > 0 && nchunks) {
        double ntok = 0, bits = 0;
        for (int i = 0; i < DF_NSYM; i++) ntok += ghist[i];
        for (int i = 0; i < DF_NSYM; i++)
            if (ghist[i]) bits += ghist[i] * (std::log2(ntok / ghist[i]) + (i > 264 ? 2. : (i > 256 ? 1. : 0.)));
        store_all = bits * 0.125 + 128. >= (shared_table ? 0.90 : 0.97) * sampled;
        if (!store_all) {
            const uint32_t nt = (uint32_t)ntok;
            int cshift = 0;
            while ((nt >> 21) >> cshift) cshift++;
            for (int i = 0; i < 512; i++) {
                uint32_t c = i < DF_NSYM ? ghist[i] : 0;
                if (c) c = std::max(c >> cshift, 1u);
                if (shared_table && i < DF_LEN_SYMS) c = c * 2u + 1u;
                if (i == 256 && c == 0) c = 1;
                B.keys[i] = c ? ((c << 9) | (uint32_t)i) : 0xffffffffu;
            }
            std::sort(B.keys, B.keys + 512);
            int n_used = 0;
            while (n_used < 512 && B.keys[n_used] != 0xffffffffu) n_used++;
            df_phase_build(B, n_used);
        }
    }
    // chunks
    uint32_t pos = 0;
    if (cap < 8) return -1;
    out[pos++] = 0x78; out[pos++] = 0x01;
    uint32_t s1 = 1, s2 = 0;
    std::vector<uint32_t> bits_of(DF_THREADS);
    for (uint32_t c = 0; c < nchunks; c++) {
        const int clen = (int)std::min<uint32_t>(DF_CHUNK, n - c * DF_CHUNK);
        stage(S.io, c * DF_CHUNK, clen);
        uint32_t ca = 0, cb = 0;
        for (int t = 0; t < DF_THREADS; t++) {
            uint32_t a, b;
            df_adler_partial(S.io, t, clen, a, b);
            ca = (ca + a % 65521u) % 65521u; cb = (cb + b % 65521u) % 65521u;
        }
        uint32_t body_bits = 0;
        bool stored = store_all;
        bool sparse = false;
        uint32_t n_nz = 0;
        const DeflateTable &T = B.tab;
        if (!stored) {
            S.e_bad = 0;
            for (int i = 0; i < DF_THREADS; i++) df_load_table(S, T, i, DF_THREADS);
            S.header_bits = T.header_bits;
            S.overflow = 0;
            uint32_t run = 0;
            // sparse path (k_deflate_chunks: shared code, zero-run composites fit, list fits)
            sparse = false;
            if (shared_table && !S.e_bad) {
                static uint32_t nzm[DF_THREADS][2];
                n_nz = 0;
                std::vector<uint32_t> base(DF_THREADS);
                for (int t = 0; t < DF_THREADS; t++) { base[t] = n_nz; n_nz += df_nz_masks(S.io, t, df_seg_bytes(t, clen), nzm[t]); }
                sparse = n_nz <= (uint32_t)DF_STAGE_WORDS;
                if (sparse) {
                    for (int t = 0; t < DF_THREADS; t++) df_nz_scatter(S.io, S.priv, t, base[t], nzm[t]);
                    // the Adler-32 partials of a sparse chunk come from its list: must equal the byte-wise ones
                    uint32_t la = 0, lb = 0;
                    for (int t = 0; t < DF_THREADS; t++) {
                        uint32_t a, b;
                        bits_of[t] = df_sparse_bits(S, S.priv, t, n_nz, clen, a, b);
                        la = (la + a % 65521u) % 65521u; lb = (lb + b % 65521u) % 65521u;
                    }
                    if (la != ca || lb != cb) return -7;
                }
            }
            if (!sparse) for (int t = 0; t < DF_THREADS; t++) bits_of[t] = df_encode_segment(S, t, clen);
            for (int t = 0; t < DF_THREADS; t++) { S.tbits[t] = run; run += bits_of[t]; }
            body_bits = S.header_bits + run;
            stored = S.overflow || df_dynamic_bytes(body_bits, (int)(S.tbl[256] >> 24)) >= (uint32_t)clen + 10u;
        }
        std::vector<uint8_t> piece;
        if (!stored) {
            const int hw = (int)((T.header_bits + 31) >> 5);
            for (int i = 0; i < DF_STAGE_WORDS; i++) S.io[i] = i < hw ? T.header[i] : 0;
            if (sparse) for (int t = 0; t < DF_THREADS; t++) df_sparse_emit(S, S.priv, t, n_nz, clen, S.header_bits + S.tbits[t]);
            else for (int t = 0; t < DF_THREADS; t++) df_place_segment(S, t, bits_of[t]);
            df_phase_finish(S, body_bits);
            piece.assign(reinterpret_cast<uint8_t *>(S.io), reinterpret_cast<uint8_t *>(S.io) + S.out_bytes);
        } else {
            piece.resize(clen + 10);
            piece[0] = 0; piece[1] = clen & 0xff; piece[2] = clen >> 8; piece[3] = ~clen & 0xff; piece[4] = (~clen >> 8) & 0xff;
            memcpy(piece.data() + 5, in + c * DF_CHUNK, clen);
            piece[5 + clen] = 0; piece[6 + clen] = 0; piece[7 + clen] = 0; piece[8 + clen] = 0xff; piece[9 + clen] = 0xff;
        }
        if (pos + piece.size() + 6 > cap) return -1;
        memcpy(out + pos, piece.data(), piece.size());
        pos += (uint32_t)piece.size();
        s2 = (uint32_t)(((uint64_t)s2 + (uint64_t)clen * s1 + cb) % 65521u);
        s1 = (s1 + ca) % 65521u;
    }
    out[pos++] = 0x03; out[pos++] = 0x00;
    const uint32_t ad = (s2 << 16) | s1;
    out[pos++] = ad >> 24; out[pos++] = ad >> 16; out[pos++] = ad >> 8; out[pos++] = ad;
    return pos;
}

// serial = 1: k_inflate_serial path.  serial = 0: scan / chunks / validate path; returns -100 if the validator
// would fall back to the serial path.
extern "C" long host_inflate_stream(const uint8_t *in, uint32_t n, uint8_t *out, uint32_t cap, int serial)
{
    static IfTables T;
    if (n < 8) return -1;
    if ((in[0] & 0x0f) != 8 || ((in[0] << 8) | in[1]) % 31 != 0 || (in[1] & 0x20)) return -1;
    if (serial) {
        // serial = 1: plain output; serial = 2: through the 32 KiB history ring (k_inflate_serial)
        static uint8_t ring[IF_RING_BYTES];
        IfOut O;
        if (serial == 2) O.init_ring(out, cap, ring); else O.init(out, cap);
        uint64_t end = 0;
        const int code = if_inflate(in, n, 2, O, T, false, &end);
        if (serial == 2) O.flush(O.n);
        if (code != IF_END_FINAL || end + 4 > n) return code == IF_ERR_OUT ? -2 : -1;
        const uint32_t a = (1u + O.s1 % 65521u) % 65521u;
        const uint32_t b = (uint32_t)(((uint64_t)O.n + O.s2) % 65521u);
        const uint32_t want = ((uint32_t)in[end] << 24) | (in[end + 1] << 16) | (in[end + 2] << 8) | in[end + 3];
        if (want != ((b << 16) | a)) return -3;
        return (long)O.n;
    }
    std::vector<uint32_t> cand{2};
    const uint32_t last = n - 6;
    for (uint32_t q = 2; q + 4 <= last; q++)
        if (in[q] == 0 && in[q + 1] == 0 && in[q + 2] == 0xff && in[q + 3] == 0xff) cand.push_back(q + 4);
    uint32_t pos = 2, total = 0, s1 = 1, s2 = 0;
    bool finished = false;
    for (size_t j = 0; j < cand.size(); j++) {
        const uint64_t ooff = (uint64_t)j * 16384;
        IfOut O;
        O.init(out + ooff, ooff < cap ? cap - ooff : 0);
        uint64_t end = 0;
        const int code = if_inflate(in, n, cand[j], O, T, true, &end);
        if (cand[j] != pos) return -100;
        if (code != IF_END_SYNC && code != IF_END_FINAL) return -100;
        if (O.n && total != j * 16384u) return -100;
        s2 = (uint32_t)(((uint64_t)s2 + (uint64_t)O.n * s1 + O.s2 % 65521u) % 65521u);
        s1 = (s1 + O.s1 % 65521u) % 65521u;
        total += (uint32_t)O.n;
        pos = (uint32_t)end;
        if (code == IF_END_FINAL) { finished = j + 1 == cand.size(); break; }
    }
    if (!finished || pos + 4 > n) return -100;
    const uint32_t want = ((uint32_t)in[pos] << 24) | (in[pos + 1] << 16) | (in[pos + 2] << 8) | in[pos + 3];
    if (want != ((s2 << 16) | s1)) return -100;
    return total;
}

// Huffman builder stress: returns 0 if the produced lengths form a complete prefix code within maxbits
extern "C" int host_huffman_check(const uint32_t *freq, int n, int maxbits)
{
    std::vector<uint32_t> keys;
    for (int i = 0; i < n; i++) if (freq[i]) keys.push_back((freq[i] << 9) | (uint32_t)i);
    std::sort(keys.begin(), keys.end());
    std::vector<uint8_t> len(n);
    std::vector<uint32_t> nw(DF_NSYM);
    std::vector<uint16_t> np(2 * DF_NSYM);
    std::vector<uint8_t> nd(DF_NSYM);
    df_build_lengths(keys.data(), (int)keys.size(), maxbits, len.data(), n, nw.data(), np.data(), nd.data());
    if (keys.size() < 2) return 0;
    uint64_t kraft = 0;
    for (int i = 0; i < n; i++) {
        if ((freq[i] != 0) != (len[i] != 0)) return -1;
        if (len[i] > maxbits) return -2;
        if (len[i]) kraft += 1ull << (maxbits - len[i]);
    }
    return kraft == (1ull << maxbits) ? 0 : -3;
}

// closed-form length / distance bases of inflate_core.cuh against the tables of RFC 1951 3.2.5; returns 0 when equal
extern "C" int host_base_tables_check()
{
    static const uint16_t lbase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
    static const uint8_t lext[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    static const uint16_t dbase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
    static const uint8_t dext[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
    for (int i = 0; i < 29; i++) {
        int e = -1;
        if (if_len_base(i, e) != lbase[i] || e != lext[i]) return 1 + i;
    }
    for (int i = 0; i < 30; i++) {
        int e = -1;
        if (if_dist_base(i, e) != dbase[i] || e != dext[i]) return 100 + i;
    }
    return 0;
}
