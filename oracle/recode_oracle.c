/*
 * recode_oracle.c -- CPU restatement of pyReCoDe's per-frame reduce / unpack
 * arithmetic.  TEST INFRASTRUCTURE ONLY: nothing under pyrecode_b200/ may
 * import, link or call this file.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs use it, as the checker or the
 * timed CPU baseline, never as the product path.
 *
 * Every function cites the reference lines (paths relative to /root/reference)
 * whose behaviour it restates.  The restatement is pinned against the live
 * reference by oracle/make_golden.py (run in the build container, where the
 * reference is importable) and against the committed fixtures in tests/golden/
 * by tests/test_oracle.py.
 *
 * Parity status:
 *   L1 / L3 ......... pinned against the executing reference (writer + reader).
 *   label8 .......... pinned against scipy.ndimage.label (the reference's call).
 *   L4 centroids .... pinned against the executing get_centroids_2D_nb.
 *   L2 stats, L4 map  "parity unpinned": the reference functions do not execute
 *                     (SURVEY 0.1); restated from the cited lines' intent.
 *
 * Plain C99, no dependencies.  Build: see oracle/Makefile.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

/* ------------------------------------------------------------------------- */
/* threshold frame: thr = dark + eps evaluated in the source dtype            */
/* pyrecode/recode_writer.py:126-127,132-137 (numpy >= 2: the sum stays in    */
/* the dark array's dtype and wraps).                                         */
/* ------------------------------------------------------------------------- */
void orc_threshold_u16(const uint16_t *dark, uint64_t eps, uint16_t *thr, size_t n)
{
    for (size_t i = 0; i < n; i++) thr[i] = (uint16_t)(dark[i] + (uint16_t)eps);
}

void orc_threshold_u8(const uint8_t *dark, uint64_t eps, uint8_t *thr, size_t n)
{
    for (size_t i = 0; i < n; i++) thr[i] = (uint8_t)(dark[i] + (uint8_t)eps);
}

/* ------------------------------------------------------------------------- */
/* binarise: binary = frame > thr (strict).  recode_writer.py:437             */
/* returns the number of foreground pixels.                                   */
/* ------------------------------------------------------------------------- */
size_t orc_binarize_u16(const uint16_t *frame, const uint16_t *thr, uint8_t *binary, size_t n)
{
    size_t c = 0;
    for (size_t i = 0; i < n; i++) { binary[i] = frame[i] > thr[i]; c += binary[i]; }
    return c;
}

/* ------------------------------------------------------------------------- */
/* L1 gather: vals = frame[binary] - thr[binary], raster order.               */
/* recode_writer.py:440                                                       */
/* ------------------------------------------------------------------------- */
size_t orc_l1_values_u16(const uint16_t *frame, const uint16_t *thr, const uint8_t *binary,
                         size_t n, uint16_t *vals)
{
    size_t c = 0;
    for (size_t i = 0; i < n; i++)
        if (binary[i]) vals[c++] = (uint16_t)(frame[i] - thr[i]);
    return c;
}

/* ------------------------------------------------------------------------- */
/* binary-map bit packing: pixel i -> byte i/8, bit i%8 (LSB first).          */
/* recode_writer.py:622-634 (_pack_binary_frame). out has ceil(n/8) bytes.    */
/* ------------------------------------------------------------------------- */
void orc_pack_map(const uint8_t *binary, size_t n, uint8_t *out)
{
    size_t nb = (n + 7) / 8;
    memset(out, 0, nb);
    for (size_t i = 0; i < n; i++)
        if (binary[i] == 1) out[i >> 3] |= (uint8_t)(1u << (i & 7));
}

/* ------------------------------------------------------------------------- */
/* variable bit depth packing: value j occupies stream bits [j*b,(j+1)*b),    */
/* LSB first, only the low b bits are kept.  recode_writer.py:637-652         */
/* (_bit_pack) == c_extensions/reader.h:105-140.  Returns ceil(n*b/8).        */
/* ------------------------------------------------------------------------- */
size_t orc_bit_pack_u16(const uint16_t *vals, size_t n, int b, uint8_t *out)
{
    size_t nb = (n * (size_t)b + 7) / 8;
    memset(out, 0, nb);
    size_t bit = 0;
    for (size_t j = 0; j < n; j++) {
        uint32_t v = vals[j];
        for (int i = 0; i < b; i++, bit++)
            if (v & (1u << i)) out[bit >> 3] |= (uint8_t)(1u << (bit & 7));
    }
    return nb;
}

/* inverse of orc_bit_pack_u16; intent of reader.h:74-99 with the loop bug    */
/* (reader.h:86) fixed.  Output widened to uint64 like the reference.         */
void orc_bit_unpack(const uint8_t *packed, size_t n, int b, uint64_t *out)
{
    size_t bit = 0;
    for (size_t j = 0; j < n; j++) {
        uint64_t v = 0;
        for (int i = 0; i < b; i++, bit++)
            if (packed[bit >> 3] & (1u << (bit & 7))) v |= 1ull << i;
        out[j] = v;
    }
}

/* ------------------------------------------------------------------------- */
/* 8-connected component labelling; labels 1..k numbered by the raster order  */
/* of each component's first pixel.  Restates scipy.ndimage.label with the    */
/* 3x3 structure as called at recode_writer.py:166,443.  Returns k.           */
/* ------------------------------------------------------------------------- */
static int32_t uf_find(int32_t *parent, int32_t x)
{
    while (parent[x] != x) { parent[x] = parent[parent[x]]; x = parent[x]; }
    return x;
}

static void uf_union(int32_t *parent, int32_t a, int32_t b)
{
    a = uf_find(parent, a); b = uf_find(parent, b);
    if (a == b) return;
    if (a < b) parent[b] = a; else parent[a] = b;
}

int32_t orc_label8(const uint8_t *binary, int ny, int nx, int32_t *labels)
{
    size_t n = (size_t)ny * (size_t)nx;
    int32_t *parent = (int32_t *)malloc(n * sizeof(int32_t));
    if (!parent) return -1;
    for (int r = 0; r < ny; r++) {
        for (int c = 0; c < nx; c++) {
            size_t p = (size_t)r * nx + c;
            if (!binary[p]) { parent[p] = -1; continue; }
            parent[p] = (int32_t)p;
            if (c > 0 && binary[p - 1]) uf_union(parent, (int32_t)p, (int32_t)(p - 1));
            if (r > 0) {
                if (c > 0 && binary[p - nx - 1]) uf_union(parent, (int32_t)p, (int32_t)(p - nx - 1));
                if (binary[p - nx]) uf_union(parent, (int32_t)p, (int32_t)(p - nx));
                if (c + 1 < nx && binary[p - nx + 1]) uf_union(parent, (int32_t)p, (int32_t)(p - nx + 1));
            }
        }
    }
    /* roots are minimum linear indices, so raster order of roots == label order */
    int32_t k = 0;
    for (size_t p = 0; p < n; p++) {
        if (parent[p] < 0) { labels[p] = 0; continue; }
        int32_t root = uf_find(parent, (int32_t)p);
        if ((size_t)root == p) labels[p] = ++k;      /* first pixel of its component */
        else labels[p] = labels[root];               /* root < p, already numbered  */
    }
    free(parent);
    return k;
}

/* ------------------------------------------------------------------------- */
/* L2 summary statistic per puddle, ascending label order.  Intent of         */
/* pyrecode/utils/converters.py:262-297 as called at recode_writer.py:446     */
/* with the RAW frame (not dark subtracted); method 0 = max, 1 = sum.  The    */
/* sum is taken modulo 2^16 (source dtype); the packer later keeps b bits.    */
/* The reference function does not execute (SURVEY B-4): parity unpinned.     */
/* ------------------------------------------------------------------------- */
void orc_l2_stats_u16(const int32_t *labels, const uint16_t *frame, size_t n, int32_t k,
                      int method, uint16_t *stats)
{
    uint32_t *acc = (uint32_t *)calloc((size_t)k + 1, sizeof(uint32_t));
    for (size_t p = 0; p < n; p++) {
        int32_t L = labels[p];
        if (!L) continue;
        uint32_t v = frame[p];
        if (method == 1) acc[L] += v;
        else if (v > acc[L]) acc[L] = v;
    }
    for (int32_t L = 1; L <= k; L++) stats[L - 1] = (uint16_t)acc[L];
    free(acc);
}

/* ------------------------------------------------------------------------- */
/* L4 centroids.  converters.py:157-197 (_get_centroids_2d_nb_w): per puddle  */
/* acc = float32[3] = [sum v*r, sum v*c, sum v]; v is the raw frame value as  */
/* float64, every += is evaluated in float64 and rounded to float32, raster   */
/* order; centroid = [acc0/acc2, acc1/acc2] in float32, ascending label.      */
/* mode 0/1 weighted (the only one reachable in the reference, SURVEY 6');    */
/* mode 2 max pixel (converters.py:229-259), mode 3 unweighted (:200-226).    */
/* out: k x 2 float32 (row, col).                                             */
/* ------------------------------------------------------------------------- */
void orc_l4_centroids_u16(const int32_t *labels, const uint16_t *frame, int ny, int nx,
                          int32_t k, int mode, float *out)
{
    float *acc = (float *)calloc(((size_t)k + 1) * 4, sizeof(float));
    uint8_t *seen = (uint8_t *)calloc((size_t)k + 1, 1);
    for (int r = 0; r < ny; r++) {
        for (int c = 0; c < nx; c++) {
            size_t p = (size_t)r * nx + c;
            int32_t L = labels[p];
            if (!L) continue;
            float *a = acc + (size_t)L * 4;
            double v = (double)frame[p];
            if (mode == 2) {
                if (!seen[L] || v > (double)a[2]) { a[0] = (float)r; a[1] = (float)c; a[2] = (float)v; }
            } else if (mode == 3) {
                a[0] = (float)((double)a[0] + (double)r);
                a[1] = (float)((double)a[1] + (double)c);
                a[2] = (float)((double)a[2] + 1.0);
            } else {
                a[0] = (float)((double)a[0] + v * (double)r);
                a[1] = (float)((double)a[1] + v * (double)c);
                a[2] = (float)((double)a[2] + v);
            }
            seen[L] = 1;
        }
    }
    for (int32_t L = 1; L <= k; L++) {
        float *a = acc + (size_t)L * 4;
        if (mode == 2) { out[2 * (L - 1)] = a[0]; out[2 * (L - 1) + 1] = a[1]; }
        else { out[2 * (L - 1)] = a[0] / a[2]; out[2 * (L - 1) + 1] = a[1] / a[2]; }
    }
    free(acc); free(seen);
}

/* intent of converters.py:300-309 (make_binary_map): zero [ny,nx] image,     */
/* img[round(row_c), round(col_c)] = 1 with round-half-to-even.  The          */
/* reference function does not execute (SURVEY B-5): parity unpinned.         */
void orc_centroid_map(const float *centroids, int32_t k, int ny, int nx, uint8_t *binary)
{
    memset(binary, 0, (size_t)ny * nx);
    for (int32_t i = 0; i < k; i++) {
        long r = lrintf(centroids[2 * i]);       /* default rounding mode: half to even */
        long c = lrintf(centroids[2 * i + 1]);
        if (r >= 0 && r < ny && c >= 0 && c < nx) binary[(size_t)r * nx + c] = 1;
    }
}

/* ------------------------------------------------------------------------- */
/* read side: (row, col, value) triples for every set bit in raster order.    */
/* c_extensions/reader.h:10-68 (_unpack_frame_sparse): level 1 value = b bits */
/* at stream bit rank*b, every other level emits value 1.  Returns n.         */
/* ------------------------------------------------------------------------- */
int64_t orc_unpack_sparse(int nx, int ny, int b, const uint8_t *map, const uint8_t *vals,
                          uint64_t *out, int level)
{
    uint64_t n = 0;
    for (int r = 0; r < ny; r++) {
        for (int c = 0; c < nx; c++) {
            size_t p = (size_t)r * nx + c;
            if (!(map[p >> 3] & (1u << (p & 7)))) continue;
            uint64_t v = 1;
            if (level == 1) {
                v = 0;
                for (int i = 0; i < b; i++) {
                    size_t bit = n * (size_t)b + i;
                    if (vals[bit >> 3] & (1u << (bit & 7))) v |= 1ull << i;
                }
            }
            out[3 * n] = (uint64_t)r; out[3 * n + 1] = (uint64_t)c; out[3 * n + 2] = v;
            n++;
        }
    }
    return (int64_t)n;
}

/* dense reconstruction: what user code gets from coo_matrix(...).todense()   */
/* (recode_reader.py:464-471, tests/minimal_read_write_test.py:91).           */
void orc_unpack_dense_u16(int nx, int ny, int b, const uint8_t *map, const uint8_t *vals,
                          uint16_t *dense, int level)
{
    size_t n = 0, P = (size_t)nx * ny;
    for (size_t p = 0; p < P; p++) {
        if (!(map[p >> 3] & (1u << (p & 7)))) { dense[p] = 0; continue; }
        uint32_t v = 1;
        if (level == 1) {
            v = 0;
            for (int i = 0; i < b; i++) {
                size_t bit = n * (size_t)b + i;
                if (vals[bit >> 3] & (1u << (bit & 7))) v |= 1u << i;
            }
        }
        dense[p] = (uint16_t)v;
        n++;
    }
}

/* ------------------------------------------------------------------------- */
/* one-call frame reduction used by the CPU baseline: returns sizes through   */
/* out_sizes = {n_fg_or_k, map_bytes, packed_bytes}.  Buffers are caller      */
/* owned; scratch must hold n*(1+4) bytes + n*2 bytes + k*8 bytes worst case  */
/* (callers allocate n*16).  Levels 1..4 as recode_writer.py:436-477.         */
/* ------------------------------------------------------------------------- */
int orc_reduce_frame_u16(const uint16_t *frame, const uint16_t *thr, int ny, int nx, int level,
                         int bit_depth, int l2_method, int l4_mode,
                         uint8_t *map_out, uint8_t *packed_out, uint8_t *scratch, uint64_t *out_sizes)
{
    size_t n = (size_t)ny * nx;
    uint8_t *binary = scratch;                         /* n bytes   */
    int32_t *labels = (int32_t *)(scratch + ((n + 7) & ~(size_t)7));        /* n * 4     */
    uint16_t *vals = (uint16_t *)((uint8_t *)labels + n * 4);               /* n * 2     */
    float *cent = (float *)((uint8_t *)vals + ((n * 2 + 7) & ~(size_t)7));  /* k * 8     */
    size_t nfg = orc_binarize_u16(frame, thr, binary, n);
    out_sizes[0] = nfg; out_sizes[1] = (n + 7) / 8; out_sizes[2] = 0;
    if (level == 1) {
        orc_l1_values_u16(frame, thr, binary, n, vals);
        orc_pack_map(binary, n, map_out);
        out_sizes[2] = orc_bit_pack_u16(vals, nfg, bit_depth, packed_out);
    } else if (level == 2) {
        int32_t k = orc_label8(binary, ny, nx, labels);
        if (k < 0) return -1;
        orc_l2_stats_u16(labels, frame, n, k, l2_method, vals);
        orc_pack_map(binary, n, map_out);
        out_sizes[0] = (uint64_t)k;
        out_sizes[2] = orc_bit_pack_u16(vals, (size_t)k, bit_depth, packed_out);
    } else if (level == 3) {
        orc_pack_map(binary, n, map_out);
    } else if (level == 4) {
        int32_t k = orc_label8(binary, ny, nx, labels);
        if (k < 0) return -1;
        orc_l4_centroids_u16(labels, frame, ny, nx, k, l4_mode, cent);
        orc_centroid_map(cent, k, ny, nx, binary);
        orc_pack_map(binary, n, map_out);
        out_sizes[0] = (uint64_t)k;
    } else return -2;
    return 0;
}
