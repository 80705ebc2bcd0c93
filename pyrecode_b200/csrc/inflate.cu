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
constexpr int INF_WARPS = 4;

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

__global__ void __launch_bounds__(INF_WARPS * 32)
k_inflate_chunks(const uint8_t *__restrict__ in, const uint64_t *__restrict__ in_off,
                 const uint32_t *__restrict__ in_bytes, int n_streams, uint32_t cmax,
                 const uint32_t *__restrict__ cand, const uint32_t *__restrict__ task_base,
                 uint32_t *__restrict__ counters, uint8_t *__restrict__ out, size_t out_stride,
                 InfTask *__restrict__ tasks)
{
    __shared__ IfTables s_tab[INF_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t total = task_base[n_streams];
    while (true) {
        uint32_t ti = 0;
        if (lane == 0) ti = atomicAdd(&counters[0], 1u);
        ti = __shfl_sync(0xffffffffu, ti, 0);
        if (ti >= total) break;
        // stream / candidate of this task (all lanes)
        int s;
        {
            int lo = 0, hi = n_streams;
            while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (task_base[mid] <= ti) lo = mid; else hi = mid; }
            s = lo;
        }
        const uint32_t j = ti - task_base[s];
        const uint64_t ooff = (uint64_t)j * INF_CHUNK;
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
                const uint64_t cap = ooff < out_stride ? out_stride - ooff : 0;
                fast = len > 0 && (len ^ 0xffffu) == nlen && m + 5 <= nb && len <= cap && p[m] == 0 && p[m + 1] == 0 &&
                       p[m + 2] == 0 && p[m + 3] == 0xff && p[m + 4] == 0xff;
            }
            if (fast) {
                const uint8_t *src = p + st0 + 5;
                uint8_t *dst = out + (size_t)s * out_stride + ooff;
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
        if (lane == 0) {
            IfOut O;
            O.out = out + (size_t)s * out_stride + ooff;
            O.cap = ooff < out_stride ? out_stride - ooff : 0;
            O.n = 0; O.s1 = 0; O.s2 = 0;
            uint64_t end = 0;
            const int code = if_inflate(in + in_off[s], in_bytes[s], cand[(size_t)s * cmax + j], O, s_tab[warp], true, &end);
            InfTask r;
            r.end = (uint32_t)end; r.out_len = (uint32_t)O.n; r.s1 = O.s1 % 65521u; r.s2 = O.s2 % 65521u; r.code = code;
            tasks[ti] = r;
#ifdef RC_DEBUG
            printf("[chunks] ti=%u s=%d j=%u start=%u code=%d end=%u out=%u\n", ti, s, j, cand[(size_t)s * cmax + j], code, r.end, r.out_len);
#endif
        }
        __syncwarp();
    }
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
    __shared__ IfTables s_tab;
    const int s = blockIdx.x;
    if (!need_serial[s] || threadIdx.x != 0) return;
    IfOut O;
    O.out = out + (size_t)s * out_stride; O.cap = out_stride; O.n = 0; O.s1 = 0; O.s2 = 0;
    uint64_t end = 0;
    const int code = if_inflate(in + in_off[s], in_bytes[s], 2, O, s_tab, false, &end);
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

    k_inflate_scan<<<n_streams, 256, 0, st>>>(in, in_off, in_bytes, cmax, cand, ncand, status);
    RC_LAUNCH_CHECK(ctx, "k_inflate_scan");
    k_scan_u32<<<1, 256, 0, st>>>(ncand, n_streams, cmax, task_base, counters);
    RC_LAUNCH_CHECK(ctx, "k_scan_u32");
    size_t blocks = ((size_t)n_streams * cmax + INF_WARPS - 1) / INF_WARPS;
    const size_t cap = (size_t)ctx->sm_count * 12;
    if (blocks > cap) blocks = cap;
    k_inflate_chunks<<<(unsigned)blocks, INF_WARPS * 32, 0, st>>>(in, in_off, in_bytes, n_streams, cmax, cand, task_base,
                                                                 counters, out, out_stride, tasks);
    RC_LAUNCH_CHECK(ctx, "k_inflate_chunks");
    k_inflate_validate<<<(n_streams + 127) / 128, 128, 0, st>>>(in, in_off, in_bytes, n_streams, cmax, cand, ncand,
                                                               task_base, tasks, out_bytes, status, need_serial);
    RC_LAUNCH_CHECK(ctx, "k_inflate_validate");
    k_inflate_serial<<<n_streams, 32, 0, st>>>(in, in_off, in_bytes, need_serial, out, out_stride, out_bytes, status);
    RC_LAUNCH_CHECK(ctx, "k_inflate_serial");
    return 0;
}
