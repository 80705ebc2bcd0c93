// calibrate.cu -- per-pixel median and standard deviation of a stack of flat-field frames.
//
// Replaces _median_std_nb (pyrecode/utils/calibration.py:48-57: numba loops calling np.median / np.std per pixel,
// results stored as float32), the dominant cost of make_calibration_frames (:87-138), which turns a dark / gain
// reference stack into the threshold frames the writer starts from.
//
// One thread per pixel, the frames of the stack are walked in order, so every load instruction of a warp reads 32
// consecutive pixels of one frame (coalesced); the stack is read twice.  HBM-bound: 2 * n_frames * itemsize bytes per
// pixel.
//   k_cal_moments   sum and sum of squares as 64-bit integers (exact) -> population standard deviation
//                   sqrt((N sum x^2 - (sum x)^2) / N^2) in float64, stored as float32; the rounded mean becomes the
//                   centre of the pixel's histogram window
//   k_cal_median    per-thread 64-bin histogram (one count wide bins, in shared memory, bin-major so that the lanes
//                   of a warp never share a bank) around that centre plus the number of values below the window:
//                   the two middle order statistics are read off the counts; median = their mean (np.median).
//                   A pixel whose median falls outside its window (hot pixels, wide distributions) is queued for
//   k_cal_bisect    one warp per queued pixel: binary search over the value range on "how many values <= v".
#include "common.cuh"
#include "kernels.cuh"

constexpr int CAL_THREADS = 256;
constexpr int CAL_BINS = 64;

// two adjacent pixels per thread: one 2 * sizeof(T)-byte load per frame (a warp request covers 64 pixels)
template <typename T> struct Pair;
template <> struct Pair<uint16_t> {
    typedef uint32_t L;
    static __device__ __forceinline__ void split(uint32_t w, uint32_t &a, uint32_t &b) { a = w & 0xffffu; b = w >> 16; }
};
template <> struct Pair<uint8_t> {
    typedef uint16_t L;
    static __device__ __forceinline__ void split(uint16_t w, uint32_t &a, uint32_t &b) { a = w & 0xffu; b = w >> 8; }
};

template <typename T>
__global__ void __launch_bounds__(CAL_THREADS)
k_cal_moments(const T *__restrict__ stack, size_t P, int N, float *__restrict__ std_out, uint32_t *__restrict__ centre)
{
    typedef typename Pair<T>::L L;
    const size_t p = ((size_t)blockIdx.x * CAL_THREADS + threadIdx.x) * 2;      // P is even on this path
    if (p >= P) return;
    unsigned long long s1[2] = {0, 0}, s2[2] = {0, 0};
    int f = 0;
    for (; f + 8 <= N; f += 8) {
        L w[8];
#pragma unroll
        for (int k = 0; k < 8; k++) w[k] = *reinterpret_cast<const L *>(stack + (size_t)(f + k) * P + p);
#pragma unroll
        for (int k = 0; k < 8; k++) {
            uint32_t a, b;
            Pair<T>::split(w[k], a, b);
            s1[0] += a; s2[0] += (unsigned long long)a * a;
            s1[1] += b; s2[1] += (unsigned long long)b * b;
        }
    }
    for (; f < N; f++) {
        uint32_t a, b;
        Pair<T>::split(*reinterpret_cast<const L *>(stack + (size_t)f * P + p), a, b);
        s1[0] += a; s2[0] += (unsigned long long)a * a;
        s1[1] += b; s2[1] += (unsigned long long)b * b;
    }
#pragma unroll
    for (int q = 0; q < 2; q++) {
        const unsigned long long num = (unsigned long long)N * s2[q] - s1[q] * s1[q];     // N^2 * variance, exact
        std_out[p + q] = (float)sqrt((double)num / ((double)N * (double)N));
        centre[p + q] = (uint32_t)((s1[q] + (unsigned long long)(N / 2)) / (unsigned long long)N);
    }
}

// one pixel per thread (odd pixel counts / unaligned stacks)
template <typename T>
__global__ void __launch_bounds__(CAL_THREADS)
k_cal_moments1(const T *__restrict__ stack, size_t P, int N, float *__restrict__ std_out, uint32_t *__restrict__ centre)
{
    const size_t p = (size_t)blockIdx.x * CAL_THREADS + threadIdx.x;
    if (p >= P) return;
    unsigned long long s1 = 0, s2 = 0;
    for (int f = 0; f < N; f++) {
        const uint32_t v = stack[(size_t)f * P + p];
        s1 += v; s2 += (unsigned long long)v * v;
    }
    const unsigned long long num = (unsigned long long)N * s2 - s1 * s1;
    std_out[p] = (float)sqrt((double)num / ((double)N * (double)N));
    centre[p] = (uint32_t)((s1 + (unsigned long long)(N / 2)) / (unsigned long long)N);
}

// VEC = 2: CAL_THREADS / 2 threads per CTA, two adjacent pixels (histogram columns t and CAL_THREADS / 2 + t) per thread
template <typename T, int VEC>
__global__ void __launch_bounds__(CAL_THREADS)
k_cal_median(const T *__restrict__ stack, size_t P, int N, const uint32_t *__restrict__ centre,
             float *__restrict__ med_out, uint32_t *__restrict__ queue, uint32_t *__restrict__ n_queued)
{
    typedef typename Pair<T>::L L;
    // VEC == 1: [CAL_BINS][CAL_THREADS] 32-bit counts.  VEC == 2: [CAL_BINS][CAL_THREADS / 2] words, the thread's two
    // pixels counted in the two 16-bit halves of its own word (n_frames < 65536): half the shared memory per pixel,
    // i.e. twice the pixels -- and loads -- in flight per SM
    extern __shared__ uint32_t s_hist[];
    constexpr int ROW = CAL_THREADS / VEC;
    const int t = threadIdx.x;
    const size_t p = ((size_t)blockIdx.x * ROW + t) * VEC;
#pragma unroll
    for (int b = 0; b < CAL_BINS; b++) s_hist[b * ROW + t] = 0;
    if (p >= P) return;                                  // (no barrier below: the histogram columns are private)
    uint32_t lo[VEC], below[VEC];
#pragma unroll
    for (int q = 0; q < VEC; q++) {
        const uint32_t c = centre[p + q];
        lo[q] = c >= CAL_BINS / 2 ? c - CAL_BINS / 2 : 0u;
        below[q] = 0;
    }
    auto count = [&](int q, uint32_t v) {
        const uint32_t d = v - lo[q];                    // wraps for v < lo
        if (v < lo[q]) below[q]++;
        else if (d < (uint32_t)CAL_BINS) s_hist[d * ROW + t] += q ? 0x10000u : 1u;
    };
    int f = 0;
    for (; f + 8 <= N; f += 8) {
        L w[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (VEC == 2) w[k] = *reinterpret_cast<const L *>(stack + (size_t)(f + k) * P + p);
            else w[k] = (L)stack[(size_t)(f + k) * P + p];
        }
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (VEC == 2) {
                uint32_t a, b;
                Pair<T>::split(w[k], a, b);
                count(0, a);
                count(VEC - 1, b);
            } else {
                count(0, (uint32_t)w[k]);
            }
        }
    }
    for (; f < N; f++) {
        if (VEC == 2) {
            uint32_t a, b;
            Pair<T>::split(*reinterpret_cast<const L *>(stack + (size_t)f * P + p), a, b);
            count(0, a);
            count(VEC - 1, b);
        } else {
            count(0, (uint32_t)stack[(size_t)f * P + p]);
        }
    }
    // order statistics k1 = (N - 1) / 2 and k2 = N / 2 (0-based): np.median = mean of the two
    const uint32_t k1 = (uint32_t)(N - 1) / 2, k2 = (uint32_t)N / 2;
#pragma unroll
    for (int q = 0; q < VEC; q++) {
        uint32_t acc = below[q], v1 = 0xffffffffu, v2 = 0xffffffffu;
        if (below[q] <= k1) {
#pragma unroll 4
            for (int b = 0; b < CAL_BINS; b++) {
                const uint32_t hw = s_hist[b * ROW + t];
                acc += VEC == 2 ? (q ? hw >> 16 : hw & 0xffffu) : hw;
                if (v1 == 0xffffffffu && acc > k1) v1 = lo[q] + b;
                if (v2 == 0xffffffffu && acc > k2) v2 = lo[q] + b;
            }
        }
        if (v2 != 0xffffffffu) med_out[p + q] = (float)(((double)v1 + (double)v2) * 0.5);
        else queue[atomicAdd(n_queued, 1u)] = (uint32_t)(p + q);
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
k_cal_bisect(const T *__restrict__ stack, size_t P, int N, const uint32_t *__restrict__ queue,
             const uint32_t *__restrict__ n_queued, float *__restrict__ med_out)
{
    const uint32_t nq = *n_queued;
    const int lane = threadIdx.x & 31;
    for (uint32_t q = blockIdx.x * 8 + (threadIdx.x >> 5); q < nq; q += gridDim.x * 8) {
        const size_t p = queue[q];
        uint32_t res[2];
#pragma unroll
        for (int which = 0; which < 2; which++) {
            const uint32_t k = which == 0 ? (uint32_t)(N - 1) / 2 : (uint32_t)N / 2;
            // smallest v with #(x <= v) >= k + 1
            uint32_t lo = 0, hi = (1u << (8 * sizeof(T))) - 1u;
            while (lo < hi) {
                const uint32_t mid = (lo + hi) >> 1;
                uint32_t cnt = 0;
                for (int f = lane; f < N; f += 32) cnt += (uint32_t)stack[(size_t)f * P + p] <= mid;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
                if (cnt >= k + 1) hi = mid; else lo = mid + 1;
            }
            res[which] = lo;
        }
        if (lane == 0) med_out[p] = (float)(((double)res[0] + (double)res[1]) * 0.5);
    }
}

// ---- per-pixel "accurate" thresholds (pyrecode/utils/calibration.py:27-45, _get_pixel_thresh_2) -------------------
// Of the values of a pixel that lie above its threshold t (the float32 median), the k + 1 largest are kept (with
// multiplicity); the result is the mean of the two smallest of those, i.e. of the (k + 1)-th and k-th largest value.
// Missing values count as the float32 minimum, exactly like the reference's initial fill, so a pixel with fewer than
// k + 1 values above t yields the same -inf / -1.7e38 the reference stores.  One thread per pixel, the k + 1 largest
// values in a small sorted register / local array (k is the expected number of events per pixel: a handful).
// as_run = 1 reproduces what the reference actually computes when run (numba 0.65): its "remove the maximum found"
// step stores the float32 minimum into a list of unsigned integers, an out-of-range conversion that leaves the list
// unchanged, so every round finds the same maximum and the result is the largest value above t, whatever k.
constexpr int CAL_TOPK_MAX = 32;

template <typename T>
__global__ void __launch_bounds__(CAL_THREADS)
k_cal_topk(const T *__restrict__ stack, size_t P, int N, const float *__restrict__ thr, int k, int as_run,
           float *__restrict__ out)
{
    const size_t p = (size_t)blockIdx.x * CAL_THREADS + threadIdx.x;
    if (p >= P) return;
    const float t = thr[p];
    const float FMIN = -3.4028234663852886e38f;          // np.finfo(np.float32).min
    float top[CAL_TOPK_MAX + 1];                         // ascending: top[0] is the smallest of the kept values
    const int m = k + 1;
    for (int i = 0; i < m; i++) top[i] = FMIN;
    for (int f = 0; f < N; f++) {
        const float v = (float)stack[(size_t)f * P + p];
        if (v > t && v > top[0]) {
            // replace the smallest kept value and restore the order
            int i = 0;
            while (i + 1 < m && top[i + 1] < v) { top[i] = top[i + 1]; i++; }
            top[i] = v;
        }
    }
    if (as_run) top[0] = top[1] = top[m - 1];            // k + 1 copies of the maximum
    out[p] = (float)(((double)(top[0] + top[1])) / 2.0);  // float32 sum (may overflow to -inf), then the halving
}

int launch_cal_topk(rc_ctx *ctx, int itemsize, const void *stack, size_t P, int N, const float *thr, int k, int as_run,
                    float *out, cudaStream_t st)
{
    if (N <= 0 || P == 0) return 0;
    if (k < 1 || k > CAL_TOPK_MAX - 1) RC_FAIL(ctx, -1, "expected_n_events must be 1..%d (got %d)", CAL_TOPK_MAX - 1, k);
    const unsigned grid = (unsigned)((P + CAL_THREADS - 1) / CAL_THREADS);
    if (itemsize == 2) k_cal_topk<uint16_t><<<grid, CAL_THREADS, 0, st>>>((const uint16_t *)stack, P, N, thr, k, as_run, out);
    else k_cal_topk<uint8_t><<<grid, CAL_THREADS, 0, st>>>((const uint8_t *)stack, P, N, thr, k, as_run, out);
    RC_LAUNCH_CHECK(ctx, "k_cal_topk");
    return 0;
}

size_t median_std_workspace_bytes(size_t P) { return round_up(P * 4, 256) * 2 + 256; }

template <typename T>
static int launch_median_std_t(rc_ctx *ctx, const void *stack, size_t P, int N, float *med, float *sd, void *ws,
                               cudaStream_t st)
{
    uint8_t *w = (uint8_t *)ws;
    uint32_t *centre = (uint32_t *)w;
    uint32_t *queue = (uint32_t *)(w + round_up(P * 4, 256));
    uint32_t *n_queued = (uint32_t *)(w + 2 * round_up(P * 4, 256));
    const bool vec = P % 2 == 0 && (uintptr_t)stack % 4 == 0 && N < 65536;
    const unsigned grid = (unsigned)((P + CAL_THREADS - 1) / CAL_THREADS);
    RC_CUDA(ctx, cudaMemsetAsync(n_queued, 0, sizeof(uint32_t), st));
    const int smem = CAL_BINS * CAL_THREADS * (int)sizeof(uint32_t) / (vec ? 2 : 1);
    if (vec) {
        k_cal_moments<T><<<(unsigned)((P / 2 + CAL_THREADS - 1) / CAL_THREADS), CAL_THREADS, 0, st>>>(
            (const T *)stack, P, N, sd, centre);
        RC_LAUNCH_CHECK(ctx, "k_cal_moments");
        RC_CUDA(ctx, cudaFuncSetAttribute(k_cal_median<T, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        k_cal_median<T, 2><<<grid, CAL_THREADS / 2, smem, st>>>((const T *)stack, P, N, centre, med, queue, n_queued);
    } else {
        k_cal_moments1<T><<<grid, CAL_THREADS, 0, st>>>((const T *)stack, P, N, sd, centre);
        RC_LAUNCH_CHECK(ctx, "k_cal_moments");
        RC_CUDA(ctx, cudaFuncSetAttribute(k_cal_median<T, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        k_cal_median<T, 1><<<grid, CAL_THREADS, smem, st>>>((const T *)stack, P, N, centre, med, queue, n_queued);
    }
    RC_LAUNCH_CHECK(ctx, "k_cal_median");
    k_cal_bisect<T><<<(unsigned)ctx->sm_count * 4, 256, 0, st>>>((const T *)stack, P, N, queue, n_queued, med);
    RC_LAUNCH_CHECK(ctx, "k_cal_bisect");
    return 0;
}

int launch_median_std(rc_ctx *ctx, int itemsize, const void *stack, size_t P, int N, float *med, float *sd, void *ws,
                      cudaStream_t st)
{
    if (N <= 0 || P == 0) return 0;
    if (itemsize == 2) return launch_median_std_t<uint16_t>(ctx, stack, P, N, med, sd, ws, st);
    return launch_median_std_t<uint8_t>(ctx, stack, P, N, med, sd, ws, st);
}
