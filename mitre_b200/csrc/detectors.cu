// detectors.cu -- detector evaluation (replaces DiscreteRulePopulation.mine_rules,
// mitre/rules.py:325-436): window sample ranges, per-window means / slopes, thresholds,
// packed truth rows.  This file holds the general, phase-by-phase path (any N):
//
//   window_ranges_kernel      (w, s) -> first sample index, sample count         rules.py:100, 405-414
//   detector_values_kernel    TMA-staged series -> values[G, N]                  rules.py:80-141, 578-600
//   thresholds_*_kernel       sort, neighbour midpoints, unique -> thr, K        rules.py:513-576
//   rows_*_kernel             v > thr / v <= thr -> packed rows                  rules.py:490-497
//
// Arithmetic contract: 'average' reproduces numpy's pairwise summation order exactly (np.mean,
// rules.py:141) -- explicit __dadd_rn / __ddiv_rn, no FMA contraction -- so means, their
// thresholds and their truth rows are bit-identical to the reference.  'slope' is the centred
// closed form of the same least-squares problem the reference hands to LAPACK gelsd
// (rules.py:110-114); truth rows agree whenever subjects' slopes are not within rounding of
// each other ("slope parity").
#include <math.h>

#include "common.cuh"
#include "scan.cuh"

namespace mitre {

// ---------------------------------------------------------------------------------------------
// numpy pairwise summation (numpy/_core/src/umath/loops_utils.h.src: PW_BLOCKSIZE 128, 8-way
// unrolled accumulators, -0.0 start for n < 8).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double pw_leaf(const double *a, int n)
{
    if (n < 8) {
        double res = -0.0;
        for (int i = 0; i < n; ++i) res = __dadd_rn(res, a[i]);
        return res;
    }
    double r0 = a[0], r1 = a[1], r2 = a[2], r3 = a[3], r4 = a[4], r5 = a[5], r6 = a[6], r7 = a[7];
    int i = 8;
    const int lim = n - (n % 8);
    for (; i < lim; i += 8) {
        r0 = __dadd_rn(r0, a[i + 0]);
        r1 = __dadd_rn(r1, a[i + 1]);
        r2 = __dadd_rn(r2, a[i + 2]);
        r3 = __dadd_rn(r3, a[i + 3]);
        r4 = __dadd_rn(r4, a[i + 4]);
        r5 = __dadd_rn(r5, a[i + 5]);
        r6 = __dadd_rn(r6, a[i + 6]);
        r7 = __dadd_rn(r7, a[i + 7]);
    }
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r0, r1), __dadd_rn(r2, r3)),
                           __dadd_rn(__dadd_rn(r4, r5), __dadd_rn(r6, r7)));
    for (; i < n; ++i) res = __dadd_rn(res, a[i]);
    return res;
}

__device__ double pairwise_sum(const double *a, int n)
{
    if (n <= 128) return pw_leaf(a, n);
    // explicit post-order walk of numpy's recursion: node(off, n) = node(off, n2) + node(off+n2, n-n2),
    // n2 = n/2 rounded down to a multiple of 8.
    int st_off[48], st_n[48];
    bool st_open[48];
    double vals[48];
    int top = 0, vtop = 0;
    st_off[0] = 0; st_n[0] = n; st_open[0] = false; top = 1;
    while (top > 0) {
        const int off = st_off[top - 1], m = st_n[top - 1];
        if (m <= 128) {
            --top;
            vals[vtop++] = pw_leaf(a + off, m);
        } else if (!st_open[top - 1]) {
            st_open[top - 1] = true;
            int n2 = m / 2;
            n2 -= n2 % 8;
            st_off[top] = off + n2; st_n[top] = m - n2; st_open[top] = false; ++top;
            st_off[top] = off; st_n[top] = n2; st_open[top] = false; ++top;
        } else {
            --top;
            const double right = vals[--vtop];
            const double left = vals[--vtop];
            vals[vtop++] = __dadd_rn(left, right);
        }
    }
    return vals[0];
}

// Closed-form OLS slope of x on (t - mid) with intercept, centred for accuracy.
__device__ __forceinline__ double ols_slope(const double *x, const double *t, int n, double mid)
{
    double st = 0.0, sx = 0.0;
    for (int i = 0; i < n; ++i) {
        st += t[i] - mid;
        sx += x[i];
    }
    const double tb = st / n, xb = sx / n;
    double sxy = 0.0, sxx = 0.0;
    for (int i = 0; i < n; ++i) {
        const double dt = (t[i] - mid) - tb;
        sxy = fma(dt, x[i] - xb, sxy);
        sxx = fma(dt, dt, sxx);
    }
    return sxy / sxx;
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
window_ranges_kernel(const double *__restrict__ times, const int32_t *__restrict__ t_offsets, int N,
                     const double *__restrict__ windows, int W, int32_t *__restrict__ lo,
                     int32_t *__restrict__ cnt)
{
    const int total = W * N;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int w = e / N, s = e % N;
        const double t0 = windows[2 * w], t1 = windows[2 * w + 1];
        const int b = t_offsets[s], eend = t_offsets[s + 1];
        int first = eend, c = 0;
        for (int j = b; j < eend; ++j) {
            const double t = times[j];
            if (t >= t0 && t <= t1) {
                if (c == 0) first = j;
                ++c;
            }
        }
        lo[e] = first;
        cnt[e] = c;
    }
}

// ---------------------------------------------------------------------------------------------
// values: CTA = (variable v, chunk of `spb` subjects).  The chunk's samples of variable v and its
// sample times are staged into shared memory by two 1-D bulk (TMA) copies; every thread then
// takes (window, subject) pairs.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
detector_values_kernel(const double *__restrict__ series, int64_t series_stride,
                       const double *__restrict__ times, const int32_t *__restrict__ t_offsets, int N,
                       int V, const double *__restrict__ windows, const int32_t *__restrict__ lo,
                       const int32_t *__restrict__ cnt, const int64_t *__restrict__ group_base,
                       const uint8_t *__restrict__ slope_ok, int W, int spb, int cap, int use_tma,
                       double *__restrict__ values)
{
    extern __shared__ __align__(16) double smem[];
    double *sx = smem;        // [cap]
    double *stime = smem + cap;  // [cap]
    __shared__ __align__(8) uint64_t bar;

    const int chunks = (N + spb - 1) / spb;
    const int v = blockIdx.x / chunks;
    const int c = blockIdx.x % chunks;
    const int s0 = c * spb;
    const int s1 = min(N, s0 + spb);
    const int g0 = t_offsets[s0], g1 = t_offsets[s1];
    const int a0 = g0 & ~1;                 // 16-byte aligned start (elements)
    const int a1 = (g1 + 1) & ~1;
    const int n_el = a1 - a0;
    const double *xrow = series + (size_t)v * series_stride;

    if (use_tma && n_el > 0) {
        if (threadIdx.x == 0) {
            mbar_init(&bar, 1);
            mbar_fence_init();
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            mbar_arrive_expect_tx(&bar, 2u * n_el * sizeof(double));
            bulk_g2s(sx, xrow + a0, n_el * sizeof(double), &bar);
            bulk_g2s(stime, times + a0, n_el * sizeof(double), &bar);
        }
        mbar_wait(&bar, 0);
    } else {
        for (int j = threadIdx.x; j < g1 - a0; j += blockDim.x) {
            sx[j] = (a0 + j >= g0) ? xrow[a0 + j] : 0.0;
            stime[j] = (a0 + j >= g0) ? times[a0 + j] : 0.0;
        }
        __syncthreads();
    }

    const int ns = s1 - s0;
    for (int item = threadIdx.x; item < W * ns; item += blockDim.x) {
        const int w = item / ns, s = s0 + item % ns;
        const int first = lo[w * N + s] - a0;
        const int n = cnt[w * N + s];
        const int has_slope = slope_ok[w];
        const int64_t g = group_base[w] + (int64_t)v * (1 + has_slope);
        const double mean = __ddiv_rn(pairwise_sum(sx + first, n), (double)n);
        if (has_slope) {
            const double mid = (windows[2 * w] + windows[2 * w + 1]) * 0.5;   // rules.py:43
            values[g * N + s] = ols_slope(sx + first, stime + first, n, mid);
            values[(g + 1) * N + s] = mean;
        } else {
            values[g * N + s] = mean;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// thresholds, N <= 64: one warp per group.  Lane l holds sorted elements l and l+32 (bitonic
// network over order-preserving uint64 keys, +inf padding), forms neighbour midpoints, keeps
// the first of every run of equal midpoints (np.unique on an already sorted vector).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t shfl_xor_u64(uint64_t v, int m)
{
    return (uint64_t)__shfl_xor_sync(0xffffffffu, (unsigned long long)v, m);
}
__device__ __forceinline__ double key_f64(uint64_t k)
{
    const uint64_t b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

// Sort 64 keys held as (a = element lane, b = element lane+32) ascending across the warp.
__device__ __forceinline__ void warp_sort64(uint64_t &a, uint64_t &b, int lane)
{
#pragma unroll
    for (int k = 2; k <= 64; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j == 32) {
                // partner of element l is l+32 (same lane); k == 64 here: ascending
                const uint64_t mn = a < b ? a : b, mx = a < b ? b : a;
                a = mn; b = mx;
            } else {
                const uint64_t pa = shfl_xor_u64(a, j), pb = shfl_xor_u64(b, j);
                const bool lower = (lane & j) == 0;
                const bool up_a = (k == 64) ? true : ((lane & k) == 0);
                const bool up_b = (k == 64) ? true : (k == 32 ? false : ((lane & k) == 0));
                a = (lower == up_a) ? (a < pa ? a : pa) : (a < pa ? pa : a);
                b = (lower == up_b) ? (b < pb ? b : pb) : (b < pb ? pb : b);
            }
        }
    }
}

__global__ void __launch_bounds__(256)
thresholds_warp_kernel(const double *__restrict__ values, int64_t G, int N, double *__restrict__ thresholds,
                       int32_t *__restrict__ K)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const uint64_t kInf = ~0ull;
    for (int64_t g = warp0; g < G; g += nwarps) {
        const double *vg = values + g * N;
        uint64_t a = lane < N ? f64_key(vg[lane]) : kInf;
        uint64_t b = lane + 32 < N ? f64_key(vg[lane + 32]) : kInf;
        warp_sort64(a, b, lane);
        const double va = key_f64(a), vb = key_f64(b);
        // next sorted element
        double na = __shfl_down_sync(0xffffffffu, va, 1);
        const double b0 = __shfl_sync(0xffffffffu, vb, 0);
        if (lane == 31) na = b0;
        const double nb = __shfl_down_sync(0xffffffffu, vb, 1);
        const bool valid_a = lane < N - 1, valid_b = lane + 32 < N - 1;
        const double ma = __dmul_rn(0.5, __dadd_rn(va, na));
        const double mb = __dmul_rn(0.5, __dadd_rn(vb, nb));
        // previous midpoint
        double pa = __shfl_up_sync(0xffffffffu, ma, 1);
        double pb = __shfl_up_sync(0xffffffffu, mb, 1);
        const double a31 = __shfl_sync(0xffffffffu, ma, 31);
        if (lane == 0) pb = a31;
        const bool ua = valid_a && (lane == 0 || ma != pa);
        const bool ub = valid_b && (mb != pb);
        const unsigned balA = __ballot_sync(0xffffffffu, ua), balB = __ballot_sync(0xffffffffu, ub);
        const unsigned lt = (1u << lane) - 1u;
        double *tg = thresholds + g * (N - 1);
        if (ua) tg[__popc(balA & lt)] = ma;
        if (ub) tg[__popc(balA) + __popc(balB & lt)] = mb;
        if (lane == 0) K[g] = __popc(balA) + __popc(balB);
    }
}

// Slope guard band as a separate pass (fused_detectors.cuh explains the criterion): one warp per slope
// group for N <= 64, one CTA per slope group beyond.
constexpr double kGuard = 1e-10;

__global__ void __launch_bounds__(256)
slope_guard_warp_kernel(const double *__restrict__ values, int64_t G, int N, const uint8_t *__restrict__ group_type,
                        uint8_t *__restrict__ group_flags, unsigned long long *__restrict__ flagged)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const uint64_t kInf = ~0ull;
    for (int64_t g = warp0; g < G; g += nwarps) {
        if (group_type[g] != 0) {
            if (lane == 0 && group_flags) group_flags[g] = 0;
            continue;
        }
        const double *vg = values + g * N;
        const double xa = lane < N ? vg[lane] : 0.0, xb = lane + 32 < N ? vg[lane + 32] : 0.0;
        bool flag = false;
        if (g + 1 < G) {       // zero slope, non-zero window mean: a constant non-zero series
            if (lane < N && xa == 0.0) flag |= values[(g + 1) * N + lane] != 0.0;
            if (lane + 32 < N && xb == 0.0) flag |= values[(g + 1) * N + lane + 32] != 0.0;
        }
        uint64_t a = lane < N ? f64_key(xa) : kInf;
        uint64_t b = lane + 32 < N ? f64_key(xb) : kInf;
        warp_sort64(a, b, lane);
        const double va = key_f64(a), vb = key_f64(b);
        double na = __shfl_down_sync(0xffffffffu, va, 1);
        const double b0 = __shfl_sync(0xffffffffu, vb, 0);
        if (lane == 31) na = b0;
        const double nb = __shfl_down_sync(0xffffffffu, vb, 1);
        // largest |value|: the sorted extremes (positions 0 and N-1)
        const double first = __shfl_sync(0xffffffffu, va, 0);
        const double last = (N - 1 < 32) ? __shfl_sync(0xffffffffu, va, (N - 1) & 31)
                                         : __shfl_sync(0xffffffffu, vb, (N - 1 - 32) & 31);
        const double scale = fmax(fabs(first), fabs(last));
        if (lane < N - 1) flag |= (va != na) && (na - va) <= kGuard * scale;
        if (lane + 32 < N - 1) flag |= (vb != nb) && (nb - vb) <= kGuard * scale;
        const bool any = __any_sync(0xffffffffu, flag);
        if (lane == 0) {
            if (group_flags) group_flags[g] = any ? 1 : 0;
            if (any && flagged) atomicAdd(flagged, 1ull);
        }
    }
}

__global__ void __launch_bounds__(256)
slope_guard_block_kernel(const double *__restrict__ values, int64_t G, int N, int npow2,
                         const uint8_t *__restrict__ group_type, uint8_t *__restrict__ group_flags,
                         unsigned long long *__restrict__ flagged)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t *keys = reinterpret_cast<uint64_t *>(smem_raw);            // [npow2]
    __shared__ int s_flag;
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int64_t g = blockIdx.x; g < G; g += gridDim.x) {
        if (group_type[g] != 0) {
            if (tid == 0 && group_flags) group_flags[g] = 0;
            continue;
        }
        const double *vg = values + g * N;
        if (tid == 0) s_flag = 0;
        bool flag = false;
        for (int i = tid; i < npow2; i += nt) {
            const double x = i < N ? vg[i] : 0.0;
            keys[i] = i < N ? f64_key(x) : ~0ull;
            if (i < N && x == 0.0 && g + 1 < G) flag |= values[(g + 1) * N + i] != 0.0;
        }
        __syncthreads();
        for (int k = 2; k <= npow2; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = tid; i < npow2; i += nt) {
                    const int ixj = i ^ j;
                    if (ixj > i) {
                        const uint64_t x = keys[i], y = keys[ixj];
                        const bool up = (i & k) == 0;
                        if ((x > y) == up) { keys[i] = y; keys[ixj] = x; }
                    }
                }
                __syncthreads();
            }
        }
        const double scale = fmax(fabs(key_f64(keys[0])), fabs(key_f64(keys[N - 1])));
        for (int i = tid; i < N - 1; i += nt) {
            const double a = key_f64(keys[i]), b = key_f64(keys[i + 1]);
            flag |= (a != b) && (b - a) <= kGuard * scale;
        }
        if (flag) atomicOr(&s_flag, 1);
        __syncthreads();
        if (tid == 0) {
            if (group_flags) group_flags[g] = s_flag ? 1 : 0;
            if (s_flag && flagged) atomicAdd(flagged, 1ull);
        }
        __syncthreads();
    }
}

// thresholds, any N: one CTA per group, bitonic sort in shared memory.
__global__ void __launch_bounds__(256)
thresholds_block_kernel(const double *__restrict__ values, int64_t G, int N, int npow2,
                        double *__restrict__ thresholds, int32_t *__restrict__ K)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t *keys = reinterpret_cast<uint64_t *>(smem_raw);            // [npow2]
    __shared__ int s_warp_tot[8];
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int64_t g = blockIdx.x; g < G; g += gridDim.x) {
        const double *vg = values + g * N;
        for (int i = tid; i < npow2; i += nt) keys[i] = i < N ? f64_key(vg[i]) : ~0ull;
        __syncthreads();
        for (int k = 2; k <= npow2; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = tid; i < npow2; i += nt) {
                    const int ixj = i ^ j;
                    if (ixj > i) {
                        const uint64_t x = keys[i], y = keys[ixj];
                        const bool up = (i & k) == 0;
                        if ((x > y) == up) { keys[i] = y; keys[ixj] = x; }
                    }
                }
                __syncthreads();
            }
        }
        // midpoints + unique flags, processed in tiles of nt with a running base
        double *tg = thresholds + g * (N - 1);
        int base = 0;
        for (int i0 = 0; i0 < N - 1; i0 += nt) {
            const int i = i0 + tid;
            bool u = false;
            double m = 0.0;
            if (i < N - 1) {
                m = __dmul_rn(0.5, __dadd_rn(key_f64(keys[i]), key_f64(keys[i + 1])));
                u = (i == 0) || (m != __dmul_rn(0.5, __dadd_rn(key_f64(keys[i - 1]), key_f64(keys[i]))));
            }
            const unsigned bal = __ballot_sync(0xffffffffu, u);
            const int lane = tid & 31, wid = tid >> 5;
            if (lane == 0) s_warp_tot[wid] = __popc(bal);
            __syncthreads();
            int pos = base + __popc(bal & ((1u << lane) - 1u));
            int tot = 0;
            for (int w = 0; w < (nt >> 5); ++w) {
                if (w < wid) pos += s_warp_tot[w];
                tot += s_warp_tot[w];
            }
            if (u) tg[pos] = m;
            base += tot;
            __syncthreads();
        }
        if (tid == 0) K[g] = base;
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// rows, N <= 64: one warp per group; lane k owns thresholds k and k+32, walks the subjects
// (values broadcast by shuffle) and writes the (above, below) row pair as one 16-byte store.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
rows_warp_kernel(const double *__restrict__ values, const double *__restrict__ thresholds,
                 const int32_t *__restrict__ K, const int64_t *__restrict__ row_offsets, int64_t G, int N,
                 uint64_t *__restrict__ packed)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const uint64_t valid = (N == 64) ? ~0ull : (~0ull << (64 - N));
    for (int64_t g = warp0; g < G; g += nwarps) {
        const double *vg = values + g * N;
        const double va = lane < N ? vg[lane] : 0.0;
        const double vb = lane + 32 < N ? vg[lane + 32] : 0.0;
        const int k_g = K[g];
        const double *tg = thresholds + g * (N - 1);
        const double th0 = lane < k_g ? tg[lane] : 0.0;
        const double th1 = lane + 32 < k_g ? tg[lane + 32] : 0.0;
        uint64_t m0 = 0, m1 = 0;
        for (int s = 0; s < N; ++s) {
            const double v = __shfl_sync(0xffffffffu, s < 32 ? va : vb, s & 31);
            const uint64_t bit = 1ull << (63 - s);
            if (v > th0) m0 |= bit;
            if (v > th1) m1 |= bit;
        }
        ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(packed + row_offsets[g]);
        if (lane < k_g) dst[lane] = make_ulonglong2(m0, ~m0 & valid);
        if (lane + 32 < k_g) dst[lane + 32] = make_ulonglong2(m1, ~m1 & valid);
    }
}

// rows, any N: one CTA per group.
//   1. rank_i = #{k : theta_k < v_i} by binary search (v_i > theta_k  <=>  k < rank_i; the
//      thresholds are ascending and unique), one subject per thread;
//   2. the ranks are bit-sliced with warp ballots: plane b, word w holds bit b of the ranks of
//      subjects 64w .. 64w+63 in the packed-row bit order;
//   3. thread = (threshold k, word w): `rank > k` for the word's 64 subjects at once, as a
//      bit-serial comparison over the B = ceil(log2(K+1)) planes (~6 logic ops per plane instead
//      of 64 fp64 compares per word), then one coalesced store of the (above, below) words.
__global__ void __launch_bounds__(256)
rows_block_kernel(const double *__restrict__ values, const double *__restrict__ thresholds,
                  const int32_t *__restrict__ K, const int64_t *__restrict__ row_offsets, int64_t G, int N,
                  int W64, uint64_t *__restrict__ packed)
{
    extern __shared__ __align__(16) uint64_t planes[];   // [B <= 32][W64]
    uint32_t *planes32 = reinterpret_cast<uint32_t *>(planes);
    const int lane = threadIdx.x & 31;
    const int n_chunks = (N + 31) >> 5;
    for (int64_t g = blockIdx.x; g < G; g += gridDim.x) {
        const int k_g = K[g];
        const double *tg = thresholds + g * (N - 1);
        int B = 0;
        while ((1 << B) <= k_g) ++B;                     // ranks are in [0, k_g]
        // -- 1 + 2: ranks and their bit planes; a warp takes 32 consecutive subjects at a time --
        for (int c = threadIdx.x >> 5; c < n_chunks; c += blockDim.x >> 5) {
            const int i = c * 32 + lane;
            int rank = 0;
            if (i < N) {
                const double v = values[g * N + i];
                int lo = 0, hi = k_g;                    // first index with theta >= v (NaN: 0)
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (tg[mid] < v) lo = mid + 1; else hi = mid;
                }
                rank = lo;
            }
            for (int b = 0; b < B; ++b) {
                const uint32_t bal = __ballot_sync(0xffffffffu, (rank >> b) & 1);
                // subject 64w + j sits at bit 63 - j: the first 32 subjects of a word are its high half
                if (lane == 0) planes32[(size_t)b * 2 * W64 + (c ^ 1)] = __brev(bal);
            }
        }
        if ((n_chunks & 1) && threadIdx.x < B) planes32[(size_t)threadIdx.x * 2 * W64 + (n_chunks ^ 1)] = 0;
        __syncthreads();
        // -- 3: emit --
        uint64_t *dst = packed + row_offsets[g] * W64;
        const int last_bits = N - 64 * (W64 - 1);
        for (int item = threadIdx.x; item < k_g * W64; item += blockDim.x) {
            const int k = item / W64, w = item - k * W64;
            uint64_t gt = 0, eq = ~0ull;
            for (int b = B - 1; b >= 0; --b) {
                const uint64_t p = planes[(size_t)b * W64 + w];
                const uint64_t kb = ((k >> b) & 1) ? ~0ull : 0ull;
                gt |= eq & p & ~kb;
                eq &= ~(p ^ kb);
            }
            const uint64_t valid = (w == W64 - 1 && last_bits < 64) ? (~0ull << (64 - last_bits)) : ~0ull;
            dst[(size_t)(2 * k) * W64 + w] = gt;
            dst[(size_t)(2 * k + 1) * W64 + w] = ~gt & valid;
        }
        __syncthreads();
    }
}

}  // namespace mitre

using namespace mitre;

static unsigned capped_grid(int64_t blocks, int per_sm)
{
    const int64_t cap = (int64_t)sm_count() * per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (unsigned)blocks;
}

extern "C" int mitre_window_ranges(void *stream, const double *times, const int32_t *t_offsets, int N,
                                   const double *windows, int W, int32_t *lo, int32_t *cnt)
{
    if (N <= 0 || W < 0 || !times || !t_offsets || (W > 0 && (!windows || !lo || !cnt))) return MITRE_ERR_INVALID;
    if (W == 0) return MITRE_OK;
    window_ranges_kernel<<<capped_grid(ceil_div64((int64_t)W * N, 256), 8), 256, 0, as_stream(stream)>>>(
        times, t_offsets, N, windows, W, lo, cnt);
    MITRE_LAUNCH_CHECK();
    return MITRE_OK;
}

// `max_T`: the largest number of samples any subject has (host knows the offsets it built).
extern "C" int mitre_detector_values(void *stream, const double *series, int64_t series_stride,
                                        const double *times, const int32_t *t_offsets, int N, int V,
                                        const double *windows, const int32_t *lo, const int32_t *cnt,
                                        const int64_t *group_base, const uint8_t *slope_ok, int W,
                                        int max_T, double *values)
{
    if (N <= 0 || V <= 0 || W < 0 || max_T <= 0) return MITRE_ERR_INVALID;
    if (!series || !times || !t_offsets || !values) return MITRE_ERR_INVALID;
    if (W == 0) return MITRE_OK;
    if (!windows || !lo || !cnt || !group_base || !slope_ok) return MITRE_ERR_INVALID;
    // subjects per CTA: as many as fit ~24 KB of (x, t) staging, so several CTAs share an SM
    const int budget = 1536;  // samples per array
    int spb = budget / max_T;
    if (spb < 1) spb = 1;
    if (spb > N) spb = N;
    const int cap = (spb * max_T + 2 + 1) & ~1;   // even => both staging arrays 16-byte aligned
    const size_t smem = 2 * (size_t)cap * sizeof(double);
    if (smem > smem_optin()) return MITRE_ERR_UNSUPPORTED;
    auto kern = detector_values_kernel;
    MITRE_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int chunks = (N + spb - 1) / spb;
    const int64_t blocks = (int64_t)V * chunks;
    if (blocks > 0x7fffffffll) return MITRE_ERR_UNSUPPORTED;
    const int use_tma = ((reinterpret_cast<uintptr_t>(series) | reinterpret_cast<uintptr_t>(times)) % 16 == 0) &&
                        (series_stride % 2 == 0);
    kern<<<(unsigned)blocks, 256, smem, as_stream(stream)>>>(series, series_stride, times, t_offsets, N, V,
                                                             windows, lo, cnt, group_base, slope_ok, W, spb,
                                                             cap, use_tma, values);
    MITRE_LAUNCH_CHECK();
    return MITRE_OK;
}

namespace mitre {
__global__ void __launch_bounds__(256)
count_flags_kernel(const uint8_t *__restrict__ flags, int64_t G, unsigned long long *__restrict__ out)
{
    unsigned long long c = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < G; i += stride) c += flags[i] != 0;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

int launch_count_flags(cudaStream_t st, const uint8_t *group_flags, int64_t G, unsigned long long *flagged)
{
    if (G <= 0) return MITRE_OK;
    count_flags_kernel<<<capped_grid(ceil_div64(G, 256), 4), 256, 0, st>>>(group_flags, G, flagged);
    MITRE_LAUNCH_CHECK();
    return MITRE_OK;
}

int launch_slope_guard(cudaStream_t st, const double *values, int64_t G, int N, const uint8_t *group_type,
                       uint8_t *group_flags, unsigned long long *flagged)
{
    if (G <= 0 || N < 2) return MITRE_OK;
    if (N <= 64) {
        slope_guard_warp_kernel<<<capped_grid(ceil_div64(G, 8), 8), 256, 0, st>>>(values, G, N, group_type,
                                                                                 group_flags, flagged);
    } else {
        int npow2 = 1;
        while (npow2 < N) npow2 <<= 1;
        const size_t smem = (size_t)npow2 * sizeof(uint64_t);
        if (smem > smem_optin()) return MITRE_ERR_UNSUPPORTED;
        auto kern = slope_guard_block_kernel;
        MITRE_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<capped_grid(G, 8), 256, smem, st>>>(values, G, N, npow2, group_type, group_flags, flagged);
    }
    MITRE_LAUNCH_CHECK();
    return MITRE_OK;
}
}  // namespace mitre

extern "C" int mitre_slope_guard(void *stream, const double *values, int64_t G, int N, const uint8_t *group_type,
                                 uint8_t *group_flags, int64_t *flagged)
{
    if (G < 0 || N <= 0 || (G > 0 && (!values || !group_type))) return MITRE_ERR_INVALID;
    if (flagged) MITRE_CUDA_TRY(cudaMemsetAsync(flagged, 0, sizeof(int64_t), as_stream(stream)));
    return launch_slope_guard(as_stream(stream), values, G, N, group_type, group_flags,
                              reinterpret_cast<unsigned long long *>(flagged));
}

extern "C" int mitre_detector_thresholds(void *stream, const double *values, int64_t G, int N,
                                         double *thresholds, int32_t *K)
{
    if (G < 0 || N <= 0 || (G > 0 && (!values || !K))) return MITRE_ERR_INVALID;
    if (G == 0) return MITRE_OK;
    if (N > 1 && !thresholds) return MITRE_ERR_INVALID;
    cudaStream_t st = as_stream(stream);
    if (N <= 64) {
        thresholds_warp_kernel<<<capped_grid(ceil_div64(G, 8), 8), 256, 0, st>>>(values, G, N, thresholds, K);
    } else {
        int npow2 = 1;
        while (npow2 < N) npow2 <<= 1;
        const size_t smem = (size_t)npow2 * sizeof(uint64_t);
        if (smem > smem_optin()) return MITRE_ERR_UNSUPPORTED;
        auto kern = thresholds_block_kernel;
        MITRE_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<capped_grid(G, 8), 256, smem, st>>>(values, G, N, npow2, thresholds, K);
    }
    MITRE_LAUNCH_CHECK();
    return MITRE_OK;
}

extern "C" size_t mitre_scan_workspace_bytes(int64_t n) { return scan_workspace_bytes(n); }

extern "C" int mitre_detector_row_offsets(void *stream, const int32_t *K, int64_t G, int64_t *row_offsets,
                                          void *workspace, size_t workspace_bytes)
{
    if (G < 0 || !row_offsets || (G > 0 && !K) || !workspace) return MITRE_ERR_INVALID;
    if (workspace_bytes < scan_workspace_bytes(G)) return MITRE_ERR_WORKSPACE;
    return exclusive_scan_i32(as_stream(stream), K, G, 2, row_offsets, workspace);
}

extern "C" int mitre_detector_rows(void *stream, const double *values, const double *thresholds,
                                   const int32_t *K, const int64_t *row_offsets, int64_t G, int N,
                                   uint64_t *packed_truth)
{
    if (G < 0 || N <= 0) return MITRE_ERR_INVALID;
    if (G == 0) return MITRE_OK;
    if (!values || !K || !row_offsets || !packed_truth || (N > 1 && !thresholds)) return MITRE_ERR_INVALID;
    cudaStream_t st = as_stream(stream);
    if (N <= 64) {
        rows_warp_kernel<<<capped_grid(ceil_div64(G, 8), 8), 256, 0, st>>>(values, thresholds, K, row_offsets,
                                                                          G, N, packed_truth);
    } else {
        const size_t smem = (size_t)32 * words_for(N) * sizeof(uint64_t);   // up to 32 rank bit planes
        if (smem > smem_optin()) return MITRE_ERR_UNSUPPORTED;
        auto kern = rows_block_kernel;
        MITRE_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<capped_grid(G, 8), 256, smem, st>>>(values, thresholds, K, row_offsets, G, N, words_for(N),
                                                   packed_truth);
    }
    MITRE_LAUNCH_CHECK();
    return MITRE_OK;
}
