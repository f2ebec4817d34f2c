// fused_detectors.cuh -- single-launch detector evaluation for N <= 63 subjects
// (DiscreteRulePopulation.calculate_values + thresholds_from_values + update_base_rules,
// mitre/rules.py:465-600, in one kernel).
//
// One THREAD owns one (window, variable, type) group: a warp is 32 consecutive variables of one
// window and type.  Every per-group quantity -- the N subject values, their sort, the neighbour
// midpoints, the unique thresholds and the 2K truth rows -- lives in that thread's registers, so
// the only HBM traffic is the algorithmic output (values, thresholds, K, rows) plus the series,
// which is read through L2 from a sample-major transposed copy (series_t[j][v]: all 32 lanes of a
// warp read the same sample j of consecutive variables = one coalesced 256-byte request).
//
//   sort      Batcher merge-exchange network over NB = N rounded up to 8 (+inf padding),
//             generated at compile time, fully unrolled, values + subject ids in registers
//   rows      with the values sorted, {s : v_s > theta_k} is a suffix of the sorted order:
//             S[k] = OR of subject bits at ranks > k, with the run of values EQUAL to the
//             midpoint excluded (T[k], see below) -- identical to comparing every value with
//             the threshold as the reference does (rules.py:494-497), ties and midpoint
//             rounding included
//   output    staged per warp in shared memory and written back as contiguous 8-byte-per-lane
//             runs (full sectors), rows in the PADDED layout rows[g][2(N-1)]: the first 2K[g]
//             slots of a group hold (above, below) pairs in threshold order, the rest hold the
//             sentinel ~0 so that a stable sort of the padded table leaves the valid rows first,
//             in the reference's enumeration order among equals.
//
// Compiled with -fmad=false: 'average' must reproduce numpy's pairwise summation bit for bit.
#pragma once
#include <math.h>

#include <utility>

#include "common.cuh"

namespace mitre {

extern int g_detector_split;  // see mitre_set_detector_split
extern int g_emit_blocks_per_sm;   // > 0: cap on resident sort/emit blocks per SM for the next launch (pipelined call)
extern int g_sort_emit_rolled; // 1: sort_emit_rolled_kernel instead of sort_emit_staged_kernel (mitre_set_sort_emit_rolled)

// ---- slope guard band (SURVEY hard part 1) ------------------------------------------------------------
// 'slope' values come from a closed form here and from an SVD least-squares solve in the reference
// (rules.py:110-114); they agree to ~1e-12 relative, and truth rows depend only on the ORDER and TIE
// structure of a group's N values.  A group is flagged as "rows may differ from the reference" when
//   * two neighbours of its sorted values are unequal but closer than kSlopeGuard x the group's
//     largest |value| (their order is not certain), or
//   * a subject's slope is exactly 0 while its window mean is not: a constant non-zero series, for
//     which the reference returns LAPACK rounding noise (+-1e-20) instead of 0, i.e. a different tie
//     structure against the all-zero series of other subjects.
// group_type uint8[G]: 0 = slope, 1 = average (only slope groups are examined; a slope group g is
// followed by its window's average group g + 1).  group_flags uint8[G] (optional) receives 0/1 per
// group, *flagged is incremented once per flagged group.
constexpr double kSlopeGuard = 1e-10;
struct SlopeGuardArgs {
    const uint8_t *group_type;
    uint8_t *group_flags;
    unsigned long long *flagged;
};
extern SlopeGuardArgs g_slope_guard;   // set by the extern "C" entry for the launch it is about to make

__device__ __forceinline__ bool near_tie(double a, double b, double scale)   // a <= b: sorted neighbours
{
    return a != b && (b - a) <= kSlopeGuard * scale;
}

// shared-memory stores through a 32-bit shared address that the emission loops advance themselves (the
// generic pointer + index form was re-derived from the window base on every iteration)
__device__ __forceinline__ void sts_f64(uint32_t addr, double v)
{
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ void sts_u64(uint32_t addr, uint64_t v)
{
    asm volatile("st.shared.u64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}

// ---- compile-time Batcher merge-exchange network for n inputs ------------------------------------
template <int N>
struct Network {
    int a[N * 12];
    int b[N * 12];
    int count;
};

template <int N>
constexpr Network<N> make_network()
{
    Network<N> net{};
    int c = 0;
    for (int p = 1; p < N; p *= 2)
        for (int k = p; k >= 1; k /= 2)
            for (int j = k % p; j + k < N; j += 2 * k) {
                const int lim = (k < N - j - k) ? k : (N - j - k);
                for (int i = 0; i < lim; ++i)
                    if ((i + j) / (2 * p) == (i + j + k) / (2 * p)) {
                        net.a[c] = i + j;
                        net.b[c] = i + j + k;
                        ++c;
                    }
            }
    net.count = c;
    return net;
}

template <int N>
struct NetHolder {
    static constexpr Network<N> net = make_network<N>();
};

__device__ __forceinline__ void cmp_exchange(double &x, double &y, int &ix, int &iy)
{
    const bool sw = x > y;
    const double lo = sw ? y : x, hi = sw ? x : y;
    const int ilo = sw ? iy : ix, ihi = sw ? ix : iy;
    x = lo; y = hi; ix = ilo; iy = ihi;
}

template <int N, size_t... I>
__device__ __forceinline__ void run_network(double (&v)[N], int (&id)[N], std::index_sequence<I...>)
{
    (cmp_exchange(v[NetHolder<N>::net.a[I]], v[NetHolder<N>::net.b[I]], id[NetHolder<N>::net.a[I]],
                  id[NetHolder<N>::net.b[I]]),
     ...);
}

// ---- 32-bit key variant: sort keys = order-preserving bits of the value rounded toward zero to
// float32, with the subject id in the low 6 bits.  A compare-exchange is one unsigned min and one
// max (2 instructions instead of 7, one register per element instead of three).  The key orders
// every pair whose float32 images differ above their 6 lowest mantissa bits; exact values are then
// fetched by id and the few groups with a mis-ordered near-tie are repaired by odd-even
// transposition passes on the exact (value, id) pairs.
__device__ __forceinline__ uint32_t sort_key32(double v, int s)
{
    uint32_t u = __float_as_uint(__double2float_rz(v));
    u ^= (uint32_t)((int32_t)u >> 31) | 0x80000000u;
    return (u & ~63u) | (uint32_t)s;
}

__device__ __forceinline__ void cmp_exchange_u32(uint32_t &x, uint32_t &y)
{
    const uint32_t lo = min(x, y), hi = max(x, y);
    x = lo; y = hi;
}

template <int N, size_t... I>
__device__ __forceinline__ void run_network_u32(uint32_t (&k)[N], std::index_sequence<I...>)
{
    (cmp_exchange_u32(k[NetHolder<N>::net.a[I]], k[NetHolder<N>::net.b[I]]), ...);
}

// The launchers instantiate NB = N rounded up to a multiple of 4, so NB - 4 < N <= NB: only the last three
// elements / four thresholds of the unrolled loops need a run-time comparison with N; for every other index
// the bound folds at compile time (a run-time test per index was a compare and a branch per iteration and
// kept the midpoints from being shared between iterations).
#define MITRE_LT_N(s) ((s) < NB - 3 || (s) < N)
#define MITRE_LT_M(k) ((k) < NB - 4 || (k) < M)
#define MITRE_EQ_M(k) ((k) >= NB - 4 && (k) == M)

// Sorted (value, id) pairs of one group from its values row `vin` (shared memory, N doubles).
template <int NB>
__device__ __forceinline__ void sort_group_key32(const double *vin, int N, double (&val)[NB], int (&id)[NB])
{
    uint32_t key[NB];
#pragma unroll
    for (int s = 0; s < NB; ++s) key[s] = MITRE_LT_N(s) ? sort_key32(vin[s], s) : (0xFFFFFFC0u | (uint32_t)s);
    run_network_u32<NB>(key, std::make_index_sequence<NetHolder<NB>::net.count>{});
    bool bad = false;
#pragma unroll
    for (int k = 0; k < NB; ++k) {
        id[k] = (int)(key[k] & 63u);
        val[k] = MITRE_LT_N(k) ? vin[id[k]] : INFINITY;
        if (k > 0 && MITRE_LT_N(k)) bad |= val[k - 1] > val[k];
    }
    while (bad) {   // rare: near-ties below the key's resolution
        bad = false;
#pragma unroll
        for (int par = 0; par < 2; ++par) {
#pragma unroll
            for (int k = par; k + 1 < NB; k += 2) {
                if (MITRE_LT_N(k + 1)) {
                    bad |= val[k] > val[k + 1];
                    cmp_exchange(val[k], val[k + 1], id[k], id[k + 1]);
                }
            }
        }
    }
}

// ---- numpy pairwise sum over a strided column --------------------------------------------------
struct Column {
    const double *base;   // &series_t[lo * stride + v]
    int64_t stride;
    __device__ __forceinline__ double operator()(int j) const { return __ldg(base + (int64_t)j * stride); }
};

template <class Col>
__device__ __forceinline__ double pw_leaf_col(const Col &x, int off, int n)
{
    if (n < 8) {
        double res = -0.0;
        for (int i = 0; i < n; ++i) res = __dadd_rn(res, x(off + i));
        return res;
    }
    double r0 = x(off + 0), r1 = x(off + 1), r2 = x(off + 2), r3 = x(off + 3);
    double r4 = x(off + 4), r5 = x(off + 5), r6 = x(off + 6), r7 = x(off + 7);
    int i = 8;
    const int lim = n - (n % 8);
    for (; i < lim; i += 8) {
        r0 = __dadd_rn(r0, x(off + i + 0));
        r1 = __dadd_rn(r1, x(off + i + 1));
        r2 = __dadd_rn(r2, x(off + i + 2));
        r3 = __dadd_rn(r3, x(off + i + 3));
        r4 = __dadd_rn(r4, x(off + i + 4));
        r5 = __dadd_rn(r5, x(off + i + 5));
        r6 = __dadd_rn(r6, x(off + i + 6));
        r7 = __dadd_rn(r7, x(off + i + 7));
    }
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r0, r1), __dadd_rn(r2, r3)),
                           __dadd_rn(__dadd_rn(r4, r5), __dadd_rn(r6, r7)));
    for (; i < n; ++i) res = __dadd_rn(res, x(off + i));
    return res;
}

template <class Col>
static __device__ __noinline__ double pw_tree_col(const Col &x, int n)
{
    int st_off[40], st_n[40];
    bool st_open[40];
    double vals[40];
    int top = 1, vtop = 0;
    st_off[0] = 0; st_n[0] = n; st_open[0] = false;
    while (top > 0) {
        const int off = st_off[top - 1], m = st_n[top - 1];
        if (m <= 128) {
            --top;
            vals[vtop++] = pw_leaf_col(x, off, m);
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

template <class Col>
__device__ __forceinline__ double pairwise_sum_col(const Col &x, int n)
{
    return n <= 128 ? pw_leaf_col(x, 0, n) : pw_tree_col(x, n);
}

// ---- per-window tables shared by every variable ----------------------------------------------
// One thread per window.  For window w (row stride `ld` = S_pad for the per-sample lists):
//   lo, cnt [W, N]       first global sample index / number of samples with t0 <= t <= t1
//   start   [W, N]       position of the subject's first sample in the window's flat sample list
//   inv_n   [W, N]       1/cnt (correctly rounded)      inv_sxx [W, N]  1 / sum ((t-mid) - tbar)^2
//   idx     [W, ld]      flat list of the window's global sample indices TIMES vstride (element
//                        offsets into series_t), subjects concatenated
//   dt      [W, ld]      (t - mid) - tbar for the same positions (the variable-independent half of
//                        the centred least-squares slope; same arithmetic as ols_slope in detectors.cu)
static __global__ void __launch_bounds__(128)
window_lists_kernel(const double *__restrict__ times, const int32_t *__restrict__ t_offsets, int N,
                    const double *__restrict__ windows, int W, int64_t ld, int64_t vstride,
                    int32_t *__restrict__ lo, int32_t *__restrict__ cnt, int32_t *__restrict__ start, double *__restrict__ inv_n,
                    double *__restrict__ inv_sxx, int32_t *__restrict__ idx, double *__restrict__ dt)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= W) return;
    const double t0 = windows[2 * w], t1 = windows[2 * w + 1];
    const double mid = (t0 + t1) * 0.5;
    int pos = 0;
    for (int s = 0; s < N; ++s) {
        const int b = t_offsets[s], eend = t_offsets[s + 1];
        int first = eend, c = 0;
        for (int j = b; j < eend; ++j) {
            const double t = times[j];
            if (t >= t0 && t <= t1) {
                if (c == 0) first = j;
                ++c;
            }
        }
        double st = 0.0;
        for (int j = 0; j < c; ++j) st += times[first + j] - mid;
        const double tb = c > 0 ? st / c : 0.0;
        double sx = 0.0;
        for (int j = 0; j < c; ++j) {
            const double d = (times[first + j] - mid) - tb;
            sx = fma(d, d, sx);
            idx[(int64_t)w * ld + pos + j] = (int32_t)((first + j) * vstride);
            dt[(int64_t)w * ld + pos + j] = d;
        }
        lo[w * N + s] = first;
        cnt[w * N + s] = c;
        start[w * N + s] = pos;
        inv_n[w * N + s] = c > 0 ? 1.0 / c : 0.0;
        inv_sxx[w * N + s] = 1.0 / sx;
        pos += c;
    }
}

// ---- the fused kernel -------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async8(void *dst_smem, const void *src_gmem)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// a / n for a small positive integer n with y = RN(1/n): q = RN(a y), r = a - n q (exact, one
// fma), q' = RN(q + r y) is the correctly rounded quotient (Markstein); the guard sends values
// near the ends of the exponent range to the IEEE divide.  Bit-identical to __ddiv_rn
// (tests/test_detectors_gpu.py::test_exact_division_selftest).
__device__ __forceinline__ double div_small_int(double a, double n, double y)
{
    const double aa = fabs(a);
    const double q = __dmul_rn(a, y);
    const double r = __fma_rn(-q, n, a);
    const double fast = __fma_rn(r, y, q);
    if (aa == 0.0) return a;                                   // +-0 / n
    if (aa > 1e-280 && aa < 1e280) return fast;
    return __ddiv_rn(a, n);
}

// The same quotient with the out-of-range case behind a real branch to a cold call: the if-converted form
// above makes every caller pay for the full IEEE divide sequence.
static __device__ __noinline__ double div_cold(double a, double n) { return __ddiv_rn(a, n); }
__device__ __forceinline__ double div_small_int_hot(double a, double n, double y)
{
    const double aa = fabs(a);
    if (aa == 0.0) return a;                                   // +-0 / n
    if (!(aa > 1e-280 && aa < 1e280)) return div_cold(a, n);
    const double q = __dmul_rn(a, y);
    const double r = __fma_rn(-q, n, a);
    return __fma_rn(r, y, q);
}

// numpy pairwise sum of this lane's column x[0..n) staged as x[j * 32] (n <= kFusedCap <= 128)
__device__ __forceinline__ double pw_sum_staged(const double *x, int n)
{
    if (n < 8) {
        double res = -0.0;
#pragma unroll
        for (int i = 0; i < 7; ++i)
            if (i < n) res = __dadd_rn(res, x[i * 32]);
        return res;
    }
    double r0 = x[0], r1 = x[32], r2 = x[64], r3 = x[96], r4 = x[128], r5 = x[160], r6 = x[192], r7 = x[224];
    int i = 8;
    const int lim = n & ~7;
    for (; i < lim; i += 8) {
        const double *xi = x + i * 32;
        r0 = __dadd_rn(r0, xi[0]);
        r1 = __dadd_rn(r1, xi[32]);
        r2 = __dadd_rn(r2, xi[64]);
        r3 = __dadd_rn(r3, xi[96]);
        r4 = __dadd_rn(r4, xi[128]);
        r5 = __dadd_rn(r5, xi[160]);
        r6 = __dadd_rn(r6, xi[192]);
        r7 = __dadd_rn(r7, xi[224]);
    }
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r0, r1), __dadd_rn(r2, r3)),
                           __dadd_rn(__dadd_rn(r4, r5), __dadd_rn(r6, r7)));
#pragma unroll
    for (int t = 0; t < 7; ++t)
        if (i + t < n) res = __dadd_rn(res, x[(i + t) * 32]);
    return res;
}

// launch bound 256 threads => up to 255 registers/thread: the NB values + ids stay in registers
// (a 128-register bound spills the network and measured slower)
constexpr int kFusedThreads = 256;
constexpr int kFusedCap = 40;   // samples per staging chunk ([cap][32 lanes] doubles = 10 KB per warp)

// Oversized subject (more in-window samples than a staging chunk): straight from global memory.
static __device__ __noinline__ double subject_value_global(const double *series_t, int64_t vstride, int lo_s,
                                                            int vc, int n, int type, const double *dts,
                                                            double inv_n, double inv_sxx)
{
    const Column x{series_t + (int64_t)lo_s * vstride + vc, vstride};
    if (type == 1) return __ddiv_rn(pairwise_sum_col(x, n), (double)n);
    double sx = 0.0;
    for (int j = 0; j < n; ++j) sx += x(j);
    const double xb = sx * inv_n;
    double sxy = 0.0;
    for (int j = 0; j < n; ++j) sxy = fma(__ldg(dts + j), x(j) - xb, sxy);
    return sxy * inv_sxx;
}

template <int NB>
__global__ void __launch_bounds__(kFusedThreads, 1)
fused_detectors_kernel(const double *__restrict__ series_t, int64_t vstride, int N, int V,
                       const int32_t *__restrict__ wlo, const int32_t *__restrict__ wcnt,
                       const int32_t *__restrict__ wstart, const double *__restrict__ winv_n,
                       const double *__restrict__ winv_sxx, const int32_t *__restrict__ widx,
                       const double *__restrict__ wdt, int64_t ld, const int64_t *__restrict__ group_base,
                       const uint8_t *__restrict__ slope_ok, const int32_t *__restrict__ task_window,
                       const int8_t *__restrict__ task_type, int n_wt, double *__restrict__ values,
                       double *__restrict__ thresholds, int32_t *__restrict__ K,
                       uint64_t *__restrict__ rows)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int warps_per_cta = blockDim.x >> 5;
    const int L = 2 * (N - 1);                 // row slots per group
    const int vs_stride = N | 1;               // odd 8-byte stride: conflict-free per-lane columns
    const size_t per_warp = (size_t)kFusedCap * 32 + (size_t)32 * vs_stride + 3 * NB;   // 8-byte units
    double *samples = reinterpret_cast<double *>(smem_raw) + (size_t)wid * per_warp;   // [cap][32]
    double *vals = samples + kFusedCap * 32;                                          // [32][vs_stride]
    double *s_invn = vals + 32 * vs_stride;                                           // [NB]
    double *s_invsxx = s_invn + NB;                                                   // [NB]
    int *s_cnt = reinterpret_cast<int *>(s_invsxx + NB);                              // [NB]
    int *s_start = s_cnt + NB;                                                        // [NB]

    const int n_vtiles = (V + 31) >> 5;
    const int64_t n_tasks = (int64_t)n_wt * n_vtiles;
    const int64_t warp0 = (int64_t)blockIdx.x * warps_per_cta + wid;
    const int64_t nwarps = (int64_t)gridDim.x * warps_per_cta;
    const uint64_t valid = ~0ull << (64 - N);

    for (int64_t task = warp0; task < n_tasks; task += nwarps) {
        const int wt = (int)(task / n_vtiles);
        const int vt = (int)(task % n_vtiles);
        const int w = task_window[wt];
        const int type = task_type[wt];                 // 0 slope, 1 average
        const int nt = 1 + slope_ok[w];
        const int v = vt * 32 + lane;
        const bool live = v < V;
        const int vc = live ? v : V - 1;               // dead lanes recompute a live column, store nothing
        const int64_t g = group_base[w] + (nt == 2 ? type : 0) + (int64_t)vc * nt;
        const double *col = series_t + vc;
        const int32_t *idx_w = widx + (int64_t)w * ld;
        const double *dt_w = wdt + (int64_t)w * ld;

        __syncwarp();
        for (int s = lane; s < N; s += 32) {
            s_cnt[s] = wcnt[w * N + s];
            s_start[s] = wstart[w * N + s];
            s_invn[s] = winv_n[w * N + s];
            s_invsxx[s] = winv_sxx[w * N + s];
        }
        __syncwarp();

        // ---- values (rules.py:80-141): subjects in chunks whose samples fit the staging buffer,
        //      fetched with cp.async (no register round trip, every load of a chunk in flight) ----
        for (int s = 0; s < N;) {
            int e = s, tot = 0;
            while (e < N && tot + s_cnt[e] <= kFusedCap) { tot += s_cnt[e]; ++e; }
            if (e == s) {
                vals[lane * vs_stride + s] =
                    subject_value_global(series_t, vstride, wlo[w * N + s], vc, s_cnt[s], type,
                                         dt_w + s_start[s], s_invn[s], s_invsxx[s]);
                ++s;
                continue;
            }
            const int p0 = s_start[s];
            for (int j0 = 0; j0 < tot; j0 += 32) {
                const int mine = (j0 + lane < tot) ? __ldg(idx_w + p0 + j0 + lane) : 0;   // coalesced
                const int nn = min(32, tot - j0);
                double *dst = samples + j0 * 32 + lane;
#pragma unroll 8
                for (int j = 0; j < nn; ++j)
                    cp_async8(dst + j * 32, col + __shfl_sync(0xffffffffu, mine, j));
            }
            cp_async_wait_all();
            __syncwarp();
            for (int q = s; q < e; ++q) {
                const int n = s_cnt[q];
                const int off = s_start[q] - p0;
                const double *x = samples + off * 32 + lane;
                double vq;
                if (type == 1) {
                    vq = div_small_int(pw_sum_staged(x, n), (double)n, s_invn[q]);
                } else {
                    double sx = 0.0;
                    for (int j = 0; j < n; ++j) sx += x[j * 32];
                    const double xb = sx * s_invn[q];
                    const double *dq = dt_w + p0 + off;
                    double sxy = 0.0;
                    for (int j = 0; j < n; ++j) sxy = fma(__ldg(dq + j), x[j * 32] - xb, sxy);
                    vq = sxy * s_invsxx[q];
                }
                vals[lane * vs_stride + q] = vq;
            }
            __syncwarp();
            s = e;
        }
        double val[NB];
        int id[NB];
#pragma unroll
        for (int s = 0; s < NB; ++s) {
            id[s] = s;
            val[s] = (s < N) ? vals[lane * vs_stride + s] : INFINITY;
        }
        if (live) {
            double *vg = values + g * N;
#pragma unroll
            for (int s = 0; s < NB; ++s)
                if (s < N) vg[s] = val[s];
        }

        // ---- sort (rules.py:536) ----
        run_network<NB>(val, id, std::make_index_sequence<NetHolder<NB>::net.count>{});

        // ---- thresholds (rules.py:539, 576): neighbour midpoints, first of every run ----
        int kcount = 0;
        {
            double *tg = thresholds + g * (N - 1);
            double prev = 0.0;
#pragma unroll
            for (int k = 0; k < NB - 1; ++k) {
                if (k < N - 1) {
                    const double m = __dmul_rn(0.5, __dadd_rn(val[k], val[k + 1]));
                    const bool u = (k == 0) || (m != prev);
                    if (u && live) tg[kcount] = m;
                    kcount += u;
                    prev = m;
                }
            }
            if (live) {
                for (int k2 = kcount; k2 < N - 1; ++k2) tg[k2] = 0.0;
                K[g] = kcount;
            }
        }

        // ---- rows (rules.py:490-497), walking the sorted order downwards:
        //      acc = subjects at ranks > k; Tn = subjects after the run of values equal to val[k+1].
        //      {s : v_s > m_k} = acc unless the midpoint rounds onto val[k+1], then Tn. ----
        {
            ulonglong2 *rg = reinterpret_cast<ulonglong2 *>(rows + g * L);
            uint64_t acc = 0, Tn = 0;
            int pos = kcount - 1;
#pragma unroll
            for (int k = NB - 1; k >= 0; --k) {
                if (k == N - 1) {
                    acc = 1ull << (63 - id[k]);
                    Tn = 0;
                } else if (k < N - 1) {
                    const double m = __dmul_rn(0.5, __dadd_rn(val[k], val[k + 1]));
                    bool uniq = true;
                    if (k > 0) uniq = m != __dmul_rn(0.5, __dadd_rn(val[k - 1], val[k]));
                    const uint64_t above = (m == val[k + 1]) ? Tn : acc;
                    if (uniq && live) rg[pos] = make_ulonglong2(above, ~above & valid);
                    pos -= uniq;
                    Tn = (val[k] == val[k + 1]) ? Tn : acc;
                    acc |= 1ull << (63 - id[k]);
                }
            }
            if (live)
                for (int k2 = kcount; k2 < N - 1; ++k2) rg[k2] = make_ulonglong2(~0ull, ~0ull);
        }
    }
}

// =================================================================================================
// Split path: (1) values with one thread per (window, type, subject, variable), (2) sort + emit
// with one thread per group.  The value reduction is a short, occupancy-friendly kernel (no shared
// memory, ~40 registers, coalesced reads of series_t with the 32 lanes on consecutive variables);
// the sorting network then runs in a kernel that holds nothing but the group's values and ids.
// Costs one re-read of values[G, N] (an output anyway) in exchange for much better latency hiding
// than the single fused kernel, whose 220-register threads also carried the staging loop.
// =================================================================================================
#ifdef MITRE_FUSED_MAIN
static __global__ void __launch_bounds__(256)
values_flat_kernel(const double *__restrict__ series_t, int64_t vstride, int N, int V,
                   const int32_t *__restrict__ wlo, const int32_t *__restrict__ wcnt,
                   const int32_t *__restrict__ wstart, const double *__restrict__ winv_n,
                   const double *__restrict__ winv_sxx, const double *__restrict__ wdt, int64_t ld,
                   const int64_t *__restrict__ group_base, const uint8_t *__restrict__ slope_ok,
                   const int32_t *__restrict__ task_window, const int8_t *__restrict__ task_type,
                   double *__restrict__ values)
{
    // block = (window-type pair, subject, slab of variables)
    const int wt = blockIdx.x / N, s = blockIdx.x % N;
    const int w = task_window[wt];
    const int type = task_type[wt];
    const int nt = 1 + slope_ok[w];
    const int n = wcnt[w * N + s];
    const int lo_s = wlo[w * N + s];
    const double inv_n = winv_n[w * N + s], inv_sxx = winv_sxx[w * N + s];
    const double *dts = wdt + (int64_t)w * ld + wstart[w * N + s];
    const int64_t gb = group_base[w] + (nt == 2 ? type : 0);
    const double nd = (double)n;
    for (int v = blockIdx.y * blockDim.x + threadIdx.x; v < V; v += gridDim.y * blockDim.x) {
        const Column x{series_t + (int64_t)lo_s * vstride + v, vstride};
        double val;
        if (type == 1) {
            val = div_small_int(pairwise_sum_col(x, n), nd, inv_n);
        } else {
            double sx = 0.0;
            for (int j = 0; j < n; ++j) sx += x(j);
            const double xb = sx * inv_n;
            double sxy = 0.0;
            for (int j = 0; j < n; ++j) sxy = fma(__ldg(dts + j), x(j) - xb, sxy);
            val = sxy * inv_sxx;
        }
        values[(gb + (int64_t)v * nt) * N + s] = val;
    }
}


// Tile variant of the values kernel: block = (window, tile of 32 variables), both detector types of
// the window in one pass over the samples.  A warp takes one subject at a time with its 32 lanes on
// consecutive variables (256-byte coalesced reads of series_t); results go through a shared-memory
// tile [32 * nt][N] so that the block's slice of values -- which is one contiguous range of
// 32 * nt * N doubles -- is written with full-line stores.  Windows of up to kTileCap samples keep
// the samples in registers (all loads in flight at once); longer ones walk the column.
constexpr int kTileCap = 16;

// Mean (numpy pairwise order, pw_leaf_col for n <= 16) and centred least-squares slope of NN samples
// held in registers; NN is a compile-time constant so that nothing is predicated off at run time.
template <int NN, class Col, class Dt>
__device__ __forceinline__ void window_stats(const Col &x, const Dt &dt, double inv_n, double inv_sxx,
                                             bool with_slope, double &mean, double &slope)
{
    double xr[NN];
#pragma unroll
    for (int j = 0; j < NN; ++j) xr[j] = x(j);
    double sum;
    if (NN < 8) {
        sum = -0.0;
#pragma unroll
        for (int j = 0; j < NN; ++j) sum = __dadd_rn(sum, xr[j]);
    } else {
        double r[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) r[k] = (NN == 16) ? __dadd_rn(xr[k], xr[(8 + k) % NN]) : xr[k];
        sum = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                        __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
        if (NN < 16) {
#pragma unroll
            for (int j = 8; j < NN; ++j) sum = __dadd_rn(sum, xr[j]);
        }
    }
    mean = div_small_int(sum, (double)NN, inv_n);
    if (with_slope) {
        // sequential sum from +0.0; for NN < 8 that is the mean's sum except for the sign of a zero
        double sx;
        if (NN < 8) {
            sx = __dadd_rn(sum, 0.0);
        } else {
            sx = 0.0;
#pragma unroll
            for (int j = 0; j < NN; ++j) sx = __dadd_rn(sx, xr[j]);
        }
        const double xb = __dmul_rn(sx, inv_n);
        double sxy = 0.0;
#pragma unroll
        for (int j = 0; j < NN; ++j) sxy = __fma_rn(dt(j), __dadd_rn(xr[j], -xb), sxy);
        slope = __dmul_rn(sxy, inv_sxx);
    }
}

// all sample counts: compile-time specialisations up to kTileCap, the column walk beyond
template <class Col, class Dt>
__device__ __forceinline__ void window_stats_any(int n, const Col &x, const Dt &dt, double inv_n, double inv_sxx,
                                                 bool ws, double &mean, double &slope)
{
#define MITRE_WS(NN) case NN: window_stats<NN>(x, dt, inv_n, inv_sxx, ws, mean, slope); break;
    if (n >= 1 && n <= kTileCap) {
        switch (n) {
            MITRE_WS(1) MITRE_WS(2) MITRE_WS(3) MITRE_WS(4) MITRE_WS(5) MITRE_WS(6) MITRE_WS(7) MITRE_WS(8)
            MITRE_WS(9) MITRE_WS(10) MITRE_WS(11) MITRE_WS(12) MITRE_WS(13) MITRE_WS(14) MITRE_WS(15)
            MITRE_WS(16)
        default: break;
        }
    } else {
        mean = div_small_int(pairwise_sum_col(x, n), (double)n, inv_n);
        if (ws) {
            double sx = 0.0;
            for (int j = 0; j < n; ++j) sx = __dadd_rn(sx, x(j));
            const double xb = __dmul_rn(sx, inv_n);
            double sxy = 0.0;
            for (int j = 0; j < n; ++j) sxy = __fma_rn(dt(j), __dadd_rn(x(j), -xb), sxy);
            slope = __dmul_rn(sxy, inv_sxx);
        }
    }
#undef MITRE_WS
}

struct LdgVec {
    const double *p;
    __device__ __forceinline__ double operator()(int j) const { return __ldg(p + j); }
};
struct SmemCol32 {   // column of a [rows][32] shared-memory tile
    const double *p;
    __device__ __forceinline__ double operator()(int j) const { return p[j * 32]; }
};
struct SmemVec {
    const double *p;
    __device__ __forceinline__ double operator()(int j) const { return p[j]; }
};

static __global__ void __launch_bounds__(256)
values_tile_kernel(const double *__restrict__ series_t, int64_t vstride, int N, int V, int n_vtiles,
                   const int32_t *__restrict__ wlo, const int32_t *__restrict__ wcnt,
                   const int32_t *__restrict__ wstart, const double *__restrict__ winv_n,
                   const double *__restrict__ winv_sxx, const double *__restrict__ wdt, int64_t ld,
                   const int64_t *__restrict__ group_base, const uint8_t *__restrict__ slope_ok,
                   double *__restrict__ values, uint8_t *__restrict__ group_flags)
{
    extern __shared__ double tile[];   // [32 * nt][Np], Np = N | 1
    const int w = blockIdx.x / n_vtiles, vt = blockIdx.x % n_vtiles;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    const int Np = N | 1;
    const int nt = 1 + slope_ok[w];
    const int v0 = vt * 32;
    const int v = min(v0 + lane, V - 1);   // lanes past V recompute the last variable; never stored
    for (int s = warp; s < N; s += n_warps) {
        const int n = __ldg(wcnt + w * N + s);
        const int lo_s = __ldg(wlo + w * N + s);
        const double inv_n = __ldg(winv_n + w * N + s);
        const double *px = series_t + (int64_t)lo_s * vstride + v;
        const double *dts = wdt + (int64_t)w * ld + __ldg(wstart + w * N + s);
        double mean = 0.0, slope = 0.0;
        window_stats_any(n, Column{px, vstride}, LdgVec{dts}, inv_n, __ldg(winv_sxx + w * N + s), nt == 2, mean,
                         slope);
        tile[(lane * nt + (nt - 1)) * Np + s] = mean;
        if (nt == 2) {
            tile[(lane * 2) * Np + s] = slope;
            // slope guard band: slope exactly 0 with a non-zero mean = constant non-zero series
            if (group_flags && slope == 0.0 && mean != 0.0) group_flags[group_base[w] + (int64_t)v * 2] = 1;
        }
    }
    __syncthreads();
    const int n_rows = min(32, V - v0) * nt;
    double *out = values + (group_base[w] + (int64_t)v0 * nt) * N;
    for (int r = warp; r < n_rows; r += n_warps)
        for (int s = lane; s < N; s += 32) out[(int64_t)r * N + s] = tile[r * Np + s];
}

// Resident variant: block = tile of 32 variables for ALL windows.  The tile's samples
// [S rows][32 variables] are loaded into shared memory once, so series_t is read from HBM exactly
// once; windows then run back to back out of shared memory.  Per window, the warps split the
// subjects, results go to one of two staging tiles, and after the single barrier of the iteration
// the block streams the finished tile (a contiguous range of values) to HBM.  The per-window
// tables (lo, cnt, start, 1/n, 1/sxx, centred times) are prefetched one window ahead with cp.async.
__device__ __forceinline__ void cp_async4(void *dst_smem, const void *src_gmem)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem));
}

struct alignas(16) SubjectRec {   // one subject's share of a window, as the kernel wants to read it
    int32_t lo, cnt, start, pad;
    double inv_n, inv_sxx;
};

__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src_gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem));
}

__host__ __device__ inline size_t resident_meta_bytes(int N, int64_t ld)
{
    return sizeof(SubjectRec) * (size_t)N + sizeof(double) * (size_t)ld;   // ld is even: a multiple of 16
}

// window w's tables -> one of the two shared-memory copies; dt16: the centred times may move 16 bytes at a time
__device__ __forceinline__ void resident_prefetch(char *dst, int w, int N, int S, int64_t ld, bool dt16,
                                                  const int32_t *__restrict__ wlo, const int32_t *__restrict__ wcnt,
                                                  const int32_t *__restrict__ wstart,
                                                  const double *__restrict__ winv_n,
                                                  const double *__restrict__ winv_sxx, const double *__restrict__ wdt)
{
    const int tid = threadIdx.x, nthr = blockDim.x;
    SubjectRec *rec = reinterpret_cast<SubjectRec *>(dst);
    double *dt = reinterpret_cast<double *>(rec + N);
    const int64_t o = (int64_t)w * N;
    if (tid < N) {
        cp_async4(&rec[tid].lo, wlo + o + tid);
        cp_async4(&rec[tid].cnt, wcnt + o + tid);
        cp_async4(&rec[tid].start, wstart + o + tid);
        cp_async8(&rec[tid].inv_n, winv_n + o + tid);
        cp_async8(&rec[tid].inv_sxx, winv_sxx + o + tid);
    }
    const double *src = wdt + (int64_t)w * ld;
    if (dt16) {
#pragma unroll 1
        for (int i = tid; 2 * i < S; i += nthr) cp_async16(dt + 2 * i, src + 2 * i);
    } else {
#pragma unroll 1
        for (int i = tid; i < S; i += nthr) cp_async8(dt + i, src + i);
    }
}

static __global__ void __launch_bounds__(512, 1)
values_resident_kernel(const double *__restrict__ series_t, int64_t vstride, int N, int V, int W,
                       const int32_t *__restrict__ wlo, const int32_t *__restrict__ wcnt,
                       const int32_t *__restrict__ wstart, const double *__restrict__ winv_n,
                       const double *__restrict__ winv_sxx, const double *__restrict__ wdt, int64_t ld,
                       const int64_t *__restrict__ group_base, const uint8_t *__restrict__ slope_ok,
                       int windows_per_block, double *__restrict__ values, uint8_t *__restrict__ group_flags)
{
    extern __shared__ __align__(16) double rsm[];
    __shared__ int next_subject[2];
    const int S = (int)ld;
    double *ser = rsm;                                   // [S][32]
    double *stage = ser + (size_t)S * 32;                // [2][64 * N]
    char *meta_base = reinterpret_cast<char *>(stage + 2 * 64 * N);
    const size_t meta_bytes = resident_meta_bytes(N, ld);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n_warps = blockDim.x >> 5;
    const int v0 = blockIdx.x * 32;
    const int vc = min(v0 + lane, V - 1);
    const bool dt16 = (reinterpret_cast<uintptr_t>(wdt) & 15) == 0 && (ld & 1) == 0;
    const int w_begin = blockIdx.y * windows_per_block, w_end = min(W, w_begin + windows_per_block);
    if (w_begin >= w_end) return;
    if (tid < 2) next_subject[tid] = 0;
    resident_prefetch(meta_base, w_begin, N, S, ld, dt16, wlo, wcnt, wstart, winv_n, winv_sxx, wdt);
    for (int r = warp; r < S; r += n_warps) cp_async8(ser + r * 32 + lane, series_t + (int64_t)r * vstride + vc);
    int nt_next = 1 + slope_ok[w_begin];
    int64_t gb_next = group_base[w_begin];
    cp_async_wait_all();
    __syncthreads();

    for (int w = w_begin; w < w_end; ++w) {
        const int b = (w - w_begin) & 1;
        const int nt = nt_next;
        const int64_t gb = gb_next;
        if (w + 1 < w_end) {
            resident_prefetch(meta_base + (b ^ 1) * meta_bytes, w + 1, N, S, ld, dt16, wlo, wcnt, wstart, winv_n,
                              winv_sxx, wdt);
            nt_next = 1 + slope_ok[w + 1];   // consumed one barrier from now
            gb_next = group_base[w + 1];
        }
        if (tid == 0) next_subject[b ^ 1] = 0;   // last used two barriers ago
        const SubjectRec *rec = reinterpret_cast<const SubjectRec *>(meta_base + b * meta_bytes);
        const double *dt = reinterpret_cast<const double *>(rec + N);
        double *st = stage + b * 64 * N;
        double *st_lane = st + lane * nt * N;
        const uint32_t ctr_addr = smem_u32(&next_subject[b]);
        // warps take subjects as they come free: sample counts differ a lot between subjects
        for (;;) {
            int s = 0;
            if (lane == 0)   // plain PTX: the compiler's warp-aggregated atomicAdd costs three times as much here
                asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(s) : "r"(ctr_addr) : "memory");
            s = __shfl_sync(0xffffffffu, s, 0);
            if (s >= N) break;
            const int4 ri = *reinterpret_cast<const int4 *>(&rec[s].lo);          // lo, cnt, start
            const double2 rd = *reinterpret_cast<const double2 *>(&rec[s].inv_n);  // 1/n, 1/sxx
            double mean = 0.0, slope = 0.0;
            window_stats_any(ri.y, SmemCol32{ser + ri.x * 32 + lane}, SmemVec{dt + ri.z}, rd.x, rd.y, nt == 2, mean,
                             slope);
            st_lane[(nt - 1) * N + s] = mean;
            if (nt == 2) {
                st_lane[s] = slope;
                // slope guard band: slope exactly 0 with a non-zero mean = constant non-zero series
                if (group_flags && slope == 0.0 && mean != 0.0) group_flags[gb + (int64_t)vc * 2] = 1;
            }
        }
        cp_async_wait_all();
        __syncthreads();
        // stream the finished tile out: rows [0, n_rows) x N of the staging tile are one contiguous
        // range of values
        const int n_out = min(32, V - v0) * nt * N;
        double *out = values + (gb + (int64_t)v0 * nt) * N;
        if (((reinterpret_cast<uintptr_t>(out) | (uintptr_t)(n_out & 1) * 8) & 15) == 0) {
            double2 *out2 = reinterpret_cast<double2 *>(out);
            const double2 *st2 = reinterpret_cast<const double2 *>(st);
#pragma unroll 1
            for (int i = tid; 2 * i < n_out; i += blockDim.x) out2[i] = st2[i];
        } else {
#pragma unroll 1
            for (int i = tid; i < n_out; i += blockDim.x) out[i] = st[i];
        }
    }
}

// =================================================================================================
// Walk-forward values kernel (default from round 2).
//
// Windows are all pairs (t_i, t_j) of one grid of time points (rules.py:391-402), enumerated start-major,
// so the admitted windows that share a start t_i are consecutive ("start group") and differ only in how
// far they reach: the samples of window (i, j + 1) are those of (i, j) plus the ones between t_j and
// t_{j+1}.  The kernels above re-read and re-sum every window from scratch (~200 warp-instructions per
// window x subject x 32 variables); here a thread owns one (variable, subject) pair and WALKS the
// subject's samples once per start group, keeping running sums and emitting both detector values every
// time the sample count reaches a window's count:
//   'average'  numpy's pairwise summation order is prefix-stable -- for n < 8 a sequential sum; from 8
//              on eight strided accumulators r[k] that absorb each completed block of 8, the combine
//              ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), then the tail added one element at a time
//              (pw_leaf_col above) -- so the mean of every window of the walk is bit-identical to
//              np.mean of its own masked copy (windows beyond 128 samples take pairwise_sum_col);
//   'slope'    with d_j = x_j - x_first and tau_j = t_j - t_i:  slope = (sum tau_j d_j - taubar sum d_j) / Sxx,
//              taubar and 1/Sxx per (window, subject) from window tables.  Differencing against the first
//              sample keeps a constant series at slope exactly 0.
// Thread layout: tid = v_local * N + s, so the N subjects of a variable are consecutive lanes and every
// emission is a coalesced store of values[g][0..N) -- no staging tile, no barrier after the tile load.
// The tile's series (kWalk variables x S samples, variable-major) sits in shared memory; sample counts
// differ between the lanes of a warp, so the inner while loop is divergent (each lane advances to its
// own count) while the emissions are converged.
// =================================================================================================
constexpr int kWalkMaxThreads = 640;
constexpr int kWalk2MaxThreads = 384;   // two variables per thread: ceil(VB / 2) * N <= 320 + N / 2, rounded to warps

__host__ __device__ inline int walk_vars_per_block(int N, int cap = 16)
{
    int v = kWalkMaxThreads / N;
    return v > cap ? cap : (v < 1 ? 1 : v);
}

// shared memory of values_walk_kernel, in doubles: series tile, sample times, 1/n table, then for ALL the
// windows of the block's start groups the (window, subject) tables (taubar, 1/Sxx, counts as int32) and the
// per-window group offsets / type counts -- loaded once, so that the walk itself has no barrier
__host__ __device__ inline size_t walk_smem_doubles(int N, int S, int max_block_windows, int VB_)
{
    const size_t VB = VB_;
    const size_t m = (size_t)max_block_windows * N;
    return VB * (S | 1) + S + 129 + 2 * m + (m + 1) / 2 + (size_t)max_block_windows + ((size_t)max_block_windows + 1) / 2;
}

static __global__ void __launch_bounds__(kWalkMaxThreads)
values_walk_kernel(const double *__restrict__ series, int64_t series_stride, const double *__restrict__ times,
                   const int32_t *__restrict__ t_offsets, int S, int N, int V,
                   const double *__restrict__ windows, const int32_t *__restrict__ wlo,
                   const int32_t *__restrict__ wcnt, const double *__restrict__ winv_sxx,
                   const double *__restrict__ wtbar, const int64_t *__restrict__ group_base,
                   const uint8_t *__restrict__ slope_ok, const int32_t *__restrict__ sg_first, int sg_base,
                   int n_sg, int sg_per_block, int max_block_windows, double *__restrict__ values,
                   uint8_t *__restrict__ group_flags, int VB)
{
    extern __shared__ __align__(16) double wsm[];
    const int Sp = S | 1;                               // odd row length: different variables, different banks
    const int M = max_block_windows * N;
    double *ser = wsm;                                  // [VB][Sp]
    double *st = ser + (size_t)VB * Sp;                 // [S] sample times
    double *rcp = st + S;                               // [129] 1/n
    double *m_tb = rcp + 129;                           // [block windows][N] taubar
    double *m_is = m_tb + M;                            // [block windows][N] 1/Sxx
    int32_t *m_cnt = reinterpret_cast<int32_t *>(m_is + M);                           // [block windows][N]
    long long *m_gb = reinterpret_cast<long long *>(m_cnt + ((M + 1) & ~1));          // [block windows] group_base * N
    int32_t *m_nt = reinterpret_cast<int32_t *>(m_gb + max_block_windows);            // [block windows]
    // this launch covers start groups [sg_base, n_sg), sg_per_block of them per blockIdx.y
    const int sg_begin = sg_base + blockIdx.y * sg_per_block, sg_end = min(n_sg, sg_begin + sg_per_block);
    if (sg_begin >= sg_end) return;
    const int wb0 = sg_first[sg_begin], nbw = sg_first[sg_end] - wb0;      // the block's windows
    // the (window, subject) tables depend on the start groups only: a CTA loads them once and then takes
    // variable tiles blockIdx.x, blockIdx.x + gridDim.x, ...  By default gridDim.x = the number of tiles (one
    // tile per CTA, placed dynamically by the hardware); a persistent grid of SMs / chunks CTAs saves two
    // thirds of the start-up bytes and measured SLOWER (2.00 vs 1.62 ms for the pair at S5): tile times differ
    // and a static deal of 8.4 tiles per CTA loses more than the table loads cost.
    for (int i = threadIdx.x; i < S; i += blockDim.x) st[i] = times[i];
    for (int i = threadIdx.x; i <= 128; i += blockDim.x) rcp[i] = i ? __ddiv_rn(1.0, (double)i) : 0.0;
    for (int i = threadIdx.x; i < nbw * N; i += blockDim.x) {
        m_cnt[i] = wcnt[wb0 * N + i];
        m_tb[i] = wtbar[wb0 * N + i];
        m_is[i] = winv_sxx[wb0 * N + i];
    }
    for (int i = threadIdx.x; i < nbw; i += blockDim.x) {
        m_gb[i] = group_base[wb0 + i] * N;               // element offset of the window's first group
        m_nt[i] = 1 + slope_ok[wb0 + i];
    }
    const int vl = threadIdx.x / N, s = threadIdx.x - vl * N;
    const int n_tiles = (V + VB - 1) / VB;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int v0 = tile * VB;
    const int nv = min(VB, V - v0);
    __syncthreads();                                     // every thread is done with the previous tile's samples
    for (int i = threadIdx.x; i < nv * S; i += blockDim.x) {
        const int vi = i / S, j = i - vi * S;
        ser[vi * Sp + j] = series[(int64_t)(v0 + vi) * series_stride + j];
    }
    __syncthreads();
    if (vl >= nv) continue;
    const double *x = ser + vl * Sp;
    const int64_t vNs = (int64_t)(v0 + vl) * N + s;      // (variable, subject) offset inside a window's groups
    for (int sg = sg_begin; sg < sg_end; ++sg) {
        const int w_begin = sg_first[sg], nw = sg_first[sg + 1] - w_begin;
        const int mb = (w_begin - wb0) * N + s;         // this thread's column in the block's tables
        const double t0 = windows[2 * w_begin];
        const int a = wlo[w_begin * N + s];            // first sample at or after t0: the same for the whole group
        const double x0 = x[a];
        int n = 0;
        double res = -0.0, Sd = 0.0, Std = 0.0;
        double r0 = 0, r1 = 0, r2 = 0, r3 = 0, r4 = 0, r5 = 0, r6 = 0, r7 = 0;
        for (int wi = 0; wi < nw; ++wi) {
            const int target = m_cnt[mb + wi * N];
            while (n < target) {
                const double xv = x[a + n];
                const double d = __dadd_rn(xv, -x0);
                Sd = __dadd_rn(Sd, d);
                Std = __fma_rn(__dadd_rn(st[a + n], -t0), d, Std);
                ++n;
                if (n < 8 || (n & 7)) {
                    res = __dadd_rn(res, xv);          // sequential head (n < 8) or the tail after the last block
                } else {
                    const double *b = x + a + n - 8;    // a block of 8 is complete
                    if (n == 8) {
                        r0 = b[0]; r1 = b[1]; r2 = b[2]; r3 = b[3]; r4 = b[4]; r5 = b[5]; r6 = b[6]; r7 = b[7];
                    } else {
                        r0 = __dadd_rn(r0, b[0]); r1 = __dadd_rn(r1, b[1]); r2 = __dadd_rn(r2, b[2]);
                        r3 = __dadd_rn(r3, b[3]); r4 = __dadd_rn(r4, b[4]); r5 = __dadd_rn(r5, b[5]);
                        r6 = __dadd_rn(r6, b[6]); r7 = __dadd_rn(r7, b[7]);
                    }
                    res = __dadd_rn(__dadd_rn(__dadd_rn(r0, r1), __dadd_rn(r2, r3)),
                                    __dadd_rn(__dadd_rn(r4, r5), __dadd_rn(r6, r7)));
                }
            }
            const int wl = w_begin - wb0 + wi;
            const int nt = m_nt[wl];
            // element offset of this thread's slope value (nt == 2) / mean (nt == 1): no 64-bit multiplies here
            double *out = values + (m_gb[wl] + (nt == 2 ? 2 * vNs - s : vNs));
            double mean;
            if (n <= 128) mean = div_small_int_hot(res, (double)n, rcp[n]);
            else mean = __ddiv_rn(pairwise_sum_col(SmemVec{x + a}, n), (double)n);   // numpy splits above 128
            if (nt == 2) {
                const double slope =
                    __dmul_rn(__dadd_rn(Std, -__dmul_rn(m_tb[mb + wi * N], Sd)), m_is[mb + wi * N]);
                out[0] = slope;
                out[N] = mean;
                // slope guard band: slope exactly 0 with a non-zero mean = constant non-zero series
                if (group_flags && slope == 0.0 && mean != 0.0)
                    group_flags[group_base[w_begin + wi] + (int64_t)(v0 + vl) * 2] = 1;
            } else {
                out[0] = mean;
            }
        }
    }
    }   // tiles
}

// VPT = variables per thread (1 or 2).  With two, a thread walks the SAME subject's samples for two
// neighbouring variables at once: the sample counts, the window tables, the centred times and all of the
// control flow of the walk and of the emission are shared, and the two variables' fp64 chains interleave
// (the kernel is bound by instruction issue -- under a fifth of its instructions are fp64 arithmetic -- and
// by the latency of those chains, not by memory).  Per-variable arithmetic is the same sequence of
// operations in the same order for either VPT.
template <int VPT>
static __global__ void __launch_bounds__(VPT == 1 ? kWalkMaxThreads : kWalk2MaxThreads)
values_walkn_kernel(const double *__restrict__ series, int64_t series_stride, const double *__restrict__ times,
                   const int32_t *__restrict__ t_offsets, int S, int N, int V,
                   const double *__restrict__ windows, const int32_t *__restrict__ wlo,
                   const int32_t *__restrict__ wcnt, const double *__restrict__ winv_sxx,
                   const double *__restrict__ wtbar, const int64_t *__restrict__ group_base,
                   const uint8_t *__restrict__ slope_ok, const int32_t *__restrict__ sg_first, int sg_base,
                   int n_sg, int sg_per_block, int max_block_windows, double *__restrict__ values,
                   uint8_t *__restrict__ group_flags, int VB)
{
    extern __shared__ __align__(16) double wsm[];
    const int Sp = S | 1;                               // odd row length: different variables, different banks
    const int M = max_block_windows * N;
    double *ser = wsm;                                  // [VB][Sp]
    double *st = ser + (size_t)VB * Sp;                 // [S] sample times
    double *rcp = st + S;                               // [129] 1/n
    double *m_tb = rcp + 129;                           // [block windows][N] taubar
    double *m_is = m_tb + M;                            // [block windows][N] 1/Sxx
    int32_t *m_cnt = reinterpret_cast<int32_t *>(m_is + M);                           // [block windows][N]
    long long *m_gb = reinterpret_cast<long long *>(m_cnt + ((M + 1) & ~1));          // [block windows] group_base * N
    int32_t *m_nt = reinterpret_cast<int32_t *>(m_gb + max_block_windows);            // [block windows]
    const int v0 = blockIdx.x * VB;
    const int nv = min(VB, V - v0);
    // this launch covers start groups [sg_base, n_sg), sg_per_block of them per blockIdx.y
    const int sg_begin = sg_base + blockIdx.y * sg_per_block, sg_end = min(n_sg, sg_begin + sg_per_block);
    if (sg_begin >= sg_end) return;
    const int wb0 = sg_first[sg_begin], nbw = sg_first[sg_end] - wb0;      // the block's windows
    for (int i = threadIdx.x; i < nv * S; i += blockDim.x) {
        const int vl = i / S, j = i - vl * S;
        ser[vl * Sp + j] = series[(int64_t)(v0 + vl) * series_stride + j];
    }
    for (int i = threadIdx.x; i < S; i += blockDim.x) st[i] = times[i];
    for (int i = threadIdx.x; i <= 128; i += blockDim.x) rcp[i] = i ? __ddiv_rn(1.0, (double)i) : 0.0;
    for (int i = threadIdx.x; i < nbw * N; i += blockDim.x) {
        m_cnt[i] = wcnt[wb0 * N + i];
        m_tb[i] = wtbar[wb0 * N + i];
        m_is[i] = winv_sxx[wb0 * N + i];
    }
    for (int i = threadIdx.x; i < nbw; i += blockDim.x) {
        m_gb[i] = group_base[wb0 + i] * N;               // element offset of the window's first group
        m_nt[i] = 1 + slope_ok[wb0 + i];
    }
    __syncthreads();
    const int vt = threadIdx.x / N, s = threadIdx.x - vt * N;
    const int vl = vt * VPT;                             // first of this thread's variables
    if (vl >= nv) return;
    // an odd tile leaves the last thread one variable: its second slot re-walks the first and stores nothing
    const bool has2 = VPT == 2 && vl + 1 < nv;
    const double *x[VPT];
#pragma unroll
    for (int u = 0; u < VPT; ++u) x[u] = ser + (vl + ((u && has2) ? 1 : 0)) * Sp;
    const int64_t vNs = (int64_t)(v0 + vl) * N + s;      // (variable, subject) offset inside a window's groups
    double *const out_two = values + (2 * vNs - s);      // windows with both types: slope row, then the mean row
    double *const out_one = values + vNs;                // mean only
    for (int sg = sg_begin; sg < sg_end; ++sg) {
        const int w_begin = sg_first[sg], nw = sg_first[sg + 1] - w_begin;
        const int mb = (w_begin - wb0) * N + s;         // this thread's column in the block's tables
        const double t0 = windows[2 * w_begin];
        const int a = wlo[w_begin * N + s];            // first sample at or after t0: the same for the whole group
        const double *ta = st + a;
        const double *xa[VPT];
        double x0[VPT], res[VPT], Sd[VPT], Std[VPT], r[VPT][8];
#pragma unroll
        for (int u = 0; u < VPT; ++u) {
            xa[u] = x[u] + a;
            x0[u] = xa[u][0];
            res[u] = -0.0;
            Sd[u] = 0.0;
            Std[u] = 0.0;
#pragma unroll
            for (int q = 0; q < 8; ++q) r[u][q] = 0.0;
        }
        int n = 0;
        for (int wi = 0; wi < nw; ++wi) {
            const int target = m_cnt[mb + wi * N];
            while (n < target) {
                const double tau = __dadd_rn(ta[n], -t0);
                double xv[VPT];
#pragma unroll
                for (int u = 0; u < VPT; ++u) {
                    xv[u] = xa[u][n];
                    const double d = __dadd_rn(xv[u], -x0[u]);
                    Sd[u] = __dadd_rn(Sd[u], d);
                    Std[u] = __fma_rn(tau, d, Std[u]);
                }
                ++n;
                if (n < 8 || (n & 7)) {
#pragma unroll
                    for (int u = 0; u < VPT; ++u)
                        res[u] = __dadd_rn(res[u], xv[u]);   // sequential head (n < 8) or the tail after the last block
                } else {
#pragma unroll
                    for (int u = 0; u < VPT; ++u) {
                        const double *b = xa[u] + n - 8;     // a block of 8 is complete
                        if (n == 8) {
#pragma unroll
                            for (int q = 0; q < 8; ++q) r[u][q] = b[q];
                        } else {
#pragma unroll
                            for (int q = 0; q < 8; ++q) r[u][q] = __dadd_rn(r[u][q], b[q]);
                        }
                        res[u] = __dadd_rn(__dadd_rn(__dadd_rn(r[u][0], r[u][1]), __dadd_rn(r[u][2], r[u][3])),
                                           __dadd_rn(__dadd_rn(r[u][4], r[u][5]), __dadd_rn(r[u][6], r[u][7])));
                    }
                }
            }
            const int wl = w_begin - wb0 + wi;
            const bool two = m_nt[wl] == 2;
            double *out = (two ? out_two : out_one) + m_gb[wl];
            const int vstep = two ? 2 * N : N;            // from one variable's rows to the next variable's
            double tb = 0.0, is = 0.0;
            if (two) {
                tb = m_tb[mb + wi * N];
                is = m_is[mb + wi * N];
            }
            const double nd = (double)n, y = rcp[n <= 128 ? n : 0];
#pragma unroll
            for (int u = 0; u < VPT; ++u) {
                if (u && !has2) break;
                double mean;
                if (n <= 128) mean = div_small_int_hot(res[u], nd, y);
                else mean = __ddiv_rn(pairwise_sum_col(SmemVec{xa[u]}, n), nd);   // numpy splits above 128
                if (two) {
                    const double slope = __dmul_rn(__dadd_rn(Std[u], -__dmul_rn(tb, Sd[u])), is);
                    out[0] = slope;
                    out[N] = mean;
                    // slope guard band: slope exactly 0 with a non-zero mean = constant non-zero series
                    if (group_flags && slope == 0.0 && mean != 0.0)
                        group_flags[group_base[w_begin + wi] + (int64_t)(v0 + vl + u) * 2] = 1;
                } else {
                    out[0] = mean;
                }
                out += vstep;
            }
        }
    }
}

// taubar[w][s] = mean over the window's samples of (t - t0_w): the variable-independent half of the slope
static __global__ void __launch_bounds__(128)
window_tbar_kernel(const double *__restrict__ times, int N, const double *__restrict__ windows, int W,
                   const int32_t *__restrict__ wlo, const int32_t *__restrict__ wcnt, double *__restrict__ tbar)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= W * N) return;
    const int w = i / N;
    const double t0 = windows[2 * w];
    const int lo = wlo[i], c = wcnt[i];
    double acc = 0.0;
    for (int j = 0; j < c; ++j) acc = __dadd_rn(acc, __dadd_rn(times[lo + j], -t0));
    tbar[i] = c > 0 ? __ddiv_rn(acc, (double)c) : 0.0;
}
#endif  // MITRE_FUSED_MAIN

template <int NB, int THREADS, int MINB, bool SYNC>
__global__ void __launch_bounds__(THREADS, MINB)
sort_emit_kernel(const double *__restrict__ values, int64_t G, int N, double *__restrict__ thresholds,
                 int32_t *__restrict__ K, uint64_t *__restrict__ rows)
{
    const int L = 2 * (N - 1);
    const uint64_t valid = ~0ull << (64 - N);
    const int64_t stride = (int64_t)gridDim.x * THREADS;
    for (int64_t base = (int64_t)blockIdx.x * THREADS; base < G; base += stride) {
        if (SYNC) __syncthreads();   // keeps the block's warps on the same stretch of this long code
        const int64_t g = base + threadIdx.x;
        if (g >= G) continue;
        const double *vg = values + g * N;
        double val[NB];
        int id[NB];
#pragma unroll
        for (int s = 0; s < NB; ++s) {
            id[s] = s;
            val[s] = (s < N) ? __ldg(vg + s) : INFINITY;
        }
        run_network<NB>(val, id, std::make_index_sequence<NetHolder<NB>::net.count>{});
        int kcount = 0;
        {
            double *tg = thresholds + g * (N - 1);
            double prev = 0.0;
#pragma unroll
            for (int k = 0; k < NB - 1; ++k) {
                if (k < N - 1) {
                    const double m = __dmul_rn(0.5, __dadd_rn(val[k], val[k + 1]));
                    const bool u = (k == 0) || (m != prev);
                    if (u) tg[kcount] = m;
                    kcount += u;
                    prev = m;
                }
            }
            for (int k2 = kcount; k2 < N - 1; ++k2) tg[k2] = 0.0;
            K[g] = kcount;
        }
        {
            ulonglong2 *rg = reinterpret_cast<ulonglong2 *>(rows + g * L);
            uint64_t acc = 0, Tn = 0;
            int pos = kcount - 1;
#pragma unroll
            for (int k = NB - 1; k >= 0; --k) {
                if (k == N - 1) {
                    acc = 1ull << (63 - id[k]);
                    Tn = 0;
                } else if (k < N - 1) {
                    const double m = __dmul_rn(0.5, __dadd_rn(val[k], val[k + 1]));
                    bool uniq = true;
                    if (k > 0) uniq = m != __dmul_rn(0.5, __dadd_rn(val[k - 1], val[k]));
                    const uint64_t above = (m == val[k + 1]) ? Tn : acc;
                    if (uniq) rg[pos] = make_ulonglong2(above, ~above & valid);
                    pos -= uniq;
                    Tn = (val[k] == val[k + 1]) ? Tn : acc;
                    acc |= 1ull << (63 - id[k]);
                }
            }
            for (int k2 = kcount; k2 < N - 1; ++k2) rg[k2] = make_ulonglong2(~0ull, ~0ull);
        }
    }
}

// Staged variant of sort_emit_kernel: one thread still owns one group, but nothing it reads or
// writes goes to global memory with a per-thread stride.  A block of kStageThreads groups
// bulk-loads its slice of values (one contiguous range) into shared memory through the TMA unit;
// every thread sorts its group by 32-bit keys, re-reads the exact values in sorted order and emits
// thresholds and the `above` row of every threshold into its own rows of two shared-memory tiles.
// The thresholds tile goes back as one bulk store; the rows tile is expanded to (above, below)
// pairs on the way out, 16 bytes per thread with consecutive threads on consecutive addresses.
// Per-thread strided 8/16-byte global accesses (32 sectors per warp instruction -- the L2 request
// rate that bounded the direct version) become full-line transfers, and keeping only one row of
// each pair on chip leaves room for six blocks per SM.
constexpr int kStageThreads = 64;

__host__ __device__ inline size_t sort_emit_stage_bytes(int N)
{
    return (size_t)kStageThreads * (2 * (size_t)N - 1) * 8;   // [T][N] values / above rows + [T][N-1] thresholds
}

template <int NB>
__global__ void __launch_bounds__(kStageThreads, (NB <= 40) ? 6 : 3)
sort_emit_staged_kernel(const double *__restrict__ values, int64_t G, int N, double *__restrict__ thresholds,
                        int32_t *__restrict__ K, uint64_t *__restrict__ rows, int dry, SlopeGuardArgs sg)
{
    extern __shared__ __align__(128) unsigned char stage_smem[];
    __shared__ uint64_t bar;
    constexpr int T = kStageThreads;
    const int M = N - 1;
    const int tid = threadIdx.x;
    double *sval = reinterpret_cast<double *>(stage_smem);            // [T][N] on the way in
    uint64_t *sabove = reinterpret_cast<uint64_t *>(stage_smem);      // [T][M] on the way out (same bytes)
    double *sthr = sval + (size_t)T * N;                              // [T][M]
    const uint64_t valid = ~0ull << (64 - N);
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    uint32_t phase = 0;
    const int64_t n_tiles = ceil_div64(G, T);
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t g0 = tile * T;
        const int cnt = (int)((G - g0 < T) ? (G - g0) : T);
        const bool full = cnt == T;
        // ---- values tile -> shared memory ----
        if (full) {
            if (tid == 0) {
                mbar_arrive_expect_tx(&bar, (uint32_t)(T * N * 8));
                bulk_g2s(sval, values + g0 * N, (uint32_t)(T * N * 8), &bar);
            }
            mbar_wait(&bar, phase);
            phase ^= 1;
        } else {
            for (int i = tid; i < cnt * N; i += T) sval[i] = values[g0 * N + i];
            __syncthreads();
        }
        double val[NB];
        int id[NB];
        // threads past the end of the last tile sort whatever the buffer holds; nothing of it is stored
        sort_group_key32<NB>(sval + tid * N, N, val, id);
        if (sg.group_type && tid < cnt && sg.group_type[g0 + tid] == 0) {
            // slope guard band, near-tie half: the sorted values are in registers (compile-time indices only:
            // a run-time index would push val[] into local memory).  The constant-series half of the
            // criterion is checked by the values kernel; group_flags was cleared before the launches.
            double last = val[0];
#pragma unroll
            for (int k = (NB > 4 ? NB - 4 : 1); k < NB; ++k)
                if (k == M) last = val[k];
            const double tol = kSlopeGuard * fmax(fabs(val[0]), fabs(last));
            bool flag = false;
            // past M the neighbours are (value, +inf) or (+inf, +inf): neither is a near-tie
#pragma unroll
            for (int k = 0; k < NB - 1; ++k) flag |= (val[k] != val[k + 1]) && (val[k + 1] - val[k] <= tol);
            if (flag) sg.group_flags[g0 + tid] = 1;
        }
        if (tid == 0) bulk_wait_read0();   // the previous tile's thresholds have left their buffer
        __syncthreads();                   // every thread holds its values: their tile becomes the rows tile
        if (!dry) {
        int kcount = 0;
        {
            uint32_t ta = smem_u32(sthr + tid * M);   // slot of the threshold being written (ascending)
            double prev = 0.0;
#pragma unroll
            for (int k = 0; k < NB - 1; ++k) {
                if (MITRE_LT_M(k)) {
                    const double m = __dmul_rn(0.5, __dadd_rn(val[k], val[k + 1]));
                    const bool u = (k == 0) || (m != prev);
                    // a repeated midpoint lands on its (equal) predecessor
                    if (k > 0 && u) ta += 8;
                    sts_f64(ta, m);
                    kcount += u;
                    prev = m;
                }
            }
            double *tg = sthr + tid * M;
            for (int k2 = kcount; k2 < M; ++k2) tg[k2] = 0.0;
            if (tid < cnt) K[g0 + tid] = kcount;
        }
        {
            uint64_t *rg = sabove + (size_t)tid * M;
            uint64_t acc = 0, Tn = 0;
            uint32_t ra = smem_u32(rg + (kcount - 1));   // slot of the row being emitted (descending)
            double m = 0.0;                              // midpoint k, carried down from iteration k + 1
#pragma unroll
            for (int k = NB - 1; k >= 0; --k) {
                if (MITRE_EQ_M(k)) {
                    acc = 1ull << (63 - id[k]);
                    Tn = 0;
                    if (k > 0) m = __dmul_rn(0.5, __dadd_rn(val[k - 1], val[k]));
                } else if (MITRE_LT_M(k)) {
                    double m_below = 0.0;
                    bool uniq = true;
                    if (k > 0) {
                        m_below = __dmul_rn(0.5, __dadd_rn(val[k - 1], val[k]));
                        uniq = m != m_below;
                    }
                    sts_u64(ra, (m == val[k + 1]) ? Tn : acc);   // same row for a repeated midpoint
                    if (uniq) ra -= 8;
                    Tn = (val[k] == val[k + 1]) ? Tn : acc;
                    acc |= 1ull << (63 - id[k]);
                    m = m_below;
                }
            }
            for (int k2 = kcount; k2 < M; ++k2) rg[k2] = ~0ull;   // sentinel: no subject set is all ones
        }
        } else if (val[0] == 12345.678) K[g0 + tid] = id[3];
        // ---- finished tiles -> global ----
        if (full) fence_async_smem();
        __syncthreads();
        if (full) {
            if (tid == 0) {
                bulk_s2g(thresholds + g0 * M, sthr, (uint32_t)(T * M * 8));
                bulk_commit();
            }
        } else {
            for (int i = tid; i < cnt * M; i += T) thresholds[g0 * M + i] = sthr[i];
        }
        {
            ulonglong2 *out = reinterpret_cast<ulonglong2 *>(rows) + g0 * M;
            const int n_out = cnt * M;
#pragma unroll 2
            for (int i = tid; i < n_out; i += T) {
                const uint64_t above = sabove[i];
                out[i] = make_ulonglong2(above, above == ~0ull ? ~0ull : (~above & valid));
            }
        }
        __syncthreads();   // the rows tile is free for the next values
    }
    if (tid == 0) bulk_wait0();
}

// ---------------------------------------------------------------------------------------------------------
// Rolled variant of the staged sort/emit kernel (mitre_set_detector_split(6)).  sort_emit_staged_kernel keeps
// the sorted values and ids in registers through fully unrolled threshold / row loops: 150 registers (12 warps
// per SM) and 4432 SASS instructions = 71 KB of code against a 32 KB instruction cache.  Here only the sorting
// network itself (32-bit keys) is unrolled; the sorted values and ids go back to shared memory and every later
// phase is a rolled loop over them:
//   A [T][N]  values in (TMA)            -> thresholds out, [T][M] flat (bulk store)
//   B [T][N]  sorted values              -> `above` rows, in place
//   ID [NB][T] sorted subject ids (bytes)
// rows: iteration k (descending) reads B[k-1], B[k], B[k+1] and then overwrites slot k+1 -- no longer needed --
// with threshold k's row or, for a repeated midpoint, the marker kDupRow (bit 0 = subject 63, never set for
// N <= 63); an ascending pass compacts the rows to slots 0 .. K-1 and pads with the ~0 sentinel.
// ---------------------------------------------------------------------------------------------------------
constexpr uint64_t kDupRow = 1ull;

__host__ __device__ inline size_t sort_emit_rolled_bytes(int N, int NB)
{
    return (size_t)kStageThreads * (2 * (size_t)N * 8 + (size_t)NB);
}

template <int NB>
__global__ void __launch_bounds__(kStageThreads, 8)
sort_emit_rolled_kernel(const double *__restrict__ values, int64_t G, int N, double *__restrict__ thresholds,
                        int32_t *__restrict__ K, uint64_t *__restrict__ rows, SlopeGuardArgs sg)
{
    extern __shared__ __align__(128) unsigned char stage_smem[];
    __shared__ uint64_t bar;
    constexpr int T = kStageThreads;
    const int M = N - 1;
    const int tid = threadIdx.x;
    double *sA = reinterpret_cast<double *>(stage_smem);              // [T][N] values in; [T][M] thresholds out
    double *sB = sA + (size_t)T * N;                                  // [T][N] sorted values, then rows
    uint64_t *sBr = reinterpret_cast<uint64_t *>(sB);
    uint8_t *sID = reinterpret_cast<uint8_t *>(sB + (size_t)T * N);   // [NB][T]
    const uint64_t valid = ~0ull << (64 - N);
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    uint32_t phase = 0;
    const int64_t n_tiles = ceil_div64(G, T);
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t g0 = tile * T;
        const int cnt = (int)((G - g0 < T) ? (G - g0) : T);
        const bool full = cnt == T;
        if (full) {
            if (tid == 0) {
                bulk_wait_read0();       // the previous tile's thresholds have left A
                mbar_arrive_expect_tx(&bar, (uint32_t)(T * N * 8));
                bulk_g2s(sA, values + g0 * N, (uint32_t)(T * N * 8), &bar);
            }
            mbar_wait(&bar, phase);
            phase ^= 1;
        } else {
            if (tid == 0) bulk_wait_read0();
            __syncthreads();
            for (int i = tid; i < cnt * N; i += T) sA[i] = values[g0 * N + i];
            __syncthreads();
        }
        double *a = sA + tid * N, *b = sB + tid * N;
        uint64_t *br = sBr + tid * N;
        {
            // sort by 32-bit keys (registers), then sorted values and ids back to shared memory
            uint32_t key[NB];
#pragma unroll
            for (int s = 0; s < NB; ++s) key[s] = (s < N) ? sort_key32(a[s], s) : (0xFFFFFFC0u | (uint32_t)s);
            run_network_u32<NB>(key, std::make_index_sequence<NetHolder<NB>::net.count>{});
#pragma unroll
            for (int k = 0; k < NB; ++k) {
                const int id = (int)(key[k] & 63u);
                sID[k * T + tid] = (uint8_t)id;
                if (k < N) b[k] = a[id];
            }
        }
        // near-ties below the key's resolution: odd-even transposition on the exact values (rare)
        {
            bool bad = false;
            for (int k = 0; k + 1 < N; ++k) bad |= b[k] > b[k + 1];
            while (bad) {
                bad = false;
                for (int par = 0; par < 2; ++par)
                    for (int k = par; k + 1 < N; k += 2) {
                        const double x = b[k], y = b[k + 1];
                        if (x > y) {
                            bad = true;
                            b[k] = y; b[k + 1] = x;
                            const uint8_t i0 = sID[k * T + tid];
                            sID[k * T + tid] = sID[(k + 1) * T + tid];
                            sID[(k + 1) * T + tid] = i0;
                        }
                    }
            }
        }
        if (sg.group_type && tid < cnt && sg.group_type[g0 + tid] == 0) {
            // slope guard band, near-tie half (the constant-series half is the values kernel's)
            const double tol = kSlopeGuard * fmax(fabs(b[0]), fabs(b[M]));
            bool flag = false;
            for (int k = 0; k < M; ++k) flag |= (b[k] != b[k + 1]) && (b[k + 1] - b[k] <= tol);
            if (flag) sg.group_flags[g0 + tid] = 1;
        }
        __syncthreads();      // every thread has read its values from A: A becomes the thresholds tile
        int kcount = 0;
        {
            double *tg = sA + tid * M;
            double prev = 0.0, lo = b[0];
            for (int k = 0; k < M; ++k) {
                const double hi = b[k + 1];
                const double m = __dmul_rn(0.5, __dadd_rn(lo, hi));
                const bool u = (k == 0) || (m != prev);
                tg[kcount] = m;   // a repeated midpoint is overwritten by its successor
                kcount += u;
                prev = m;
                lo = hi;
            }
            for (int k2 = kcount; k2 < M; ++k2) tg[k2] = 0.0;
            if (tid < cnt) K[g0 + tid] = kcount;
        }
        {
            // rows, walking the sorted order downwards: acc = subjects at ranks > k; Tn = subjects after the
            // run of values equal to v[k+1]
            uint64_t acc = 1ull << (63 - sID[M * T + tid]), Tn = 0;
            double vk1 = b[M], vk = b[M - 1];
            double mk = __dmul_rn(0.5, __dadd_rn(vk, vk1));
            for (int k = M - 1; k >= 0; --k) {
                const double vkm = k > 0 ? b[k - 1] : 0.0;
                const double mkm = __dmul_rn(0.5, __dadd_rn(vkm, vk));
                const bool uniq = (k == 0) || (mk != mkm);
                const uint64_t above = (mk == vk1) ? Tn : acc;
                br[k + 1] = uniq ? above : kDupRow;
                Tn = (vk == vk1) ? Tn : acc;
                acc |= 1ull << (63 - sID[k * T + tid]);
                vk1 = vk; vk = vkm; mk = mkm;
            }
            int pos = 0;
            for (int j = 1; j <= M; ++j) {
                const uint64_t r = br[j];
                if (r != kDupRow) br[pos++] = r;
            }
            for (int k2 = pos; k2 < M; ++k2) br[k2] = ~0ull;
        }
        // ---- finished tiles -> global ----
        if (full) fence_async_smem();
        __syncthreads();
        if (full) {
            if (tid == 0) {
                bulk_s2g(thresholds + g0 * M, sA, (uint32_t)(T * M * 8));
                bulk_commit();
            }
        } else {
            for (int i = tid; i < cnt * M; i += T) thresholds[g0 * M + i] = sA[i];
        }
        {
            // rows tile has row stride N, the output is [groups][M] flat: (group, slot) advance without a division
            ulonglong2 *out = reinterpret_cast<ulonglong2 *>(rows) + g0 * M;
            const int n_out = cnt * M;
            const int dg = T / M, dj = T - dg * M;
            int g = tid / M, j = tid - g * M;
            for (int i = tid; i < n_out; i += T) {
                const uint64_t above = sBr[g * N + j];
                out[i] = make_ulonglong2(above, above == ~0ull ? ~0ull : (~above & valid));
                g += dg; j += dj;
                if (j >= M) { j -= M; ++g; }
            }
        }
        __syncthreads();   // B and ID are free for the next tile
    }
    if (tid == 0) bulk_wait0();
}

template <int NB>
static int launch_sort_emit_rolled(cudaStream_t st, const double *values, int64_t G, int N, double *thresholds,
                                   int32_t *K, uint64_t *rows)
{
    auto kern = sort_emit_rolled_kernel<NB>;
    const size_t smem = sort_emit_rolled_bytes(N, NB);
    MITRE_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    MITRE_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kStageThreads, smem));
    if (per_sm < 1) return MITRE_ERR_UNSUPPORTED;
    int64_t blocks = (int64_t)sm_count() * per_sm;
    const int64_t need = ceil_div64(G, kStageThreads);
    if (need < blocks) blocks = need;
    if (blocks < 1) blocks = 1;
    kern<<<(unsigned)blocks, kStageThreads, smem, st>>>(values, G, N, thresholds, K, rows, g_slope_guard);
    MITRE_LAUNCH_CHECK();
    return MITRE_OK;
}

template <int NB>
static int launch_sort_emit_staged(cudaStream_t st, const double *values, int64_t G, int N, double *thresholds,
                                   int32_t *K, uint64_t *rows)
{
    auto kern = sort_emit_staged_kernel<NB>;
    const size_t smem = sort_emit_stage_bytes(N);
    MITRE_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    MITRE_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kStageThreads, smem));
    if (per_sm < 1) return MITRE_ERR_UNSUPPORTED;
    if (g_emit_blocks_per_sm > 0 && g_emit_blocks_per_sm < per_sm) per_sm = g_emit_blocks_per_sm;
    int64_t blocks = (int64_t)sm_count() * per_sm;
    const int64_t need = ceil_div64(G, kStageThreads);
    if (need < blocks) blocks = need;
    if (blocks < 1) blocks = 1;
    kern<<<(unsigned)blocks, kStageThreads, smem, st>>>(values, G, N, thresholds, K, rows, g_detector_split == 5,
                                                        g_slope_guard);
    MITRE_LAUNCH_CHECK();
    return MITRE_OK;
}

template <int NB, int THREADS, int MINB, bool SYNC>
static int launch_sort_emit(cudaStream_t st, const double *values, int64_t G, int N, double *thresholds,
                            int32_t *K, uint64_t *rows)
{
    auto kern = sort_emit_kernel<NB, THREADS, MINB, SYNC>;
    int per_sm = 1;
    MITRE_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, 0));
    if (per_sm < 1) return MITRE_ERR_UNSUPPORTED;
    int64_t blocks = (int64_t)sm_count() * per_sm;
    const int64_t need = ceil_div64(G, THREADS);
    if (need < blocks) blocks = need;
    if (blocks < 1) blocks = 1;
    kern<<<(unsigned)blocks, THREADS, 0, st>>>(values, G, N, thresholds, K, rows);
    MITRE_LAUNCH_CHECK();
    return MITRE_OK;
}

template <int NB>
static int launch_split_impl(cudaStream_t st, const double *series_t, int64_t vstride, int N, int V,
                             const int32_t *wlo, const int32_t *wcnt, const int32_t *wstart,
                             const double *winv_n, const double *winv_sxx, const int32_t *widx,
                             const double *wdt, int64_t ld, const int64_t *group_base,
                             const uint8_t *slope_ok, const int32_t *task_window, const int8_t *task_type,
                             int n_wt, int64_t G, double *values, double *thresholds, int32_t *K,
                             uint64_t *rows)
{
    (void)widx;
    (void)series_t; (void)vstride; (void)V; (void)wlo; (void)wcnt; (void)wstart; (void)winv_n; (void)winv_sxx;
    (void)wdt; (void)ld; (void)group_base; (void)slope_ok; (void)task_window; (void)task_type; (void)n_wt;
    if (g_detector_split == 4) return launch_sort_emit<NB, 256, 1, false>(st, values, G, N, thresholds, K, rows);
    if (g_sort_emit_rolled && N >= 2) return launch_sort_emit_rolled<NB>(st, values, G, N, thresholds, K, rows);
    return launch_sort_emit_staged<NB>(st, values, G, N, thresholds, K, rows);
}

template <int NB>
static int launch_fused_impl(cudaStream_t st, const double *series_t, int64_t vstride, int N, int V,
                             const int32_t *wlo, const int32_t *wcnt, const int32_t *wstart,
                             const double *winv_n, const double *winv_sxx, const int32_t *widx,
                             const double *wdt, int64_t ld, const int64_t *group_base,
                             const uint8_t *slope_ok, const int32_t *task_window, const int8_t *task_type,
                             int n_wt, double *values, double *thresholds, int32_t *K, uint64_t *rows)
{
    const int vs_stride = N | 1;
    const size_t per_warp = ((size_t)kFusedCap * 32 + (size_t)32 * vs_stride + 3 * NB) * sizeof(double);
    int warps = (int)((smem_optin() - 1024) / per_warp);
    if (warps > kFusedThreads / 32) warps = kFusedThreads / 32;
    if (warps < 1) return MITRE_ERR_UNSUPPORTED;
    const size_t smem = per_warp * warps;
    auto kern = fused_detectors_kernel<NB>;
    MITRE_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    MITRE_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, warps * 32, smem));
    if (per_sm < 1) return MITRE_ERR_UNSUPPORTED;
    const int64_t n_tasks = (int64_t)n_wt * ((V + 31) / 32);
    int64_t blocks = (int64_t)sm_count() * per_sm;
    const int64_t need = ceil_div64(n_tasks, warps);
    if (need < blocks) blocks = need;
    if (blocks < 1) blocks = 1;
    kern<<<(unsigned)blocks, warps * 32, smem, st>>>(series_t, vstride, N, V, wlo, wcnt, wstart, winv_n,
                                                     winv_sxx, widx, wdt, ld, group_base, slope_ok,
                                                     task_window, task_type, n_wt, values, thresholds, K,
                                                     rows);
    MITRE_LAUNCH_CHECK();
    return MITRE_OK;
}

// test hook: div_small_int against the IEEE divide
static __global__ void selftest_divide_kernel(const double *a, const int32_t *n, int64_t count, double *fast,
                                              double *exact)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const double nd = (double)n[i];
    fast[i] = div_small_int(a[i], nd, __ddiv_rn(1.0, nd));
    exact[i] = __ddiv_rn(a[i], nd);
}

}  // namespace mitre
