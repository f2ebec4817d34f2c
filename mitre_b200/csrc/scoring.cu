// scoring.cu -- all-candidate marginal-likelihood scoring (replaces the reference's one native
// component, mitre/efficient_likelihoods.pyx:15-169) and the batched dense likelihood
// (mitre/mvnpdf_cholesky.py:53-74 as called from logit_rules.py:171-177).
//
// Algebra (Woodbury + matrix determinant lemma, kappa = y - 1/2 so Omega z = kappa):
//   Sigma   = Omega^-1 + s2 X X'               A = I/s2 + X' Omega X = L L'      b = X' kappa
//   logdet Sigma = -sum log omega + p log s2 + 2 sum log L_jj
//   z' Sigma^-1 z = sum kappa_i^2/omega_i - |L^-1 b|^2
// Appending the candidate column v (a subject bitmask) to X:
//   w = L^-1 X' Omega v = sum_{i in v} w_i ,   w_i = L^-1 (omega_i x_i)
//   c = sum_{i in v} omega_i ,   r = sum_{i in v} (kappa_i - w_i . L^-1 b)
//   s = 1/s2 + c - |w|^2                       (Schur complement, >= ~1/s2)
//   loglik_k = loglik_base - 1/2 log(s2 s) + 1/2 r^2 / s
// so a candidate is a masked sum of the per-subject vectors q_i = (omega_i, r_i, w_i) followed by
// p+3 flops, one log and one divide.  The reference instead factors the N x N Sigma once and
// does an O(N^2) Givens rank-1 update per candidate; both are the same number to ~1e-14 rel.
//
// Kernels:
//   score_setup_kernel   1 CTA: base system, Cholesky, q_i (masked-out subjects get q_i = 0)
//   score_w1_kernel      N <= 64: one thread per candidate; the masked sum is ceil(N/CB) lookups
//                        in a shared-memory table of all 2^CB subset sums per CB-subject chunk
//   score_wide_kernel    N  > 64: one warp per candidate, lanes own 64-subject words
#include <math.h>

#include "common.cuh"

namespace mitre {

// workspace layout (doubles): header[8] then q[N][EP]
//   header[0] = base log-likelihood C0, [1] = 1/sigma2, [2] = sigma2
constexpr int kHdr = 8;
constexpr int kTicketSlot = 6;         // header[6] (as uint32): ticket of the candidate kernel's last-CTA reduction
constexpr int kMaxStatCtas = 2048;   // per-CTA (max, sum-exp) partials live at the end of the workspace
__host__ __device__ inline int ep_for(int p) { return (p + 2 + 1) & ~1; }  // even, >= p+2

// Table t covers subjects [table16_start, +table16_bits): the split is aligned to the LAST subject
// (table 0 takes the N mod 16 remainder), because neighbouring rows of a lexsorted table differ in
// their trailing subjects: all of a warp's divergent lookups then fall into the last table and the
// leading tables are warp-uniform loads.
// The same split with B subjects per table (the duplicate-aware kernel uses B = 11: the last table is then
// 2048 entries = 96 KB at p = 4 and stays in L1, where a 16-subject table (3 MB) is an L2 round trip per row).
__host__ __device__ inline int n_tables(int N, int B) { return (N + B - 1) / B; }
__host__ __device__ inline int table_bits(int N, int B, int t)
{
    return t == 0 ? N - B * (n_tables(N, B) - 1) : B;
}
__host__ __device__ inline int table_start(int N, int B, int t)
{
    return t == 0 ? 0 : table_bits(N, B, 0) + B * (t - 1);
}
__host__ __device__ inline size_t tables_doubles(int N, int B, int stride)
{
    size_t n = 0;
    for (int t = 0; t < n_tables(N, B); ++t) n += ((size_t)1 << table_bits(N, B, t)) * stride;
    return n;
}
constexpr int kMaxTables = 6;          // N <= 64, B >= 11
__host__ __device__ inline int n_tables16(int N) { return (N + 15) >> 4; }
__host__ __device__ inline int table16_bits(int N, int t)
{
    const int first = N - 16 * (n_tables16(N) - 1);
    return t == 0 ? first : 16;
}
__host__ __device__ inline int table16_start(int N, int t)
{
    return t == 0 ? 0 : table16_bits(N, 0) + 16 * (t - 1);
}
__host__ __device__ inline size_t tables16_doubles(int N, int EP)
{
    size_t n = 0;
    for (int t = 0; t < n_tables16(N); ++t) n += ((size_t)1 << table16_bits(N, t)) * EP;
    return n;
}


// ---------------------------------------------------------------------------------------------
// Table-driven log / reciprocal / exp for the per-candidate epilogue.  The stock fp64 log and
// divide cost ~70 instructions per candidate; here s = 2^e * m, m in [1,2) is reduced against
// c_i = 1 + (i + 1/2)/128 (i = top 7 mantissa bits): r = m/c_i - 1, |r| <= 2^-8, and
//   log s = e ln2 + log c_i + log1p(r)          (degree-6 series, truncation 2e-18)
//   1/s   = 2^-e (1/c_i) (1-r)(1+r^2)(1+r^4)    (= sum_{k<8} (-r)^k, truncation 5e-20)
// Absolute error ~3e-16 in the log, ~4 ulp in the reciprocal: far inside the 1e-9 contract.
// ---------------------------------------------------------------------------------------------
constexpr int kLogTab = 128;
constexpr int kExpTab = 64;

__device__ __forceinline__ void build_math_tables(double2 *logtab, double *exptab)
{
    for (int i = threadIdx.x; i < kLogTab; i += blockDim.x) {
        const double c = 1.0 + (i + 0.5) / kLogTab;
        logtab[i] = make_double2(1.0 / c, log(c));
    }
    for (int j = threadIdx.x; j < kExpTab; j += blockDim.x) exptab[j] = exp2((double)j / kExpTab);
}

// s must be a positive normal double.  half_rcp = 1/(2s).
__device__ __forceinline__ void log_and_half_rcp(double s, const double2 *__restrict__ logtab, double &log_s,
                                                 double &half_rcp)
{
    const int hi = __double2hiint(s), lo = __double2loint(s);
    const int e = (hi >> 20) - 1023;
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, lo);
    const double2 t = logtab[(hi >> 13) & (kLogTab - 1)];
    const double r = fma(m, t.x, -1.0);
    const double r2 = r * r;
    double p = fma(r, -1.0 / 6.0, 0.2);
    p = fma(r, p, -0.25);
    p = fma(r, p, 1.0 / 3.0);
    p = fma(r, p, -0.5);
    const double l1p = fma(r2, p, r);
    log_s = fma((double)e, 0.69314718055994530942, t.y) + l1p;
    const double r4 = r2 * r2;
    const double q = ((1.0 - r) * (1.0 + r2)) * (1.0 + r4);
    half_rcp = (t.x * q) * __hiloint2double((1022 - e) << 20, 0);
}

// exp(x) for x <= 0 (returns 0 below the double range); 2^(j/64) table + degree-5 series.
__device__ __forceinline__ double exp_nonpos(double x, const double *__restrict__ exptab)
{
    if (!(x > -700.0)) return 0.0;
    const double t = fma(x, 92.332482616893656877, 6755399441055744.0);   // 64/ln2, 1.5*2^52
    const int n = __double2loint(t);
    const double nd = t - 6755399441055744.0;
    double r = fma(nd, -0.010830424667801708, x);                          // ln2/64 hi (29 bits)
    r = fma(nd, -2.8447437476627285e-11, r);                               // ln2/64 lo
    double p = fma(r, 1.0 / 120.0, 1.0 / 24.0);
    p = fma(r, p, 1.0 / 6.0);
    p = fma(r, p, 0.5);
    p = fma(r * r, p, r);                                                   // exp(r) - 1
    const double b = exptab[n & (kExpTab - 1)];
    const double v = fma(b, p, b);
    return __hiloint2double(__double2hiint(v) + ((n >> 6) << 20), __double2loint(v));
}

// (max, sum exp(w - max)) pairs; empty = (-inf, 0)
struct MaxSum {
    double m, s;
};
__device__ __forceinline__ MaxSum ms_merge(MaxSum x, MaxSum y)
{
    if (y.s == 0.0) return x;
    if (x.s == 0.0) return y;
    MaxSum r;
    if (x.m >= y.m) { r.m = x.m; r.s = x.s + y.s * exp(y.m - x.m); }
    else { r.m = y.m; r.s = y.s + x.s * exp(x.m - y.m); }
    return r;
}
__device__ __forceinline__ void ms_push(MaxSum &a, double w, const double *__restrict__ exptab)
{
    if (w == -INFINITY) return;
    const double e = exp_nonpos(-fabs(w - a.m), exptab);
    if (w > a.m) { a.s = fma(a.s, e, 1.0); a.m = w; }
    else a.s += e;
}

__global__ void __launch_bounds__(256)
score_stats_final_kernel(const double *__restrict__ part, int nparts, double *__restrict__ stats)
{
    __shared__ double sm[256], ss[256];
    MaxSum acc = {-INFINITY, 0.0};
    for (int i = threadIdx.x; i < nparts; i += 256) acc = ms_merge(acc, MaxSum{part[2 * i], part[2 * i + 1]});
    sm[threadIdx.x] = acc.m; ss[threadIdx.x] = acc.s;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            const MaxSum x = ms_merge(MaxSum{sm[threadIdx.x], ss[threadIdx.x]},
                                      MaxSum{sm[threadIdx.x + s], ss[threadIdx.x + s]});
            sm[threadIdx.x] = x.m; ss[threadIdx.x] = x.s;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { stats[0] = sm[0]; stats[1] = ss[0]; }
}

// ---------------------------------------------------------------------------------------------
// Base system for one design matrix.  Called by every thread of a CTA; smem scratch:
//   sA[p*p] (in: nothing, out: L strictly-lower-triangular, diagonal = 1 / L_jj), sb[p] (out: yb = L^-1 b), sred[blockDim]
// Returns the base log-likelihood log N(kappa/omega; 0, Omega^-1 + s2 X X') in every thread;
// *ok = 0 if A is not positive definite.
// ---------------------------------------------------------------------------------------------
__device__ double base_system(const double *__restrict__ X, int ldx, const double *__restrict__ omega,
                              const double *__restrict__ kappa, int N, int p, double sigma2,
                              double *sA, double *sb, double *sred, int *s_ok)
{
    const int tid = threadIdx.x, nt = blockDim.x;
    // A = I/s2 + X' Omega X, b = X' kappa : one (j,k) entry per thread, fixed i order.
    for (int e = tid; e < p * p + p; e += nt) {
        if (e < p * p) {
            const int j = e / p, k = e % p;
            double acc = 0.0;
            for (int i = 0; i < N; ++i) acc += omega[i] * X[(size_t)i * ldx + j] * X[(size_t)i * ldx + k];
            sA[e] = acc + (j == k ? 1.0 / sigma2 : 0.0);
        } else {
            const int j = e - p * p;
            double acc = 0.0;
            for (int i = 0; i < N; ++i) acc += X[(size_t)i * ldx + j] * kappa[i];
            sb[j] = acc;
        }
    }
    // sum_i ( kappa_i^2/omega_i - log omega_i ): strided partials, fixed-order tree.
    double part = 0.0;
    for (int i = tid; i < N; i += nt) part += kappa[i] * kappa[i] / omega[i] - log(omega[i]);
    sred[tid] = part;
    if (tid == 0) *s_ok = 1;
    __syncthreads();
    for (int s = nt >> 1; s > 0; s >>= 1) {
        if (tid < s) sred[tid] += sred[tid + s];
        __syncthreads();
    }
    if (tid < 32) {
        // Cholesky of the p x p system by one warp, lane = row: per column the pivot lane forms d_j, every
        // lane below it its entry -- a dependency depth of ~p^2 multiply-adds instead of the p^3/3 + divides
        // + square roots of a single thread (that serial stretch was most of the prepare launch).  The
        // reciprocal square root of the pivot replaces every later division: L_ij = s * rs_j,
        // L_jj = d_j * rs_j, yb_i = s * rs_i, log L_jj = log(d_j) / 2; the diagonal itself is stored as its
        // reciprocal (sA[j][j] = 1 / L_jj), which is what subject_vector multiplies by.
        const int lane = tid;
        int ok = 1;
        double half_logd = 0.0;                      // lane j: log(L_jj)
        for (int j = 0; j < p; ++j) {
            double d = 0.0;
            if (lane == j) {
                d = sA[j * p + j];
                for (int k = 0; k < j; ++k) d -= sA[j * p + k] * sA[j * p + k];
            }
            d = __shfl_sync(0xffffffffu, d, j);
            if (!(d > 0.0)) { ok = 0; break; }        // warp-uniform
            const double rs = rsqrt(d);
            if (lane == j) {
                half_logd = 0.5 * log(d);
                sA[j * p + j] = rs;
            } else if (lane > j && lane < p) {
                double v = sA[lane * p + j];
                for (int k = 0; k < j; ++k) v -= sA[lane * p + k] * sA[j * p + k];
                sA[lane * p + j] = v * rs;
            }
            __syncwarp();
        }
        double logdetL = 0.0;
        for (int j = 0; j < p; ++j) logdetL += __shfl_sync(0xffffffffu, half_logd, j);   // fixed order
        if (lane == 0) {
            double yy = 0.0;
            if (ok) {
                for (int i = 0; i < p; ++i) {
                    double v = sb[i];
                    for (int k = 0; k < i; ++k) v -= sA[i * p + k] * sb[k];
                    v *= sA[i * p + i];
                    sb[i] = v;
                    yy += v * v;
                }
            }
            *s_ok = ok;
            // sred[0] = sum kappa^2/omega - sum log omega
            const double quad_plus_logdet = sred[0] - yy + p * log(sigma2) + 2.0 * logdetL;
            sred[0] = -0.5 * quad_plus_logdet - 0.5 * N * log(2.0 * M_PI);
        }
    }
    __syncthreads();
    return sred[0];
}

// q_i = (omega_i, r_i, w_i) of one subject (zeros outside the rule mask, NaN when the base system failed)
__device__ __forceinline__ void subject_vector(const double *__restrict__ X, const double *__restrict__ omega,
                                               const double *__restrict__ kappa, const uint64_t *__restrict__ mask,
                                               int i, int p, int EP, const double *sA, const double *sb, int ok,
                                               double *qi)
{
    const bool in_mask = (mask == nullptr) || (mask[i >> 6] & subject_bit(i));
    if (!in_mask || !ok) {
        for (int e = 0; e < EP; ++e) qi[e] = ok ? 0.0 : nan("");
        return;
    }
    const double om = omega[i];
    double w[MITRE_MAX_P];
    double dot = 0.0;
    for (int j = 0; j < p; ++j) {
        double s = om * X[(size_t)i * p + j];
        for (int k = 0; k < j; ++k) s -= sA[j * p + k] * w[k];
        s *= sA[j * p + j];                           // the diagonal holds 1 / L_jj
        w[j] = s;
        dot += s * sb[j];
    }
    qi[0] = om;
    qi[1] = kappa[i] - dot;
    for (int j = 0; j < p; ++j) qi[2 + j] = w[j];
    for (int e = p + 2; e < EP; ++e) qi[e] = 0.0;
}

__global__ void __launch_bounds__(256)
score_setup_kernel(const double *__restrict__ X, const double *__restrict__ omega,
                   const double *__restrict__ kappa, const uint64_t *__restrict__ mask, int N, int p,
                   double sigma2, double *__restrict__ ws, int32_t *status)
{
    __shared__ double sA[MITRE_MAX_P * MITRE_MAX_P];
    __shared__ double sb[MITRE_MAX_P];
    __shared__ double sred[256];
    __shared__ int s_ok;
    const int EP = ep_for(p);
    const double C0 = base_system(X, p, omega, kappa, N, p, sigma2, sA, sb, sred, &s_ok);
    if (threadIdx.x == 0) {
        ws[0] = C0;
        ws[1] = 1.0 / sigma2;
        ws[2] = sigma2;
        ws[3] = C0 - 0.5 * log(sigma2);   // base log-likelihood with the -1/2 log sigma2 of the new column folded in
        if (!s_ok && status) atomicOr(reinterpret_cast<unsigned int *>(status), MITRE_STATUS_BASE_NOT_PD);
    }
    double *q = ws + kHdr;
    for (int i = threadIdx.x; i < N; i += blockDim.x)
        subject_vector(X, omega, kappa, mask, i, p, EP, sA, sb, s_ok, q + (size_t)i * EP);
}

// One launch for everything a large-table sweep needs before the candidate kernel: every CTA redoes the
// (tiny) base system in shared memory and builds its share of the 16-subject subset-sum tables straight
// from it; CTA 0 also writes the header and q to the workspace and clears the ticket of the candidate
// kernel's final reduction.  Replaces score_setup_kernel + score_build_tables16_kernel (two dependent
// launches, ~30 us) where the step itself is a few hundred microseconds (sharded sweeps).
__global__ void __launch_bounds__(256)
score_prepare16_kernel(const double *__restrict__ X, const double *__restrict__ omega,
                       const double *__restrict__ kappa, const uint64_t *__restrict__ mask, int N, int p,
                       double sigma2, double *__restrict__ ws, int32_t *status, int stride, int B,
                       double *__restrict__ tables)
{
    __shared__ double sA[MITRE_MAX_P * MITRE_MAX_P];
    __shared__ double sb[MITRE_MAX_P];
    __shared__ double sred[256];
    __shared__ int s_ok;
    extern __shared__ double sq[];                  // [N][EP], then the staged inputs X[N][p], omega[N], kappa[N]
    const int EP = ep_for(p);
    // stage the sweep inputs once (one round of global latency instead of one per loop iteration)
    double *sX = sq + (size_t)N * EP, *som = sX + (size_t)N * p, *ska = som + N;
    for (int i = threadIdx.x; i < N * p; i += blockDim.x) sX[i] = X[i];
    for (int i = threadIdx.x; i < N; i += blockDim.x) { som[i] = omega[i]; ska[i] = kappa[i]; }
    __syncthreads();
    const double C0 = base_system(sX, p, som, ska, N, p, sigma2, sA, sb, sred, &s_ok);
    for (int i = threadIdx.x; i < N; i += blockDim.x)
        subject_vector(sX, som, ska, mask, i, p, EP, sA, sb, s_ok, sq + (size_t)i * EP);
    __syncthreads();
    if (blockIdx.x == 0) {
        if (threadIdx.x == 0) {
            ws[0] = C0;
            ws[1] = 1.0 / sigma2;
            ws[2] = sigma2;
            ws[3] = C0 - 0.5 * log(sigma2);
            *reinterpret_cast<unsigned int *>(ws + kTicketSlot) = 0u;
            if (!s_ok && status) atomicOr(reinterpret_cast<unsigned int *>(status), MITRE_STATUS_BASE_NOT_PD);
        }
        for (int i = threadIdx.x; i < N * EP; i += blockDim.x) ws[kHdr + i] = sq[i];
    }
    // tables: one thread per subset pattern forms all EP components of the entry (ascending subject order, like
    // score_build_tables16_kernel, so both layouts hold bit-identical sums); a thread per (pattern, component)
    // spent most of its instructions on index arithmetic and re-testing the pattern bits
    const int nt = n_tables(N, B);
    size_t base = 0;
    for (int t = 0; t < nt; ++t) {
        const int bits = table_bits(N, B, t), start = table_start(N, B, t);
        const unsigned npat = 1u << bits;
        for (unsigned pat = blockIdx.x * blockDim.x + threadIdx.x; pat < npat; pat += gridDim.x * blockDim.x) {
            double acc[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] = 0.0;
            for (int j = 0; j < bits; ++j) {
                if ((pat >> (bits - 1 - j)) & 1) {
                    const double *qj = sq + (start + j) * EP;
#pragma unroll
                    for (int e = 0; e < 8; ++e)
                        if (e < EP) acc[e] += qj[e];
                }
            }
            double *dst = tables + base + (size_t)pat * stride;
#pragma unroll
            for (int e = 0; e < 8; ++e)
                if (e < stride) dst[e] = acc[e];
        }
        base += (size_t)npat * stride;
    }
}

// Finish one candidate from its masked sums acc = (c, r, w_0..).
template <int EP>
__device__ __forceinline__ double finish_candidate(const double (&acc)[EP], double C0, double inv_s2,
                                                   double s2, bool &bad)
{
    double ww = 0.0;
#pragma unroll
    for (int e = 2; e < EP; ++e) ww = fma(acc[e], acc[e], ww);
    const double s = (inv_s2 + acc[0]) - ww;
    bad = bad || !(s > 0.0);
    return C0 - 0.5 * log(s2 * s) + 0.5 * (acc[1] * acc[1]) / s;
}

// ---------------------------------------------------------------------------------------------
// N <= 64: thread per candidate, subset-sum table per chunk of CB subjects in shared memory.
// Table entry (chunk c, pattern t): sum of q_i over subjects i = c*CB + j whose pattern bit
// (CB-1-j) is set, EP doubles, added in ascending subject order.
// STATS: additionally accumulate (max, sum exp(w - max)) of w = loglik (+ log_prior) per CTA.
// ---------------------------------------------------------------------------------------------
template <int EP, int CB, bool STATS>
__global__ void __launch_bounds__(256)
score_w1_kernel(const uint64_t *__restrict__ rows, int64_t P, const double *__restrict__ ws, int N,
                double *__restrict__ out, int32_t *status, const double *__restrict__ log_prior,
                double *__restrict__ stat_partials)
{
    extern __shared__ __align__(16) double lut[];
    constexpr int PAT = 1 << CB;
    const int nchunks = (N + CB - 1) / CB;
    double2 *logtab = reinterpret_cast<double2 *>(lut + (size_t)nchunks * PAT * EP);
    double *exptab = reinterpret_cast<double *>(logtab + kLogTab);
    const double *q = ws + kHdr;
    for (int idx = threadIdx.x; idx < nchunks * PAT * EP; idx += blockDim.x) {
        const int e = idx % EP;
        const int pat = (idx / EP) % PAT;
        const int c = idx / (EP * PAT);
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < CB; ++j) {
            const int i = c * CB + j;
            if (((pat >> (CB - 1 - j)) & 1) && i < N) acc += q[(size_t)i * EP + e];
        }
        lut[idx] = acc;
    }
    build_math_tables(logtab, exptab);
    __syncthreads();
    const double inv_s2 = ws[1], C1 = ws[3];
    bool bad = false;
    MaxSum run = {-INFINITY, 0.0};
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t rnext = k < P ? ld_stream_u64(rows + k) : 0;
    for (; k < P; k += stride) {
        const uint64_t r = rnext;
        if (k + stride < P) rnext = ld_stream_u64(rows + k + stride);   // prefetch the next row
        double acc[EP];
        {
            const int pat = (int)(r >> (64 - CB));
            const double2 *e = reinterpret_cast<const double2 *>(lut + (size_t)pat * EP);
#pragma unroll
            for (int j = 0; j < EP / 2; ++j) {
                const double2 v = e[j];
                acc[2 * j] = v.x;
                acc[2 * j + 1] = v.y;
            }
        }
        for (int c = 1; c < nchunks; ++c) {
            const int pat = (int)(r >> (64 - CB * (c + 1))) & (PAT - 1);
            const double2 *e = reinterpret_cast<const double2 *>(lut + ((size_t)c * PAT + pat) * EP);
#pragma unroll
            for (int j = 0; j < EP / 2; ++j) {
                const double2 v = e[j];
                acc[2 * j] += v.x;
                acc[2 * j + 1] += v.y;
            }
        }
        double ww = 0.0;
#pragma unroll
        for (int e = 2; e < EP; ++e) ww = fma(acc[e], acc[e], ww);
        const double sc = (inv_s2 + acc[0]) - ww;          // Schur complement
        const bool ok = sc > 1e-300 && sc < 1e300;
        bad = bad || !ok;
        double log_s, half_rcp;
        log_and_half_rcp(ok ? sc : 1.0, logtab, log_s, half_rcp);
        double val = fma(-0.5, log_s, C1) + (acc[1] * acc[1]) * half_rcp;
        if (!ok) val = nan("");
        st_stream_f64(out + k, val);
        if (STATS) ms_push(run, log_prior ? val + log_prior[k] : val, exptab);
    }
    if (bad && status) atomicOr(reinterpret_cast<unsigned int *>(status), MITRE_STATUS_BAD_SCHUR);
    if (STATS) {
        __shared__ double sm[256], ss[256];
        sm[threadIdx.x] = run.m; ss[threadIdx.x] = run.s;
        __syncthreads();
        for (int s = 128; s > 0; s >>= 1) {
            if (threadIdx.x < s) {
                const MaxSum x = ms_merge(MaxSum{sm[threadIdx.x], ss[threadIdx.x]},
                                          MaxSum{sm[threadIdx.x + s], ss[threadIdx.x + s]});
                sm[threadIdx.x] = x.m; ss[threadIdx.x] = x.s;
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) { stat_partials[2 * blockIdx.x] = sm[0]; stat_partials[2 * blockIdx.x + 1] = ss[0]; }
    }
}

// ---------------------------------------------------------------------------------------------
// N <= 64, byte tables, FOUR consecutive candidates per thread.  The table is lexsorted
// (rules.py:439-462), so neighbouring rows share their leading bytes almost always: the subset
// sums of the bytes common to the four rows are fetched and added once (prefix), only the
// trailing bytes are looked up per row.  That cuts shared-memory traffic -- the binding resource
// of the one-candidate-per-thread kernel (240 B of table per candidate at 128 B/clk/SM) -- by
// ~1.8x and gives every thread four independent dependency chains.
// Requires rows / out 32-byte aligned (checked by the launcher).
// ---------------------------------------------------------------------------------------------
template <int EP>
__device__ __forceinline__ void lut_add(double (&acc)[EP], const double *__restrict__ entry)
{
    const double2 *e = reinterpret_cast<const double2 *>(entry);
#pragma unroll
    for (int j = 0; j < EP / 2; ++j) {
        const double2 v = e[j];
        acc[2 * j] += v.x;
        acc[2 * j + 1] += v.y;
    }
}

template <int EP>
__device__ __forceinline__ void lut_add_g(double (&acc)[EP], const double *__restrict__ entry)
{
    const double2 *e = reinterpret_cast<const double2 *>(entry);
#pragma unroll
    for (int j = 0; j < EP / 2; ++j) {
        const double2 v = __ldg(e + j);
        acc[2 * j] += v.x;
        acc[2 * j + 1] += v.y;
    }
}

template <int EP>
__device__ __forceinline__ double finish_tab(const double (&acc)[EP], double inv_s2, double C1,
                                             const double2 *__restrict__ logtab, bool &bad)
{
    double ww = 0.0;
#pragma unroll
    for (int e = 2; e < EP; ++e) ww = fma(acc[e], acc[e], ww);
    const double sc = (inv_s2 + acc[0]) - ww;          // Schur complement
    const bool ok = sc > 1e-300 && sc < 1e300;
    bad = bad || !ok;
    double log_s, half_rcp;
    log_and_half_rcp(ok ? sc : 1.0, logtab, log_s, half_rcp);
    const double val = fma(-0.5, log_s, C1) + (acc[1] * acc[1]) * half_rcp;
    return ok ? val : nan("");
}

template <int EP, bool STATS>
__global__ void __launch_bounds__(256)
score_w1r4_kernel(const uint64_t *__restrict__ rows, int64_t P, const double *__restrict__ ws, int N,
                  double *__restrict__ out, int32_t *status, const double *__restrict__ log_prior,
                  double *__restrict__ stat_partials)
{
    extern __shared__ __align__(16) double lut[];
    constexpr int CB = 8, PAT = 256;
    const int nchunks = (N + CB - 1) / CB;
    double2 *logtab = reinterpret_cast<double2 *>(lut + (size_t)nchunks * PAT * EP);
    double *exptab = reinterpret_cast<double *>(logtab + kLogTab);
    const double *q = ws + kHdr;
    for (int idx = threadIdx.x; idx < nchunks * PAT * EP; idx += blockDim.x) {
        const int e = idx % EP;
        const int pat = (idx / EP) % PAT;
        const int c = idx / (EP * PAT);
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < CB; ++j) {
            const int i = c * CB + j;
            if (((pat >> (CB - 1 - j)) & 1) && i < N) acc += q[(size_t)i * EP + e];
        }
        lut[idx] = acc;
    }
    build_math_tables(logtab, exptab);
    __syncthreads();
    const double inv_s2 = ws[1], C1 = ws[3];
    bool bad = false;
    MaxSum run = {-INFINITY, 0.0};
    const int64_t quads = P >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const ulonglong2 *rows2 = reinterpret_cast<const ulonglong2 *>(rows);
    double2 *out2 = reinterpret_cast<double2 *>(out);
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < quads; t += stride) {
        const ulonglong2 ra = rows2[2 * t], rb = rows2[2 * t + 1];
        const uint64_t r[4] = {ra.x, ra.y, rb.x, rb.y};
        const uint64_t diff = (r[0] ^ r[1]) | (r[0] ^ r[2]) | (r[0] ^ r[3]);
        const int shared_chunks = min(nchunks, diff ? (__clzll((long long)diff) >> 3) : 8);
        double prefix[EP], acc[4][EP];
#pragma unroll
        for (int e = 0; e < EP; ++e) {
            prefix[e] = 0.0;
            acc[0][e] = acc[1][e] = acc[2][e] = acc[3][e] = 0.0;
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            if (c < nchunks) {
                const int sh = 56 - 8 * c;
                const double *tab = lut + (size_t)c * PAT * EP;
                if (c < shared_chunks) {
                    lut_add<EP>(prefix, tab + ((r[0] >> sh) & 255) * EP);
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) lut_add<EP>(acc[i], tab + ((r[i] >> sh) & 255) * EP);
                }
            }
        }
        double val[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
            for (int e = 0; e < EP; ++e) acc[i][e] += prefix[e];
            val[i] = finish_tab<EP>(acc[i], inv_s2, C1, logtab, bad);
        }
        out2[2 * t] = make_double2(val[0], val[1]);
        out2[2 * t + 1] = make_double2(val[2], val[3]);
        if (STATS) {
#pragma unroll
            for (int i = 0; i < 4; ++i) ms_push(run, log_prior ? val[i] + log_prior[4 * t + i] : val[i], exptab);
        }
    }
    // tail: the last P % 4 rows, one thread each
    if (blockIdx.x == 0 && threadIdx.x < (P & 3)) {
        const int64_t k = (quads << 2) + threadIdx.x;
        const uint64_t r = rows[k];
        double acc[EP];
#pragma unroll
        for (int e = 0; e < EP; ++e) acc[e] = 0.0;
        for (int c = 0; c < nchunks; ++c) lut_add<EP>(acc, lut + ((size_t)c * PAT + ((r >> (56 - 8 * c)) & 255)) * EP);
        const double val = finish_tab<EP>(acc, inv_s2, C1, logtab, bad);
        out[k] = val;
        if (STATS) ms_push(run, log_prior ? val + log_prior[k] : val, exptab);
    }
    if (bad && status) atomicOr(reinterpret_cast<unsigned int *>(status), MITRE_STATUS_BAD_SCHUR);
    if (STATS) {
        __shared__ double sm[256], ss[256];
        sm[threadIdx.x] = run.m; ss[threadIdx.x] = run.s;
        __syncthreads();
        for (int s = 128; s > 0; s >>= 1) {
            if (threadIdx.x < s) {
                const MaxSum x = ms_merge(MaxSum{sm[threadIdx.x], ss[threadIdx.x]},
                                          MaxSum{sm[threadIdx.x + s], ss[threadIdx.x + s]});
                sm[threadIdx.x] = x.m; ss[threadIdx.x] = x.s;
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) { stat_partials[2 * blockIdx.x] = sm[0]; stat_partials[2 * blockIdx.x + 1] = ss[0]; }
    }
}

// ---------------------------------------------------------------------------------------------
// N <= 64, LARGE lexsorted tables: 16-subject subset-sum tables in global memory (L2 / L1
// resident: 65536 x EP doubles = 3 MB per table at p = 4) instead of 8-subject tables in shared
// memory.  ceil(N/16) lookups + adds per candidate instead of ceil(N/8); and because the rows
// are lexsorted (rules.py:439-462), the 32 candidates of a warp almost always agree on their
// leading 32 subjects, so those lookups are warp-uniform loads served from L1 -- the
// shared-memory bandwidth that bounds score_w1_kernel (240 B of table reads per candidate) is
// gone.  No shared-memory table => full occupancy (48 registers/thread).
// Costs a table build per sweep (score_build_tables16_kernel, ~3 MB written per table), so the
// launcher only takes this path for P >= kLargeTableMinRows.
// ---------------------------------------------------------------------------------------------
constexpr int64_t kLargeTableMinRows = 4 << 20;
static int64_t g_large_table_min_rows = kLargeTableMinRows;
static int g_groups_warp = 1;    // mitre_score_groups, N <= 64: 1 warp per group (default), 0 CTA per group
static int g_wide_tables = 1;    // N > 64: 1 byte tables (score_wide8_kernel), 0 per-set-bit sums (score_wide_kernel)
static int g_split_tables = 0;   // 1: last table in shared memory (score_w1s4_kernel), 0: score_w1g4_kernel (default)

// entry (table t, pattern x): sum of q_i over subjects i = 16 t + j whose pattern bit (bits-1-j)
// is set, in ascending subject order.
// `stride` doubles per entry: EP (packed, one lane reads a whole entry) or 8 (64-byte entries whose
// 16-byte quarters are read by four cooperating lanes, score_w1d_kernel); components >= EP are zero.
__global__ void __launch_bounds__(256)
score_build_tables16_kernel(const double *__restrict__ ws, int N, int EP, int stride, double *__restrict__ tables)
{
    const double *q = ws + kHdr;
    const int nt = n_tables16(N);
    size_t base = 0;
    for (int t = 0; t < nt; ++t) {
        const int bits = table16_bits(N, t);
        const size_t count = ((size_t)1 << bits) * stride;
        for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < count;
             idx += (size_t)gridDim.x * blockDim.x) {
            const int e = (int)(idx % stride);
            const unsigned pat = (unsigned)(idx / stride);
            double acc = 0.0;
            if (e < EP)
                for (int j = 0; j < bits; ++j)
                    if ((pat >> (bits - 1 - j)) & 1) acc += q[(size_t)(table16_start(N, t) + j) * EP + e];
            tables[base + idx] = acc;
        }
        base += count;
    }
}

template <int EP, bool STATS>
__global__ void __launch_bounds__(256)
score_w1g_kernel(const uint64_t *__restrict__ rows, int64_t P, const double *__restrict__ ws, int N,
                 const double *__restrict__ tables, double *__restrict__ out, int32_t *status,
                 const double *__restrict__ log_prior, double *__restrict__ stat_partials)
{
    __shared__ double2 logtab[kLogTab];
    __shared__ double exptab[kExpTab];
    build_math_tables(logtab, exptab);
    __syncthreads();
    const double inv_s2 = ws[1], C1 = ws[3];
    const int nt = n_tables16(N);
    // table bases and shifts (nt <= 4)
    const double *tb[4];
    int shift[4];
    unsigned msk[4];
    {
        size_t base = 0;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int bits = t < nt ? table16_bits(N, t) : 0;
            tb[t] = tables + base;
            shift[t] = 64 - (t < nt ? table16_start(N, t) : 0) - bits;
            msk[t] = (1u << bits) - 1u;
            base += ((size_t)1 << bits) * EP;
        }
    }
    bool bad = false;
    MaxSum run = {-INFINITY, 0.0};
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t rnext = k < P ? ld_stream_u64(rows + k) : 0;
    for (; k < P; k += stride) {
        const uint64_t r = rnext;
        if (k + stride < P) rnext = ld_stream_u64(rows + k + stride);   // prefetch the next row
        double acc[EP];
        {
            const double2 *e = reinterpret_cast<const double2 *>(tb[0] + (size_t)((unsigned)(r >> shift[0]) & msk[0]) * EP);
#pragma unroll
            for (int j = 0; j < EP / 2; ++j) {
                const double2 v = __ldg(e + j);
                acc[2 * j] = v.x;
                acc[2 * j + 1] = v.y;
            }
        }
#pragma unroll
        for (int t = 1; t < 4; ++t) {
            if (t < nt) {
                const double2 *e =
                    reinterpret_cast<const double2 *>(tb[t] + (size_t)((unsigned)(r >> shift[t]) & msk[t]) * EP);
#pragma unroll
                for (int j = 0; j < EP / 2; ++j) {
                    const double2 v = __ldg(e + j);
                    acc[2 * j] += v.x;
                    acc[2 * j + 1] += v.y;
                }
            }
        }
        const double val = finish_tab<EP>(acc, inv_s2, C1, logtab, bad);
        st_stream_f64(out + k, val);
        if (STATS) ms_push(run, log_prior ? val + log_prior[k] : val, exptab);
    }
    if (bad && status) atomicOr(reinterpret_cast<unsigned int *>(status), MITRE_STATUS_BAD_SCHUR);
    if (STATS) {
        __shared__ double sm[256], ss[256];
        sm[threadIdx.x] = run.m; ss[threadIdx.x] = run.s;
        __syncthreads();
        for (int s = 128; s > 0; s >>= 1) {
            if (threadIdx.x < s) {
                const MaxSum x = ms_merge(MaxSum{sm[threadIdx.x], ss[threadIdx.x]},
                                          MaxSum{sm[threadIdx.x + s], ss[threadIdx.x + s]});
                sm[threadIdx.x] = x.m; ss[threadIdx.x] = x.s;
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) { stat_partials[2 * blockIdx.x] = sm[0]; stat_partials[2 * blockIdx.x + 1] = ss[0]; }
    }
}

// ---------------------------------------------------------------------------------------------
// N > 64: warp per candidate.  Lane l owns words l, l+32, ...; q is read through L1/L2
// (N*EP doubles, shared by every warp).  Fixed-order butterfly reduction => deterministic.
// ---------------------------------------------------------------------------------------------
template <int EP>
__global__ void __launch_bounds__(256)
score_wide_kernel(const uint64_t *__restrict__ rows, int64_t P, int W64, const double *__restrict__ ws,
                  int N, double *__restrict__ out, int32_t *status)
{
    const double *q = ws + kHdr;
    const double C0 = ws[0], inv_s2 = ws[1], s2 = ws[2];
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    bool bad = false;
    for (int64_t k = warp0; k < P; k += nwarps) {
        double acc[EP];
#pragma unroll
        for (int e = 0; e < EP; ++e) acc[e] = 0.0;
        for (int w = lane; w < W64; w += 32) {
            uint64_t bits = rows[k * W64 + w];
            while (bits) {
                const int b = __clzll((long long)bits);
                bits &= ~(1ull << (63 - b));
                const double2 *qi = reinterpret_cast<const double2 *>(q + ((size_t)w * 64 + b) * EP);
#pragma unroll
                for (int j = 0; j < EP / 2; ++j) {
                    const double2 v = __ldg(qi + j);
                    acc[2 * j] += v.x;
                    acc[2 * j + 1] += v.y;
                }
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
            for (int e = 0; e < EP; ++e) acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], off);
        }
        if (lane == 0) out[k] = finish_candidate<EP>(acc, C0, inv_s2, s2, bad);
    }
    if (bad && status) atomicOr(reinterpret_cast<unsigned int *>(status), MITRE_STATUS_BAD_SCHUR);
}

// ---------------------------------------------------------------------------------------------
// N > 64, byte tables: table b holds the 256 subset sums of subjects [8b, 8b + 8) (3 MB at
// N = 2000, p = 4: L2 resident, rebuilt per sweep).  A candidate is ceil(N/8) lookups instead of
// one q_i fetch + EP adds per SET BIT (~N/2 of them): ~4x fewer loads and adds.  A group of
// LPC = 2^k lanes owns one candidate (LPC = 32 from N = 249 up; smaller N put several candidates
// in a warp), lane j of the group takes bytes j, j + LPC, ...; fixed-order butterfly => deterministic.
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline size_t tables8_doubles(int N, int EP) { return (size_t)((N + 7) / 8) * 256 * EP; }

__global__ void __launch_bounds__(256)
score_build_tables8_kernel(const double *__restrict__ ws, int N, int EP, double *__restrict__ tables)
{
    const double *q = ws + kHdr;
    const size_t count = tables8_doubles(N, EP);
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < count;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int e = (int)(idx % EP);
        const unsigned pat = (unsigned)((idx / EP) & 255);
        const int b = (int)(idx / ((size_t)EP * 256));
        double acc = 0.0;
        for (int j = 0; j < 8; ++j) {
            const int subject = 8 * b + j;
            if (subject < N && ((pat >> (7 - j)) & 1)) acc += q[(size_t)subject * EP + e];
        }
        tables[idx] = acc;
    }
}

template <int EP>
__global__ void __launch_bounds__(256)
score_wide8_kernel(const uint64_t *__restrict__ rows, int64_t P, int W64, const double *__restrict__ ws,
                   int N, const double *__restrict__ tables, int lpc_log2, double *__restrict__ out,
                   int32_t *status)
{
    const double C0 = ws[0], inv_s2 = ws[1], s2 = ws[2];
    const int nbytes = (N + 7) >> 3;
    const int LPC = 1 << lpc_log2, cpw = 32 >> lpc_log2;
    const int lane = threadIdx.x & 31, sub = lane & (LPC - 1), slot = lane >> lpc_log2;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    bool bad = false;
    for (int64_t k0 = warp0 * cpw; k0 < P; k0 += nwarps * cpw) {      // warp-uniform trip count
        const int64_t k = k0 + slot;
        double acc[EP];
#pragma unroll
        for (int e = 0; e < EP; ++e) acc[e] = 0.0;
        if (k < P) {
            const uint64_t *row = rows + k * W64;
            for (int b = sub; b < nbytes; b += LPC) {
                const unsigned pat = (unsigned)(__ldg(row + (b >> 3)) >> (56 - 8 * (b & 7))) & 255u;
                if (pat) lut_add_g<EP>(acc, tables + ((size_t)b * 256 + pat) * EP);
            }
        }
        for (int off = LPC >> 1; off > 0; off >>= 1) {
#pragma unroll
            for (int e = 0; e < EP; ++e) acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], off);
        }
        if (sub == 0 && k < P) out[k] = finish_candidate<EP>(acc, C0, inv_s2, s2, bad);
    }
    if (bad && status) atomicOr(reinterpret_cast<unsigned int *>(status), MITRE_STATUS_BAD_SCHUR);
}

// ---------------------------------------------------------------------------------------------
// Group scoring: candidates in ENUMERATION order, straight from the detector values -- no truth
// rows at all.  Within one (window, variable, type) group the 2K candidates are nested sets along
// the sorted order of the group's N values ('above theta_k' = subjects whose rank exceeds k,
// 'below' = the rest; rules.py:465-511), so every masked sum is a suffix sum of the per-subject
// vectors q_i in rank order: O(N p) per GROUP instead of O(N p) (or N/8 table lookups) per
// candidate.  One CTA per group:
//   rank_i = #{k : theta_k < v_i} (binary search)  ->  bitonic sort of (rank, subject) keys in
//   shared memory  ->  chunked suffix scan of q over the sorted order  ->  at every rank boundary
//   finish the candidates above(k) = suffix, below(k) = total - suffix.
// Used by the streaming sweep (streaming.py), where tables exist only tile by tile.
// Deterministic: fixed sort, fixed scan order.
// ---------------------------------------------------------------------------------------------
template <int EP>
__global__ void __launch_bounds__(256)
score_groups_kernel(const double *__restrict__ values, const double *__restrict__ thresholds,
                    const int32_t *__restrict__ K, const int64_t *__restrict__ row_offsets, int64_t G, int N,
                    int NP2, const double *__restrict__ ws, const double *__restrict__ group_prior,
                    double *__restrict__ out, int32_t *status)
{
    extern __shared__ __align__(16) uint32_t gkeys[];          // [NP2] (rank << 16 | subject), sorted in place
    double *warp_tot = reinterpret_cast<double *>(gkeys + NP2);  // [8][EP]
    double *total = warp_tot + 8 * EP;                           // [EP]
    const double *q = ws + kHdr;
    const double C0 = ws[0], inv_s2 = ws[1], s2 = ws[2];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int chunk = (N + 255) >> 8;
    bool bad = false;
    for (int64_t g = blockIdx.x; g < G; g += gridDim.x) {
        const int k_g = K[g];
        const double *tg = thresholds + g * (N - 1);
        const double *vg = values + g * N;
        for (int i = tid; i < NP2; i += 256) {
            uint32_t key = 0xffffffffu;
            if (i < N) {
                const double v = vg[i];
                int lo = 0, hi = k_g;
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (tg[mid] < v) lo = mid + 1; else hi = mid;
                }
                key = ((uint32_t)lo << 16) | (uint32_t)i;
            }
            gkeys[i] = key;
        }
        __syncthreads();
        for (int size = 2; size <= NP2; size <<= 1) {
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                for (int t = tid; t < (NP2 >> 1); t += 256) {
                    const int lo_i = 2 * t - (t & (stride - 1));
                    const int hi_i = lo_i + stride;
                    const bool up = (lo_i & size) == 0;
                    const uint32_t a = gkeys[lo_i], b = gkeys[hi_i];
                    if ((a > b) == up) { gkeys[lo_i] = b; gkeys[hi_i] = a; }
                }
                __syncthreads();
            }
        }
        // chunked suffix scan over sorted positions [0, N): thread t owns [t * chunk, (t + 1) * chunk)
        const int j0 = tid * chunk, j1 = min(N, j0 + chunk);
        double loc[EP];
#pragma unroll
        for (int e = 0; e < EP; ++e) loc[e] = 0.0;
        for (int j = j1 - 1; j >= j0; --j) {
            const double *qi = q + (size_t)(gkeys[j] & 0xffffu) * EP;
#pragma unroll
            for (int e = 0; e < EP; ++e) loc[e] += qi[e];
        }
        // exclusive suffix over threads: sum of the chunks of all higher threads
        double suf[EP];
#pragma unroll
        for (int e = 0; e < EP; ++e) {
            double x = loc[e];
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const double y = __shfl_down_sync(0xffffffffu, x, off);
                if (lane + off < 32) x += y;
            }
            suf[e] = x - loc[e];                       // higher lanes of this warp
            if (lane == 0) warp_tot[wid * EP + e] = x;
        }
        __syncthreads();
#pragma unroll
        for (int e = 0; e < EP; ++e) {
            double hi_sum = 0.0;
            for (int w = 7; w > wid; --w) hi_sum += warp_tot[w * EP + e];
            suf[e] += hi_sum;
            if (tid == 0) total[e] = suf[e] + loc[e];
        }
        __syncthreads();
        double tot[EP];
#pragma unroll
        for (int e = 0; e < EP; ++e) tot[e] = total[e];
        double *og = out + row_offsets[g];
        const double gp = group_prior ? group_prior[g] : 0.0;
        for (int j = j1 - 1; j >= j0; --j) {
            const uint32_t key = gkeys[j];
            const double *qi = q + (size_t)(key & 0xffffu) * EP;
#pragma unroll
            for (int e = 0; e < EP; ++e) suf[e] += qi[e];          // suffix including position j
            if (j == 0) break;
            const int r = (int)(key >> 16), r_prev = (int)(gkeys[j - 1] >> 16);
            for (int k = r_prev; k < r; ++k) {                     // thresholds between the two ranks
                double below[EP];
#pragma unroll
                for (int e = 0; e < EP; ++e) below[e] = tot[e] - suf[e];
                og[2 * k] = finish_candidate<EP>(suf, C0, inv_s2, s2, bad) + gp;
                og[2 * k + 1] = finish_candidate<EP>(below, C0, inv_s2, s2, bad) + gp;
            }
        }
        __syncthreads();
    }
    if (bad && status) atomicOr(reinterpret_cast<unsigned int *>(status), MITRE_STATUS_BAD_SCHUR);
}

// Warp-per-group variant for N <= 64 (two subjects per lane): the same ranks -> sort -> suffix
// sums -> finish, with the sort a 64-key bitonic network over warp shuffles, the suffix sums two
// shuffle scans per component, q in shared memory and the table-driven log / reciprocal of the
// row kernels.  Sorted position j lives in lane j % 32, slot j / 32.
template <int EP>
__global__ void __launch_bounds__(256)
score_groups_warp_kernel(const double *__restrict__ values, const double *__restrict__ thresholds,
                         const int32_t *__restrict__ K, const int64_t *__restrict__ row_offsets, int64_t G, int N,
                         const double *__restrict__ ws, const double *__restrict__ group_prior,
                         double *__restrict__ out, int32_t *status)
{
    extern __shared__ __align__(16) double sq[];       // [N][EP]
    __shared__ double2 logtab[kLogTab];
    __shared__ double exptab[kExpTab];
    build_math_tables(logtab, exptab);
    for (int i = threadIdx.x; i < N * EP; i += blockDim.x) sq[i] = ws[kHdr + i];
    __syncthreads();
    const double inv_s2 = ws[1], C1 = ws[3];
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    bool bad = false;
    for (int64_t g = warp0; g < G; g += nwarps) {
        const int k_g = K[g];
        const double *tg = thresholds + g * (N - 1);
        const double *vg = values + g * N;
        // keys (rank << 6 | subject); padding sorts last
        uint32_t key[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int i = lane + 32 * h;
            key[h] = 0xffffffffu;
            if (i < N) {
                const double v = vg[i];
                int lo = 0, hi = k_g;
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (__ldg(tg + mid) < v) lo = mid + 1; else hi = mid;
                }
                key[h] = ((uint32_t)lo << 6) | (uint32_t)i;
            }
        }
        // bitonic sort of 64 keys, element index j = lane + 32 * slot
#pragma unroll
        for (int size = 2; size <= 64; size <<= 1) {
#pragma unroll
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                if (stride == 32) {
                    // partner is the other slot of the same lane; size == 64 here: ascending everywhere
                    const uint32_t a = min(key[0], key[1]), b = max(key[0], key[1]);
                    key[0] = a; key[1] = b;
                } else {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int j = lane + 32 * h;
                        const uint32_t other = __shfl_xor_sync(0xffffffffu, key[h], stride);
                        const bool up = (j & size) == 0;          // ascending block?
                        const bool lower = (lane & stride) == 0;   // this element is the lower index of the pair
                        const uint32_t mn = min(key[h], other), mx = max(key[h], other);
                        key[h] = (lower == up) ? mn : mx;
                    }
                }
            }
        }
        // suffix sums of q over sorted positions: slot 1 (positions 32..63) first, then slot 0
        double suf[2][EP], tot[EP];
#pragma unroll
        for (int h = 1; h >= 0; --h) {
            const uint32_t k32 = key[h];
            const bool live = k32 != 0xffffffffu;
            const double *qi = sq + (live ? (k32 & 63u) : 0u) * EP;
#pragma unroll
            for (int e = 0; e < EP; ++e) {
                double x = live ? qi[e] : 0.0;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const double y = __shfl_down_sync(0xffffffffu, x, off);
                    if (lane + off < 32) x += y;
                }
                if (h == 0) x += tot[e];                     // everything in slot 1 comes after slot 0
                suf[h][e] = x;
                tot[e] = __shfl_sync(0xffffffffu, x, 0);     // after h == 0: the group total
            }
        }
        // candidates at rank boundaries
        double *og = out + row_offsets[g];
        const double gp = group_prior ? group_prior[g] : 0.0;
        const int r0 = (int)(key[0] >> 6), r1 = (int)(key[1] >> 6);
        const int prev0 = __shfl_up_sync(0xffffffffu, r0, 1);            // rank at position lane - 1
        const int last0 = __shfl_sync(0xffffffffu, r0, 31);              // rank at position 31
        const int prev1s = __shfl_up_sync(0xffffffffu, r1, 1);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const bool live = key[h] != 0xffffffffu;
            const int j = lane + 32 * h;
            if (!live || j == 0) continue;
            const int r = h == 0 ? r0 : r1;
            const int rp = h == 0 ? prev0 : (lane == 0 ? last0 : prev1s);
            for (int k = rp; k < r; ++k) {
                double below[EP];
#pragma unroll
                for (int e = 0; e < EP; ++e) below[e] = tot[e] - suf[h][e];
                og[2 * k] = finish_tab<EP>(suf[h], inv_s2, C1, logtab, bad) + gp;
                og[2 * k + 1] = finish_tab<EP>(below, inv_s2, C1, logtab, bad) + gp;
            }
        }
    }
    if (bad && status) atomicOr(reinterpret_cast<unsigned int *>(status), MITRE_STATUS_BAD_SCHUR);
}

template <int EP>
static int launch_groups(cudaStream_t st, const double *values, const double *thresholds, const int32_t *K,
                         const int64_t *row_offsets, int64_t G, int N, const double *ws, const double *group_prior,
                         double *out, int32_t *status)
{
    if (N <= 64 && g_groups_warp) {
        auto wkern = score_groups_warp_kernel<EP>;
        const size_t wsmem = sizeof(double) * (size_t)N * EP;
        int per_sm = 1;
        MITRE_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, wkern, 256, wsmem));
        int64_t blocks = (int64_t)sm_count() * (per_sm > 0 ? per_sm : 1);
        const int64_t need = ceil_div64(G, 8);
        if (need < blocks) blocks = need;
        wkern<<<(unsigned)blocks, 256, wsmem, st>>>(values, thresholds, K, row_offsets, G, N, ws, group_prior, out,
                                                    status);
        MITRE_LAUNCH_CHECK();
        return MITRE_OK;
    }
    int NP2 = 2;
    while (NP2 < N) NP2 <<= 1;
    const size_t smem = sizeof(uint32_t) * (size_t)NP2 + sizeof(double) * 9 * EP;
    if (smem > smem_optin()) return MITRE_ERR_UNSUPPORTED;
    auto kern = score_groups_kernel<EP>;
    MITRE_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    MITRE_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, smem));
    int64_t blocks = (int64_t)sm_count() * (per_sm > 0 ? per_sm : 1);
    if (G < blocks) blocks = G;
    kern<<<(unsigned)blocks, 256, smem, st>>>(values, thresholds, K, row_offsets, G, N, NP2, ws, group_prior, out,
                                              status);
    MITRE_LAUNCH_CHECK();
    return MITRE_OK;
}

// ---------------------------------------------------------------------------------------------
// Batched dense likelihood: one CTA per design matrix.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
dense_batch_kernel(const double *__restrict__ Xs, const int32_t *__restrict__ p_b, int pmax,
                   const double *__restrict__ omega, const double *__restrict__ kappa, int N,
                   double sigma2, double *__restrict__ out, int32_t *status)
{
    __shared__ double sA[MITRE_MAX_P * MITRE_MAX_P];
    __shared__ double sb[MITRE_MAX_P];
    __shared__ double sred[128];
    __shared__ int s_ok;
    const int b = blockIdx.x;
    const int p = p_b ? p_b[b] : pmax;
    const double *X = Xs + (size_t)b * N * pmax;
    const double ll = base_system(X, pmax, omega, kappa, N, p, sigma2, sA, sb, sred, &s_ok);
    if (threadIdx.x == 0) {
        out[b] = s_ok ? ll : nan("");
        if (!s_ok && status) atomicOr(reinterpret_cast<unsigned int *>(status), MITRE_STATUS_BASE_NOT_PD);
    }
}

// ---------------------------------------------------------------------------------------------
// mc_logpdf_np for an arbitrary dense covariance: single-CTA right-looking Cholesky on a scratch
// copy, forward substitution, log-density.  (Compatibility entry; the model path never builds
// the N x N covariance.)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
mc_logpdf_kernel(const double *__restrict__ x, const double *__restrict__ cov, int N,
                 double *__restrict__ A, double *__restrict__ out, int32_t *status)
{
    __shared__ double sred[1024];
    __shared__ int s_ok;
    const int tid = threadIdx.x, nt = blockDim.x;
    for (size_t e = tid; e < (size_t)N * N; e += nt) A[e] = cov[e];
    if (tid == 0) s_ok = 1;
    __syncthreads();
    for (int j = 0; j < N; ++j) {
        const double d2 = A[(size_t)j * N + j];
        if (!(d2 > 0.0)) {
            if (tid == 0) s_ok = 0;
            break;  // uniform: every thread reads the same d2
        }
        const double d = sqrt(d2);
        __syncthreads();  // everyone has read the pivot before it is overwritten
        if (tid == 0) A[(size_t)j * N + j] = d;
        for (int i = j + 1 + tid; i < N; i += nt) A[(size_t)i * N + j] /= d;
        __syncthreads();
        // trailing update of the lower triangle: A[i][k] -= L[i][j] L[k][j], j < k <= i
        const int m = N - j - 1;
        for (int ib = tid >> 5; ib < m; ib += (nt >> 5)) {
            const double lij = A[(size_t)(j + 1 + ib) * N + j];
            for (int kb = tid & 31; kb <= ib; kb += 32)
                A[(size_t)(j + 1 + ib) * N + (j + 1 + kb)] -= lij * A[(size_t)(j + 1 + kb) * N + j];
        }
        __syncthreads();
    }
    __syncthreads();
    if (!s_ok) {
        if (tid == 0) {
            out[0] = nan("");
            if (status) atomicOr(reinterpret_cast<unsigned int *>(status), MITRE_STATUS_COV_NOT_PD);
        }
        return;
    }
    // y = L^-1 x by column sweeps; y is kept in the strict upper triangle's first row scratch:
    // reuse A's upper part row 0 is unsafe (N may be 1) -> use sred for N <= 1024, else global tail.
    double *y = (N <= 1024) ? sred : nullptr;
    if (y) {
        for (int i = tid; i < N; i += nt) y[i] = x[i];
        __syncthreads();
        for (int j = 0; j < N; ++j) {
            if (tid == 0) y[j] /= A[(size_t)j * N + j];
            __syncthreads();
            const double yj = y[j];
            for (int i = j + 1 + tid; i < N; i += nt) y[i] -= A[(size_t)i * N + j] * yj;
            __syncthreads();
        }
        if (tid == 0) {
            double yy = 0.0, ld = 0.0;
            for (int i = 0; i < N; ++i) {
                yy += y[i] * y[i];
                ld += log(A[(size_t)i * N + i]);
            }
            out[0] = -0.5 * yy - 0.5 * (N * log(2.0 * M_PI) + 2.0 * ld);
        }
    } else {
        // large N: store y in the (unused) strict upper triangle, row 0 .. via column N-1 of rows < N-1
        // y_i lives at A[i][N-1] for i < N-1 and in sred[0] for i = N-1.
        for (int i = tid; i < N; i += nt) {
            if (i < N - 1) A[(size_t)i * N + (N - 1)] = x[i];
            else sred[0] = x[i];
        }
        __syncthreads();
        for (int j = 0; j < N; ++j) {
            if (tid == 0) {
                if (j < N - 1) A[(size_t)j * N + (N - 1)] /= A[(size_t)j * N + j];
                else sred[0] /= A[(size_t)j * N + j];
            }
            __syncthreads();
            const double yj = (j < N - 1) ? A[(size_t)j * N + (N - 1)] : sred[0];
            for (int i = j + 1 + tid; i < N; i += nt) {
                if (i < N - 1) A[(size_t)i * N + (N - 1)] -= A[(size_t)i * N + j] * yj;
                else sred[0] -= A[(size_t)i * N + j] * yj;
            }
            __syncthreads();
        }
        if (tid == 0) {
            double yy = 0.0, ld = 0.0;
            for (int i = 0; i < N; ++i) {
                const double yi = (i < N - 1) ? A[(size_t)i * N + (N - 1)] : sred[0];
                yy += yi * yi;
                ld += log(A[(size_t)i * N + i]);
            }
            out[0] = -0.5 * yy - 0.5 * (N * log(2.0 * M_PI) + 2.0 * ld);
        }
    }
}

// ---- launch helpers --------------------------------------------------------------------------

template <int EP, int CB>
static int launch_w1(cudaStream_t st, const uint64_t *rows, int64_t P, double *ws, int N, double *out,
                     int32_t *status, const double *log_prior, double *stats)
{
    const int nchunks = (N + CB - 1) / CB;
    const size_t smem = (size_t)nchunks * (1 << CB) * EP * sizeof(double) + kLogTab * sizeof(double2) +
                        kExpTab * sizeof(double);
    int per_sm = 1;
    int64_t blocks = 0;
    double *partials = ws + kHdr + (size_t)N * EP;
#define MITRE_W1_LAUNCH(STATSV)                                                                          \
    {                                                                                                    \
        auto kern = score_w1_kernel<EP, CB, STATSV>;                                                     \
        MITRE_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        MITRE_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, smem));         \
        if (per_sm < 1) return MITRE_ERR_UNSUPPORTED;                                                    \
        blocks = (int64_t)sm_count() * per_sm;                                                           \
        const int64_t need = ceil_div64(P, 256);                                                         \
        if (need < blocks) blocks = need;                                                                \
        if (blocks < 1) blocks = 1;                                                                      \
        if (blocks > kMaxStatCtas) blocks = kMaxStatCtas;                                                \
        kern<<<(unsigned)blocks, 256, smem, st>>>(rows, P, ws, N, out, status, log_prior, partials);     \
        MITRE_LAUNCH_CHECK();                                                                            \
    }
    if (stats) {
        MITRE_W1_LAUNCH(true)
        score_stats_final_kernel<<<1, 256, 0, st>>>(partials, (int)blocks, stats);
        MITRE_LAUNCH_CHECK();
    } else {
        MITRE_W1_LAUNCH(false)
    }
#undef MITRE_W1_LAUNCH
    return MITRE_OK;
}

// Four consecutive candidates per thread over the 16-subject global tables: the lookups of every
// table but the last are shared by the four rows whenever they agree on those subjects (checked
// with one XOR/OR/shift; lexsorted neighbours nearly always do), so each thread fetches them once
// -- the register-fill bandwidth of the L1 (48 B per lookup per lane), which bounds
// score_w1g_kernel, drops from 3 to ~1.5 lookups per candidate.
template <int EP, bool STATS>
__global__ void __launch_bounds__(256)
score_w1g4_kernel(const uint64_t *__restrict__ rows, int64_t P, const double *__restrict__ ws, int N,
                  const double *__restrict__ tables, double *__restrict__ out, int32_t *status,
                  const double *__restrict__ log_prior, double *__restrict__ stat_partials)
{
    __shared__ double2 logtab[kLogTab];
    __shared__ double exptab[kExpTab];
    build_math_tables(logtab, exptab);
    __syncthreads();
    const double inv_s2 = ws[1], C1 = ws[3];
    const int nt = n_tables16(N);
    const double *tb[4];
    int shift[4];
    unsigned msk[4];
    {
        size_t base = 0;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int bits = t < nt ? table16_bits(N, t) : 0;
            tb[t] = tables + base;
            shift[t] = 64 - (t < nt ? table16_start(N, t) : 0) - bits;
            msk[t] = (1u << bits) - 1u;
            base += ((size_t)1 << bits) * EP;
        }
    }
    const int last_shift = shift[nt - 1];          // bits below the leading tables + padding
    const int lead_bits = 64 - last_shift - table16_bits(N, nt - 1);   // subjects covered by leading tables
    bool bad = false;
    MaxSum run = {-INFINITY, 0.0};
    const int64_t quads = P >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const ulonglong2 *rows2 = reinterpret_cast<const ulonglong2 *>(rows);
    double2 *out2 = reinterpret_cast<double2 *>(out);
    int64_t q4 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    ulonglong2 ra = make_ulonglong2(0, 0), rb = ra;
    if (q4 < quads) { ra = rows2[2 * q4]; rb = rows2[2 * q4 + 1]; }
    for (; q4 < quads; q4 += stride) {
        // software prefetch: the next quad's rows are in flight while this one is scored
        const int64_t nq = q4 + stride;
        ulonglong2 na = make_ulonglong2(0, 0), nb = na;
        if (nq < quads) { na = rows2[2 * nq]; nb = rows2[2 * nq + 1]; }
        const uint64_t r[4] = {ra.x, ra.y, rb.x, rb.y};
        ra = na; rb = nb;
        const uint64_t diff = (r[0] ^ r[1]) | (r[0] ^ r[2]) | (r[0] ^ r[3]);
        const bool share = lead_bits == 0 || (diff >> (64 - lead_bits)) == 0;
        double val[4];
        if (share) {
            double prefix[EP];
#pragma unroll
            for (int e = 0; e < EP; ++e) prefix[e] = 0.0;
#pragma unroll
            for (int t = 0; t < 3; ++t)
                if (t < nt - 1) lut_add_g<EP>(prefix, tb[t] + (size_t)((unsigned)(r[0] >> shift[t]) & msk[t]) * EP);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                double acc[EP];
#pragma unroll
                for (int e = 0; e < EP; ++e) acc[e] = prefix[e];
                lut_add_g<EP>(acc, tb[nt - 1] + (size_t)((unsigned)(r[i] >> last_shift) & msk[nt - 1]) * EP);
                val[i] = finish_tab<EP>(acc, inv_s2, C1, logtab, bad);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                double acc[EP];
#pragma unroll
                for (int e = 0; e < EP; ++e) acc[e] = 0.0;
#pragma unroll
                for (int t = 0; t < 4; ++t)
                    if (t < nt) lut_add_g<EP>(acc, tb[t] + (size_t)((unsigned)(r[i] >> shift[t]) & msk[t]) * EP);
                val[i] = finish_tab<EP>(acc, inv_s2, C1, logtab, bad);
            }
        }
        out2[2 * q4] = make_double2(val[0], val[1]);
        out2[2 * q4 + 1] = make_double2(val[2], val[3]);
        if (STATS) {
#pragma unroll
            for (int i = 0; i < 4; ++i) ms_push(run, log_prior ? val[i] + log_prior[4 * q4 + i] : val[i], exptab);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < (P & 3)) {      // tail rows
        const int64_t k = (quads << 2) + threadIdx.x;
        const uint64_t r = rows[k];
        double acc[EP];
#pragma unroll
        for (int e = 0; e < EP; ++e) acc[e] = 0.0;
#pragma unroll
        for (int t = 0; t < 4; ++t)
            if (t < nt) lut_add_g<EP>(acc, tb[t] + (size_t)((unsigned)(r >> shift[t]) & msk[t]) * EP);
        const double val = finish_tab<EP>(acc, inv_s2, C1, logtab, bad);
        out[k] = val;
        if (STATS) ms_push(run, log_prior ? val + log_prior[k] : val, exptab);
    }
    if (bad && status) atomicOr(reinterpret_cast<unsigned int *>(status), MITRE_STATUS_BAD_SCHUR);
    if (STATS) {
        __shared__ double sm[256], ss[256];
        sm[threadIdx.x] = run.m; ss[threadIdx.x] = run.s;
        __syncthreads();
        for (int s = 128; s > 0; s >>= 1) {
            if (threadIdx.x < s) {
                const MaxSum x = ms_merge(MaxSum{sm[threadIdx.x], ss[threadIdx.x]},
                                          MaxSum{sm[threadIdx.x + s], ss[threadIdx.x + s]});
                sm[threadIdx.x] = x.m; ss[threadIdx.x] = x.s;
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) { stat_partials[2 * blockIdx.x] = sm[0]; stat_partials[2 * blockIdx.x + 1] = ss[0]; }
    }
}

// ---------------------------------------------------------------------------------------------
// Duplicate-aware kernel (default for large lexsorted tables).  In a lexsorted candidate table
// equal rows are adjacent, and they are the rule, not the exception: a subject's abundance level
// persists over time, so the rank order of the subjects by a variable's window mean repeats from
// window to window and with it the truth rows (S5 shape: 81 % of the rows equal their predecessor,
// ~27 distinct rows per 128).  The reference exploits the same thing -- a candidate identical to
// its predecessor has divergence index N and costs nothing (efficient_likelihoods.pyx:70-101).
//
// A warp takes a tile of 128 consecutive rows (4 per lane, two 16-byte loads), flags the rows that
// differ from their predecessor, compacts those D distinct rows into shared memory with a warp
// prefix sum, scores them, and every lane reads back the values of its four rows.
//
// Scoring the distinct rows: what bounds the per-row kernels is the L1 wavefront rate of the
// divergent table lookups -- one lane fetching a 48-byte entry is three 16-byte loads, each a
// separate pass over up to 32 different cache lines (~2 cycles per line per instruction).  Here the
// entries are 64 bytes (one half line) and FOUR lanes fetch one entry with ONE load instruction:
// a warp instruction covers 8 rows and touches 8 lines instead of 32.  Four such sub-rounds load
// 32 rows; a 4x4 transpose over shuffles inside every group of four lanes then gives each lane all
// the components of one row, and the epilogue (Schur complement, table log / reciprocal) runs once
// for 32 rows.
//
// STATS: per tile (max, sum exp(w - max)) of w = loglik + log_prior -- max by warp reduction, one
// exp per candidate against the tile max, no running rescale -- written as the pair the draw
// consumes (draw.cu) and merged into the CTA's partial.  Deterministic for a given grid.
// ---------------------------------------------------------------------------------------------
static int g_dedupe = 1;
static int g_dedupe_bits = 11;      // subjects per subset-sum table of the duplicate-aware kernel (11..16)
constexpr int kCoopStride = 8;      // doubles per table entry (64 bytes)
constexpr int kPrefetchTiles = 3;   // L2 prefetch distance of the row stream, in tiles per warp

// 4x4 transpose of double2 elements inside each group of four lanes: a[r] of lane j <-> a[j] of lane r
__device__ __forceinline__ double2 shfl_xor_d2(double2 v, int m)
{
    return make_double2(__shfl_xor_sync(0xffffffffu, v.x, m), __shfl_xor_sync(0xffffffffu, v.y, m));
}
__device__ __forceinline__ void transpose4(double2 (&a)[4], int j)
{
    {
        const bool odd = j & 1;
        const double2 x = shfl_xor_d2(odd ? a[0] : a[1], 1), y = shfl_xor_d2(odd ? a[2] : a[3], 1);
        if (odd) { a[0] = x; a[2] = y; } else { a[1] = x; a[3] = y; }
    }
    {
        const bool hi = j & 2;
        const double2 x = shfl_xor_d2(hi ? a[0] : a[2], 2), y = shfl_xor_d2(hi ? a[1] : a[3], 2);
        if (hi) { a[0] = x; a[1] = y; } else { a[2] = x; a[3] = y; }
    }
}

template <int EP, bool STATS, bool COOP>
__global__ void __launch_bounds__(256, STATS ? 3 : 4)
score_w1d_kernel(const uint64_t *__restrict__ rows, int64_t P, const double *__restrict__ ws, int N,
                 const double *__restrict__ tables, int B, double *__restrict__ out, int32_t *status,
                 const double *__restrict__ log_prior, double *__restrict__ stat_partials,
                 double *__restrict__ tile_pairs, unsigned int *ticket, double *__restrict__ stats)
{
    static_assert(EP <= kCoopStride, "entries are 64 bytes");
    __shared__ double2 logtab[kLogTab];
    __shared__ double exptab[kExpTab];
    __shared__ __align__(16) unsigned long long sbuf[8][kDrawTileRows];
    build_math_tables(logtab, exptab);
    __syncthreads();
    const double inv_s2 = ws[1], C1 = ws[3];
    const int nt = n_tables(N, B);
    const double2 *tb[kMaxTables];
    int shift[kMaxTables];
    unsigned msk[kMaxTables];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int g = lane >> 2, j = lane & 3;      // group of four lanes, quarter of the entry
    {
        size_t base = 0;
#pragma unroll
        for (int t = 0; t < kMaxTables; ++t) {
            const int bits = t < nt ? table_bits(N, B, t) : 0;
            tb[t] = reinterpret_cast<const double2 *>(tables + base) + (COOP ? j : 0);
            shift[t] = 64 - (t < nt ? table_start(N, B, t) : 0) - bits;
            msk[t] = (1u << bits) - 1u;
            base += ((size_t)1 << bits) * (COOP ? kCoopStride : EP);
        }
    }
    const int last_shift = shift[nt - 1];
    const int lead_bits = 64 - last_shift - table_bits(N, B, nt - 1);   // subjects covered by the leading tables
    unsigned long long *sb = sbuf[wid];
    const int64_t ntiles = draw_ntiles(P);
    const int64_t nwarps = (int64_t)gridDim.x * 8;
    const ulonglong2 *rows2 = reinterpret_cast<const ulonglong2 *>(rows);
    bool bad = false;
    MaxSum run = {-INFINITY, 0.0};

    // this lane's four rows of a tile; past the end of the table the last row repeats (no new distinct rows)
    auto load_rows = [&](int64_t tile, uint64_t (&r)[4]) {
        const int64_t k0 = tile * kDrawTileRows + 4 * lane;
        if (k0 + 3 < P) {
            const ulonglong2 a = ld_stream_u64x2(rows2 + (k0 >> 1)), b = ld_stream_u64x2(rows2 + (k0 >> 1) + 1);
            r[0] = a.x; r[1] = a.y; r[2] = b.x; r[3] = b.y;
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) r[i] = rows[(k0 + i < P) ? (k0 + i) : (P - 1)];
        }
    };

    int64_t t = (int64_t)blockIdx.x * 8 + wid;
    uint64_t rn[4] = {0, 0, 0, 0};
    if (t < ntiles) load_rows(t, rn);
    for (; t < ntiles; t += nwarps) {
        const uint64_t r[4] = {rn[0], rn[1], rn[2], rn[3]};
        if (t + nwarps < ntiles) load_rows(t + nwarps, rn);     // next tile's rows in flight
        const int64_t k0 = t * kDrawTileRows + 4 * lane;
        {
            // ... and the tiles after it on their way from HBM to L2: with one tile of registers in flight per
            // warp the row loads were the kernel's long-scoreboard stall (HBM latency under load exceeds a
            // tile's processing time)
            const int64_t kp = k0 + (int64_t)kPrefetchTiles * nwarps * kDrawTileRows;
            if (kp + 3 < P) {
                prefetch_l2(rows + kp);
                if (STATS && log_prior) prefetch_l2(log_prior + kp);
            }
        }
        const bool full = k0 + 3 < P;
        const uint64_t prev = __shfl_up_sync(0xffffffffu, r[3], 1);
        const int h0 = (lane == 0) || (r[0] != prev);
        const int h1 = r[1] != r[0], h2 = r[2] != r[1], h3 = r[3] != r[2];
        const int c = h0 + h1 + h2 + h3;
        int incl = c;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += y;
        }
        const int D = __shfl_sync(0xffffffffu, incl, 31);
        const int i0 = incl - c + h0 - 1, i1 = i0 + h1, i2 = i1 + h2, i3 = i2 + h3;
        __syncwarp();                                   // the previous tile's read-back is over
        if (h0) sb[i0] = r[0];
        if (h1) sb[i1] = r[1];
        if (h2) sb[i2] = r[2];
        if (h3) sb[i3] = r[3];
        __syncwarp();
        if (!COOP) {
            // one lane per distinct row: the whole (packed, EP-double) entry of every table
            // (two rows per lane per pass for tiles with D > 32 was measured slower: 1.62 vs 1.44 ms)
            for (int jj = lane; jj < D; jj += 32) {
                const uint64_t rr = sb[jj];
                double acc[EP];
#pragma unroll
                for (int e = 0; e < EP; ++e) acc[e] = 0.0;
#pragma unroll
                for (int tt = 0; tt < kMaxTables; ++tt)
                    if (tt < nt)
                        lut_add_g<EP>(acc, reinterpret_cast<const double *>(tb[tt]) +
                                               (size_t)((unsigned)(rr >> shift[tt]) & msk[tt]) * EP);
                sb[jj] = (unsigned long long)__double_as_longlong(finish_tab<EP>(acc, inv_s2, C1, logtab, bad));
            }
        }
        if (COOP) {
            // leading tables: when every row of the tile agrees on the leading subjects (lexsorted neighbours
            // nearly always do) their entries are fetched once per tile, one quarter per lane
            const uint64_t first = __shfl_sync(0xffffffffu, r[0], 0);
            const uint64_t lead_mask = lead_bits ? (~0ull << (64 - lead_bits)) : 0ull;
            const bool uniform =
                __all_sync(0xffffffffu, ((((r[0] ^ first) | (r[1] ^ first)) | ((r[2] ^ first) | (r[3] ^ first))) & lead_mask) == 0);
            double2 lead = make_double2(0.0, 0.0);
            if (uniform) {
#pragma unroll
                for (int tt = 0; tt < kMaxTables - 1; ++tt) {
                    if (tt < nt - 1) {
                        const double2 v2 = __ldg(tb[tt] + (size_t)((unsigned)(first >> shift[tt]) & msk[tt]) * (kCoopStride / 2));
                        lead.x += v2.x;
                        lead.y += v2.y;
                    }
                }
            }
            for (int base = 0; base < D; base += 32) {
                double2 a[4];
#pragma unroll
                for (int sr = 0; sr < 4; ++sr) {
                    a[sr] = lead;
                    const int idx = base + 4 * g + sr;       // after the transpose lane L holds row base + L
                    const uint64_t rr = sb[idx < D ? idx : D - 1];
                    if (!uniform) {
#pragma unroll
                        for (int tt = 0; tt < kMaxTables - 1; ++tt) {
                            if (tt < nt - 1) {
                                const double2 v2 = __ldg(tb[tt] + (size_t)((unsigned)(rr >> shift[tt]) & msk[tt]) * (kCoopStride / 2));
                                a[sr].x += v2.x;
                                a[sr].y += v2.y;
                            }
                        }
                    }
                    const double2 v2 = __ldg(tb[nt - 1] + (size_t)((unsigned)(rr >> last_shift) & msk[nt - 1]) * (kCoopStride / 2));
                    a[sr].x += v2.x;
                    a[sr].y += v2.y;
                }
                transpose4(a, j);
                double acc[EP];
#pragma unroll
                for (int e = 0; e < EP; ++e) acc[e] = (e & 1) ? a[e >> 1].y : a[e >> 1].x;
                const double val = finish_tab<EP>(acc, inv_s2, C1, logtab, bad);
                __syncwarp();                               // every lane has read its rows of this batch
                if (base + lane < D) sb[base + lane] = (unsigned long long)__double_as_longlong(val);
            }
        }
        __syncwarp();
        double v[4];
        v[0] = __longlong_as_double((long long)sb[i0]);
        v[1] = __longlong_as_double((long long)sb[i1]);
        v[2] = __longlong_as_double((long long)sb[i2]);
        v[3] = __longlong_as_double((long long)sb[i3]);
        if (STATS) {
            double w[4] = {v[0], v[1], v[2], v[3]};
            if (full) {
                if (log_prior) {
                    const double2 *lp2 = reinterpret_cast<const double2 *>(log_prior);
                    const double2 a = ld_stream_f64x2(lp2 + (k0 >> 1)), b = ld_stream_f64x2(lp2 + (k0 >> 1) + 1);
                    w[0] += a.x; w[1] += a.y; w[2] += b.x; w[3] += b.y;
                }
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (k0 + i >= P) w[i] = -INFINITY;
                    else if (log_prior) w[i] += log_prior[k0 + i];
                }
            }
            double m = fmax(fmax(w[0], w[1]), fmax(w[2], w[3]));
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, off));
            double sum = 0.0;
            if (m > -INFINITY && m < INFINITY) {
#pragma unroll
                for (int i = 0; i < 4; ++i) sum += exp_nonpos(w[i] - m, exptab);
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
            if (lane == 0 && tile_pairs) reinterpret_cast<double2 *>(tile_pairs)[t] = make_double2(m, sum);
            run = ms_merge(run, MaxSum{m, sum});
        }
        if (full) {
            double2 *o2 = reinterpret_cast<double2 *>(out) + (k0 >> 1);
            st_stream_f64x2(o2, v[0], v[1]);
            st_stream_f64x2(o2 + 1, v[2], v[3]);
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (k0 + i < P) out[k0 + i] = v[i];
        }
    }
    if (bad && status) atomicOr(reinterpret_cast<unsigned int *>(status), MITRE_STATUS_BAD_SCHUR);
    if (STATS) {
        __shared__ double sm[256], ss[256];
        __shared__ int s_last;
        if (lane == 0) { sm[wid] = run.m; ss[wid] = run.s; }
        __syncthreads();
        if (threadIdx.x == 0) {
            MaxSum x = {sm[0], ss[0]};
            for (int w8 = 1; w8 < 8; ++w8) x = ms_merge(x, MaxSum{sm[w8], ss[w8]});
            stat_partials[2 * blockIdx.x] = x.m;
            stat_partials[2 * blockIdx.x + 1] = x.s;
            // the CTA that takes the last ticket merges all partials (in index order: the result does not
            // depend on which CTA that is) -- no separate reduction launch
            __threadfence();
            s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
        }
        __syncthreads();
        if (s_last) {
            __threadfence();
            MaxSum acc = {-INFINITY, 0.0};
            for (int i = threadIdx.x; i < (int)gridDim.x; i += 256)
                acc = ms_merge(acc, MaxSum{__ldcg(stat_partials + 2 * i), __ldcg(stat_partials + 2 * i + 1)});
            sm[threadIdx.x] = acc.m; ss[threadIdx.x] = acc.s;
            __syncthreads();
            for (int s2 = 128; s2 > 0; s2 >>= 1) {
                if (threadIdx.x < s2) {
                    const MaxSum x = ms_merge(MaxSum{sm[threadIdx.x], ss[threadIdx.x]},
                                              MaxSum{sm[threadIdx.x + s2], ss[threadIdx.x + s2]});
                    sm[threadIdx.x] = x.m; ss[threadIdx.x] = x.s;
                }
                __syncthreads();
            }
            if (threadIdx.x == 0) {
                stats[0] = sm[0];
                stats[1] = ss[0];
                *ticket = 0u;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Split tables: the LAST `LB` subjects' subset-sum table lives in shared memory, the leading
// N - LB subjects keep 16-subject tables in global memory.  score_w1g4_kernel fetches one 48-byte
// entry of a 3 MB table per candidate from L2 (divergent: one L1 tag per lane per 16 bytes and
// ~48 B of L2 traffic per candidate, the two units ncu shows saturated); here the per-candidate
// lookup is a shared-memory read and only the leading lookups -- shared by the four consecutive
// rows of a thread whenever they agree on the leading subjects, which lexsorted neighbours nearly
// always do -- go to L1/L2.
// Measured (S5, 2.94e8 candidates, p = 4): 1.94 ms against 1.97 ms for score_w1g4_kernel, and
// 2.46 ms against 2.33 ms with the fused stats: both variants sit on the same unit, the L1 data
// pipe's wavefront rate (87 % of peak in ncu) -- a 48-byte entry per lane costs ~25 wavefronts per
// warp from shared memory (bank conflicts of random 16-byte reads) where the divergent global
// read cost tags instead.  Kept selectable (mitre_set_split_tables), not the default.
// ---------------------------------------------------------------------------------------------
constexpr int kSplitThreads = 384;

__host__ __device__ inline int split_last_bits(int N, int EP)
{
    const int lb = (EP <= 6) ? 11 : 10;     // 2^LB * EP * 8 bytes <= ~98 KB: two blocks per SM
    return N < lb ? N : lb;
}
// leading tables: 16 subjects each from subject 0, the remainder (R mod 16) in the LAST leading
// table, so that rows which agree on everything but their trailing subjects disagree only there
__host__ __device__ inline int split_n_lead(int R) { return (R + 15) >> 4; }
__host__ __device__ inline int split_lead_bits(int R, int t) { return (16 * (t + 1) <= R) ? 16 : R - 16 * t; }
__host__ __device__ inline size_t split_tables_doubles(int N, int EP)
{
    const int LB = split_last_bits(N, EP), R = N - LB;
    size_t n = ((size_t)1 << LB) * EP;
    for (int t = 0; t < split_n_lead(R); ++t) n += ((size_t)1 << split_lead_bits(R, t)) * EP;
    return n;
}

// leading tables over subjects [16 t, 16 t + bits), then the last table over subjects [R, N)
__global__ void __launch_bounds__(256)
score_build_tables_split_kernel(const double *__restrict__ ws, int N, int EP, double *__restrict__ tables)
{
    const double *q = ws + kHdr;
    const int LB = split_last_bits(N, EP), R = N - LB;
    const int nt = split_n_lead(R);
    size_t base = 0;
    for (int t = 0; t <= nt; ++t) {
        const int bits = t < nt ? split_lead_bits(R, t) : LB;
        const int start = t < nt ? 16 * t : R;
        const size_t count = ((size_t)1 << bits) * EP;
        for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < count;
             idx += (size_t)gridDim.x * blockDim.x) {
            const int e = (int)(idx % EP);
            const unsigned pat = (unsigned)(idx / EP);
            double acc = 0.0;
            for (int j = 0; j < bits; ++j)
                if ((pat >> (bits - 1 - j)) & 1) acc += q[(size_t)(start + j) * EP + e];
            tables[base + idx] = acc;
        }
        base += count;
    }
}

template <int EP, bool STATS>
__global__ void __launch_bounds__(kSplitThreads, 2)
score_w1s4_kernel(const uint64_t *__restrict__ rows, int64_t P, const double *__restrict__ ws, int N,
                  const double *__restrict__ tables, double *__restrict__ out, int32_t *status,
                  const double *__restrict__ log_prior, double *__restrict__ stat_partials)
{
    extern __shared__ __align__(16) double last_tab[];    // [2^LB][EP]
    __shared__ double2 logtab[kLogTab];
    __shared__ double exptab[kExpTab];
    build_math_tables(logtab, exptab);
    const int LB = split_last_bits(N, EP), R = N - LB;
    const int nt = split_n_lead(R);
    const double inv_s2 = ws[1], C1 = ws[3];
    const double *tb[4];
    int shift[4];
    unsigned msk[4];
    size_t base = 0;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int bits = t < nt ? split_lead_bits(R, t) : 0;
        tb[t] = tables + base;
        shift[t] = 64 - 16 * t - bits;
        msk[t] = (1u << bits) - 1u;
        base += t < nt ? ((size_t)1 << bits) * EP : 0;
    }
    // subjects covered by all leading tables but the last one: the second sharing level
    const int R1 = nt > 1 ? 16 * (nt - 1) : 0;
    {
        const double2 *src = reinterpret_cast<const double2 *>(tables + base);
        double2 *dst = reinterpret_cast<double2 *>(last_tab);
        const int n2 = (EP << LB) >> 1;
        for (int i = threadIdx.x; i < n2; i += blockDim.x) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    const int last_shift = 64 - N;
    const unsigned last_msk = (1u << LB) - 1u;
    bool bad = false;
    MaxSum run = {-INFINITY, 0.0};
    const int64_t quads = P >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const ulonglong2 *rows2 = reinterpret_cast<const ulonglong2 *>(rows);
    double2 *out2 = reinterpret_cast<double2 *>(out);
    int64_t q4 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    ulonglong2 ra = make_ulonglong2(0, 0), rb = ra;
    if (q4 < quads) { ra = rows2[2 * q4]; rb = rows2[2 * q4 + 1]; }
    for (; q4 < quads; q4 += stride) {
        const int64_t nq = q4 + stride;
        ulonglong2 na = make_ulonglong2(0, 0), nb = na;
        if (nq < quads) { na = rows2[2 * nq]; nb = rows2[2 * nq + 1]; }
        const uint64_t r[4] = {ra.x, ra.y, rb.x, rb.y};
        ra = na; rb = nb;
        const uint64_t diff = (r[0] ^ r[1]) | (r[0] ^ r[2]) | (r[0] ^ r[3]);
        // how many leading tables the four rows share: all of them (lexsorted neighbours nearly
        // always), all but the last (short) one, or none
        const int n_shared = (R == 0 || (diff >> (64 - R)) == 0) ? nt
                             : (R1 > 0 && (diff >> (64 - R1)) == 0) ? nt - 1 : 0;
        double val[4];
        double prefix[EP];
#pragma unroll
        for (int e = 0; e < EP; ++e) prefix[e] = 0.0;
#pragma unroll
        for (int t = 0; t < 4; ++t)
            if (t < n_shared) lut_add_g<EP>(prefix, tb[t] + (size_t)((unsigned)(r[0] >> shift[t]) & msk[t]) * EP);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            double acc[EP];
#pragma unroll
            for (int e = 0; e < EP; ++e) acc[e] = prefix[e];
            if (n_shared < nt) {
#pragma unroll
                for (int t = 0; t < 4; ++t)
                    if (t >= n_shared && t < nt)
                        lut_add_g<EP>(acc, tb[t] + (size_t)((unsigned)(r[i] >> shift[t]) & msk[t]) * EP);
            }
            lut_add<EP>(acc, last_tab + ((unsigned)(r[i] >> last_shift) & last_msk) * EP);
            val[i] = finish_tab<EP>(acc, inv_s2, C1, logtab, bad);
        }
        out2[2 * q4] = make_double2(val[0], val[1]);
        out2[2 * q4 + 1] = make_double2(val[2], val[3]);
        if (STATS) {
#pragma unroll
            for (int i = 0; i < 4; ++i) ms_push(run, log_prior ? val[i] + log_prior[4 * q4 + i] : val[i], exptab);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < (P & 3)) {      // tail rows
        const int64_t k = (quads << 2) + threadIdx.x;
        const uint64_t r = rows[k];
        double acc[EP];
#pragma unroll
        for (int e = 0; e < EP; ++e) acc[e] = 0.0;
#pragma unroll
        for (int t = 0; t < 4; ++t)
            if (t < nt) lut_add_g<EP>(acc, tb[t] + (size_t)((unsigned)(r >> shift[t]) & msk[t]) * EP);
        lut_add<EP>(acc, last_tab + ((unsigned)(r >> last_shift) & last_msk) * EP);
        const double val = finish_tab<EP>(acc, inv_s2, C1, logtab, bad);
        out[k] = val;
        if (STATS) ms_push(run, log_prior ? val + log_prior[k] : val, exptab);
    }
    if (bad && status) atomicOr(reinterpret_cast<unsigned int *>(status), MITRE_STATUS_BAD_SCHUR);
    if (STATS) {
        __shared__ double sm[kSplitThreads], ss[kSplitThreads];
        sm[threadIdx.x] = run.m; ss[threadIdx.x] = run.s;
        __syncthreads();
        for (int s = 256; s > 0; s >>= 1) {
            if (threadIdx.x < s && threadIdx.x + s < kSplitThreads) {
                const MaxSum x = ms_merge(MaxSum{sm[threadIdx.x], ss[threadIdx.x]},
                                          MaxSum{sm[threadIdx.x + s], ss[threadIdx.x + s]});
                sm[threadIdx.x] = x.m; ss[threadIdx.x] = x.s;
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) { stat_partials[2 * blockIdx.x] = sm[0]; stat_partials[2 * blockIdx.x + 1] = ss[0]; }
    }
}

template <int EP>
static int launch_w1s4(cudaStream_t st, const uint64_t *rows, int64_t P, double *ws, int N, double *out,
                       int32_t *status, const double *log_prior, double *stats)
{
    double *partials = ws + kHdr + (size_t)N * EP;
    double *tables = partials + 2 * kMaxStatCtas;
    const size_t entries = split_tables_doubles(N, EP);
    int64_t bblocks = ceil_div64((int64_t)entries, 256);
    if (bblocks > 8 * sm_count()) bblocks = 8 * sm_count();
    score_build_tables_split_kernel<<<(unsigned)bblocks, 256, 0, st>>>(ws, N, EP, tables);
    MITRE_LAUNCH_CHECK();
    const size_t smem = (sizeof(double) * EP) << split_last_bits(N, EP);
    int per_sm = 1;
    int64_t blocks = 0;
#define MITRE_S4_LAUNCH(STATSV)                                                                          \
    {                                                                                                    \
        auto kern = score_w1s4_kernel<EP, STATSV>;                                                       \
        MITRE_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        MITRE_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kSplitThreads, smem)); \
        if (per_sm < 1) return MITRE_ERR_UNSUPPORTED;                                                    \
        blocks = (int64_t)sm_count() * per_sm;                                                           \
        const int64_t need = ceil_div64(P >> 2, kSplitThreads);                                          \
        if (need < blocks) blocks = need;                                                                \
        if (blocks < 1) blocks = 1;                                                                      \
        if (blocks > kMaxStatCtas) blocks = kMaxStatCtas;                                                \
        kern<<<(unsigned)blocks, kSplitThreads, smem, st>>>(rows, P, ws, N, tables, out, status, log_prior, \
                                                            partials);                                   \
        MITRE_LAUNCH_CHECK();                                                                            \
    }
    if (stats) {
        MITRE_S4_LAUNCH(true)
        score_stats_final_kernel<<<1, 256, 0, st>>>(partials, (int)blocks, stats);
        MITRE_LAUNCH_CHECK();
    } else {
        MITRE_S4_LAUNCH(false)
    }
#undef MITRE_S4_LAUNCH
    return MITRE_OK;
}

template <int EP>
static int launch_w1r4(cudaStream_t st, const uint64_t *rows, int64_t P, double *ws, int N, double *out,
                       int32_t *status, const double *log_prior, double *stats)
{
    const int nchunks = (N + 7) / 8;
    const size_t smem = (size_t)nchunks * 256 * EP * sizeof(double) + kLogTab * sizeof(double2) +
                        kExpTab * sizeof(double);
    int per_sm = 1;
    int64_t blocks = 0;
    double *partials = ws + kHdr + (size_t)N * EP;
#define MITRE_R4_LAUNCH(STATSV)                                                                          \
    {                                                                                                    \
        auto kern = score_w1r4_kernel<EP, STATSV>;                                                       \
        MITRE_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        MITRE_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, smem));         \
        if (per_sm < 1) return MITRE_ERR_UNSUPPORTED;                                                    \
        blocks = (int64_t)sm_count() * per_sm;                                                           \
        const int64_t need = ceil_div64(P >> 2, 256);                                                    \
        if (need < blocks) blocks = need;                                                                \
        if (blocks < 1) blocks = 1;                                                                      \
        if (blocks > kMaxStatCtas) blocks = kMaxStatCtas;                                                \
        kern<<<(unsigned)blocks, 256, smem, st>>>(rows, P, ws, N, out, status, log_prior, partials);     \
        MITRE_LAUNCH_CHECK();                                                                            \
    }
    if (stats) {
        MITRE_R4_LAUNCH(true)
        score_stats_final_kernel<<<1, 256, 0, st>>>(partials, (int)blocks, stats);
        MITRE_LAUNCH_CHECK();
    } else {
        MITRE_R4_LAUNCH(false)
    }
#undef MITRE_R4_LAUNCH
    return MITRE_OK;
}

// The duplicate-aware kernel: 4 rows per lane as two 16-byte loads, entries of at most 8 doubles.
static bool dedupe_eligible(int EP, int64_t P, const uint64_t *rows, const double *out, const double *log_prior,
                            const double *tile_pairs)
{
    return g_dedupe && EP <= 8 && P >= 1024 &&
           ((reinterpret_cast<uintptr_t>(rows) | reinterpret_cast<uintptr_t>(out)) & 31) == 0 &&
           ((reinterpret_cast<uintptr_t>(log_prior) | reinterpret_cast<uintptr_t>(tile_pairs)) & 15) == 0;
}
static double *tables16_ptr(double *ws, int N, int EP) { return ws + kHdr + (size_t)N * EP + 2 * kMaxStatCtas; }

template <int EP>
static int launch_w1g(cudaStream_t st, const uint64_t *rows, int64_t P, double *ws, int N, double *out,
                      int32_t *status, const double *log_prior, double *stats, double *tile_pairs, bool *pairs_done)
{
    double *partials = ws + kHdr + (size_t)N * EP;
    double *tables = partials + 2 * kMaxStatCtas;
    const bool quad = EP <= 8 && P >= 1024 &&
                      ((reinterpret_cast<uintptr_t>(rows) | reinterpret_cast<uintptr_t>(out)) & 31) == 0;
    const bool dedupe = dedupe_eligible(EP, P, rows, out, log_prior, tile_pairs);
    const bool coop = dedupe && g_dedupe == 2;
    const int stride = coop ? kCoopStride : EP;
    if (!dedupe) {      // the duplicate-aware path got its tables from score_prepare16_kernel
        const size_t entries = tables16_doubles(N, stride);
        int64_t bblocks = ceil_div64((int64_t)entries, 256);
        if (bblocks > 8 * sm_count()) bblocks = 8 * sm_count();
        score_build_tables16_kernel<<<(unsigned)bblocks, 256, 0, st>>>(ws, N, EP, stride, tables);
        MITRE_LAUNCH_CHECK();
    }
    int per_sm = 1;
    int64_t blocks = 0;
    if (dedupe) {
#define MITRE_D_LAUNCH(STATSV)                                                                           \
    {                                                                                                    \
        auto kern = coop ? score_w1d_kernel<EP <= 8 ? EP : 8, STATSV, true>                              \
                         : score_w1d_kernel<EP <= 8 ? EP : 8, STATSV, false>;                            \
        MITRE_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, 0));            \
        if (per_sm < 1) return MITRE_ERR_UNSUPPORTED;                                                    \
        blocks = (int64_t)sm_count() * per_sm;                                                           \
        const int64_t need = ceil_div64(draw_ntiles(P), 8);                                              \
        if (need < blocks) blocks = need;                                                                \
        if (blocks > kMaxStatCtas) blocks = kMaxStatCtas;                                                \
        kern<<<(unsigned)blocks, 256, 0, st>>>(rows, P, ws, N, tables, g_dedupe_bits, out, status, log_prior, \
                                               partials, tile_pairs,                                      \
                                               reinterpret_cast<unsigned int *>(ws + kTicketSlot), stats); \
        MITRE_LAUNCH_CHECK();                                                                            \
    }
        if (stats) {
            MITRE_D_LAUNCH(true)
            *pairs_done = tile_pairs != nullptr;
        } else {
            MITRE_D_LAUNCH(false)
        }
#undef MITRE_D_LAUNCH
        return MITRE_OK;
    }
#define MITRE_G_LAUNCH(STATSV)                                                                           \
    {                                                                                                    \
        auto kern = quad ? score_w1g4_kernel<EP <= 8 ? EP : 8, STATSV> : score_w1g_kernel<EP, STATSV>;   \
        MITRE_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, 0));            \
        if (per_sm < 1) return MITRE_ERR_UNSUPPORTED;                                                    \
        blocks = (int64_t)sm_count() * per_sm;                                                           \
        const int64_t need = ceil_div64(quad ? (P >> 2) : P, 256);                                       \
        if (need < blocks) blocks = need;                                                                \
        if (blocks > kMaxStatCtas) blocks = kMaxStatCtas;                                                \
        kern<<<(unsigned)blocks, 256, 0, st>>>(rows, P, ws, N, tables, out, status, log_prior, partials); \
        MITRE_LAUNCH_CHECK();                                                                            \
    }
    if (stats) {
        MITRE_G_LAUNCH(true)
        score_stats_final_kernel<<<1, 256, 0, st>>>(partials, (int)blocks, stats);
        MITRE_LAUNCH_CHECK();
    } else {
        MITRE_G_LAUNCH(false)
    }
#undef MITRE_G_LAUNCH
    return MITRE_OK;
}

template <int EP>
static int launch_wide(cudaStream_t st, const uint64_t *rows, int64_t P, int W64, double *ws,
                       int N, double *out, int32_t *status)
{
    int per_sm = 1;
    if (g_wide_tables) {
        double *tables = ws + kHdr + (size_t)N * EP + 2 * kMaxStatCtas;
        int64_t bblocks = ceil_div64((int64_t)tables8_doubles(N, EP), 256);
        if (bblocks > 8 * sm_count()) bblocks = 8 * sm_count();
        score_build_tables8_kernel<<<(unsigned)bblocks, 256, 0, st>>>(ws, N, EP, tables);
        MITRE_LAUNCH_CHECK();
        const int nbytes = (N + 7) / 8;
        int lpc_log2 = 0;
        while (lpc_log2 < 5 && (1 << lpc_log2) < nbytes) ++lpc_log2;
        auto kern = score_wide8_kernel<EP>;
        MITRE_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, 0));
        int64_t blocks = (int64_t)sm_count() * (per_sm > 0 ? per_sm : 1);
        const int64_t need = ceil_div64(P, 8 * (32 >> lpc_log2));
        if (need < blocks) blocks = need;
        if (blocks < 1) blocks = 1;
        kern<<<(unsigned)blocks, 256, 0, st>>>(rows, P, W64, ws, N, tables, lpc_log2, out, status);
        MITRE_LAUNCH_CHECK();
        return MITRE_OK;
    }
    auto kern = score_wide_kernel<EP>;
    MITRE_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, 0));
    int64_t blocks = (int64_t)sm_count() * (per_sm > 0 ? per_sm : 1);
    const int64_t need = ceil_div64(P, 8);
    if (need < blocks) blocks = need;
    if (blocks < 1) blocks = 1;
    kern<<<(unsigned)blocks, 256, 0, st>>>(rows, P, W64, ws, N, out, status);
    MITRE_LAUNCH_CHECK();
    return MITRE_OK;
}

template <int EP>
static int launch_ep(cudaStream_t st, const uint64_t *rows, int64_t P, int W64, double *ws, int N, double *out,
                     int32_t *status, const double *log_prior, double *stats, bool *stats_done, double *tile_pairs,
                     bool *pairs_done)
{
    if (W64 == 1) {
        *stats_done = true;
        const size_t lut8 = (size_t)((N + 7) / 8) * 256 * EP * sizeof(double);
        // keep >= 2 CTAs per SM where the byte table allows it, else fall to nibble tables
        if (P >= g_large_table_min_rows) {
            const bool aligned = ((reinterpret_cast<uintptr_t>(rows) | reinterpret_cast<uintptr_t>(out)) & 31) == 0;
            if (g_split_tables && EP <= 8 && aligned)
                return launch_w1s4<(EP <= 8 ? EP : 8)>(st, rows, P, ws, N, out, status, log_prior, stats);
            return launch_w1g<EP>(st, rows, P, ws, N, out, status, log_prior, stats, tile_pairs, pairs_done);
        }
        const bool aligned32 = ((reinterpret_cast<uintptr_t>(rows) | reinterpret_cast<uintptr_t>(out)) & 31) == 0;
        if (EP <= 8 && lut8 <= 100 * 1024 && aligned32 && P >= 64)
            return launch_w1r4<EP>(st, rows, P, ws, N, out, status, log_prior, stats);
        if (lut8 <= 100 * 1024) return launch_w1<EP, 8>(st, rows, P, ws, N, out, status, log_prior, stats);
        return launch_w1<EP, 4>(st, rows, P, ws, N, out, status, log_prior, stats);
    }
    *stats_done = false;
    return launch_wide<EP>(st, rows, P, W64, ws, N, out, status);
}

}  // namespace mitre

using namespace mitre;

extern "C" size_t mitre_score_workspace_bytes(int N, int p)
{
    if (N <= 0 || p <= 0 || p > MITRE_MAX_P) return 0;
    size_t n = (size_t)kHdr + (size_t)N * ep_for(p) + 2 * (size_t)kMaxStatCtas;
    if (N <= 64) {   // tables of the large-table kernels (any layout)
        const int stride = ep_for(p) > kCoopStride ? ep_for(p) : kCoopStride;
        const size_t a = tables16_doubles(N, stride), b = split_tables_doubles(N, ep_for(p));
        n += a > b ? a : b;
    } else {
        n += tables8_doubles(N, ep_for(p));                // byte tables of score_wide8_kernel
    }
    return sizeof(double) * n;
}

// Test / tuning hook: the candidate count from which the large-table kernel is used (default 4 Mi).
extern "C" int mitre_set_large_table_min_rows(int64_t rows)
{
    if (rows < 0) return MITRE_ERR_INVALID;
    g_large_table_min_rows = rows;
    return MITRE_OK;
}

extern "C" int mitre_set_split_tables(int enable)
{
    g_split_tables = enable ? 1 : 0;
    return MITRE_OK;
}

extern "C" int mitre_set_groups_warp(int enable)
{
    g_groups_warp = enable ? 1 : 0;
    return MITRE_OK;
}

extern "C" int mitre_set_wide_tables(int enable)
{
    g_wide_tables = enable ? 1 : 0;
    return MITRE_OK;
}

extern "C" int mitre_weight_stats(void *stream, const double *a, const double *b, int64_t P, double *stats,
                                  void *workspace, size_t workspace_bytes);
extern "C" size_t mitre_draw_workspace_bytes(int64_t P);
namespace mitre {
int launch_tile_pairs(cudaStream_t st, const double *a, const double *b, int64_t P, double *pairs);   // draw.cu
}

extern "C" int mitre_set_dedupe_table_bits(int bits)
{
    if (bits < 11 || bits > 16) return MITRE_ERR_INVALID;
    g_dedupe_bits = bits;
    return MITRE_OK;
}

extern "C" int mitre_set_dedupe(int mode)
{
    if (mode < 0 || mode > 2) return MITRE_ERR_INVALID;
    g_dedupe = mode;
    return MITRE_OK;
}

extern "C" int mitre_score_all_stats(void *stream, const uint64_t *packed_truth, int64_t P, int W64,
                                     const uint64_t *mask, const double *X, const double *omega,
                                     const double *kappa, int N, int p, double sigma2, void *workspace,
                                     size_t workspace_bytes, double *out, int32_t *status,
                                     const double *log_prior, double *stats, void *stats_workspace,
                                     size_t stats_workspace_bytes)
{
    if (P < 0 || N <= 0 || p <= 0 || !X || !omega || !kappa || !workspace) return MITRE_ERR_INVALID;
    if (P > 0 && (!packed_truth || !out)) return MITRE_ERR_INVALID;
    if (W64 != words_for(N)) return MITRE_ERR_INVALID;
    if (p > MITRE_MAX_P) return MITRE_ERR_UNSUPPORTED;
    if (!(sigma2 > 0.0)) return MITRE_ERR_INVALID;
    if (workspace_bytes < mitre_score_workspace_bytes(N, p)) return MITRE_ERR_WORKSPACE;
    if (log_prior && !stats) return MITRE_ERR_INVALID;
    if (stats_workspace && stats_workspace_bytes < mitre_draw_workspace_bytes(P)) return MITRE_ERR_WORKSPACE;
    cudaStream_t st = as_stream(stream);
    double *ws = static_cast<double *>(workspace);
    // per-tile (max, sum-exp) pairs for mitre_draw_from_tiles, when the caller passes a draw workspace
    double *tile_pairs = (stats && stats_workspace) ? static_cast<double *>(stats_workspace) + draw_off_pairs() : nullptr;
    const int EPr = ep_for(p);
    if (P > 0 && W64 == 1 && P >= g_large_table_min_rows && !g_split_tables &&
        dedupe_eligible(EPr, P, packed_truth, out, log_prior, tile_pairs)) {
        // base system + tables in one launch
        const int stride = g_dedupe == 2 ? kCoopStride : EPr;
        const size_t patterns = tables_doubles(N, g_dedupe_bits, 1);
        int64_t pblocks = ceil_div64((int64_t)patterns, 256);
        if (pblocks > 2 * sm_count()) pblocks = 2 * sm_count();
        score_prepare16_kernel<<<(unsigned)pblocks, 256, sizeof(double) * ((size_t)N * (EPr + p + 2)), st>>>(
            X, omega, kappa, mask, N, p, sigma2, ws, status, stride, g_dedupe_bits, tables16_ptr(ws, N, EPr));
    } else {
        score_setup_kernel<<<1, 256, 0, st>>>(X, omega, kappa, mask, N, p, sigma2, ws, status);
    }
    MITRE_LAUNCH_CHECK();
    if (P == 0) return MITRE_OK;
    bool stats_done = false, pairs_done = false;
    int rc = MITRE_ERR_UNSUPPORTED;
    switch (ep_for(p)) {
#define MITRE_EP_CASE(E)                                                                                  \
    case E:                                                                                               \
        rc = launch_ep<E>(st, packed_truth, P, W64, ws, N, out, status, log_prior, stats, &stats_done,    \
                          tile_pairs, &pairs_done);                                                       \
        break;
        MITRE_EP_CASE(4)
        MITRE_EP_CASE(6)
        MITRE_EP_CASE(8)
        MITRE_EP_CASE(10)
        MITRE_EP_CASE(12)
        MITRE_EP_CASE(14)
        MITRE_EP_CASE(16)
        MITRE_EP_CASE(18)
        MITRE_EP_CASE(20)
        MITRE_EP_CASE(22)
        MITRE_EP_CASE(24)
        MITRE_EP_CASE(26)
#undef MITRE_EP_CASE
    }
    if (rc != MITRE_OK) return rc;
    if (stats && !stats_done) {
        // wide path: separate reduction over the scored slice
        if (!stats_workspace) return MITRE_ERR_WORKSPACE;
        rc = mitre_weight_stats(stream, out, log_prior, P, stats, stats_workspace, stats_workspace_bytes);
        if (rc != MITRE_OK) return rc;
    }
    if (tile_pairs && !pairs_done) return launch_tile_pairs(st, out, log_prior, P, tile_pairs);
    return MITRE_OK;
}

extern "C" int mitre_score_groups(void *stream, const double *values, const double *thresholds,
                                  const int32_t *K, const int64_t *row_offsets, int64_t G, const uint64_t *mask,
                                  const double *X, const double *omega, const double *kappa, int N, int p,
                                  double sigma2, void *workspace, size_t workspace_bytes,
                                  const double *group_prior, double *out, int32_t *status)
{
    if (G < 0 || N <= 1 || p <= 0 || !X || !omega || !kappa || !workspace) return MITRE_ERR_INVALID;
    if (G > 0 && (!values || !thresholds || !K || !row_offsets || !out)) return MITRE_ERR_INVALID;
    if (p > MITRE_MAX_P || N > 65535) return MITRE_ERR_UNSUPPORTED;
    if (!(sigma2 > 0.0)) return MITRE_ERR_INVALID;
    if (workspace_bytes < mitre_score_workspace_bytes(N, p)) return MITRE_ERR_WORKSPACE;
    cudaStream_t st = as_stream(stream);
    double *ws = static_cast<double *>(workspace);
    score_setup_kernel<<<1, 256, 0, st>>>(X, omega, kappa, mask, N, p, sigma2, ws, status);
    MITRE_LAUNCH_CHECK();
    if (G == 0) return MITRE_OK;
    switch (ep_for(p)) {
#define MITRE_EP_CASE(E) \
    case E: return launch_groups<E>(st, values, thresholds, K, row_offsets, G, N, ws, group_prior, out, status);
        MITRE_EP_CASE(4)
        MITRE_EP_CASE(6)
        MITRE_EP_CASE(8)
        MITRE_EP_CASE(10)
        MITRE_EP_CASE(12)
        MITRE_EP_CASE(14)
        MITRE_EP_CASE(16)
        MITRE_EP_CASE(18)
        MITRE_EP_CASE(20)
        MITRE_EP_CASE(22)
        MITRE_EP_CASE(24)
        MITRE_EP_CASE(26)
#undef MITRE_EP_CASE
    }
    return MITRE_ERR_UNSUPPORTED;
}

extern "C" int mitre_score_all(void *stream, const uint64_t *packed_truth, int64_t P, int W64,
                               const uint64_t *mask, const double *X, const double *omega,
                               const double *kappa, int N, int p, double sigma2, void *workspace,
                               size_t workspace_bytes, double *out, int32_t *status)
{
    return mitre_score_all_stats(stream, packed_truth, P, W64, mask, X, omega, kappa, N, p, sigma2, workspace,
                                 workspace_bytes, out, status, nullptr, nullptr, nullptr, 0);
}

extern "C" int mitre_likelihood_z_given_X_batch(void *stream, const double *Xs, const int32_t *p_b,
                                                int pmax, const double *omega, const double *kappa,
                                                int N, int B, double sigma2, double *out,
                                                int32_t *status)
{
    if (B < 0 || N <= 0 || pmax <= 0 || !Xs || !omega || !kappa || !out) return MITRE_ERR_INVALID;
    if (pmax > MITRE_MAX_P) return MITRE_ERR_UNSUPPORTED;
    if (!(sigma2 > 0.0)) return MITRE_ERR_INVALID;
    if (B == 0) return MITRE_OK;
    dense_batch_kernel<<<B, 128, 0, as_stream(stream)>>>(Xs, p_b, pmax, omega, kappa, N, sigma2, out,
                                                         status);
    MITRE_LAUNCH_CHECK();
    return MITRE_OK;
}

extern "C" int mitre_mc_logpdf(void *stream, const double *x, const double *cov, int N, double *scratch,
                               double *out, int32_t *status)
{
    if (N <= 0 || !x || !cov || !scratch || !out) return MITRE_ERR_INVALID;
    mc_logpdf_kernel<<<1, 1024, 0, as_stream(stream)>>>(x, cov, N, scratch, out, status);
    MITRE_LAUNCH_CHECK();
    return MITRE_OK;
}
