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
__host__ __device__ inline int ep_for(int p) { return (p + 2 + 1) & ~1; }  // even, >= p+2

// ---------------------------------------------------------------------------------------------
// Base system for one design matrix.  Called by every thread of a CTA; smem scratch:
//   sA[p*p] (in: nothing, out: L lower-triangular), sb[p] (out: yb = L^-1 b), sred[blockDim]
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
    if (tid == 0) {
        // unblocked Cholesky, in place (lower), then yb = L^-1 b
        int ok = 1;
        for (int j = 0; j < p && ok; ++j) {
            double d = sA[j * p + j];
            for (int k = 0; k < j; ++k) d -= sA[j * p + k] * sA[j * p + k];
            if (!(d > 0.0)) { ok = 0; break; }
            d = sqrt(d);
            sA[j * p + j] = d;
            for (int i = j + 1; i < p; ++i) {
                double s = sA[i * p + j];
                for (int k = 0; k < j; ++k) s -= sA[i * p + k] * sA[j * p + k];
                sA[i * p + j] = s / d;
            }
        }
        double logdetL = 0.0, yy = 0.0;
        if (ok) {
            for (int i = 0; i < p; ++i) {
                double s = sb[i];
                for (int k = 0; k < i; ++k) s -= sA[i * p + k] * sb[k];
                s /= sA[i * p + i];
                sb[i] = s;
                yy += s * s;
                logdetL += log(sA[i * p + i]);
            }
        }
        *s_ok = ok;
        // sred[0] = sum kappa^2/omega - sum log omega
        const double quad_plus_logdet = sred[0] - yy + p * log(sigma2) + 2.0 * logdetL;
        sred[0] = -0.5 * quad_plus_logdet - 0.5 * N * log(2.0 * M_PI);
    }
    __syncthreads();
    return sred[0];
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
        if (!s_ok && status) atomicOr(reinterpret_cast<unsigned int *>(status), MITRE_STATUS_BASE_NOT_PD);
    }
    double *q = ws + kHdr;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const bool in_mask = (mask == nullptr) || (mask[i >> 6] & subject_bit(i));
        double *qi = q + (size_t)i * EP;
        if (!in_mask || !s_ok) {
            for (int e = 0; e < EP; ++e) qi[e] = s_ok ? 0.0 : nan("");
            continue;
        }
        const double om = omega[i];
        double w[MITRE_MAX_P];
        double dot = 0.0;
        for (int j = 0; j < p; ++j) {
            double s = om * X[(size_t)i * p + j];
            for (int k = 0; k < j; ++k) s -= sA[j * p + k] * w[k];
            s /= sA[j * p + j];
            w[j] = s;
            dot += s * sb[j];
        }
        qi[0] = om;
        qi[1] = kappa[i] - dot;
        for (int j = 0; j < p; ++j) qi[2 + j] = w[j];
        for (int e = p + 2; e < EP; ++e) qi[e] = 0.0;
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
// ---------------------------------------------------------------------------------------------
template <int EP, int CB>
__global__ void __launch_bounds__(256)
score_w1_kernel(const uint64_t *__restrict__ rows, int64_t P, const double *__restrict__ ws, int N,
                double *__restrict__ out, int32_t *status)
{
    extern __shared__ __align__(16) double lut[];
    constexpr int PAT = 1 << CB;
    const int nchunks = (N + CB - 1) / CB;
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
    __syncthreads();
    const double C0 = ws[0], inv_s2 = ws[1], s2 = ws[2];
    bool bad = false;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < P; k += stride) {
        const uint64_t r = ld_stream_u64(rows + k);
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
        st_stream_f64(out + k, finish_candidate<EP>(acc, C0, inv_s2, s2, bad));
    }
    if (bad && status) atomicOr(reinterpret_cast<unsigned int *>(status), MITRE_STATUS_BAD_SCHUR);
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
static int launch_w1(cudaStream_t st, const uint64_t *rows, int64_t P, const double *ws, int N,
                     double *out, int32_t *status)
{
    const int nchunks = (N + CB - 1) / CB;
    const size_t smem = (size_t)nchunks * (1 << CB) * EP * sizeof(double);
    auto kern = score_w1_kernel<EP, CB>;
    MITRE_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    MITRE_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, smem));
    if (per_sm < 1) return MITRE_ERR_UNSUPPORTED;
    int64_t blocks = (int64_t)sm_count() * per_sm;
    const int64_t need = ceil_div64(P, 256);
    if (need < blocks) blocks = need;
    if (blocks < 1) blocks = 1;
    kern<<<(unsigned)blocks, 256, smem, st>>>(rows, P, ws, N, out, status);
    MITRE_LAUNCH_CHECK();
    return MITRE_OK;
}

template <int EP>
static int launch_wide(cudaStream_t st, const uint64_t *rows, int64_t P, int W64, const double *ws,
                       int N, double *out, int32_t *status)
{
    auto kern = score_wide_kernel<EP>;
    int per_sm = 1;
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
static int launch_ep(cudaStream_t st, const uint64_t *rows, int64_t P, int W64, const double *ws, int N,
                     double *out, int32_t *status)
{
    if (W64 == 1) {
        const size_t lut8 = (size_t)((N + 7) / 8) * 256 * EP * sizeof(double);
        // keep >= 2 CTAs per SM where the byte table allows it, else fall to nibble tables
        if (lut8 <= 100 * 1024) return launch_w1<EP, 8>(st, rows, P, ws, N, out, status);
        return launch_w1<EP, 4>(st, rows, P, ws, N, out, status);
    }
    return launch_wide<EP>(st, rows, P, W64, ws, N, out, status);
}

}  // namespace mitre

using namespace mitre;

extern "C" size_t mitre_score_workspace_bytes(int N, int p)
{
    if (N <= 0 || p <= 0 || p > MITRE_MAX_P) return 0;
    return sizeof(double) * ((size_t)kHdr + (size_t)N * ep_for(p));
}

extern "C" int mitre_score_all(void *stream, const uint64_t *packed_truth, int64_t P, int W64,
                               const uint64_t *mask, const double *X, const double *omega,
                               const double *kappa, int N, int p, double sigma2, void *workspace,
                               size_t workspace_bytes, double *out, int32_t *status)
{
    if (P < 0 || N <= 0 || p <= 0 || !X || !omega || !kappa || !workspace) return MITRE_ERR_INVALID;
    if (P > 0 && (!packed_truth || !out)) return MITRE_ERR_INVALID;
    if (W64 != words_for(N)) return MITRE_ERR_INVALID;
    if (p > MITRE_MAX_P) return MITRE_ERR_UNSUPPORTED;
    if (!(sigma2 > 0.0)) return MITRE_ERR_INVALID;
    if (workspace_bytes < mitre_score_workspace_bytes(N, p)) return MITRE_ERR_WORKSPACE;
    cudaStream_t st = as_stream(stream);
    double *ws = static_cast<double *>(workspace);
    score_setup_kernel<<<1, 256, 0, st>>>(X, omega, kappa, mask, N, p, sigma2, ws, status);
    MITRE_LAUNCH_CHECK();
    if (P == 0) return MITRE_OK;
    switch (ep_for(p)) {
#define MITRE_EP_CASE(E) \
    case E:              \
        return launch_ep<E>(st, packed_truth, P, W64, ws, N, out, status);
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
