// gibbs.cu -- the omega / beta Gibbs inner loop of one MCMC step on the device.
//
// Reference: LogisticRuleSampler.step runs 10 x { update_omega (logit_rules.py:753-760:
// psi = X beta, omega_i ~ PG(1, psi_i) through pypolyagamma), update_beta (762-787:
// beta ~ N(V X' kappa, V), V = (I/v + X' Omega X)^-1) }.  pypolyagamma is a third-party C++
// dependency that is neither vendored nor pinned (setup.py:57); here both draws are one
// single-CTA launch:
//   omega_i : Devroye's exact alternating-series sampler for J*(1, z/2) (Polson, Scott & Windle
//             2013), one thread per subject -- the same algorithm as mitre_b200/polyagamma.py
//   beta    : Cholesky A = I/v + X' Omega X = L L', mean = L^-T L^-1 X' kappa, beta = mean + L^-T eps
// Random numbers: Philox4x32-10 (counter-based; key = caller's seed, counter = (thread, draw)).
// This path is distributionally, not stream-for-stream, equivalent to the host sampler
// (tests/test_gibbs_gpu.py checks the PG moments and the beta posterior).
#include <math.h>

#include "common.cuh"

namespace mitre {

struct Philox {
    uint32_t key0, key1;
    uint32_t c0, c1, c2, c3;
    uint32_t out[4];
    int have;

    __device__ void init(uint64_t seed, uint32_t stream, uint32_t substream)
    {
        key0 = (uint32_t)seed; key1 = (uint32_t)(seed >> 32);
        c0 = 0; c1 = 0; c2 = stream; c3 = substream;
        have = 0;
    }
    __device__ void round(uint32_t &a, uint32_t &b, uint32_t &c, uint32_t &d, uint32_t k0, uint32_t k1)
    {
        const uint32_t hi0 = __umulhi(0xD2511F53u, a), lo0 = 0xD2511F53u * a;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c), lo1 = 0xCD9E8D57u * c;
        const uint32_t na = hi1 ^ b ^ k0, nb = lo1, nc = hi0 ^ d ^ k1, nd = lo0;
        a = na; b = nb; c = nc; d = nd;
    }
    __device__ void refill()
    {
        uint32_t a = c0, b = c1, c = c2, d = c3, k0 = key0, k1 = key1;
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            round(a, b, c, d, k0, k1);
            k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
        }
        out[0] = a; out[1] = b; out[2] = c; out[3] = d;
        if (++c0 == 0) ++c1;
        have = 4;
    }
    __device__ uint32_t next32()
    {
        if (have == 0) refill();
        return out[--have];
    }
    // uniform in (0, 1): 53 random bits, never 0
    __device__ double uniform()
    {
        const uint64_t hi = next32(), lo = next32();
        const uint64_t bits = ((hi << 32) | lo) >> 11;
        return ((double)bits + 0.5) * (1.0 / 9007199254740992.0);
    }
    __device__ double exponential() { return -log(uniform()); }
    __device__ double normal()
    {
        const double u1 = uniform(), u2 = uniform();
        return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
    }
};

constexpr double kTrunc = 0.64;

__device__ double pg_a_coef(int n, double x)
{
    const double k = (n + 0.5) * M_PI;
    if (x > kTrunc) return k * exp(-0.5 * k * k * x);
    if (x > 0.0) return exp(-1.5 * (log(0.5 * M_PI) + log(x)) + log(k) - 2.0 * (n + 0.5) * (n + 0.5) / x);
    return 0.0;
}

__device__ double pg_rtigauss(double z, Philox &rng)
{
    const double t = kTrunc;
    if (z < 1.0 / t) {
        for (;;) {
            double e1, e2;
            do {
                e1 = rng.exponential();
                e2 = rng.exponential();
            } while (e1 * e1 > 2.0 * e2 / t);
            const double x = t / ((1.0 + t * e1) * (1.0 + t * e1));
            if (rng.uniform() <= exp(-0.5 * z * z * x)) return x;
        }
    }
    const double mu = 1.0 / z;
    for (;;) {
        double y = rng.normal();
        y *= y;
        const double half_mu = 0.5 * mu, mu_y = mu * y;
        double x = mu + half_mu * mu_y - half_mu * sqrt(4.0 * mu_y + mu_y * mu_y);
        if (rng.uniform() > mu / (mu + x)) x = mu * mu / x;
        if (x <= t) return x;
    }
}

// PG(1, z)
__device__ double pg_draw(double z, Philox &rng)
{
    z = fabs(z) * 0.5;
    const double t = kTrunc;
    const double K = 0.125 * M_PI * M_PI + 0.5 * z * z;
    const double w = sqrt(1.0 / t);
    const double x0 = log(K) + K * t;
    const double pb = normcdf(w * (t * z - 1.0)), pa = normcdf(-w * (t * z + 1.0));
    const double xb = x0 - z + (pb > 0.0 ? log(pb) : -INFINITY);
    const double xa = x0 + z + (pa > 0.0 ? log(pa) : -INFINITY);
    const double ratio = 1.0 / (1.0 + 4.0 / M_PI * (exp(xb) + exp(xa)));
    for (;;) {
        double x;
        if (rng.uniform() < ratio) x = t + rng.exponential() / K;
        else x = pg_rtigauss(z, rng);
        double s = pg_a_coef(0, x);
        const double y = rng.uniform() * s;
        int n = 0;
        for (;;) {
            ++n;
            if (n & 1) {
                s -= pg_a_coef(n, x);
                if (y <= s) return 0.25 * x;
            } else {
                s += pg_a_coef(n, x);
                if (y > s) break;
            }
        }
    }
}

// One CTA.  X [N, p] row-major; beta_io [p] in: starting beta, out: final beta;
// omega_out [N]; betas_out [n_iters, p].  mode 0: full sweep (omega then beta) x n_iters;
// mode 1: beta only, given omega_out as INPUT (n_iters independent draws, for testing / the
// single update_beta calls).
__global__ void __launch_bounds__(128)
gibbs_kernel(const double *__restrict__ X, const double *__restrict__ kappa, int N, int p, double v,
             int n_iters, int mode, uint64_t seed, double *__restrict__ beta_io,
             double *__restrict__ omega_out, double *__restrict__ betas_out, int32_t *status)
{
    __shared__ double sA[MITRE_MAX_P * MITRE_MAX_P];
    __shared__ double sb[MITRE_MAX_P], sy[MITRE_MAX_P], sbeta[MITRE_MAX_P];
    extern __shared__ double somega[];    // [N]
    const int tid = threadIdx.x, nt = blockDim.x;
    Philox rng;
    rng.init(seed, (uint32_t)tid, 0);
    for (int j = tid; j < p; j += nt) {
        sbeta[j] = beta_io[j];
        double acc = 0.0;
        for (int i = 0; i < N; ++i) acc += X[(size_t)i * p + j] * kappa[i];
        sb[j] = acc;
    }
    if (mode == 1)
        for (int i = tid; i < N; i += nt) somega[i] = omega_out[i];
    __syncthreads();
    for (int it = 0; it < n_iters; ++it) {
        if (mode == 0) {
            for (int i = tid; i < N; i += nt) {
                double psi = 0.0;
                for (int j = 0; j < p; ++j) psi += X[(size_t)i * p + j] * sbeta[j];
                somega[i] = pg_draw(psi, rng);
            }
            __syncthreads();
        }
        for (int e = tid; e < p * p; e += nt) {
            const int j = e / p, k = e % p;
            double acc = 0.0;
            for (int i = 0; i < N; ++i) acc += somega[i] * X[(size_t)i * p + j] * X[(size_t)i * p + k];
            sA[e] = acc + (j == k ? 1.0 / v : 0.0);
        }
        __syncthreads();
        if (tid == 0) {
            bool ok = true;
            for (int j = 0; j < p && ok; ++j) {
                double d = sA[j * p + j];
                for (int k = 0; k < j; ++k) d -= sA[j * p + k] * sA[j * p + k];
                if (!(d > 0.0)) { ok = false; break; }
                d = sqrt(d);
                sA[j * p + j] = d;
                for (int i = j + 1; i < p; ++i) {
                    double s = sA[i * p + j];
                    for (int k = 0; k < j; ++k) s -= sA[i * p + k] * sA[j * p + k];
                    sA[i * p + j] = s / d;
                }
            }
            if (!ok) {
                if (status) atomicOr(reinterpret_cast<unsigned int *>(status), MITRE_STATUS_BASE_NOT_PD);
            } else {
                // y = L^-1 b + eps  (then beta = L^-T y = mean + L^-T eps)
                for (int i = 0; i < p; ++i) {
                    double s = sb[i];
                    for (int k = 0; k < i; ++k) s -= sA[i * p + k] * sy[k];
                    sy[i] = s / sA[i * p + i];
                }
                for (int i = 0; i < p; ++i) sy[i] += rng.normal();
                for (int i = p - 1; i >= 0; --i) {
                    double s = sy[i];
                    for (int k = i + 1; k < p; ++k) s -= sA[k * p + i] * sbeta[k];
                    sbeta[i] = s / sA[i * p + i];
                }
            }
            for (int j = 0; j < p; ++j) betas_out[(size_t)it * p + j] = sbeta[j];
        }
        __syncthreads();
    }
    for (int j = tid; j < p; j += nt) beta_io[j] = sbeta[j];
    if (mode == 0)
        for (int i = tid; i < N; i += nt) omega_out[i] = somega[i];
}

// PG(1, z) draws only (test hook for the sampler's moments)
__global__ void __launch_bounds__(256)
pg_draw_kernel(const double *__restrict__ z, int64_t n, uint64_t seed, double *__restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Philox rng;
    rng.init(seed, (uint32_t)i, (uint32_t)(i >> 32) + 1u);
    out[i] = pg_draw(z[i], rng);
}

}  // namespace mitre

using namespace mitre;

extern "C" int mitre_gibbs_omega_beta(void *stream, const double *X, const double *kappa, int N, int p,
                                      double prior_variance, int n_iters, int mode, uint64_t seed,
                                      double *beta_io, double *omega, double *betas_out, int32_t *status)
{
    if (N <= 0 || p <= 0 || n_iters <= 0 || !X || !kappa || !beta_io || !omega || !betas_out)
        return MITRE_ERR_INVALID;
    if (p > MITRE_MAX_P || (mode != 0 && mode != 1)) return MITRE_ERR_UNSUPPORTED;
    if (!(prior_variance > 0.0)) return MITRE_ERR_INVALID;
    const size_t smem = sizeof(double) * (size_t)N;
    if (smem > 160 * 1024) return MITRE_ERR_UNSUPPORTED;
    if (smem > 40 * 1024)
        MITRE_CUDA_TRY(cudaFuncSetAttribute(gibbs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gibbs_kernel<<<1, 128, smem, as_stream(stream)>>>(X, kappa, N, p, prior_variance, n_iters, mode, seed, beta_io,
                                                      omega, betas_out, status);
    MITRE_LAUNCH_CHECK();
    return MITRE_OK;
}

extern "C" int mitre_pg_draw(void *stream, const double *z, int64_t n, uint64_t seed, double *out)
{
    if (n < 0 || (n > 0 && (!z || !out))) return MITRE_ERR_INVALID;
    if (n == 0) return MITRE_OK;
    pg_draw_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, as_stream(stream)>>>(z, n, seed, out);
    MITRE_LAUNCH_CHECK();
    return MITRE_OK;
}
