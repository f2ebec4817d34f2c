/*
 * oracle/scoring_oracle.c -- CPU restatement of the reference scoring path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in mitre_b200/ may link, load or call this file; it is
 * the checker used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.
 *
 * Follows (citations are into /root/reference/):
 *   oracle_likelihoods_of_all_options  <- mitre/efficient_likelihoods.pyx:15-169
 *   oracle_mc_logpdf                   <- mitre/mvnpdf_cholesky.py:53-74   (mc_logpdf_np)
 *   oracle_likelihood_z_given_X        <- mitre/logit_rules.py:171-177
 *   oracle_choose                      <- mitre/rules.py:1261-1264
 * The LAPACK/BLAS pieces the reference reaches through numpy/scipy (potrf, trsv, drotg)
 * are restated from their published definitions (reference BLAS DROTG, unblocked
 * right-looking Cholesky, forward substitution).
 *
 * Parity pin: checked against the unmodified .pyx built into oracle/_ref/ and against the
 * committed golden vectors tests/golden/scoring_*.npz (tests/test_oracle.py).
 *
 * Plain C99, no dependencies:  gcc -O2 -shared -fPIC scoring_oracle.c -o _build/liboracle.so -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* Lower Cholesky factor of the n x n row-major SPD matrix a, in place (lower triangle).
 * Returns 0, or k+1 if the k-th pivot is not positive (numpy raises LinAlgError there,
 * efficient_likelihoods.pyx:42 / mvnpdf_cholesky.py:66). */
static int chol_lower(double *a, int n)
{
    for (int j = 0; j < n; ++j) {
        double d = a[j * n + j];
        for (int k = 0; k < j; ++k) d -= a[j * n + k] * a[j * n + k];
        if (!(d > 0.0)) return j + 1;
        d = sqrt(d);
        a[j * n + j] = d;
        for (int i = j + 1; i < n; ++i) {
            double s = a[i * n + j];
            for (int k = 0; k < j; ++k) s -= a[i * n + k] * a[j * n + k];
            a[i * n + j] = s / d;
        }
    }
    return 0;
}

/* y = L^{-1} x, L lower, row-major (scipy.linalg.solve_triangular(lower=True)). */
static void fwd_solve(const double *L, int n, const double *x, double *y)
{
    for (int i = 0; i < n; ++i) {
        double s = x[i];
        for (int k = 0; k < i; ++k) s -= L[i * n + k] * y[k];
        y[i] = s / L[i * n + i];
    }
}

/* Reference-BLAS DROTG: on return *c,*s hold the rotation, *a = r, *b = z.
 * (called at efficient_likelihoods.pyx:109-114 through scipy.linalg.cython_blas) */
static void drotg(double *a, double *b, double *c, double *s)
{
    double roe = *b;
    if (fabs(*a) > fabs(*b)) roe = *a;
    double scale = fabs(*a) + fabs(*b);
    double r, z;
    if (scale == 0.0) {
        *c = 1.0; *s = 0.0; r = 0.0; z = 0.0;
    } else {
        double as = *a / scale, bs = *b / scale;
        r = scale * sqrt(as * as + bs * bs);
        if (roe < 0.0) r = -r;
        *c = *a / r;
        *s = *b / r;
        z = 1.0;
        if (fabs(*a) > fabs(*b)) z = *s;
        if (fabs(*b) >= fabs(*a) && *c != 0.0) z = 1.0 / *c;
    }
    *a = r; *b = z;
}

/*
 * efficient_likelihoods.pyx:15-169.
 *   X      [k, p]  row-major (k = number of subjects)
 *   omega  [k], response [k]  (response = z = kappa/omega, logit_rules.py:222-223)
 *   truth  [n_options, k] row-major doubles (0.0/1.0), the caller's `this_rule_all_options`
 *   out    [n_options]
 * Returns 0 or the Cholesky failure code.
 */
int oracle_likelihoods_of_all_options(const double *X, const double *omega,
                                      const double *response, const double *truth,
                                      double prior_coefficient_variance,
                                      int k, int p, int64_t n_options, double *out)
{
    const double sqrt_pcv = sqrt(prior_coefficient_variance);
    double *base_R = (double *)malloc(sizeof(double) * k * k);   /* lower factor */
    double *base_y = (double *)malloc(sizeof(double) * k);
    double *top = (double *)calloc((size_t)(k + 1) * k, sizeof(double));
    double *diag_lsq = (double *)calloc(k + 1, sizeof(double));
    double *y_pseudo = (double *)calloc(k + 1, sizeof(double));
    double *ydoty = (double *)calloc(k + 1, sizeof(double));
    double *gc = (double *)calloc(k, sizeof(double));
    double *gs = (double *)calloc(k, sizeof(double));
    double *v = (double *)calloc(k, sizeof(double));
    double *last_v = (double *)calloc(k, sizeof(double));

    /* pyx:38-41  base_covariance_z = diag(1/omega) + pcv * X X^T */
    for (int i = 0; i < k; ++i)
        for (int j = 0; j < k; ++j) {
            double s = 0.0;
            for (int c = 0; c < p; ++c) s += X[i * p + c] * X[j * p + c];
            base_R[i * k + j] = prior_coefficient_variance * s + (i == j ? 1.0 / omega[i] : 0.0);
        }
    int info = chol_lower(base_R, k);                    /* pyx:42 */
    if (info) goto done;
    fwd_solve(base_R, k, response, base_y);              /* pyx:44-45 */
    /* base_R_T[r, c] = base_R[c, r]  (pyx:43) */
#define RT(r, c) base_R[(size_t)(c) * k + (r)]
#define TOP(r, c) top[(size_t)(r) * k + (c)]
    const double normalization = 0.5 * k * log(2.0 * M_PI);   /* pyx:47 */

    for (int64_t opt = 0; opt < n_options; ++opt) {      /* pyx:66 */
        for (int i = 0; i < k; ++i) v[i] = truth[opt * k + i];
        int d = 0;
        if (opt != 0) {                                  /* pyx:70-77 */
            d = k;
            for (int i = 0; i < k; ++i)
                if (v[i] != last_v[i]) { d = i; break; }
        }
        memcpy(last_v, v, sizeof(double) * k);
        for (int i = d; i < k; ++i) TOP(0, i) = sqrt_pcv * v[i];   /* pyx:82-83 */
        for (int col = d; col < k; ++col)                /* pyx:92-101 catch-up */
            for (int row = 0; row < d; ++row)
                TOP(row + 1, col) = gc[row] * TOP(row, col) - gs[row] * RT(row, col);
        for (int u = d; u < k; ++u) {                    /* pyx:104-147 */
            double a = RT(u, u), b = TOP(u, u), c, s;
            drotg(&a, &b, &c, &s);
            gc[u] = c; gs[u] = s;
            for (int i = u; i < k; ++i)
                TOP(u + 1, i) = c * TOP(u, i) - s * RT(u, i);
            double dnew = s * TOP(u, u) + c * RT(u, u);
            diag_lsq[u + 1] = diag_lsq[u] + 2.0 * log(fabs(dnew));
            y_pseudo[u + 1] = c * y_pseudo[u] - s * base_y[u];
            double sol = s * y_pseudo[u] + c * base_y[u];
            ydoty[u + 1] = ydoty[u] + sol * sol;
        }
        out[opt] = -0.5 * (ydoty[k] + diag_lsq[k]) - normalization;   /* pyx:160-169 */
    }
#undef RT
#undef TOP
done:
    free(base_R); free(base_y); free(top); free(diag_lsq); free(y_pseudo); free(ydoty);
    free(gc); free(gs); free(v); free(last_v);
    return info;
}

/* mvnpdf_cholesky.py:53-74  mc_logpdf_np(x, cov); cov [k,k] row-major (not modified). */
int oracle_mc_logpdf(const double *x, const double *cov, int k, double *out)
{
    double *R = (double *)malloc(sizeof(double) * k * k);
    double *y = (double *)malloc(sizeof(double) * k);
    memcpy(R, cov, sizeof(double) * k * k);
    int info = chol_lower(R, k);
    if (!info) {
        double log_det = 0.0, yy = 0.0;
        for (int i = 0; i < k; ++i) log_det += log(R[i * k + i]);
        log_det *= 2.0;
        fwd_solve(R, k, x, y);
        for (int i = 0; i < k; ++i) yy += y[i] * y[i];
        *out = -0.5 * yy - 0.5 * (k * log(2.0 * M_PI) + log_det);
    }
    free(R); free(y);
    return info;
}

/* logit_rules.py:171-177  likelihood_z_given_X(omega, X) with kappa = y - 0.5. */
int oracle_likelihood_z_given_X(const double *X, const double *omega, const double *kappa,
                                double prior_coefficient_variance, int k, int p, double *out)
{
    double *cov = (double *)malloc(sizeof(double) * k * k);
    double *z = (double *)malloc(sizeof(double) * k);
    for (int i = 0; i < k; ++i) {
        z[i] = kappa[i] / omega[i];
        for (int j = 0; j < k; ++j) {
            double s = 0.0;
            for (int c = 0; c < p; ++c) s += X[i * p + c] * X[j * p + c];
            cov[i * k + j] = prior_coefficient_variance * s + (i == j ? 1.0 / omega[i] : 0.0);
        }
    }
    int info = oracle_mc_logpdf(z, cov, k, out);
    free(cov); free(z);
    return info;
}

/* rules.py:1261-1264  choose_from_relative_probabilities(p) given the uniform draw u:
 *   p = cumsum(p); marker = u * p[-1]; return sum(p < marker)
 * (np.cumsum is a sequential left-to-right accumulation.) */
int64_t oracle_choose(const double *p, int64_t n, double u)
{
    double *c = (double *)malloc(sizeof(double) * (size_t)n);
    double acc = 0.0;
    for (int64_t i = 0; i < n; ++i) { acc += p[i]; c[i] = acc; }
    double marker = u * c[n - 1];
    int64_t cnt = 0;
    for (int64_t i = 0; i < n; ++i) cnt += (c[i] < marker);
    free(c);
    return cnt;
}
