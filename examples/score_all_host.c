/* A plain-C caller of the drop-in entry point: the reference's
 *   efficient_likelihoods.likelihoods_of_all_options(X, omega, response, truth_f64, s2)
 * (efficient_likelihoods.pyx:15-21) with host buffers, no Python, no torch -- what a cgo / JNI /
 * ctypes binding of libmitre_b200.so does.
 *
 *   gcc -O2 -Iinclude examples/score_all_host.c -Lmitre_b200 -lmitre_b200 \
 *       -Wl,-rpath,$PWD/mitre_b200 -lm -o /tmp/score_all_host
 *   /tmp/score_all_host [N p P seed]
 *
 * Prints P log-likelihoods, one per line, after a header line "N p P sigma2"; the inputs come from a
 * fixed linear congruential generator so that a checker can rebuild them (tests/test_c_example_gpu.py
 * does, and compares the output with the C oracle).
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "mitre_b200.h"

static uint64_t lcg_state;
static double lcg_uniform(void)
{
    lcg_state = lcg_state * 6364136223846793005ull + 1442695040888963407ull;
    return (double)(lcg_state >> 11) * (1.0 / 9007199254740992.0);
}

int main(int argc, char **argv)
{
    const int N = argc > 1 ? atoi(argv[1]) : 35;
    const int p = argc > 2 ? atoi(argv[2]) : 3;
    const int64_t P = argc > 3 ? atoll(argv[3]) : 1000;
    lcg_state = argc > 4 ? (uint64_t)atoll(argv[4]) : 1u;
    const double sigma2 = 100.0;
    double *X = malloc(sizeof(double) * N * p), *omega = malloc(sizeof(double) * N);
    double *response = malloc(sizeof(double) * N), *truth = malloc(sizeof(double) * P * N);
    double *out = malloc(sizeof(double) * P);
    if (!X || !omega || !response || !truth || !out) return 2;
    for (int i = 0; i < N; ++i) {
        for (int j = 0; j < p; ++j) X[i * p + j] = (j == p - 1) ? 1.0 : (lcg_uniform() > 0.5 ? 1.0 : 0.0);
        omega[i] = 0.05 + 0.5 * lcg_uniform();
        response[i] = ((lcg_uniform() > 0.5 ? 1.0 : 0.0) - 0.5) / omega[i];      /* z = (y - 1/2) / omega */
    }
    for (int64_t k = 0; k < P; ++k) {
        const double density = lcg_uniform();
        for (int i = 0; i < N; ++i) truth[k * N + i] = lcg_uniform() < density ? 1.0 : 0.0;
    }
    const int rc = mitre_score_all_host(X, omega, response, truth, sigma2, N, p, P, out);
    if (rc != MITRE_OK) {
        fprintf(stderr, "mitre_score_all_host: %s (%d)\n", mitre_error_string(rc), rc);
        return 1;
    }
    printf("%d %d %lld %.17g\n", N, p, (long long)P, sigma2);
    for (int64_t k = 0; k < P; ++k) printf("%.17g\n", out[k]);
    free(X); free(omega); free(response); free(truth); free(out);
    return 0;
}
