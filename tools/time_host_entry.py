"""Throughput of the drop-in host entry (mitre_score_all_host: float64[P, N] truth matrix in pageable host
memory in, float64[P] out), the reference's exact call.  usage: python tools/time_host_entry.py [P]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from mitre_b200 import efficient_likelihoods as EL  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 2000000
N, p = 35, 4
rng = np.random.RandomState(0)
X = np.concatenate([(rng.rand(N, p - 1) > 0.5).astype(float), np.ones((N, 1))], axis=1)
omega = rng.gamma(2.0, 0.125, size=N) + 0.05
z = ((rng.rand(N) > 0.5) - 0.5) / omega
opts = (rng.rand(P, N) > 0.5).astype(np.float64)
EL.likelihoods_of_all_options(X, omega, z, opts[:1000], 100.0)
for _ in range(3):
    t0 = time.perf_counter()
    out = EL.likelihoods_of_all_options(X, omega, z, opts, 100.0)
    dt = time.perf_counter() - t0
    print("P=%d: %.1f ms, %.3g evals/s, %.1f GB/s of float64 truth matrix" % (P, dt * 1e3, P / dt, P * N * 8 / dt / 1e9))
