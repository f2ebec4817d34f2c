"""S3 slice (BASELINE configs[2] shape: N = 2000 subjects, T = 50, 12 intervals, mean + slope, full
threshold grid) for a handful of variables: the N > 64 path end to end (phase-by-phase detector
kernels, multi-word lexsort, warp-per-candidate scoring), timed with CUDA events and spot-checked
against the oracle."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, '.')
from mitre_b200 import device as D, rules as R, synthetic
import oracle

V = int(sys.argv[1]) if len(sys.argv) > 1 else 8
d = synthetic.make_shape("S3", seed=0, V=V)
N = len(d["X"])
data = R.Dataset(d["X"], d["T"], d["y"], ["v%d" % i for i in range(V)], d["variable_weights"],
                 d["experiment_start"], d["experiment_end"])
torch.cuda.synchronize()
t0 = time.perf_counter()
pop = R.DiscreteRulePopulation(None, d["tmin"], d["N_intervals"], tmax=d["tmax"], data=data)
torch.cuda.synchronize()
build_s = time.perf_counter() - t0
P = len(pop)
packed = pop.packed_truth
rng = np.random.RandomState(1)
omega = rng.gamma(2.0, 0.125, size=N) + 0.05
y = np.arange(N) % 2 == 0
cols = [pop._truth_row(int(rng.randint(P))).astype(float) for _ in range(3)]
X = np.stack(cols + [np.ones(N)], axis=1)
mask = pop._truth_row(int(rng.randint(P)))
dmask = torch.as_tensor(D.pack_mask(mask), device=D.dev())
buf = D.ScoreBuffers(P, N, 4)
out, status = D.score_all(packed, N, X, omega, y - 0.5, 100.0, mask=dmask, buffers=buf)
D.raise_on_status(status)
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
s.record()
for _ in range(3):
    D.score_all(packed, N, X, omega, y - 0.5, 100.0, mask=dmask, buffers=buf)
e.record()
torch.cuda.synchronize()
ms = s.elapsed_time(e) / 3
# parity spot check against the oracle on 40 rows
idx = np.sort(rng.choice(P, size=40, replace=False))
rows = D.unpack_rows(packed[torch.as_tensor(idx, device=D.dev())], N).cpu().numpy().astype(bool)
want = oracle.likelihoods_of_all_options(X, omega, (y - 0.5) / omega, (rows * mask).astype(float), 100.0)
got = out.cpu().numpy()[idx]
rows_all = packed.cpu().numpy().view(np.uint64)
sorted_ok = bool(np.all([tuple(rows_all[i]) <= tuple(rows_all[i + 1]) for i in range(0, min(P - 1, 20000))]))
print(json.dumps(dict(N=N, V=V, windows=len(pop.windows), groups=int(pop._G), candidates=P,
                      packed_bytes=int(P * packed.shape[1] * 8), build_wall_s=build_s,
                      score_ms=ms, evals_per_s=P / (ms * 1e-3),
                      algorithmic_gbs=(P * (packed.shape[1] * 8 + 8)) / (ms * 1e-3) / 1e9,
                      max_rel_diff_vs_oracle=float(np.max(np.abs(got - want) / np.abs(want))),
                      rows_sorted=sorted_ok)))
