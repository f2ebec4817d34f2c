"""Randomised parity sweep of the device detector path against the numpy oracle (GPU box).
usage: python tools/fuzz_detectors.py [n_cases] [seed]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from mitre_b200 import rules as R, synthetic  # noqa: E402
from oracle import detectors_oracle as DO  # noqa: E402

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.RandomState(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
t0 = time.time()
bad = 0
for case in range(n_cases):
    N = int(rng.choice([2, 3, 5, 8, 13, 21, 33, 35, 47, 63, 64, 70, 130]))
    V = int(rng.randint(2, 70))      # V = 1 makes every series the constant 1 (DESIGN "slope parity")
    tlo = int(rng.randint(2, 8))
    T = (tlo, tlo + int(rng.randint(0, 30)))
    nint = int(rng.randint(1, 7))
    zf = float(rng.choice([0.0, 0.1, 0.5]))
    seed = int(rng.randint(1 << 30))
    if V < 2:
        zf = 0.0
    d = synthetic.make_series(N=N, V=V, T=T, span=(0.0, 90.0), seed=seed, zero_fraction=zf)
    ref = DO.build_population(d["X"], d["T"], 0.0, 90.0, nint, 1.0, None)
    data = R.Dataset(d["X"], d["T"], d["y"], ["v%d" % i for i in range(V)], d["variable_weights"],
                     d["experiment_start"], d["experiment_end"])
    pop = R.DiscreteRulePopulation(None, 1.0, nint, tmax=None, data=data)
    o = ref["order"]
    ok = (len(pop) == len(o) and
          np.array_equal(pop.packed_truth.cpu().numpy().view(np.uint64), DO.pack_rows(ref["truth"][o])) and
          np.array_equal(pop._perm_host(), o))
    if ok and len(o):
        is_mean = ref["group_type"] == 1
        ok = np.array_equal(pop._values.cpu().numpy()[is_mean], ref["values"][is_mean])
    if not ok:
        bad += 1
        print("MISMATCH", dict(N=N, V=V, T=T, nint=nint, zero_fraction=zf, seed=seed))
print("%d cases, %d mismatches, %.1f s" % (n_cases, bad, time.time() - t0))
sys.exit(1 if bad else 0)
