"""cProfile of parallel.ShardedPopulation's build at the S5 shape (one rank; GPU box)."""
import cProfile
import io
import pstats
import sys
import time
import torch
sys.path.insert(0, ".")
import bench  # noqa: E402
from mitre_b200 import parallel as PL, rules as R, synthetic  # noqa: E402

V = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
d = synthetic.make_shape("S5", seed=0, V=V)
data = R.Dataset(d["X"], d["T"], d["y"], ["v%d" % i for i in range(V)], d["variable_weights"],
                 d["experiment_start"], d["experiment_end"])
sp = PL.ShardedPopulation(data, d["tmin"], d["N_intervals"], tmax=d["tmax"])
del sp
torch.cuda.synchronize()
t0 = time.perf_counter()
pr = cProfile.Profile()
pr.enable()
sp = PL.ShardedPopulation(data, d["tmin"], d["N_intervals"], tmax=d["tmax"])
torch.cuda.synchronize()
pr.disable()
print("build %.3f s, P = %d" % (time.perf_counter() - t0, sp.P))
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22)
print(s.getvalue()[:5000])
