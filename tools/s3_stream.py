"""S3 (BASELINE configs[2] shape: N = 2000 subjects, T = 50, 12 intervals, mean + slope, full threshold
grid) streamed in variable tiles (mitre_b200/streaming.py): per-tile detector evaluation + scoring +
fused (max, sum-exp), then a two-stage draw.  usage: python tools/s3_stream.py [V] [tile_variables] [shape=S3]
Times V variables and extrapolates linearly to the configuration's full variable count (tiles are independent);
with shape S5 (N = 35) it exercises the warp-per-group scoring kernel."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, '.')
from mitre_b200 import rules as R, streaming, synthetic  # noqa: E402
import oracle  # noqa: E402
from oracle import detectors_oracle as DO  # noqa: E402

import os  # noqa: E402
WORLD = int(os.environ.get("WORLD_SIZE", "1"))
RANK = int(os.environ.get("RANK", "0"))
if WORLD > 1:      # torchrun: the tiles are dealt to the ranks
    import torch.distributed as dist
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
V = int(sys.argv[1]) if len(sys.argv) > 1 else 128
TILE = int(sys.argv[2]) if len(sys.argv) > 2 else 32
SHAPE = sys.argv[3] if len(sys.argv) > 3 else "S3"
FULL_V = synthetic.SHAPES[SHAPE]["V"]
d = synthetic.make_shape(SHAPE, seed=0, V=V)
N = len(d["X"])
data = R.Dataset(d["X"], d["T"], d["y"], ["v%d" % i for i in range(V)], d["variable_weights"],
                 d["experiment_start"], d["experiment_end"])
sw = streaming.StreamingSweep(data, d["tmin"], d["N_intervals"], tmax=d["tmax"], tile_variables=TILE)
rng = np.random.RandomState(1)
omega = rng.gamma(2.0, 0.125, size=N) + 0.05
y = np.arange(N) % 2 == 0
X = np.stack([(rng.rand(N) > 0.5).astype(float) for _ in range(3)] + [np.ones(N)], axis=1)
mask = rng.rand(N) > 0.3
bv, bw = rng.normal(size=V), rng.normal(size=len(sw.windows))
sw.sweep(X, omega, y - 0.5, 100.0, mask=mask, prior_tables=(bv, bw, 0.0))      # warm-up (allocations)
torch.cuda.synchronize()
if WORLD > 1:
    dist.barrier()
t0 = time.perf_counter()
REPS = 3
for _ in range(REPS):
    log_total, n = sw.sweep(X, omega, y - 0.5, 100.0, mask=mask, prior_tables=(bv, bw, 0.0))
torch.cuda.synchronize()
if WORLD > 1:
    dist.barrier()
sweep_s = (time.perf_counter() - t0) / REPS
# phase split on one tile
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
ev[0].record()
tile = sw.build_tile(RANK % len(sw.tiles))
ev[1].record()
wts, status, stats = sw._weights_tile(tile, X, omega, y - 0.5, 100.0, mask, (bv, bw, 0.0))
ev[2].record()
out, status, stats, prior = sw._score_tile(tile, X, omega, y - 0.5, 100.0, mask, (bv, bw, 0.0))
out = out.clone()          # the sweep's output buffer is reused by draw()
torch.cuda.synchronize()
t0 = time.perf_counter()
t, k, tup = sw.draw(0.4321)
torch.cuda.synchronize()
draw_s = time.perf_counter() - t0
# parity spot check of one tile against the oracle
check = sw.build_tile(RANK % len(sw.tiles), rows=True)
idx = np.sort(rng.choice(tile['P'], size=24, replace=False))
from mitre_b200 import device as D  # noqa: E402
rows = D.unpack_rows(check['rows'][torch.as_tensor(idx, device=D.dev())], N).cpu().numpy().astype(bool)
want = oracle.likelihoods_of_all_options(X, omega, (y - 0.5) / omega, (rows * mask).astype(float), 100.0)
got = out.cpu().numpy()[idx]
# detector Gibbs at fixed structure: 2 rules x 2 detectors, full-table conditional per update
picks = []
for tt in range(min(4, len(sw.tiles))):
    tl = sw.build_tile(tt)
    picks.append(sw.describe(tl, int(np.random.RandomState(tt).randint(0, tl['P']))))
while len(picks) < 4:
    picks.append(picks[0])
gibbs = streaming.StreamingDetectorGibbs(sw, y, [[picks[0], picks[1]], [picks[2], picks[3]]], 100.0, (bv, bw, 0.0), seed=5)
gibbs.step()
torch.cuda.synchronize()
t0 = time.perf_counter()
GSTEPS = 5
for _ in range(GSTEPS):
    gibbs.step()
torch.cuda.synchronize()
gibbs_s = (time.perf_counter() - t0) / GSTEPS
if RANK == 0:
  print(json.dumps(dict(
    detector_gibbs=dict(what="update one detector from its full conditional over all candidates + 10 omega/beta sweeps, "
                             "2 rules x 2 detectors", s_per_step=gibbs_s, steps_per_s=1.0 / gibbs_s,
                        extrapolated_full_shape_s_per_step=gibbs_s * FULL_V / V),
    n_gpus=WORLD, shape=SHAPE, N=N, V=V, tile_variables=TILE, tiles=len(sw.tiles), windows=len(sw.windows), candidates=n,
    packed_bytes_never_materialised=int(n) * 8 * ((N + 63) // 64),
    sweep_wall_s=sweep_s, evals_per_s=n / sweep_s,
    tile0=dict(candidates=tile['P'], detectors_ms=ev[0].elapsed_time(ev[1]), score_ms=ev[1].elapsed_time(ev[2])),
    draw_wall_s=draw_s, drawn=[int(t), int(k), str(tup)],
    log_total=log_total, max_rel_diff_vs_oracle=float(np.max(np.abs(got - want) / np.abs(want))),
    extrapolated_full_shape=dict(variables=FULL_V, candidates=int(n * FULL_V / V), n_gpus=WORLD,
                                 sweep_s=sweep_s * FULL_V / V))))
if WORLD > 1:
    dist.destroy_process_group()
