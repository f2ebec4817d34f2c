"""A chain of the reference's full sampler step on the S3 shape (BASELINE configs[2]: N = 2000 subjects,
T = 50, 12 intervals, mean + slope, full threshold grid), streamed: streaming.make_streaming_model /
make_streaming_sampler.  usage: [torchrun ...] python tools/s3_chain.py [V=5000] [tile_variables=128] [steps=20]
Prints one JSON line: moves made, wall time per step, the sweeps behind them."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, '.')
from mitre_b200 import logit_rules as LR, polyagamma, rules as R, streaming, synthetic  # noqa: E402

WORLD = int(os.environ.get("WORLD_SIZE", "1"))
RANK = int(os.environ.get("RANK", "0"))
if WORLD > 1:
    import torch.distributed as dist
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
V = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
TILE = int(sys.argv[2]) if len(sys.argv) > 2 else 128
STEPS = int(sys.argv[3]) if len(sys.argv) > 3 else 20
d = synthetic.make_shape("S3", seed=0, V=V)
N = len(d["X"])
data = R.Dataset(d["X"], d["T"], d["y"], ["v%d" % i for i in range(V)], d["variable_weights"],
                 d["experiment_start"], d["experiment_end"])
t0 = time.perf_counter()
model = streaming.make_streaming_model(data, d["tmin"], tile_variables=TILE, N_intervals=d["N_intervals"],
                                       tmax=d["tmax"])
P = len(model.rule_population)            # one detector pass over every tile (the prior's count matrix)
torch.cuda.synchronize()
setup_s = time.perf_counter() - t0
sw = model.sweep
tile0 = sw.build_tile(0)
first = R.PrimitiveRule(*sw.describe(tile0, tile0['P'] // 3))
np.random.seed(7)                          # every rank: the same stream (the sampler is replicated, the sweeps sharded)
LR.ppg = polyagamma.PyPolyaGamma(7)
sampler = streaming.make_streaming_sampler(model, R.RuleList([[first]]))
n_sweeps = [0]
orig = sw.sweep


def counted(*a, **k):
    n_sweeps[0] += 1
    return orig(*a, **k)
sw.sweep = counted
sampler.step()                             # warm-up (allocations)
torch.cuda.synchronize()
n_sweeps[0] = 0
done = len(sampler.comments)
t0 = time.perf_counter()
sampler.sample(STEPS)
torch.cuda.synchronize()
wall = time.perf_counter() - t0
comments = sampler.comments[done:]
kind = lambda c: ('update' if 'replacing' in c else 'add' if c.startswith('add') else
                  'remove' if c.startswith('remove') or 'Nothing to remove' in c else 'move' if 'move' in c else c)
moves = {}
for c in comments:
    moves[kind(c)] = moves.get(kind(c), 0) + 1
if RANK == 0:
    print(json.dumps(dict(
        what="LogisticRuleSampler.step on a streamed S3 population (never materialised)", n_gpus=WORLD, N=N, V=V,
        tile_variables=TILE, candidates=P, packed_bytes_never_materialised=P * 8 * ((N + 63) // 64),
        setup_s=setup_s, steps=STEPS, s_per_step=wall / STEPS, sweeps=n_sweeps[0], moves=moves,
        accepted=dict(add=sampler.successes.get('add', 0), remove=sampler.successes.get('remove', 0)),
        final_rule_list=[[str(p.as_tuple()) for p in rule] for rule in sampler.current_state.rl.rules],
        comments=comments)))
if WORLD > 1:
    dist.destroy_process_group()
