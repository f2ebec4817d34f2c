"""A chain of the reference's full sampler step on the cfg5 problem (10 000 variables, 2.9e8 candidates) with the
candidate table sharded over the ranks (mitre_b200/sharded_sampler.py).
usage: [torchrun ...] python tools/s5_chain.py [V=10000] [steps=50]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, '.')
from mitre_b200 import logit_rules as LR, polyagamma, rules as R, sharded_sampler as SS, synthetic  # noqa: E402

WORLD = int(os.environ.get("WORLD_SIZE", "1"))
RANK = int(os.environ.get("RANK", "0"))
if WORLD > 1:
    import torch.distributed as dist
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
V = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 50
d = synthetic.make_shape("S5", seed=0, V=V)
data = R.Dataset(d["X"], d["T"], d["y"], ["v%d" % i for i in range(V)], d["variable_weights"],
                 d["experiment_start"], d["experiment_end"])
t0 = time.perf_counter()
model = SS.make_sharded_model(data, d["tmin"], N_intervals=d["N_intervals"], tmax=d["tmax"])
torch.cuda.synchronize()
setup_s = time.perf_counter() - t0
first = R.PrimitiveRule(*model.spop.describe(model.spop.P // 3))
np.random.seed(7)
LR.ppg = polyagamma.PyPolyaGamma(7)
sampler = SS.make_sharded_sampler(model, R.RuleList([[first]]))
sampler.sample(3)                          # warm-up: allocations, graph captures
torch.cuda.synchronize()
done = len(sampler.comments)
prof = None
if os.environ.get("PROFILE") and RANK == 0:
    import cProfile
    prof = cProfile.Profile()
    prof.enable()
t0 = time.perf_counter()
sampler.sample(STEPS)
torch.cuda.synchronize()
wall = time.perf_counter() - t0
if prof is not None:
    import io
    import pstats
    prof.disable()
    for key in ("tottime", "cumtime"):
        buf = io.StringIO()
        pstats.Stats(prof, stream=buf).sort_stats(key).print_stats(22)
        sys.stderr.write(buf.getvalue()[:4200])
comments = sampler.comments[done:]
kind = lambda c: ('update' if 'replacing' in c else 'add' if c.startswith('add') else
                  'remove' if c.startswith('remove') or 'Nothing to remove' in c else 'move' if 'move' in c else c)
moves = {}
for c in comments:
    moves[kind(c)] = moves.get(kind(c), 0) + 1
if RANK == 0:
    print(json.dumps(dict(
        what="LogisticRuleSampler.step on the cfg5 problem, candidate table sharded over the ranks", n_gpus=WORLD,
        variables=V, candidates=model.spop.P, setup_s=setup_s, steps=STEPS, s_per_step=wall / STEPS,
        samples_per_s=STEPS / wall, moves=moves,
        accepted=dict(add=sampler.successes.get('add', 0), remove=sampler.successes.get('remove', 0)),
        final_rule_list=[[str(p.as_tuple()) for p in rule] for rule in sampler.current_state.rl.rules])))
if WORLD > 1:
    dist.destroy_process_group()
