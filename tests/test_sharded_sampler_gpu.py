"""The reference's full sampler step on a candidate-SHARDED population (mitre_b200/sharded_sampler.py) against
the resident sampler: the sharded table keeps the reference's global sorted order, so with the same seed the
sharded chain must make the resident chain's moves -- same comments (move types, positions, candidate indices,
acceptance ratios), same rule lists, same hyperparameter trajectories.  One rank here; two ranks (NCCL + peer
memory exchange) when two GPUs are visible."""
import os
import socket

import numpy as np
import pytest

from mitre_b200 import synthetic

pytestmark = pytest.mark.gpu
STEPS = 60


def _data(N=24, V=11, seed=4):
    from mitre_b200 import rules as R
    d = synthetic.make_series(N=N, V=V, T=(5, 9), span=(0.0, 40.0), seed=seed, allow_constant=False)
    return R.Dataset(d["X"], d["T"], d["y"], ["v%d" % i for i in range(V)], d["variable_weights"],
                     d["experiment_start"], d["experiment_end"])


def _trace(sampler):
    states = [([[p.as_tuple() for p in rule] for rule in st.rl.rules], st.phylogeny_mean, st.phylogeny_std,
               st.window_typical_fraction, st.window_concentration, st.empty_probability) for st in sampler.states]
    return list(sampler.comments), states, [float(x) for x in sampler.priors]


def _resident_trace(seed):
    from mitre_b200 import logit_rules as LR, polyagamma, rules as R
    model = LR.LogisticRuleModel(_data(), 1.0, N_intervals=4)
    flat = model.rule_population.flat_rules
    r0 = R.RuleList([[R.PrimitiveRule(*flat[13])]])
    np.random.seed(seed)
    LR.ppg = polyagamma.PyPolyaGamma(seed)
    s = LR.LogisticRuleSampler(model, r0, device_draw=True, device_gibbs=True)
    s.sample(STEPS)
    return flat[13], _trace(s)


def _sharded_trace(seed, first, group=None):
    from mitre_b200 import logit_rules as LR, polyagamma, rules as R, sharded_sampler as SS
    model = SS.make_sharded_model(_data(), 1.0, group=group, N_intervals=4)
    r0 = R.RuleList([[R.PrimitiveRule(*first)]])
    np.random.seed(seed)
    LR.ppg = polyagamma.PyPolyaGamma(seed)
    s = SS.make_sharded_sampler(model, r0)
    s.sample(STEPS)
    return _trace(s)


def _same(a, b):
    ca, sa, pa = a
    cb, sb, pb = b
    assert ca == cb
    for x, y in zip(sa, sb):
        assert x[0] == y[0]
        assert np.allclose(x[1:], y[1:], rtol=1e-12, atol=0)
    assert np.allclose(pa, pb, rtol=1e-9)


def test_sharded_chain_on_one_rank_is_the_resident_chain():
    first, want = _resident_trace(21)
    kinds = ' '.join(want[0])
    assert 'replacing' in kinds and 'add detector' in kinds and 'remove detector' in kinds and 'moves to' in kinds
    _same(want, _sharded_trace(21, first))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, ws, port, first, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=ws, device_id=torch.device("cuda", rank))
    try:
        q.put((rank, _sharded_trace(21, first)))
    finally:
        dist.destroy_process_group()


def test_sharded_chain_on_two_gpus_is_the_resident_chain():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    first, want = _resident_trace(21)
    ws = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, ws, port, first, q)) for r in range(ws)]
    for p in procs:
        p.start()
    out = [q.get(timeout=600) for _ in range(ws)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, got in out:
        _same(want, got)
