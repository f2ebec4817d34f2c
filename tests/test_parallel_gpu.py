"""NCCL parity of candidate sharding (needs >= 2 GPUs; skipped otherwise): sharded scoring +
all-gather equals the single-GPU sweep, and the two-stage draw picks the same index.

The log-likelihoods of a slice agree with the whole-table sweep to ~3e-16 relative, not bit for bit:
the scoring kernels share table lookups between the four consecutive rows a thread owns, so the
association of a row's masked sums depends on where its quad starts (and on which kernel variant the
slice's length selects).  The bound asserted here is 1e-12; the contract is 1e-9 (BASELINE.json)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, ws, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=ws, device_id=torch.device("cuda", rank))
    try:
        from mitre_b200 import device as D, parallel as PL
        from tests.golden import make_golden as MG
        c = MG.scoring_inputs(seed=321, N=35, p=4, P=200003)
        packed = D.pack_rows(c["truth"])
        mask = torch.as_tensor(D.pack_mask(c["mask"]), device=D.dev())
        full, status = D.score_all(packed, 35, c["X"], c["omega"], c["y"] - 0.5, 100.0, mask=mask)
        full = full.clone()
        D.raise_on_status(status)
        sh = PL.ShardedScorer.from_full_table(packed, 35)
        local, status = sh.score(c["X"], c["omega"], c["y"] - 0.5, 100.0, mask=mask)
        D.raise_on_status(status)
        gathered = sh.gather(local)
        same = bool(((gathered - full).abs() <= 1e-12 * full.abs()).all())
        idxs = []
        for u in [0.0, 0.31, 0.64, 0.98]:
            want = int(D.draw_index(full, None, u, stabilise=True).item())
            got, log_total = sh.draw(local, None, u)
            idxs.append((got, want, log_total))
        lse = float(torch.logsumexp(full, 0))
        q.put((rank, same, idxs, lse))
    finally:
        dist.destroy_process_group()


def test_sharded_scoring_matches_single_gpu():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ws = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, ws, port, q)) for r in range(ws)]
    for p in procs:
        p.start()
    out = [q.get(timeout=300) for _ in range(ws)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, same, idxs, lse in out:
        assert same
        for got, want, log_total in idxs:
            assert abs(got - want) <= 1
            assert abs(log_total - lse) < 1e-9
    assert [x[0] for x in out[0][2]] == [x[0] for x in out[1][2]]


# ---- variable-sharded population build (detector evaluation) ---------------------------

def _population_case():
    from mitre_b200 import rules as R, synthetic
    d = synthetic.make_series(N=23, V=11, T=(5, 9), span=(0.0, 40.0), seed=9)
    data = R.Dataset(d["X"], d["T"], d["y"], ["v%d" % i for i in range(11)], d["variable_weights"],
                     d["experiment_start"], d["experiment_end"])
    return data, dict(tmin=1.0, N_intervals=5, tmax=None)


def test_sharded_population_world1_equals_plain_population():
    """Without a process group the sharded build is the whole build: same sorted table, same
    permutation, same primitive descriptors as DiscreteRulePopulation."""
    import torch
    from mitre_b200 import parallel as PL, rules as R
    data, cfg = _population_case()
    pop = R.DiscreteRulePopulation(None, cfg["tmin"], cfg["N_intervals"], tmax=cfg["tmax"], data=data)
    sp = PL.ShardedPopulation(data, cfg["tmin"], cfg["N_intervals"], tmax=cfg["tmax"])
    assert sp.P == len(pop) and (sp.start, sp.stop) == (0, len(pop))
    assert torch.equal(sp.local_rows, pop.packed_truth)
    assert np.array_equal(sp.local_perm.cpu().numpy(), pop._perm_host())
    for k in [0, 1, len(pop) // 3, len(pop) - 1]:
        assert sp.describe(k) == pop._tuple_at(k)


def _pop_worker(rank, ws, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=ws, device_id=torch.device("cuda", rank))
    try:
        from mitre_b200 import device as D, parallel as PL, rules as R
        data, cfg = _population_case()
        pop = R.DiscreteRulePopulation(None, cfg["tmin"], cfg["N_intervals"], tmax=cfg["tmax"], data=data)
        sp = PL.ShardedPopulation(data, cfg["tmin"], cfg["N_intervals"], tmax=cfg["tmax"])
        full = pop.packed_truth
        ok_rows = bool(torch.equal(sp.local_rows, full[sp.start:sp.stop]))
        ok_perm = bool(np.array_equal(sp.local_perm.cpu().numpy(), pop._perm_host()[sp.start:sp.stop]))
        ks = [0, 5, len(pop) // 2, len(pop) - 1]
        ok_desc = all(sp.describe(k) == pop._tuple_at(k) for k in ks)
        # the slice scores like the same slice of the single-GPU table
        from tests.golden import make_golden as MG
        c = MG.scoring_inputs(seed=5, N=data.n_subjects, p=3, P=4)
        whole, status = D.score_all(full, data.n_subjects, c["X"], c["omega"], c["y"] - 0.5, 100.0)
        whole = whole.clone()
        local, status = sp.scorer().score(c["X"], c["omega"], c["y"] - 0.5, 100.0)
        ref = whole[sp.start:sp.stop]
        ok_score = bool(((local - ref).abs() <= 1e-12 * ref.abs()).all())
        q.put((rank, sp.P, len(pop), (sp.v0, sp.v1), ok_rows, ok_perm, ok_desc, ok_score))
    finally:
        dist.destroy_process_group()


def test_sharded_population_two_gpus_matches_single_gpu_build():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ws = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_pop_worker, args=(r, ws, port, q)) for r in range(ws)]
    for p in procs:
        p.start()
    out = [q.get(timeout=300) for _ in range(ws)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(o[3] for o in out) == [(0, 6), (6, 11)]
    for rank, P, P_plain, _, ok_rows, ok_perm, ok_desc, ok_score in out:
        assert P == P_plain and ok_rows and ok_perm and ok_desc and ok_score


# ---- streaming sweep with its tiles dealt to the ranks ---------------------------------------------

def _stream_worker(rank, ws, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=ws, device_id=torch.device("cuda", rank))
    try:
        from mitre_b200 import rules as R, streaming, synthetic
        solo = [dist.new_group([r]) for r in range(ws)][rank]        # every rank must create every group
        d = synthetic.make_series(N=90, V=10, T=(5, 9), span=(0.0, 40.0), seed=21)
        data = R.Dataset(d["X"], d["T"], d["y"], ["v%d" % i for i in range(10)], d["variable_weights"],
                         d["experiment_start"], d["experiment_end"])
        rng = np.random.RandomState(4)
        N = 90
        X = np.concatenate([(rng.rand(N, 2) > 0.5).astype(float), np.ones((N, 1))], axis=1)
        omega = rng.gamma(2.0, 0.125, size=N) + 0.05
        y = rng.rand(N) > 0.5
        mask = rng.rand(N) > 0.2
        tables = (rng.normal(size=10), rng.normal(size=64), 0.3)
        res = {}
        for name, group in (("solo", solo), ("all", None)):
            sw = streaming.StreamingSweep(data, 1.0, 4, tmax=None, tile_variables=3, group=group)
            tabs = (tables[0], tables[1][:len(sw.windows)], tables[2])
            log_total, n = sw.sweep(X, omega, y - 0.5, 100.0, mask=mask, prior_tables=tabs)
            draws = [sw.draw(u) for u in (0.05, 0.5, 0.95)]
            res[name] = (log_total, n, [(t, k, repr(tup)) for t, k, tup in draws], sw.ws)
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


def test_streaming_sweep_two_gpus_matches_one():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ws = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_stream_worker, args=(r, ws, port, q)) for r in range(ws)]
    for p in procs:
        p.start()
    out = [q.get(timeout=300) for _ in range(ws)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, res in out:
        assert res["solo"][3] == 1 and res["all"][3] == 2
        assert res["all"][1] == res["solo"][1]
        assert abs(res["all"][0] - res["solo"][0]) <= 1e-12 * abs(res["solo"][0])
        assert res["all"][2] == res["solo"][2]
    assert out[0][1]["all"] == out[1][1]["all"]


# ---- ShardedSweep: score + (max, sum-exp) exchange + owner draw + index back, on every rank -----------

def _sweep_case(P=300007, seed=77):
    """A lexsorted table with runs of equal rows, a prior with -inf entries, sweep inputs."""
    from tests.golden import make_golden as MG
    rng = np.random.RandomState(seed)
    c = MG.scoring_inputs(seed=seed, N=35, p=4, P=8)
    distinct = rng.rand(P // 6, 35) > 0.5
    rows = distinct[rng.randint(distinct.shape[0], size=P)]
    rows = rows[np.lexsort(np.rot90(rows))]
    prior = rng.normal(-12, 2, size=P)
    prior[::13] = -np.inf
    return c, rows, prior


def _sweep_reference(c, rows, prior, us):
    """Single-table answer: log-likelihoods, log-sum-exp and np.cumsum draw of lik + prior."""
    import torch
    from mitre_b200 import device as D
    packed = D.pack_rows(rows)
    mask = torch.as_tensor(D.pack_mask(c["mask"]), device=D.dev())
    out, status = D.score_all(packed, 35, c["X"], c["omega"], c["y"] - 0.5, 100.0, mask=mask)
    D.raise_on_status(status)
    w = out.cpu().numpy() + prior
    cum = np.cumsum(np.exp(w - w.max()))
    lse = float(w.max() + np.log(cum[-1]))
    return packed, out.clone(), [int(np.sum(cum < u * cum[-1])) for u in us], lse


US = [0.0, 0.07, 0.31, 0.5, 0.64, 0.98, 0.999999]


@pytest.mark.parametrize("exchange", ["p2p", "nccl"])
def test_sharded_sweep_single_rank(exchange):
    """world = 1: the whole sharded step (CUDA graph replay in p2p mode) against numpy's cumsum draw."""
    import torch
    from mitre_b200 import device as D, parallel as PL
    from mitre_b200._lib import lib
    c, rows, prior = _sweep_case()
    for min_rows in (0, 4 << 20):        # duplicate-aware large-table kernel / shared-memory-table kernel
        lib().mitre_set_large_table_min_rows(min_rows)
        try:
            packed, out, want, lse = _sweep_reference(c, rows, prior, US)
            sw = PL.ShardedSweep(packed, torch.as_tensor(prior, device=D.dev()), packed.shape[0], 35, 0,
                                 exchange=exchange, max_p=4)
            for u, k in zip(US, want):
                idx, log_total = sw.step(c["X"], c["omega"], c["y"] - 0.5, c["mask"], u, 100.0)
                assert idx == k, (exchange, u, idx, k)
                assert abs(log_total - lse) < 1e-10
            assert torch.equal(sw.buffers.out[:sw.n], out)
            sw.close()
        finally:
            lib().mitre_set_large_table_min_rows(4 << 20)


def _sweep_worker(rank, ws, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=ws, device_id=torch.device("cuda", rank))
    try:
        from mitre_b200 import device as D, parallel as PL
        from mitre_b200._lib import lib
        lib().mitre_set_large_table_min_rows(0)
        c, rows, prior = _sweep_case()
        packed, out, want, lse = _sweep_reference(c, rows, prior, US)
        P = packed.shape[0]
        a, b = PL.shard_bounds(P, ws, rank)
        res = {}
        for exchange in ("p2p", "nccl"):
            sw = PL.ShardedSweep(packed[a:b].contiguous(), torch.as_tensor(prior[a:b], device=D.dev()), P, 35, a,
                                 exchange=exchange, max_p=4)
            got = [sw.step(c["X"], c["omega"], c["y"] - 0.5, c["mask"], u, 100.0) for u in US]
            same = bool(torch.equal(sw.buffers.out[:sw.n], out[a:b]))
            sw.close()
            res[exchange] = (got, same)
        q.put((rank, res, want, lse))
    finally:
        dist.destroy_process_group()


def test_sharded_sweep_two_gpus():
    """2 ranks, both exchange modes: every rank ends every step with the index np.cumsum gives on the whole
    table and the global log-sum-exp; slice log-likelihoods are bit-identical to the whole-table sweep."""
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ws = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sweep_worker, args=(r, ws, port, q)) for r in range(ws)]
    for p in procs:
        p.start()
    out = [q.get(timeout=300) for _ in range(ws)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, res, want, lse in out:
        for exchange, (got, same) in res.items():
            assert same, (rank, exchange)
            assert [g[0] for g in got] == want, (rank, exchange, got, want)
            for _, log_total in got:
                assert abs(log_total - lse) < 1e-10
