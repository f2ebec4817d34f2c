"""Host logic of the multi-GPU path under gloo, world_size 2, on CPU: shard
bounds, the uneven all-gather of log-likelihood slices, the two-stage (max, sum-exp) draw."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mitre_b200 import parallel as PL


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_shard_bounds_cover_and_balance():
    for P in [0, 1, 7, 64, 1000003]:
        for ws in [1, 2, 3, 8]:
            b = PL.all_shard_bounds(P, ws)
            assert b[0][0] == 0 and b[-1][1] == P
            assert all(b[i][1] == b[i + 1][0] for i in range(ws - 1))
            sizes = [y - x for x, y in b]
            assert max(sizes) - min(sizes) <= 1


def test_chains_for_rank_seeds():
    got = sorted(sum([PL.chains_for_rank(64, 8, r, base_seed=100) for r in range(8)], []))
    assert got == [(i, 100 + i) for i in range(64)]
    assert len(PL.chains_for_rank(64, 8, 3)) == 8


def test_owner_of_marker_matches_flat_cumsum():
    rng = np.random.RandomState(0)
    for trial in range(200):
        ws = rng.randint(1, 6)
        sizes = rng.randint(1, 50, size=ws)
        w = [rng.normal(-80, 4, size=n) for n in sizes]
        stats = np.array([[x.max(), np.exp(x - x.max()).sum()] for x in w])
        flat = np.concatenate(w)
        u = rng.rand()
        c = np.cumsum(np.exp(flat - flat.max()))
        want = int(np.sum(c < u * c[-1]))
        owner, u_local, log_total = PL.owner_of_marker(stats, u)
        starts = np.concatenate([[0], np.cumsum(sizes)])
        cl = np.cumsum(np.exp(w[owner] - w[owner].max()))
        got = int(starts[owner] + np.sum(cl < u_local * cl[-1]))
        assert abs(got - want) <= 1
        assert abs(log_total - (flat.max() + np.log(np.exp(flat - flat.max()).sum()))) < 1e-10


def test_owner_of_marker_edge_cases():
    # empty / all -inf ranks are skipped; the reference's raw-exp underflow gives index 0
    stats = np.array([[-np.inf, 0.0], [-3.0, 2.0], [-np.inf, 0.0]])
    owner, u_local, _ = PL.owner_of_marker(stats, 0.4)
    assert owner == 1 and 0.0 <= u_local < 1.0
    owner, u_local, lt = PL.owner_of_marker(np.array([[-4000.0, 5.0], [-4100.0, 7.0]]), 0.7, stabilise=False)
    assert owner == 0 and lt == -np.inf
    owner, u_local, lt = PL.owner_of_marker(np.array([[-4000.0, 5.0], [-4100.0, 7.0]]), 0.7, stabilise=True)
    assert owner == 0 and np.isfinite(lt)


def _worker(rank, ws, port, P, seed, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        rng = np.random.RandomState(seed)
        full = rng.normal(-90, 5, size=P)
        a, b = PL.shard_bounds(P, ws, rank)
        local = torch.as_tensor(full[a:b].copy())
        got = PL.allgather_loglik(local, P)
        ok_gather = bool(np.array_equal(got.numpy(), full))
        results = []
        for u in [0.0, 0.13, 0.5, 0.77, 0.999]:
            x = full[a:b]
            stats = torch.tensor([x.max(), np.exp(x - x.max()).sum()], dtype=torch.float64)

            def local_draw(u_local, x=x):
                c = np.cumsum(np.exp(x - x.max()))
                return int(np.sum(c < u_local * c[-1]))
            idx, log_total = PL.two_stage_draw(stats, local_draw, a, u)
            c = np.cumsum(np.exp(full - full.max()))
            want = int(np.sum(c < u * c[-1]))
            results.append((idx, want, log_total))
        q.put((rank, ok_gather, results))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("P", [7, 1000, 4097])
def test_gloo_world2_gather_and_two_stage_draw(P):
    ws = 2
    port = free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, ws, port, P, 11, q)) for r in range(ws)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in range(ws)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    by_rank = {r: (g, res) for r, g, res in out}
    assert by_rank[0][0] and by_rank[1][0]
    assert [x[0] for x in by_rank[0][1]] == [x[0] for x in by_rank[1][1]]       # every rank agrees
    for idx, want, log_total in by_rank[0][1]:
        assert abs(idx - want) <= 1
        assert np.isfinite(log_total)


# ---- variable-sharded build: host arithmetic ------------------------------------------------------

def _brute_force_enumeration(ok, V, K_of):
    """(window, variable, type) groups in the reference's enumeration order (rules.py:585-599:
    window, then variable, then ('slope', 'average') with slope only where allowed) and the
    enumeration index of every (group, threshold, direction) row (rules.py:465-511)."""
    groups, rows = [], []
    for w, o in enumerate(ok):
        for v in range(V):
            for t in ((0, 1) if o else (1,)):
                groups.append((w, v, t))
    e = 0
    first = []
    for g in groups:
        first.append(e)
        e += 2 * K_of[g]
    return groups, first, e


@pytest.mark.parametrize("V,ws", [(7, 2), (10, 4), (5, 8), (3, 1)])
def test_global_group_ids_and_enumeration_indices(V, ws):
    rng = np.random.RandomState(V * 10 + ws)
    ok = rng.rand(6) < 0.6
    K_of = {(w, v, t): int(rng.randint(0, 5)) for w in range(6) for v in range(V) for t in (0, 1)}
    groups, first, P = _brute_force_enumeration(ok, V, K_of)
    index_of_group = {g: i for i, g in enumerate(groups)}
    row_offsets = np.array(first + [P], dtype=np.int64)
    seen = np.zeros(P, dtype=bool)
    covered = 0
    for rank in range(ws):
        v0, v1 = PL.variable_bounds(V, ws, rank)
        g_global, GB = PL.global_group_ids(ok, V, v0, v1)
        local_groups = [(w, v, t) for w, o in enumerate(ok) for v in range(v0, v1) for t in ((0, 1) if o else (1,))]
        assert [index_of_group[g] for g in local_groups] == list(g_global)
        assert GB[-1] == len(groups)
        for g, gid in zip(local_groups, g_global):
            assert PL.describe_group(int(gid), GB, ok, V) == g
        K_local = np.array([K_of[g] for g in local_groups], dtype=np.int32)
        e = PL.global_enumeration_indices(K_local, g_global, row_offsets)
        want = np.concatenate([first[index_of_group[g]] + np.arange(2 * K_of[g]) for g in local_groups] +
                              [np.zeros(0, dtype=np.int64)]).astype(np.int64)
        assert np.array_equal(e, want)
        assert np.all(np.diff(e) > 0)            # local enumeration order is global enumeration order
        assert not seen[e].any()
        seen[e] = True
        covered += v1 - v0
    assert covered == V and seen.all()


def _ragged_worker(rank, ws, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        t1 = torch.arange(3 + 2 * rank, dtype=torch.int64) + 100 * rank
        t2 = (torch.arange((rank * 4) * 2, dtype=torch.int64) + 1000 * rank).view(-1, 2)
        p1 = PL._allgather_ragged(t1)
        p2 = PL._allgather_ragged(t2)
        q.put((rank, [x.tolist() for x in p1], [x.tolist() for x in p2]))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_ragged_allgather():
    ws = 2
    port = free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ragged_worker, args=(r, ws, port, q)) for r in range(ws)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in range(ws)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, p1, p2 in out:
        assert p1 == [[0, 1, 2], [100, 101, 102, 103, 104]]
        assert p2 == [[], [[1000, 1001], [1002, 1003], [1004, 1005], [1006, 1007]]]


def test_streaming_tile_pairs_are_taken_from_their_owner():
    from mitre_b200 import streaming
    ws, n_tiles = 3, 8
    truth = torch.arange(n_tiles * 2, dtype=torch.float64).reshape(n_tiles, 2) + 1.0
    gathered = torch.zeros((ws, n_tiles, 2), dtype=torch.float64)
    gathered[:, :, 0] = -np.inf
    for t in range(n_tiles):
        gathered[t % ws, t] = truth[t]
    assert torch.equal(streaming.combine_tile_pairs(gathered), truth)
    assert torch.equal(streaming.combine_tile_pairs(truth[None]), truth)          # world of one


def test_sharded_view_count_matrix_and_group_tables_against_brute_force():
    """sharded_sampler._ShardedPopulationView derives the prior's (window, variable) count matrix and the
    per-group (variable, window) tables from the global group layout (window-major, variable, slope before
    average where the window admits slopes): brute-force enumeration on a stub population."""
    import types
    from mitre_b200 import sharded_sampler as SS
    rng = np.random.RandomState(2)
    V, ok = 7, np.array([True, False, True, True, False])
    W = len(ok)
    groups = [(w, v, t) for w in range(W) for v in range(V) for t in ((0, 1) if ok[w] else (1,))]
    K = rng.randint(0, 6, size=len(groups))
    row_offsets = np.concatenate([[0], np.cumsum(2 * K)]).astype(np.int64)
    nt = 1 + ok.astype(np.int64)
    GB = np.concatenate([[0], np.cumsum(V * nt)]).astype(np.int64)
    spop = types.SimpleNamespace(windows=[(float(w), float(w) + 1.0) for w in range(W)], ok_slope=ok, V=V,
                                 G=len(groups), GB=GB, _row_offsets_host=row_offsets, P=int(row_offsets[-1]))
    view = SS._ShardedPopulationView.__new__(SS._ShardedPopulationView)
    view.spop, view.windows, view._counts = spop, spop.windows, None
    want = np.zeros((W, V))
    for (w, v, t), k in zip(groups, K):
        want[w, v] += 2 * k
    assert np.array_equal(view.count_matrix(), want)
    assert view.count_matrix().sum() == spop.P
    # group_tables builds device tensors: check the same arithmetic on the host
    w_of = np.repeat(np.arange(W, dtype=np.int64), V * nt)
    within = np.arange(spop.G, dtype=np.int64) - GB[w_of]
    assert [(int(a), int(b)) for a, b in zip(w_of, within // nt[w_of])] == [(w, v) for w, v, t in groups]
