"""NCCL parity of candidate sharding (needs >= 2 GPUs; skipped otherwise): sharded scoring +
all-gather equals the single-GPU sweep bit for bit, and the two-stage draw picks the same index."""
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
        same = bool(torch.equal(gathered, full))
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
