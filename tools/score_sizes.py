"""Sweep time against table size (prefix slices of the S5 bench table): fixed cost vs per-candidate cost.
usage: python tools/score_sizes.py [V]"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from mitre_b200 import device as D  # noqa: E402


def main():
    V = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
    sys.argv = sys.argv[:1]
    args = bench.parse_args()
    args.variables = V
    data, pop, marks = bench.build_population(args, 0, torch)
    N, P = data.n_subjects, len(pop)
    packed = pop.packed_truth
    state = bench.sweep_state(N, args.rules, lambda rng: pop._truth_row(int(rng.randint(P))), 1)
    dev = D.dev()
    dX = torch.as_tensor(state["X"], device=dev)
    dom = torch.as_tensor(state["omega"], device=dev)
    dka = torch.as_tensor(state["kappa"], device=dev)
    dmask = torch.as_tensor(D.pack_mask(state["mask"]), device=dev)
    prior = torch.randn(P, dtype=torch.float64, device=dev) - np.log(P)
    for frac in (1, 2, 4, 8, 16, 32, 64):
        n = (P // frac) & ~127
        rows = packed[:n]
        buffers, draw = D.ScoreBuffers(n, N, args.rules + 1), D.DrawBuffers(n)
        pr = prior[:n]

        def stats():
            D.score_all_stats(rows, N, dX, dom, dka, 100.0, mask=dmask, buffers=buffers, log_prior=pr,
                              draw_buffers=draw)

        def plain():
            D.score_all(rows, N, dX, dom, dka, 100.0, mask=dmask, buffers=buffers)
        for _ in range(5):
            stats()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            stats()
        g2 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g2):
            plain()
        ts = min(bench.cuda_time_ms(torch, g.replay, 20) for _ in range(3))
        tp = min(bench.cuda_time_ms(torch, g2.replay, 20) for _ in range(3))
        dup = float((rows[1:, 0] == rows[:-1, 0]).double().mean())
        print("P/%-3d = %10d rows  dup %.3f  stats %.4f ms (%.2f ns/kcand)  plain %.4f ms" % (
            frac, n, dup, ts, 1e9 * ts / n, tp))


if __name__ == "__main__":
    main()
