"""One scoring sweep per variant on the S5 bench population, for ncu captures (GPU box only).
usage: python tools/ncu_score.py [V] [dedupe mode] [stats 0/1]"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from mitre_b200 import device as D  # noqa: E402
from mitre_b200._lib import lib  # noqa: E402


def main():
    V = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
    mode = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    with_stats = int(sys.argv[3]) if len(sys.argv) > 3 else 0
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
    buffers = D.ScoreBuffers(P, N, args.rules + 1)
    draw = D.DrawBuffers(P)
    prior = torch.zeros(P, dtype=torch.float64, device=dev)
    lib().mitre_set_dedupe(mode)
    for _ in range(4):
        if with_stats:
            D.score_all_stats(packed, N, dX, dom, dka, 100.0, mask=dmask, buffers=buffers, log_prior=prior,
                              draw_buffers=draw)
        else:
            D.score_all(packed, N, dX, dom, dka, 100.0, mask=dmask, buffers=buffers)
    torch.cuda.synchronize()
    print("done P=%d" % P)


if __name__ == "__main__":
    main()
