"""Overlap experiment for detector evaluation on the S5 bench population (GPU box only): the values kernel
(instruction-issue bound) and the sort/emit kernel (HBM bound) pipelined over runs of start groups, with walk
CTAs small enough to share an SM with sort/emit blocks, the sort/emit grid capped per SM and its stream at
high priority.   usage: python tools/det_overlap.py [V]"""
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from mitre_b200 import device as D  # noqa: E402
from mitre_b200._lib import lib  # noqa: E402


def main():
    V = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
    sys.argv = sys.argv[:1]
    args = bench.parse_args()
    args.variables = V
    data, pop, marks = bench.build_population(args, 0, torch)
    wa = pop._walk_args
    N, S = wa[4], wa[3]
    ref = D.detectors_walk(*wa)
    base = min(bench.cuda_time_ms(torch, lambda: D.detectors_walk(*wa), 10) for _ in range(3))
    print("back to back (16 variables per walk CTA): %.3f ms" % base)
    # priority is fixed when the internal streams are created: set it before the first pipelined call
    prio = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    for vb, per_sm, stages in ((16, 0, 1), (8, 0, 1), (4, 0, 1), (8, 3, 4), (8, 3, 8), (8, 2, 8), (8, 4, 8), (4, 4, 8), (4, 3, 8),
                               (4, 5, 8), (8, 3, 16), (16, 0, 8)):
        lib().mitre_set_walk_tuning(vb, per_sm, prio)
        plan = D.values_walk_plan(N, S, V, pop._walk_first)
        if plan is None:
            print("vb %d: no plan" % vb)
            continue
        st = D.walk_stages(pop._walk_first, pop._group_base, pop._G, stages) if stages > 1 else None
        a = wa[:14] + (plan, st)
        out = D.detectors_walk(*a)
        same = all(torch.equal(x, y) for x, y in zip(ref, out))
        ms = min(bench.cuda_time_ms(torch, lambda: D.detectors_walk(*a), 10) for _ in range(3))
        print("vb %2d  sort/emit blocks per SM %d  priority %d  runs %2d  plan %s: %.3f ms  identical=%s" % (
            vb, per_sm, prio, 1 if st is None else len(st[0]) - 1, plan, ms, same))
    lib().mitre_set_walk_tuning(16, 0, 0)


if __name__ == "__main__":
    main()
