"""values_walk_kernel with smaller CTAs (several resident per SM): times (variables per CTA, start groups per
CTA) combinations on the S5 bench population.  usage: python tools/walk_plans.py [V]"""
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
    first = pop._walk_first
    n_sg = len(first) - 1
    ref = D.detectors_walk(*wa)
    base = min(bench.cuda_time_ms(torch, lambda: D.detectors_walk(*wa), 10) for _ in range(3))
    print("default plan %s: pair %.3f ms" % (wa[14], base))
    for vb, flags in ((16, 0), (16, 2)):
        lib().mitre_set_walk_tuning(vb, 0, flags)
        print("CTAs persistent over the variable tiles" if flags == 2 else "one CTA per (tile, chunk)")
        for per in (3, 4, 6, 8, 12):
            widest = max(int(first[min(k + per, n_sg)] - first[k]) for k in range(0, n_sg, per))
            smem = lib().mitre_values_walk_smem_bytes(int(N), int(S), widest)
            if not 0 < smem <= 227 * 1024:
                continue
            a = wa[:14] + ((per, widest), None)
            out = D.detectors_walk(*a)
            same = all(torch.equal(x, y) for x, y in zip(ref, out))
            ms = min(bench.cuda_time_ms(torch, lambda: D.detectors_walk(*a), 10) for _ in range(3))
            print("vb %2d  start groups per CTA %d (windows <= %3d, smem %3d KB): pair %.3f ms  identical=%s" % (
                vb, per, widest, smem // 1024, ms, same))
    lib().mitre_set_walk_tuning(16, 0, 0)


if __name__ == "__main__":
    main()
