"""Per-kernel times of sort_base_rules on the S5 bench population (GPU box only):
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/sort_launches.csv python tools/sort_time.py [V]
Without ncu it prints the CUDA-event time of the packed and of the general path."""
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
    N = pop.data.n_subjects
    out = D.detectors_walk(*pop._walk_args)
    padded, K = out[3], out[2]
    off = D.detector_row_offsets(K)
    P = int(off[-1].item())
    D.lexsort_padded(padded, N, P, off)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    D.lexsort_padded(padded, N, P, off)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    ms = min(bench.cuda_time_ms(torch, lambda: D.lexsort_padded(padded, N, P, off), 3) for _ in range(2))
    lib().mitre_set_packed_sort(0)
    ms0 = min(bench.cuda_time_ms(torch, lambda: D.lexsort_padded(padded, N, P, off), 3) for _ in range(2))
    lib().mitre_set_packed_sort(1)
    print("sort_base_rules, %d slots, %d valid rows: packed one-sweep %.2f ms, general + densify %.2f ms" % (
        padded.numel(), P, ms, ms0))
    # the values kernel with one and with two variables per thread
    for vpt in (1, 2):
        lib().mitre_set_walk_vpt(vpt)
        D.detectors_walk(*pop._walk_args)
        t = min(bench.cuda_time_ms(torch, lambda: D.detectors_walk(*pop._walk_args), 10) for _ in range(3))
        print("walk (vpt %d) + sort/emit: %.3f ms" % (vpt, t))
    lib().mitre_set_walk_vpt(1)


if __name__ == "__main__":
    main()
