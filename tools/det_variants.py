"""Times the detector-evaluation launch variants on the S5 bench population (GPU box only).
usage: python tools/det_variants.py [V]"""
import sys
import types

import numpy as np
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
    fa = pop._fused_args

    def run(split):
        lib().mitre_set_detector_split(split)
        out = D.detectors_fused(*fa)
        ms = min(bench.cuda_time_ms(torch, lambda: D.detectors_fused(*fa), 10) for _ in range(3))
        return out, ms

    for _ in range(100):    # clocks up
        D.detectors_fused(*fa)
    ref, ms0 = run(0)
    print("single fused kernel: %.3f ms" % ms0)
    for split, name in ((2, "flat values + staged sort/emit"), (3, "tile values + staged sort/emit"),
                        (4, "default values + direct sort/emit"), (5, "dry"), (1, "default")):
        out, ms = run(split)
        same = all(torch.equal(a, b) for a, b in zip(ref, out))
        print("%-36s %.3f ms  identical=%s" % (name, ms, same))
    lib().mitre_set_detector_split(1)
    if pop._walk_args is not None:
        wa = pop._walk_args
        out = D.detectors_walk(*wa)
        ms = min(bench.cuda_time_ms(torch, lambda: D.detectors_walk(*wa), 10) for _ in range(3))
        same = all(torch.equal(a, b) for a, b in zip(ref[1:4], out[1:4]))
        means_same = torch.equal(ref[0][1::2], out[0][1::2]) if bool(wa[10].all()) else None
        print("%-36s %.3f ms  thresholds/K/rows identical=%s  means identical=%s" % ("walk values + staged sort/emit", ms,
                                                                                    same, means_same))
        lib().mitre_set_sort_emit_rolled(1)
        o3 = D.detectors_walk(*wa)
        ms = min(bench.cuda_time_ms(torch, lambda: D.detectors_walk(*wa), 10) for _ in range(3))
        print("%-36s %.3f ms  identical to the unrolled sort/emit=%s" % ("walk values + ROLLED sort/emit", ms,
                                                                        all(torch.equal(a, b) for a, b in zip(out, o3))))
        lib().mitre_set_sort_emit_rolled(0)
        for ns in (1, 4):
            stages = D.walk_stages(pop._walk_first, pop._group_base, pop._G, ns) if ns > 1 else None
            args = wa[:15] + (stages,)
            o2 = D.detectors_walk(*args)
            ms = min(bench.cuda_time_ms(torch, lambda: D.detectors_walk(*args), 10) for _ in range(3))
            same = all(torch.equal(a, b) for a, b in zip(out, o2))
            print("    pipelined over %2d runs of start groups: %.3f ms  identical=%s" % (
                1 if stages is None else len(stages[0]) - 1, ms, same))
        tbar_ms = bench.cuda_time_ms(torch, lambda: D.window_tbar(wa[1], wa[6], wa[7]['lo'], wa[7]['cnt']), 10)
        import ctypes
        from mitre_b200.device import _p, _stream
        vals = out[0]
        flags = torch.zeros(wa[12], dtype=torch.uint8, device=vals.device)
        def values_only():
            lib().mitre_detector_values_walk(_stream(), _p(wa[0]), wa[0].shape[1], _p(wa[1]), _p(wa[2]), int(wa[3]),
                                             int(wa[4]), int(wa[5]), _p(wa[6]), wa[6].shape[0], _p(wa[7]['lo']),
                                             _p(wa[7]['cnt']), _p(wa[7]['inv_sxx']), _p(wa[8]), _p(wa[9]), _p(wa[10]),
                                             _p(wa[11]), wa[11].shape[0] - 1, int(wa[14][0]), int(wa[14][1]), _p(vals), _p(flags))
        print("  values_walk_kernel alone            %.3f ms  (window_tbar %.3f ms)" % (
            min(bench.cuda_time_ms(torch, values_only, 10) for _ in range(3)), tbar_ms))
    N = fa[1]
    padded = ref[3]
    D.lexsort_rows(padded.view(-1, 1), N + 1)
    ms = min(bench.cuda_time_ms(torch, lambda: D.lexsort_rows(padded.view(-1, 1), N + 1), 3) for _ in range(2))
    print("lexsort of the padded table (%d slots), general path over N + 1 bits: %.2f ms" % (padded.numel(), ms))
    K = ref[2]
    off = D.detector_row_offsets(K)
    P = int(off[-1].item())
    p1, s1 = D.lexsort_padded(padded, N, P, off)
    ms = min(bench.cuda_time_ms(torch, lambda: D.lexsort_padded(padded, N, P, off), 3) for _ in range(2))
    lib().mitre_set_packed_sort(0)
    p0, s0 = D.lexsort_padded(padded, N, P, off)
    ms0 = min(bench.cuda_time_ms(torch, lambda: D.lexsort_padded(padded, N, P, off), 3) for _ in range(2))
    lib().mitre_set_packed_sort(1)
    print("sort_base_rules from the padded table (%d valid rows): packed one-sweep %.2f ms, general + densify %.2f ms, "
          "identical=%s" % (P, ms, ms0, torch.equal(p0, p1) and torch.equal(s0, s1)))


if __name__ == "__main__":
    main()
