"""Detector evaluation over subject counts (run under ncu with -k regex:sort_emit to see how the
sort/emit kernel's issue rate depends on its code size).  usage: python tools/det_by_n.py N [N ...]"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from mitre_b200 import device as D, rules as R, synthetic  # noqa: E402


def main():
    for N in [int(a) for a in sys.argv[1:]]:
        d = synthetic.make_shape("S5", seed=0, V=10000, N=N)
        V = 10000
        data = R.Dataset(d["X"], d["T"], d["y"], ["v%d" % i for i in range(V)], d["variable_weights"],
                         d["experiment_start"], d["experiment_end"])
        pop = R.DiscreteRulePopulation(None, d["tmin"], d["N_intervals"], tmax=d["tmax"], data=data)
        for _ in range(3):
            D.detectors_fused(*pop._fused_args)
        torch.cuda.synchronize()
        print(N, pop._G)
        del pop, data
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
