"""One S5 population build (detector kernels run once) for ncu captures.  usage: python tools/ncu_det.py [V]"""
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402


def main():
    V = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
    sys.argv = sys.argv[:1]
    args = bench.parse_args()
    args.variables = V
    from mitre_b200 import rules as R, synthetic
    d = synthetic.make_shape("S5", seed=0, V=V)
    data = R.Dataset(d["X"], d["T"], d["y"], ["v%d" % i for i in range(V)], d["variable_weights"],
                     d["experiment_start"], d["experiment_end"])
    pop = R.DiscreteRulePopulation(None, d["tmin"], d["N_intervals"], tmax=d["tmax"], data=data)
    torch.cuda.synchronize()
    print("done P=%d" % len(pop))


if __name__ == "__main__":
    main()
