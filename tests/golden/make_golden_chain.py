#!/usr/bin/env python
"""Record whole-chain golden traces FROM THE REFERENCE SAMPLER (build container only).

Runs the reference's own LogisticRuleSampler (mitre/logit_rules.py:503-1420, loaded through
oracle/ref_shim.py) on its own Cython likelihood kernel (oracle/_ref) for a fixed seed, and
stores the trace: per step the sampler comment, the rule list, the recorded likelihood and prior,
plus the final PosteriorSummary.point_summarize rule list (posterior.py:492-526, restated in
mitre_b200/posterior.py since posterior.py imports matplotlib/sklearn at module level).

pypolyagamma is absent from the image and unpinned in the reference (setup.py:57), so the
Polya-Gamma draws come from mitre_b200/polyagamma.py in BOTH samplers ("shared PG sampler",
SURVEY.md 8c): the parity claim is about everything downstream of omega.

    python tests/golden/make_golden_chain.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from mitre_b200 import polyagamma, synthetic      # noqa: E402
from mitre_b200.posterior import PosteriorSummary  # noqa: E402
from oracle import build_ref, ref_shim             # noqa: E402

CHAIN_CASES = {
    "chain_small": dict(series=dict(N=24, V=5, T=(6, 10), span=(0.0, 60.0), seed=41),
                        model=dict(tmin=1.0, N_intervals=4, tmax=40.0), seed=7, steps=120),
    "chain_bokulich": dict(series=dict(N=35, V=8, T=(8, 25), span=(0.0, 375.0), seed=42),
                           model=dict(tmin=1.0, N_intervals=12, tmax=180.0), seed=11, steps=60),
}


def rl_to_list(rl):
    return [[[int(p.variable), [float(p.window[0]), float(p.window[1])], p.type_, p.direction,
              float(p.threshold)] for p in rule] for rule in rl.rules]


def initial_rule_list(rules_mod, model):
    """main.py:952-955: the single primitive with the smallest Hamming distance to y."""
    truth = np.asarray(model.rule_population.flat_truth, dtype=bool)
    dist = np.mean(truth != np.asarray(model.data.y, dtype=bool)[None, :], axis=1)
    return rules_mod.RuleList([[model.rule_population.flat_rules[int(np.argmin(dist))]]])


def run_chain(rules_mod, lr_mod, cfg):
    d = synthetic.make_series(**cfg["series"])
    V = cfg["series"]["V"]
    data = rules_mod.Dataset(d["X"], d["T"], d["y"], ["v%d" % i for i in range(V)],
                             d["variable_weights"], d["experiment_start"], d["experiment_end"])
    model = lr_mod.LogisticRuleModel(data, cfg["model"]["tmin"],
                                     N_intervals=cfg["model"]["N_intervals"],
                                     tmax=cfg["model"]["tmax"])
    np.random.seed(cfg["seed"])
    lr_mod.ppg = polyagamma.PyPolyaGamma(cfg["seed"])
    sampler = lr_mod.LogisticRuleSampler(model, initial_rule_list(rules_mod, model))
    sampler.sample(cfg["steps"])
    return model, sampler


def main():
    ref = build_ref.load()
    rules_mod, lr_mod = ref_shim.load_logit_rules(polyagamma, ref)
    for name, cfg in CHAIN_CASES.items():
        model, sampler = run_chain(rules_mod, lr_mod, cfg)
        summary = PosteriorSummary(sampler, burnin_fraction=0.05)
        summary.point_summarize()
        out = dict(
            n_primitives=len(model.rule_population),
            comments=sampler.comments,
            rule_lists=[rl_to_list(s.rl) for s in sampler.states],
            likelihoods=[float(x) for x in sampler.likelihoods],
            priors=[float(x) for x in sampler.priors],
            point=rl_to_list(summary.point.rl),
            final_numpy_random=float(np.random.rand()),
        )
        with open(os.path.join(HERE, name + ".json"), "w") as f:
            json.dump(out, f)
        moves = {}
        for c in sampler.comments:
            key = c.split(":")[0].split(" ")[0] + (" accepted" if "accepted" in c else "")
            moves[key] = moves.get(key, 0) + 1
        print(name, "P=%d" % out["n_primitives"], "steps=%d" % len(sampler.states), moves,
              "point:", out["point"])


if __name__ == "__main__":
    main()
