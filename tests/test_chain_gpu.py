"""Whole-chain parity: the mitre_b200 sampler (GPU likelihoods) against traces recorded from the
reference's own LogisticRuleSampler + Cython kernel (tests/golden/make_golden_chain.py), same
numpy seed, same Polya-Gamma sampler.  Move sequence, sampled detector indices, rule lists and
the final point summary must match; recorded likelihoods / priors within 1e-9 relative."""
import json
import os

import numpy as np
import pytest

from mitre_b200 import synthetic
from tests.golden import make_golden_chain as MC

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def same_rule_list(got, want):
    assert len(got) == len(want)
    for gr, wr in zip(got, want):
        assert len(gr) == len(wr)
        for g, w in zip(gr, wr):
            assert g[0] == w[0] and g[1] == w[1] and g[2] == w[2] and g[3] == w[3]
            if g[2] == "average":
                assert g[4] == w[4]
            else:
                assert abs(g[4] - w[4]) <= 1e-9 * abs(w[4]) + 1e-300


@pytest.mark.parametrize("device_draw", [False, True])
@pytest.mark.parametrize("name", sorted(MC.CHAIN_CASES))
def test_chain_matches_reference_trace(name, device_draw):
    from mitre_b200 import logit_rules as LR, polyagamma, rules as R
    from mitre_b200.posterior import PosteriorSummary
    cfg = MC.CHAIN_CASES[name]
    gold = json.load(open(os.path.join(GOLD, name + ".json")))
    d = synthetic.make_series(**cfg["series"])
    V = cfg["series"]["V"]
    data = R.Dataset(d["X"], d["T"], d["y"], ["v%d" % i for i in range(V)], d["variable_weights"],
                     d["experiment_start"], d["experiment_end"])
    model = LR.LogisticRuleModel(data, cfg["model"]["tmin"], N_intervals=cfg["model"]["N_intervals"],
                                 tmax=cfg["model"]["tmax"])
    assert len(model.rule_population) == gold["n_primitives"]
    np.random.seed(cfg["seed"])
    LR.ppg = polyagamma.PyPolyaGamma(cfg["seed"])
    model.fast_prior = False
    sampler = LR.LogisticRuleSampler(model, MC.initial_rule_list(R, model), device_draw=device_draw)
    sampler.sample(cfg["steps"])
    assert sampler.comments == gold["comments"]
    for s, want in zip(sampler.states, gold["rule_lists"]):
        same_rule_list(MC.rl_to_list(s.rl), want)
    lik = np.array(sampler.likelihoods)
    pri = np.array(sampler.priors)
    assert np.max(np.abs(lik - gold["likelihoods"]) / np.abs(gold["likelihoods"])) < 1e-9
    assert np.max(np.abs(pri - gold["priors"]) / np.abs(gold["priors"])) < 1e-9
    summary = PosteriorSummary(sampler, burnin_fraction=0.05)
    summary.point_summarize()
    same_rule_list(MC.rl_to_list(summary.point.rl), gold["point"])
    assert np.random.rand() == gold["final_numpy_random"]      # the RNG stream stayed aligned
