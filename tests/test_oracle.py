"""Pin the oracle: CPU restatement vs (a) the committed golden vectors generated from the
reference itself, (b) the reference's compiled Cython kernel in oracle/_ref when present,
(c) the reference's rules.py through the py3 shim when /root/reference is present, and (d) the
internal identity the reference relies on (logit_rules.py:944-961)."""
import os

import numpy as np
import pytest

import oracle
from oracle import build_ref, detectors_oracle as DO, ref_shim
from mitre_b200 import synthetic
from tests.golden import make_golden as MG

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def rel(a, b):
    return np.max(np.abs(np.asarray(a) - np.asarray(b)) / np.maximum(np.abs(np.asarray(b)), 1e-300))


@pytest.mark.parametrize("name", sorted(MG.SCORING_CASES))
def test_scoring_oracle_vs_golden(name):
    c = MG.scoring_inputs(**MG.SCORING_CASES[name])
    gold = np.load(os.path.join(GOLD, name + ".npz"))["loglik"]
    opts = (c["truth"] * c["mask"]).astype(np.float64)
    got = oracle.likelihoods_of_all_options(c["X"], c["omega"], (c["y"] - 0.5) / c["omega"], opts,
                                            c["sigma2"])
    assert got.shape == gold.shape
    assert rel(got, gold) < 1e-12


def test_dense_oracle_vs_golden():
    gold = np.load(os.path.join(GOLD, "dense.npz"))["loglik"]
    for (seed, N, p), g in zip(MG.DENSE_CASES, gold):
        c = MG.dense_inputs(seed, N, p)
        got = oracle.likelihood_z_given_X(c["omega"], c["X"], c["y"], c["sigma2"])
        assert abs(got - g) / abs(g) < 1e-12
        z = (c["y"] - 0.5) / c["omega"]
        cov = np.diag(1.0 / c["omega"]) + c["sigma2"] * c["X"].dot(c["X"].T)
        assert abs(oracle.mc_logpdf_np(z, cov) - g) / abs(g) < 1e-12
        assert abs(oracle.mc_logpdf_np_numpy(z, cov) - g) / abs(g) < 1e-13


def test_scoring_identity_with_dense():
    """likelihoods_of_all_options[k] == likelihood_z_given_X([X | v_k]) (logit_rules.py:944-961)."""
    c = MG.scoring_inputs(seed=5, N=37, p=3, P=60)
    opts = (c["truth"] * c["mask"]).astype(np.float64)
    ll = oracle.likelihoods_of_all_options(c["X"], c["omega"], (c["y"] - 0.5) / c["omega"], opts, 100.0)
    for k in range(0, 60, 7):
        Xk = np.hstack([c["X"], opts[k][:, None]])
        d = oracle.likelihood_z_given_X(c["omega"], Xk, c["y"], 100.0)
        assert abs(d - ll[k]) / abs(d) < 1e-12


def test_scoring_oracle_vs_compiled_reference():
    if build_ref.build() is None:
        pytest.skip("reference tree absent and no prebuilt oracle/_ref")
    ref = build_ref.load()
    c = MG.scoring_inputs(seed=77, N=50, p=5, P=800)
    opts = (c["truth"] * c["mask"]).astype(np.float64)
    z = (c["y"] - 0.5) / c["omega"]
    a = ref.likelihoods_of_all_options(c["X"], c["omega"], z, opts, np.float64(100.0))
    b = oracle.likelihoods_of_all_options(c["X"], c["omega"], z, opts, 100.0)
    assert rel(b, a) < 1e-12


def test_not_positive_definite_raises():
    c = MG.scoring_inputs(seed=1, N=10, p=2, P=8)
    om = c["omega"].copy()
    om[3] = -1.0
    with pytest.raises(np.linalg.LinAlgError):
        oracle.likelihoods_of_all_options(c["X"], om, (c["y"] - 0.5) / om,
                                          c["truth"].astype(np.float64), 100.0)


def test_choose_matches_numpy_definition():
    rng = np.random.RandomState(3)
    p = np.exp(rng.normal(size=1000))
    for u in [0.0, 0.1, 0.5, 0.999999]:
        c = np.cumsum(p)
        assert oracle.choose_from_relative_probabilities(p, u) == int(np.sum(c < u * c[-1]))
    assert oracle.choose_from_relative_probabilities(np.zeros(5), 0.7) == 0   # underflow edge


def test_pairwise_sum_is_numpy_mean():
    rng = np.random.RandomState(0)
    for n in list(range(1, 140)) + [200, 257, 1000, 4097]:
        x = rng.rand(n) * 10 ** rng.uniform(-8, 2, size=n)
        assert DO.mean_restated(x) == np.mean(x)


def _rebuild(name):
    cfg = MG.DETECTOR_CASES[name]
    d = synthetic.make_series(**cfg["series"])
    pop = DO.build_population(d["X"], d["T"], d["experiment_start"], d["experiment_end"],
                              cfg["pop"]["N_intervals"], cfg["pop"]["tmin"], cfg["pop"]["tmax"])
    return d, pop


@pytest.mark.parametrize("name", sorted(MG.DETECTOR_CASES))
def test_detector_oracle_vs_golden(name):
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    d, pop = _rebuild(name)
    N = int(gold["n_subjects"])
    assert np.array_equal(pop["windows"], gold["windows"])
    assert np.array_equal(pop["ok_slope"], gold["ok_slope"])
    assert np.array_equal(np.stack([pop["group_window"], pop["group_variable"], pop["group_type"]], 1),
                          gold["groups"])
    assert np.array_equal(pop["values"], gold["values"])          # bit-exact, incl. lstsq slopes
    assert np.array_equal(pop["order"], gold["order"])
    o = pop["order"]
    assert np.array_equal(DO.pack_rows(pop["truth"][o]), gold["sorted_truth_packed"])
    assert np.array_equal(pop["variable"][o], gold["rule_variable"])
    assert np.array_equal(pop["window"][o], gold["rule_window"])
    assert np.array_equal(pop["type"][o], gold["rule_type"])
    assert np.array_equal(pop["direction"][o], gold["rule_direction"])
    assert np.array_equal(pop["threshold"][o], gold["rule_threshold"])
    assert np.array_equal(DO.unpack_rows(gold["sorted_truth_packed"], N), pop["truth"][o])


def test_detector_oracle_vs_reference_rules_py():
    if not ref_shim.available():
        pytest.skip("reference tree absent")
    rules = ref_shim.load_rules()
    d = synthetic.make_series(N=17, V=4, T=(5, 11), span=(0.0, 50.0), seed=99)
    data = rules.Dataset(d["X"], d["T"], d["y"], list("abcd"), d["variable_weights"], 0.0, 50.0)
    ref_pop = rules.DiscreteRulePopulation(None, 2.0, 5, tmax=30.0, data=data)
    pop = DO.build_population(d["X"], d["T"], 0.0, 50.0, 5, 2.0, 30.0)
    assert [tuple(w) for w in pop["windows"]] == [tuple(w) for w in ref_pop.windows]
    o = pop["order"]
    assert np.array_equal(o, ref_pop._reordering_cache)
    assert np.array_equal(pop["truth"][o], ref_pop.flat_truth)
    assert np.array_equal(pop["threshold"][o], np.array([r[4] for r in ref_pop.flat_rules]))
