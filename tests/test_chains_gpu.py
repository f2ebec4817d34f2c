"""Chain-parallel entry (replicates, main.py:1913-1986): chains are seeded replicas; the same
seed reproduces the same chain, different seeds diverge."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_chain_replicas_are_seed_deterministic():
    from mitre_b200 import chains
    data, model = chains.build_model("S2", seed=3, variables=6)
    r0 = chains.initial_rule_list(model)
    a = chains.run_chain(model, 5, 25, r0)
    b = chains.run_chain(model, 5, 25, r0)
    c = chains.run_chain(model, 6, 25, r0)
    assert a.comments == b.comments
    assert np.array_equal(np.array(a.likelihoods), np.array(b.likelihoods))
    assert a.comments != c.comments
    assert isinstance(chains.summarize(a), list)


def test_model_pickle_round_trip_rebuilds_device_state():
    """Checkpoint/resume: device tensors are dropped on pickle and rebuilt on load."""
    import pickle
    from mitre_b200 import chains
    data, model = chains.build_model("S2", seed=4, variables=4)
    before = model.rule_population.packed_truth.cpu().numpy().copy()
    clone = pickle.loads(pickle.dumps(model))
    assert np.array_equal(clone.rule_population.packed_truth.cpu().numpy(), before)
    r0 = chains.initial_rule_list(model)
    a = chains.run_chain(model, 9, 10, r0)
    b = chains.run_chain(clone, 9, 10, chains.initial_rule_list(clone))
    assert a.comments == b.comments


def test_factored_prior_equals_the_per_primitive_prior():
    """flat_prior (logit_rules.py:424-457: scipy log-densities over all P primitives, log-sum-exp over
    P terms) against the factored tables (V + W evaluations, W x V matrix-vector normalisation) and the
    device vector built from them, for several hyperparameter states."""
    from mitre_b200 import chains
    data, model = chains.build_model("S2", seed=2, variables=9)
    sampler = chains.run_chain(model, 1, 3, chains.initial_rule_list(model))
    state = sampler.current_state
    pop = model.rule_population
    rng = np.random.RandomState(0)
    for _ in range(4):
        state.phylogeny_mean = float(rng.normal())
        state.phylogeny_std = float(abs(rng.normal()) + 0.3)
        state.window_concentration = float(abs(rng.normal()) * 5 + 1)
        state.window_typical_fraction = float(rng.uniform(0.1, 0.9))
        model.__dict__.pop('_flat_prior_cache', None)
        want = model.flat_prior(state)
        bv, bw, norm = model._prior_tables(state)
        got = np.array([bv[pop._tuple_at(i)[0]] + bw[pop._window_lookup[tuple(map(float, pop._tuple_at(i)[1]))]] - norm
                        for i in range(len(pop))])
        assert np.max(np.abs(got - want)) <= 1e-11 * np.max(np.abs(want))
        dev = model.flat_prior_device(state).cpu().numpy()
        assert np.max(np.abs(dev - want)) <= 1e-11 * np.max(np.abs(want))
        assert abs(np.log(np.sum(np.exp(dev))) ) <= 1e-9          # normalised over the active primitives
