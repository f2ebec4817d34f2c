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
