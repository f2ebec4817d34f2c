"""The reference's full sampler step (logit_rules.py:567-596: detector update / add / remove / move +
hyperparameter updates + omega / beta sweeps) over a STREAMED candidate population
(streaming.make_streaming_model / make_streaming_sampler) against the resident sampler on a problem that fits.

The streamed population lays its candidates out in (tile, enumeration) order, the resident one in the
reference's lexsorted order: the same candidate set with the same weights, so the same uniform draw picks
different candidates.  The chain test therefore FORCES the streamed sampler's draws to the resident chain's
choices (np.random consumed identically) and requires everything computed from them to agree: the move
sequence, every accept / reject decision with its acceptance ratio, the rule lists and the hyperparameter
trajectories.  The proposal quantities of add / remove are also compared directly, move by move."""
import re

import numpy as np
import pytest

from mitre_b200 import synthetic

pytestmark = pytest.mark.gpu


def _models(N=70, V=10, tile=4, seed=3):
    from mitre_b200 import logit_rules as LR, rules as R, streaming
    d = synthetic.make_series(N=N, V=V, T=(5, 9), span=(0.0, 40.0), seed=seed)
    mk = lambda: R.Dataset(d["X"], d["T"], d["y"], ["v%d" % i for i in range(V)], d["variable_weights"],
                           d["experiment_start"], d["experiment_end"])
    resident = LR.LogisticRuleModel(mk(), 1.0, N_intervals=4)
    streamed = streaming.make_streaming_model(mk(), 1.0, tile_variables=tile, N_intervals=4)
    return resident, streamed


def _r0(model, ks):
    from mitre_b200 import rules as R
    flat = model.rule_population.flat_rules
    return R.RuleList([[R.PrimitiveRule(*flat[k]) for k in rule] for rule in ks])


def test_streamed_population_counts_and_prior_tables():
    from mitre_b200 import logit_rules as LR
    resident, streamed = _models()
    assert len(streamed.rule_population) == len(resident.rule_population)
    assert np.array_equal(streamed._prior_count_matrix(), resident._prior_count_matrix())
    r0 = _r0(resident, [[5]])
    state = LR.LogisticRuleState(r0, np.ones(70), np.zeros(2), resident.phylogeny_lambda_l, 1.3)
    for a, b in zip(resident._prior_tables(state), streamed._prior_tables(state)):
        assert np.array_equal(np.asarray(a), np.asarray(b))
    resident.fast_prior = True
    assert streamed.rule_list_prior(state) == resident.rule_list_prior(state)
    # the memoised form the sampler uses during a step's hyperparameter updates: bit-identical to prior()
    for model in (resident, streamed):
        plain = [model.prior(state)]
        model._step_memo = {}
        memo = [model.prior(state)]
        for name, value in (("phylogeny_mean", 0.3), ("phylogeny_std", 2.2), ("window_concentration", 3.1),
                            ("window_typical_fraction", 0.21)):
            other = state.copy()
            setattr(other, name, value)
            memo.append(model.prior(other))
            model._step_memo, keep = None, model._step_memo
            plain.append(model.prior(other))
            model._step_memo = keep
        model._step_memo = None
        assert memo == plain


@pytest.mark.parametrize("rules", [[[11]], [[40, 900], [7]], [[3, 3, 250], [77], [1500, 8]]])
def test_add_remove_quantities_match_resident(rules):
    """consider_addition_to_new_rule / _to_existing_rule: (log relative prior, log likelihood ratio,
    probability of the reverse proposal) of the streamed sampler == the resident sampler's, host arithmetic."""
    from mitre_b200 import logit_rules as LR, polyagamma, streaming
    resident, streamed = _models()
    P = len(resident.rule_population)
    rules = [[k % P for k in rule] for rule in rules]
    np.random.seed(5)
    LR.ppg = polyagamma.PyPolyaGamma(5)
    rs = LR.LogisticRuleSampler(resident, _r0(resident, rules), device_draw=False)
    ss = streaming.make_streaming_sampler(streamed, _r0(resident, rules))
    rs.update_omega()
    ss.current_state.omega = rs.current_state.omega.copy()
    base = rs.current_state.rl
    for position in range(len(base) + 1):
        a = rs.consider_addition_to_new_rule(base, position)
        b = ss.consider_addition_to_new_rule(ss.current_state.rl, position)
        for x, y in zip(a[1:3] + a[4:], b[1:3] + b[4:]):
            assert abs(x - y) <= 1e-9 * max(1.0, abs(x)), (position, a[1:3], b[1:3])
    for position in range(len(base)):
        a = rs.consider_addition_to_existing_rule(base, position)
        b = ss.consider_addition_to_existing_rule(ss.current_state.rl, position)
        for x, y in zip(a[1:3] + a[4:], b[1:3] + b[4:]):
            assert abs(x - y) <= 1e-9 * max(1.0, abs(x)), (position, a[1:3] + a[4:], b[1:3] + b[4:])


def _strip(comment):
    comment = re.sub(r'adds detector .*$', 'adds detector', comment)
    return re.sub(r'replacing .* with .*$', 'replacing', comment)


def test_streamed_chain_follows_resident_chain():
    from mitre_b200 import logit_rules as LR, polyagamma, rules as R, streaming
    resident, streamed = _models()
    steps = 60
    # resident chain, recording every drawn primitive
    np.random.seed(11)
    LR.ppg = polyagamma.PyPolyaGamma(11)
    rs = LR.LogisticRuleSampler(resident, _r0(resident, [[17]]), device_draw=True, device_gibbs=True)
    drawn = []
    flat = resident.rule_population.flat_rules
    orig_candidate, orig_gibbs = rs._draw_candidate, rs._device_gibbs_draw

    def rec_candidate(dist):
        k, p = orig_candidate(dist)
        drawn.append(p.as_tuple())
        return k, p

    def rec_gibbs(*a):
        k = orig_gibbs(*a)
        drawn.append(flat[k])
        return k
    rs._draw_candidate, rs._device_gibbs_draw = rec_candidate, rec_gibbs
    rs.sample(steps)
    kinds = ' '.join(rs.comments)
    assert 'replacing' in kinds and 'add detector' in kinds and 'remove detector' in kinds and 'moves to' in kinds
    # streamed chain: same seeds, draws forced to the resident chain's choices, rules kept in the resident order
    np.random.seed(11)
    LR.ppg = polyagamma.PyPolyaGamma(11)
    ss = streaming.make_streaming_sampler(streamed, _r0(resident, [[17]]), device_gibbs=True)
    streamed.rule_population.sort_list_of_primitives = resident.rule_population.sort_list_of_primitives
    feed = iter(drawn)
    free_draws = []

    def forced(dist):
        t, k, primitive = streamed.sweep.draw(np.random.rand())     # consumes the uniform, exercises the draw
        free_draws.append(primitive)
        want = next(feed)
        assert streamed.rule_population._locate(want) is not None
        return streamed.rule_population._locate(want), R.PrimitiveRule(*want)
    ss._draw_candidate = forced
    ss.sample(steps)
    assert next(feed, None) is None and len(free_draws) == len(drawn)
    assert [_strip(c) for c in ss.comments] == [_strip(c) for c in rs.comments]
    for a, b in zip(rs.states, ss.states):
        assert [[p.as_tuple() for p in rule] for rule in a.rl.rules] == [[p.as_tuple() for p in rule] for rule in b.rl.rules]
        for name in ('phylogeny_mean', 'phylogeny_std', 'window_typical_fraction', 'window_concentration',
                     'empty_probability'):
            assert getattr(a, name) == getattr(b, name), name
        assert np.allclose(a.beta, b.beta, rtol=1e-9, atol=1e-12)
    assert np.allclose(rs.priors, ss.priors, rtol=1e-9)


def test_streamed_chain_runs_free():
    """Unforced: a streamed chain with all three structure moves keeps a valid state."""
    from mitre_b200 import logit_rules as LR, polyagamma, streaming
    resident, streamed = _models(N=40, V=9, tile=4, seed=8)
    np.random.seed(2)
    LR.ppg = polyagamma.PyPolyaGamma(2)
    ss = streaming.make_streaming_sampler(streamed, _r0(resident, [[5]]))
    ss.sample(40)
    kinds = ' '.join(ss.comments)
    assert 'replacing' in kinds and 'add detector' in kinds and 'moves to' in kinds
    assert all(np.isfinite(p) for p in ss.priors) and all(np.isfinite(l) for l in ss.likelihoods)
    for st in ss.states:
        for rule in st.rl.rules:
            for p in rule:
                assert streamed.rule_population.get_primitive_index(p) is not None
