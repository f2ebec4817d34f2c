"""Streaming sweep (mitre_b200/streaming.py; SURVEY 8d, S3 'full grid streamed in tiles') against the
resident population on problems that fit: every candidate of every tile carries the likelihood the
resident, lexsorted table gives the same primitive; the per-tile (max, sum-exp) pairs add up to the
resident log-sum-exp; the two-stage draw lands where a flat cumulative sum over the tiles'
enumeration order lands."""
import numpy as np
import pytest

from mitre_b200 import synthetic

pytestmark = pytest.mark.gpu


def _case(N, V, seed):
    from mitre_b200 import rules as R
    d = synthetic.make_series(N=N, V=V, T=(5, 9), span=(0.0, 40.0), seed=seed)
    data = R.Dataset(d["X"], d["T"], d["y"], ["v%d" % i for i in range(V)], d["variable_weights"],
                     d["experiment_start"], d["experiment_end"])
    return d, data


def _index_of(pop, tup):
    """pop._index_of, except that slope thresholds are matched to 1e-9 relative: the streaming tiles
    evaluate slopes with the phase-by-phase kernel, the resident N <= 63 population with the fused one
    (same closed form, different summation order; DESIGN.md 'slope parity')."""
    i = pop._index_of(tup)
    if i is not None or tup[2] != 'slope':
        return i
    variable, window, type_, direction, threshold = tup
    w = pop._window_lookup[(float(window[0]), float(window[1]))]
    g = int(pop._group_base[w] + variable * 2)
    K = int(pop._K_host[g])
    thr = pop._thresholds_host()[g, :K]
    k = int(np.argmin(np.abs(thr - threshold)))
    assert abs(thr[k] - threshold) <= 1e-9 * max(abs(threshold), np.abs(thr).max())
    r = int(pop._row_offsets_host[g]) + 2 * k + ('above', 'below').index(direction)
    return int(pop._inv_perm_host()[r])


@pytest.mark.parametrize("group_scoring", [True, False])
@pytest.mark.parametrize("N,V,tile", [(70, 12, 5), (20, 9, 4), (130, 6, 6)])
def test_streaming_sweep_matches_resident_population(N, V, tile, group_scoring):
    import torch
    from mitre_b200 import device as D, rules as R, streaming
    d, data = _case(N, V, seed=N + V)
    pop = R.DiscreteRulePopulation(None, 1.0, 4, tmax=None, data=data)
    sw = streaming.StreamingSweep(data, 1.0, 4, tmax=None, tile_variables=tile, group_scoring=group_scoring)
    assert sw.windows == pop.windows and len(sw.tiles) == -(-V // tile)
    rng = np.random.RandomState(N)
    P = len(pop)
    state = synthetic.make_sweep_state(N, 2, np.stack([pop._truth_row(int(rng.randint(P))) for _ in range(8)]),
                                       seed=1)
    X, omega, kappa, mask = state["X"], state["omega"], state["y"] - 0.5, state["mask"]
    bv = rng.normal(size=V)
    bw = rng.normal(size=len(pop.windows))
    prior_tables = (bv, bw, 0.7)
    log_total, n = sw.sweep(X, omega, kappa, 100.0, mask=mask, prior_tables=prior_tables)
    assert n == P
    # resident reference: likelihoods in sorted order + the same factored prior per row
    d_mask = torch.as_tensor(D.pack_mask(mask), device=D.dev())
    res, status = D.score_all(pop.packed_truth, N, X, omega, kappa, 100.0, mask=d_mask)
    D.raise_on_status(status)
    res = res.cpu().numpy()
    prior_res = np.array([bv[pop._tuple_at(i)[0]] + bw[pop._window_lookup[tuple(map(float, pop._tuple_at(i)[1]))]]
                          - 0.7 for i in range(P)])
    want_total = float(torch.logsumexp(torch.as_tensor(res + prior_res), 0))
    assert abs(log_total - want_total) <= 1e-11 * abs(want_total)
    # tile by tile: same primitive -> same likelihood and prior
    flat_w = []
    for t in range(len(sw.tiles)):
        tile_d = sw.build_tile(t)
        out, status, stats, prior = sw._score_tile(tile_d, X, omega, kappa, 100.0, mask, prior_tables)
        out = out.cpu().numpy().copy()
        prior = prior.cpu().numpy()
        flat_w.append(out + prior)
        for k in list(rng.randint(0, tile_d['P'], size=25)) + [0, tile_d['P'] - 1]:
            tup = sw.describe(tile_d, int(k))
            i = _index_of(pop, tup)
            assert i is not None, tup
            assert abs(out[k] - res[i]) <= 1e-12 * abs(res[i])
            assert abs(prior[k] - prior_res[i]) <= 1e-12
            assert np.array_equal(sw.truth_row(tile_d, int(k)), pop._truth_row(i))
    flat_w = np.concatenate(flat_w)
    # two-stage draw == flat draw over (tile-major, enumeration) order
    offs = np.concatenate([[0], np.cumsum(sw.tile_counts)])
    c = np.cumsum(np.exp(flat_w - flat_w.max()))
    for u in [0.0, 0.2, 0.55, 0.93]:
        t, k, tup = sw.draw(u)
        want = int(np.sum(c < u * c[-1]))
        assert abs(int(offs[t]) + k - want) <= 1
        assert _index_of(pop, tup) is not None


@pytest.mark.parametrize("warp", [1, 0])
@pytest.mark.parametrize("N,p", [(2, 1), (3, 2), (17, 1), (31, 2), (32, 6), (33, 3), (35, 4), (63, 9), (64, 3), (65, 2), (200, 5),
                                 (700, 4), (2000, 4)])
def test_score_groups_equals_row_scoring_and_oracle(N, p, warp):
    """mitre_score_groups (suffix sums along each group's sorted order) against the row kernels on the
    same tile, and against the C oracle on sampled candidates; ties and zero-inflated series included."""
    import torch
    import oracle
    from mitre_b200 import device as D, streaming
    from mitre_b200._lib import lib
    if N > 64 and not warp:
        pytest.skip("the CTA kernel is the only one for N > 64")
    lib().mitre_set_groups_warp(warp)          # N <= 64: warp per group (default) or CTA per group
    V = 3 if N > 300 else 7
    d, data = _case(N, V, seed=3 * N + p)
    sw = streaming.StreamingSweep(data, 1.0, 3, tmax=None, tile_variables=V)
    tile = sw.build_tile(0, rows=True)
    P = tile['P']
    if N == 2:
        assert P > 0
    rng = np.random.RandomState(N + p)
    X = np.concatenate([(rng.rand(N, p - 1) > 0.5).astype(float), np.ones((N, 1))], axis=1)
    omega = rng.gamma(2.0, 0.125, size=N) + 0.05
    y = rng.rand(N) > 0.5
    mask = rng.rand(N) > 0.25
    d_mask = torch.as_tensor(D.pack_mask(mask), device=D.dev())
    a, status = D.score_groups(tile['values'], tile['thresholds'], tile['K'], tile['row_offsets'], P, X, omega,
                               y - 0.5, 100.0, mask=d_mask)
    D.raise_on_status(status)
    a = a.cpu().numpy().copy()
    b, status = D.score_all(tile['rows'], N, X, omega, y - 0.5, 100.0, mask=d_mask)
    D.raise_on_status(status)
    b = b.cpu().numpy()
    assert np.isfinite(a).all()
    assert np.max(np.abs(a - b) / np.abs(b)) <= 1e-11
    idx = np.unique(np.concatenate([[0, 1, P - 2, P - 1], rng.randint(0, P, size=20)]))
    rows = D.unpack_rows(tile['rows'][torch.as_tensor(idx, device=D.dev())], N).cpu().numpy().astype(bool)
    want = oracle.likelihoods_of_all_options(X, omega, (y - 0.5) / omega, (rows & mask).astype(float), 100.0)
    lib().mitre_set_groups_warp(1)
    assert np.max(np.abs(a[idx] - want) / np.abs(want)) <= 1e-9


def test_streaming_locate_deltas_and_log_weight():
    """The pieces a sampler needs around the sweep: primitive tuple -> (tile, position), sparse additive
    corrections on named candidates (logit_rules.py:861-862), and one candidate's log-weight /
    log-probability under the last sweep."""
    import torch
    from mitre_b200 import streaming
    N, V = 40, 7
    d, data = _case(N, V, seed=77)
    sw = streaming.StreamingSweep(data, 1.0, 4, tmax=None, tile_variables=3)
    rng = np.random.RandomState(5)
    X = np.concatenate([(rng.rand(N, 1) > 0.5).astype(float), np.ones((N, 1))], axis=1)
    omega = rng.gamma(2.0, 0.125, size=N) + 0.05
    y = rng.rand(N) > 0.5
    tables = (rng.normal(size=V), rng.normal(size=len(sw.windows)), -0.2)
    picks, flat = [], []
    for t in range(len(sw.tiles)):
        tile = sw.build_tile(t)
        for k in rng.randint(0, tile['P'], size=2):
            tup = sw.describe(tile, int(k))
            assert sw.locate(tup) == (t, int(k)) and sw.locate(tup, tile=tile) == (t, int(k))
            picks.append(((t, int(k)), tup))
    assert sw.locate((0, sw.windows[0], 'average', 'above', 1e99)) is None
    assert sw.locate((V + 3, sw.windows[0], 'average', 'above', 0.0)) is None
    deltas = {loc: float(dv) for (loc, _), dv in zip(picks, rng.normal(size=len(picks)) * 3)}
    log_total, n = sw.sweep(X, omega, y - 0.5, 100.0, prior_tables=tables, deltas=deltas)
    offs = np.concatenate([[0], np.cumsum(sw.tile_counts)])
    for t in range(len(sw.tiles)):
        tile = sw.build_tile(t)
        out, status, stats, prior = sw._score_tile(tile, X, omega, y - 0.5, 100.0, None, tables, None)
        flat.append(out.cpu().numpy() + prior.cpu().numpy())
    flat = np.concatenate(flat)
    for (t, k), dv in deltas.items():
        flat[offs[t] + k] += dv
    want = float(torch.logsumexp(torch.as_tensor(flat), 0))
    assert abs(log_total - want) <= 1e-11 * abs(want)
    for (t, k), _ in picks[:3]:
        w, logp = sw.log_weight(t, k)
        assert abs(w - flat[offs[t] + k]) <= 1e-11 * abs(w)
        assert abs(logp - (flat[offs[t] + k] - want)) <= 1e-9
    c = np.cumsum(np.exp(flat - flat.max()))
    for u in [0.1, 0.6, 0.99]:
        t, k, tup = sw.draw(u)
        assert abs(int(offs[t]) + k - int(np.sum(c < u * c[-1]))) <= 1


def test_streaming_detector_gibbs_samples_the_resident_conditional():
    """StreamingDetectorGibbs.conditional(i, j) must weigh every candidate exactly as the resident
    sampler's detector update does (logit_rules.py:794-881: likelihood under the rule's other
    detectors as mask + flat_prior + the multinomial / relative-rate terms of the detectors already
    in the rule), and its update must return a candidate with a valid truth row."""
    import torch
    from scipy.special import gammaln
    from mitre_b200 import chains, device as D, rules as R, streaming, synthetic
    data, model = chains.build_model("S2", seed=6, variables=7)
    pop = model.rule_population
    sampler = chains.run_chain(model, 2, 2, chains.initial_rule_list(model))
    state = sampler.current_state
    cfg = synthetic.SHAPES["S2"]
    sw = streaming.StreamingSweep(data, cfg["tmin"], cfg["N_intervals"], tmax=cfg["tmax"], tile_variables=3)
    assert sw.windows == pop.windows
    rng = np.random.RandomState(11)
    P = len(pop)
    # detectors named by the streaming side; the resident population knows the same ones (its slope
    # thresholds may differ in the last bits: different kernels, see _index_of above)
    picks, picks_res = [], []
    for t in (0, 1, 2):
        tile = sw.build_tile(t)
        picks.append(sw.describe(tile, int(rng.randint(0, tile['P']))))
        picks_res.append(pop._tuple_at(_index_of(pop, picks[-1])))
    rule_list = [[picks[0], picks[1], picks[0]], [picks[2]]]          # a repeated detector in rule 0
    rule_list_res = [[picks_res[0], picks_res[1], picks_res[0]], [picks_res[2]]]
    tables = model._prior_tables(state)
    g = streaming.StreamingDetectorGibbs(sw, data.y, rule_list, model.prior_coefficient_variance, tables, seed=3)
    g.omega = np.asarray(state.omega, dtype=np.float64).copy()
    i, j = 0, 1
    log_total, n = g.conditional(i, j)
    assert n == P
    # resident weights for the same position
    rl = R.RuleList([[R.PrimitiveRule(*p) for p in rule] for rule in rule_list_res])
    lik, status = model.likelihoods_of_all_options_device(rl, g.omega, i, j)
    D.raise_on_status(status)
    w_res = lik.cpu().numpy() + model.flat_prior_device(state).cpu().numpy()
    k0 = pop._index_of(picks_res[0])
    w_res[k0] += -1.0 * gammaln(2 + 1 + 1) + np.log(1.0 + 2)           # picks[0] occurs twice besides position j
    want_total = float(torch.logsumexp(torch.as_tensor(w_res), 0))
    assert abs(log_total - want_total) <= 1e-10 * abs(want_total)
    for t in range(len(sw.tiles)):
        tile = sw.build_tile(t)
        for k in list(rng.randint(0, tile['P'], size=15)):
            tup = sw.describe(tile, int(k))
            w, logp = sw.log_weight(t, int(k))
            assert abs(w - w_res[_index_of(pop, tup)]) <= 1e-10 * abs(w)
    w0, _ = sw.log_weight(*sw.locate(picks[0]))
    assert abs(w0 - w_res[k0]) <= 1e-10 * abs(w0)
    # an update replaces the detector by a candidate and keeps the bookkeeping consistent
    new = g.update_detector(i, j)
    assert g.rules[i][j] == new and _index_of(pop, new) is not None
    assert np.array_equal(g.truth[i][j], pop._truth_row(_index_of(pop, new)))
    g.update_omega_beta(5)
    assert g.omega.shape == (data.n_subjects,) and np.all(g.omega > 0) and g.beta.shape == (3,)
    for _ in range(3):
        g.step(n_gibbs=3)
    assert len(g.log) == 4
