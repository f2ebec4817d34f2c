"""CPU-side tests of host logic that needs no device: the Polya-Gamma sampler, mask packing,
the synthetic generator, RuleList / PrimitiveRule semantics."""
import numpy as np

from mitre_b200 import polyagamma, synthetic
from mitre_b200.device import pack_mask, words_for
from mitre_b200.rules import PrimitiveRule, RuleList
from oracle.detectors_oracle import pack_rows


def test_polyagamma_moments():
    pg = polyagamma.PyPolyaGamma(3)
    for z in [0.0, 1.0, 4.0]:
        x = np.array([pg.pgdraw(1, z) for _ in range(20000)])
        mean = 0.25 if z == 0 else np.tanh(z / 2) / (2 * z)
        var = 1 / 24.0 if z == 0 else (np.sinh(z) - z) / (4 * z ** 3 * np.cosh(z / 2) ** 2)
        assert abs(x.mean() - mean) < 5 * np.sqrt(var / len(x))
        assert abs(x.var() - var) < 0.1 * var
    out = np.zeros(4)
    pg.pgdrawv(np.ones(4), np.array([0.1, -2.0, 3.0, 0.0]), out)
    assert np.all(out > 0)


def test_polyagamma_is_seed_deterministic():
    a = polyagamma.PyPolyaGamma(5)
    b = polyagamma.PyPolyaGamma(5)
    assert [a.pgdraw(1, 0.7) for _ in range(50)] == [b.pgdraw(1, 0.7) for _ in range(50)]


def test_pack_mask_layout_matches_oracle():
    rng = np.random.RandomState(0)
    for N in [1, 35, 64, 65, 200]:
        m = rng.rand(N) > 0.5
        assert np.array_equal(pack_mask(m).view(np.uint64), pack_rows(m[None])[0])
        assert words_for(N) == (N + 63) // 64


def test_synthetic_series_shape_and_no_constant_series():
    d = synthetic.make_series(N=30, V=3, T=(4, 9), span=(0.0, 50.0), seed=1, zero_fraction=0.4)
    assert len(d["X"]) == 30 and d["X"][0].shape[0] == 3
    for x, t in zip(d["X"], d["T"]):
        assert np.all(np.diff(t) > 0) and t[0] >= 0.0 and t[-1] <= 50.0
        assert (x.sum(axis=1) > 0).sum() >= 2          # never a single live variable
        assert np.allclose(x.sum(axis=0), 1.0)
    assert d["y"].sum() == 15


def test_primitive_rule_and_rule_list_semantics():
    p = PrimitiveRule(1, (0.0, 10.0), "average", "above", 0.5)
    assert PrimitiveRule(*p.as_tuple()).as_tuple() == p.as_tuple()
    X = np.array([[0.0, 0.0, 0.0], [0.2, 0.9, 1.0]])
    t = np.array([1.0, 5.0, 20.0])
    assert p.evaluate(X[1], t) == np.mean([0.2, 0.9])
    assert p.apply(X, t)
    q = PrimitiveRule(1, (0.0, 10.0), "slope", "below", 0.0)
    assert abs(q.evaluate(X[1], t) - 0.7 / 4.0) < 1e-12
    rl = RuleList([[p, q], [p.as_tuple()]])
    assert len(rl) == 2 and len(rl[0]) == 2
    assert list(rl.apply_rules(X, t)) == [False, True]
    c = rl.copy()
    del c[0]
    assert len(rl) == 2 and len(c) == 1
    try:
        RuleList([[]])
        assert False
    except ValueError:
        pass
    try:
        PrimitiveRule(0, (30.0, 40.0), "average", "above", 0.0).evaluate(X[0], t)
        assert False
    except ValueError:
        pass


def test_fastdist_is_bit_identical_to_scipy():
    """mitre_b200/fastdist.py restates the scipy evaluations of the sampler's prior
    (logit_rules.py:232-457) without scipy's argument machinery; the chain's accept/reject
    sequence only stays put if the values are the same to the last bit."""
    import scipy.special
    import scipy.stats
    from mitre_b200 import fastdist as F
    rng = np.random.RandomState(0)
    for _ in range(1500):
        x = rng.normal(size=rng.randint(1, 50)) * 3
        loc, scale = rng.normal(), abs(rng.normal()) + 1e-3
        assert np.array_equal(F.norm_logpdf(x, loc, scale), scipy.stats.norm.logpdf(x, loc=loc, scale=scale))
        xs = float(rng.normal() * 3)
        assert F.norm_logpdf(xs, loc, scale) == scipy.stats.norm.logpdf(xs, loc=loc, scale=scale)
        e = abs(rng.normal()) * 5
        assert F.expon_logpdf(e, scale) == scipy.stats.expon.logpdf(e, scale=scale)
        a, b = abs(rng.normal()) * 3 + 0.01, abs(rng.normal()) * 3 + 0.01
        u = rng.rand(rng.randint(1, 30))
        assert np.array_equal(F.beta_logpdf(u, a, b), scipy.stats.beta.logpdf(u, a, b))
        us = float(rng.rand())
        assert F.beta_logpdf(us, a, b) == scipy.stats.beta.logpdf(us, a, b)
        n, p = abs(rng.normal()) * 3 + 0.1, rng.rand() * 0.98 + 0.01
        nb, ref = F.NegativeBinomial(n, p), scipy.stats.nbinom(n, p)
        k = rng.randint(0, 12, size=rng.randint(1, 6))
        assert np.array_equal(nb.logpmf(k), ref.logpmf(k))
        assert nb.logpmf(int(k[0])) == ref.logpmf(int(k[0]))
        assert nb.logcdf(9) == ref.logcdf(9)
        v = rng.normal(size=rng.randint(1, 200)) * 20
        w = rng.randint(0, 5, size=v.size).astype(float)
        assert F.logsumexp(v) == scipy.special.logsumexp(v)
        if w.sum() > 0:
            assert F.logsumexp(v, b=w) == scipy.special.logsumexp(v, b=w)
    # edges scipy defines
    assert np.isnan(F.norm_logpdf(0.0, 0.0, -1.0)) and np.isnan(scipy.stats.norm.logpdf(0.0, 0.0, -1.0))
    assert F.expon_logpdf(-1.0, 2.0) == scipy.stats.expon.logpdf(-1.0, scale=2.0) == -np.inf
    assert F.beta_logpdf(1.5, 2.0, 2.0) == scipy.stats.beta.logpdf(1.5, 2.0, 2.0) == -np.inf
    assert F.logsumexp(np.array([-np.inf, -np.inf])) == scipy.special.logsumexp(np.array([-np.inf, -np.inf]))


def test_capped_thresholds_equal_the_reference_restatement():
    """max_thresholds (rules.py:533-576): the host clustering used by the device population against the
    oracle's restatement of thresholds_from_values, group by group, including ties and tiny groups."""
    from mitre_b200 import rules as R
    from oracle import detectors_oracle as DO
    rng = np.random.RandomState(3)
    for N, cap in [(35, 10), (12, 3), (100, 7), (9, 1), (20, 18)]:
        values = rng.normal(size=(40, N))
        values[3] = 0.0                                   # all tied: one cluster, no threshold
        values[4, : N // 2] = 0.0                         # a run of ties
        values[5] = np.round(values[5], 1)                # many ties
        thr, K = R.capped_thresholds_host(values, cap)
        for g in range(values.shape[0]):
            want = DO.thresholds_from_values(values[g], cap)
            assert K[g] == len(want) <= cap
            assert np.array_equal(thr[g, :K[g]], want)


def test_polyagamma_large_argument_does_not_overflow():
    """PG(1, z) for |z| >= ~98 used to raise OverflowError in math.exp."""
    from mitre_b200 import polyagamma
    pg = polyagamma.PyPolyaGamma(0)
    for z in (100.0, -500.0, 2000.0):
        x = pg.pgdraw(1, z)
        assert np.isfinite(x) and x > 0
        assert abs(x - 0.5 / abs(z)) < 0.5 / abs(z)      # PG(1, z) concentrates at tanh(z/2) / (2 z)


def test_split_r_hat_equals_the_reference_restatement():
    """chains.split_r_hat against a line-by-line restatement of mcmc_diagnostics.py:39-84 (divide_runs +
    compute_R_hat as generate_diagnostics calls them)."""
    from mitre_b200 import chains
    rng = np.random.RandomState(8)
    runs = [np.cumsum(rng.normal(size=(401, 3)), axis=0) * 0.05 + rng.normal(size=(401, 3)) for _ in range(3)]

    def divide_runs(runs):                                   # mcmc_diagnostics.py:39-55 (column 0 kept here)
        out = []
        for r in runs:
            spl = int(np.floor(r.shape[0] / 2))
            out.append(r[0:(spl - 1), :])
            out.append(r[spl:r.shape[0], :])
        return out

    def compute_R_hat(runs2, n):                             # mcmc_diagnostics.py:57-84
        m, nv = len(runs2), runs2[0].shape[1]
        pj, sj = np.zeros((m, nv)), np.zeros((m, nv))
        for j in range(m):
            pj[j, :] = np.mean(runs2[j][0:n, :], axis=0)
        for j in range(m):
            sj[j, :] = np.sum(np.power(runs2[j][0:n, :] - pj[j, :], 2.0), axis=0) / (n - 1)
        B = np.zeros(nv)
        pp = np.mean(pj, axis=0)
        for j in range(m):
            B = B + np.power(pj[j] - pp, 2.0)
        B = B * n / (m - 1)
        W = np.mean(sj, axis=0)
        return np.sqrt((W * (n - 1) / n + B / n) / W)

    runs2 = divide_runs(runs)
    n = min(r.shape[0] for r in runs2)
    want = compute_R_hat(runs2, n - 1)
    got = chains.split_r_hat(runs)
    assert np.allclose(got, want, rtol=1e-12, atol=0)
    same = [rng.normal(size=(300, 2)) for _ in range(4)]
    assert np.all(chains.split_r_hat(same) < 1.05)
    shifted = [rng.normal(size=(300, 2)) + 3.0 * i for i in range(4)]
    assert np.all(chains.split_r_hat(shifted) > 2.0)


def test_streamed_add_to_existing_rule_algebra():
    """streaming.make_streaming_sampler derives two of the three log-sum-exps of
    consider_addition_to_existing_rule (logit_rules.py:1005-1095) from the one sweep it runs:
    A = lse(prior + deltas) from the normalised prior, C = lse(prior + lik) from B = lse(prior + deltas + lik)
    and the weights of the few candidates that carry a sparse term.  Same identities on explicit vectors."""
    from scipy.special import logsumexp
    rng = np.random.RandomState(3)
    P = 5000
    prior = rng.normal(size=P)
    prior -= logsumexp(prior)                      # normalised over the candidates, like flat_prior
    lik = rng.normal(scale=30.0, size=P) - 400.0
    idx = np.array([17, 18, 4200])
    counts = np.array([1, 3, 2])
    deltas = -np.log(counts + 1.0)
    sparse = np.zeros(P)
    sparse[idx] = deltas
    lse_A = np.log1p(np.sum(np.exp(prior[idx]) * np.expm1(deltas)))
    assert abs(lse_A - logsumexp(prior + sparse)) < 1e-13
    lse_B = logsumexp(prior + sparse + lik)
    w = (prior + sparse + lik)[idx]                # what StreamingSweep.log_weight returns for those candidates
    lse_C = lse_B + np.log1p(np.sum(np.expm1(-deltas) * np.exp(w - lse_B)))
    assert abs(lse_C - logsumexp(prior + lik)) < 1e-12 * abs(lse_C)
    # ... also when one of the carrying candidates dominates the sum
    lik[17] += 200.0
    lse_B = logsumexp(prior + sparse + lik)
    w = (prior + sparse + lik)[idx]
    lse_C = lse_B + np.log1p(np.sum(np.expm1(-deltas) * np.exp(w - lse_B)))
    assert abs(lse_C - logsumexp(prior + lik)) < 1e-12 * abs(lse_C)


def test_values_walk_plan_cost_model():
    """device.values_walk_plan: (start groups per CTA, widest window count) minimising
    waves x (2.5 + start groups per CTA) among the splits whose tables fit in shared memory (no GPU needed: the
    shared-memory size is host arithmetic of the library)."""
    from mitre_b200 import device as D
    from mitre_b200._lib import lib
    info = {'smem_optin': 227 * 1024, 'sm_count': 148}
    # the S5 shape: 24 start groups, the k-th start has 24 - k windows (minus the inadmissible ones ~ ignore)
    counts = [23 - k for k in range(23)]
    first = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    N, S, V = 35, 561, 10000
    per, widest = D.values_walk_plan(N, S, V, first, info=info)
    n_sg = len(first) - 1
    assert widest == max(int(first[min(k + per, n_sg)] - first[k]) for k in range(0, n_sg, per))
    assert 0 < lib().mitre_values_walk_smem_bytes(N, S, widest) <= info['smem_optin']
    # it is the argmin of the model over the feasible splits
    vb = lib().mitre_walk_vars_per_block(N)
    tiles = -(-V // vb)

    def cost(p):
        w = max(int(first[min(k + p, n_sg)] - first[k]) for k in range(0, n_sg, p))
        if not 0 < lib().mitre_values_walk_smem_bytes(N, S, w) <= info['smem_optin']:
            return None
        return -(-(tiles * -(-n_sg // p)) // info['sm_count']) * (2.5 + p)
    feasible = {p: cost(p) for p in range(1, n_sg + 1) if cost(p) is not None}
    assert per in feasible and feasible[per] == min(feasible.values())
    assert per >= 8                       # fewer, larger CTAs (the previous plan took 6)
    # nothing fits: a population whose single start group already needs more shared memory than there is
    huge = np.array([0, 4000], dtype=np.int32)
    assert D.values_walk_plan(63, 3000, 100, huge, info=info) is None
    # a tiny problem: both start groups get a CTA of their own (two CTAs still are one wave)
    small = np.array([0, 2, 3], dtype=np.int32)
    assert D.values_walk_plan(10, 60, 16, small, info=info)[0] == 1
