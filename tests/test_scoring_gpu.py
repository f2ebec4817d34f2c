"""GPU parity of the scoring kernels against the oracle and the reference-generated golden
vectors.  Tolerance: <= 1e-9 relative on the log marginal likelihood (BASELINE.json
north_star); in practice the p-space algebra agrees to ~1e-13."""
import os

import numpy as np
import pytest

import oracle
from tests.golden import make_golden as MG

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
RTOL = 1e-9


def rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b)) / np.maximum(np.abs(np.asarray(b)), 1e-300)))


def gpu_score(c, truth=None):
    import torch
    from mitre_b200 import device as D
    truth = c["truth"] if truth is None else truth
    N = c["X"].shape[0]
    packed = D.pack_rows(truth)
    mask = torch.as_tensor(D.pack_mask(c["mask"]), device=D.dev())
    out, status = D.score_all(packed, N, c["X"], c["omega"], c["y"] - 0.5, c["sigma2"], mask=mask)
    res = out.cpu().numpy()
    D.raise_on_status(status)
    return res


@pytest.mark.parametrize("name", sorted(MG.SCORING_CASES))
def test_score_all_vs_reference_golden(name):
    c = MG.scoring_inputs(**MG.SCORING_CASES[name])
    gold = np.load(os.path.join(GOLD, name + ".npz"))["loglik"]
    got = gpu_score(c)
    assert got.shape == gold.shape
    assert rel(got, gold) < RTOL


@pytest.mark.parametrize("N,p,P", [(5, 1, 17), (33, 2, 1000), (63, 5, 777), (64, 9, 513), (65, 3, 300),
                                   (128, 6, 200), (129, 2, 100), (500, 4, 64), (2000, 4, 24),
                                   (35, 16, 300), (35, 24, 100)])
def test_score_all_vs_oracle_shapes(N, p, P):
    c = MG.scoring_inputs(seed=100 + N + p, N=N, p=p, P=P)
    opts = (c["truth"] * c["mask"]).astype(np.float64)
    want = oracle.likelihoods_of_all_options(c["X"], c["omega"], (c["y"] - 0.5) / c["omega"], opts,
                                             c["sigma2"])
    assert rel(gpu_score(c), want) < RTOL


def test_all_false_candidate_equals_base_likelihood():
    """An all-False column leaves the model unchanged (SURVEY 8c structural invariant)."""
    c = MG.scoring_inputs(seed=7, N=35, p=3, P=64)
    base = oracle.likelihood_z_given_X(c["omega"], np.hstack([c["X"], np.zeros((35, 1))]), c["y"], 100.0)
    c["mask"][:] = False
    got = gpu_score(c)
    assert rel(got, np.full(64, base)) < RTOL


def test_empty_candidate_set_is_ok():
    import torch
    from mitre_b200 import device as D
    c = MG.scoring_inputs(seed=8, N=20, p=2, P=4)
    packed = torch.empty((0, 1), dtype=torch.int64, device=D.dev())
    out, status = D.score_all(packed, 20, c["X"], c["omega"], c["y"] - 0.5, 100.0)
    assert out.shape[0] == 0
    D.raise_on_status(status)


def test_not_positive_definite_is_reported_as_linalgerror():
    c = MG.scoring_inputs(seed=9, N=16, p=2, P=32)
    c["omega"][3] = -50.0
    with pytest.raises(np.linalg.LinAlgError):
        gpu_score(c)


def test_dropin_module_same_signature_as_reference():
    """mitre_b200.efficient_likelihoods.likelihoods_of_all_options(X, omega, response, truth_f64, s2)
    with host arrays, exactly the call at logit_rules.py:225-228."""
    from mitre_b200 import efficient_likelihoods as EL
    for name in ["score_n35_p2", "score_n70_p4", "score_n12_p1"]:
        c = MG.scoring_inputs(**MG.SCORING_CASES[name])
        gold = np.load(os.path.join(GOLD, name + ".npz"))["loglik"]
        opts = (c["truth"] * c["mask"]).astype(np.float64)
        got = EL.likelihoods_of_all_options(c["X"].astype(np.float64), c["omega"],
                                            (c["y"] - 0.5) / c["omega"], opts, np.float64(100.0))
        assert isinstance(got, np.ndarray) and got.dtype == np.float64
        assert rel(got, gold) < RTOL
    with pytest.raises(np.linalg.LinAlgError):
        om = c["omega"].copy()
        om[0] = -3.0
        EL.likelihoods_of_all_options(c["X"], om, (c["y"] - 0.5) / om, opts, 100.0)


def test_dropin_host_entry_chunked_pipeline():
    """More rows than two staging chunks (set to 64 Ki rows here) exercises the double-buffered path, ragged
    last chunk."""
    from mitre_b200 import efficient_likelihoods as EL
    from mitre_b200._lib import lib
    c = MG.scoring_inputs(seed=55, N=35, p=2, P=300001)
    opts = (c["truth"] * c["mask"]).astype(np.float64)
    z = (c["y"] - 0.5) / c["omega"]
    lib().mitre_set_host_chunk_bytes(8 * 65536)
    try:
        got = EL.likelihoods_of_all_options(c["X"], c["omega"], z, opts, 100.0)
    finally:
        lib().mitre_set_host_chunk_bytes(0)
    sel = np.r_[0:500, 65300:65800, 131000:131400, 299500:300001]
    want = oracle.likelihoods_of_all_options(c["X"], c["omega"], z, opts[sel], 100.0)
    assert rel(got[sel], want) < RTOL
    # every row equal to a row in the checked set must have the same value (pure function of the row)
    assert np.isfinite(got).all()


def test_dense_batch_vs_golden_and_oracle():
    from mitre_b200 import device as D
    gold = np.load(os.path.join(GOLD, "dense.npz"))["loglik"]
    for (seed, N, p), g in zip(MG.DENSE_CASES, gold):
        c = MG.dense_inputs(seed, N, p)
        out, status = D.likelihood_z_given_X_batch(c["X"][None], c["omega"], c["y"] - 0.5, c["sigma2"])
        assert abs(float(out[0]) - g) / abs(g) < RTOL
        D.raise_on_status(status)
    # ragged batch: different column counts in one launch (move_primitive, logit_rules.py:1375-1401)
    c = MG.dense_inputs(40, 35, 6)
    Xs = np.zeros((5, 35, 6))
    ps = [1, 3, 6, 2, 5]
    for b, pb in enumerate(ps):
        Xs[b, :, :pb] = c["X"][:, 6 - pb:]
    out, status = D.likelihood_z_given_X_batch(Xs, c["omega"], c["y"] - 0.5, 100.0, p_b=ps)
    out = out.cpu().numpy()
    for b, pb in enumerate(ps):
        want = oracle.likelihood_z_given_X(c["omega"], Xs[b, :, :pb], c["y"], 100.0)
        assert abs(out[b] - want) / abs(want) < RTOL


def test_mc_logpdf_np_dropin():
    from mitre_b200.mvnpdf_cholesky import mc_logpdf_np
    gold = np.load(os.path.join(GOLD, "dense.npz"))["loglik"]
    for (seed, N, p), g in zip(MG.DENSE_CASES, gold):
        c = MG.dense_inputs(seed, N, p)
        z = (c["y"] - 0.5) / c["omega"]
        cov = np.diag(1.0 / c["omega"]) + c["sigma2"] * c["X"].dot(c["X"].T)
        assert abs(mc_logpdf_np(z, cov) - g) / abs(g) < RTOL
    rng = np.random.RandomState(0)
    A = rng.normal(size=(1100, 1100))
    cov = A.dot(A.T) / 1100 + np.eye(1100)
    x = rng.normal(size=1100)
    assert abs(mc_logpdf_np(x, cov) - oracle.mc_logpdf_np_numpy(x, cov)) / 1500 < 1e-10
    with pytest.raises(np.linalg.LinAlgError):
        mc_logpdf_np(np.ones(3), -np.eye(3))


def test_pack_unpack_round_trip():
    import torch
    from mitre_b200 import device as D
    from oracle.detectors_oracle import pack_rows
    rng = np.random.RandomState(2)
    for N in [1, 31, 32, 33, 64, 65, 130]:
        t = rng.rand(257, N) > 0.5
        packed = D.pack_rows(t)
        assert np.array_equal(packed.cpu().numpy().view(np.uint64), pack_rows(t))
        assert np.array_equal(D.pack_rows(t.astype(np.float64)).cpu().numpy(), packed.cpu().numpy())
        assert np.array_equal(D.unpack_rows(packed, N).cpu().numpy().astype(bool), t)
        assert np.array_equal(D.pack_mask(t[0]).view(np.uint64), pack_rows(t[:1])[0])


def test_draw_index_and_stats():
    import torch
    from mitre_b200 import device as D
    rng = np.random.RandomState(4)
    for P in [1, 7, 2048, 2049, 100003]:
        a = rng.normal(-80, 3, size=P)
        b = rng.normal(-10, 1, size=P)
        da, db = torch.as_tensor(a, device=D.dev()), torch.as_tensor(b, device=D.dev())
        st = D.weight_stats(da, db).cpu().numpy()
        w = a + b
        assert st[0] == w.max()
        assert abs(st[0] + np.log(st[1]) - (w.max() + np.log(np.exp(w - w.max()).sum()))) < 1e-10
        for u in [0.0, 0.25, 0.5, 0.93]:
            want = oracle.choose_from_relative_probabilities(np.exp(w), u)
            got = int(D.draw_index(da, db, u).item())
            got_s = int(D.draw_index(da, db, u, stabilise=True).item())
            assert got == want and got_s == want, (P, u, got, got_s, want)
    # the reference's underflow edge (SURVEY hard part 5): exp() of all weights is 0 -> index 0
    a = torch.full((5000,), -4000.0, dtype=torch.float64, device=D.dev())
    assert int(D.draw_index(a, None, 0.7).item()) == 0
    assert 0 < int(D.draw_index(a, None, 0.7, stabilise=True).item()) < 5000


def test_score_all_stats_fused_logsumexp():
    """mitre_score_all_stats: (max, sum-exp) of loglik (+ log prior) reduced inside the scoring
    kernel must match a float64 log-sum-exp of the same vector, for N <= 64 and the wide path."""
    import torch
    from mitre_b200 import device as D
    for N, p, P in [(35, 4, 300001), (64, 2, 5000), (130, 3, 700), (35, 4, 1)]:
        c = MG.scoring_inputs(seed=900 + N, N=N, p=p, P=P)
        packed = D.pack_rows(c["truth"])
        mask = torch.as_tensor(D.pack_mask(c["mask"]), device=D.dev())
        rng = np.random.RandomState(1)
        prior = rng.normal(-12, 2, size=P)
        prior[::7] = -np.inf if P > 10 else prior[::7]
        dprior = torch.as_tensor(prior, device=D.dev())
        out, status, stats = D.score_all_stats(packed, N, c["X"], c["omega"], c["y"] - 0.5, 100.0, mask=mask,
                                               log_prior=dprior)
        ll = out.cpu().numpy()
        D.raise_on_status(status)
        st = stats.cpu().numpy()
        w = ll + prior
        want = float(torch.logsumexp(torch.as_tensor(w), 0))
        assert abs(st[0] + np.log(st[1]) - want) < 1e-11 * max(1.0, abs(want))
        assert st[0] == w.max()
        out2, status2, stats2 = D.score_all_stats(packed, N, c["X"], c["omega"], c["y"] - 0.5, 100.0, mask=mask)
        st2 = stats2.cpu().numpy()
        want2 = float(torch.logsumexp(torch.as_tensor(ll), 0))
        assert abs(st2[0] + np.log(st2[1]) - want2) < 1e-11 * abs(want2)


@pytest.mark.parametrize("split,dedupe", [(0, 1), (0, 2), (0, 0), (1, 0)])
def test_large_table_kernel_parity(split, dedupe):
    """The large-table kernels (default for P >= 4 Mi rows) forced onto small cases: reference golden
    vectors, odd shapes, masks and the fused stats.  dedupe=1: score_w1d_kernel (the default: runs of
    equal rows scored once); dedupe=0, split=0: every table in L2, one walk per row (score_w1g4_kernel);
    split=1: last subjects' table in shared memory (score_w1s4_kernel)."""
    import torch
    from mitre_b200 import device as D
    from mitre_b200._lib import lib
    assert lib().mitre_set_large_table_min_rows(0) == 0
    assert lib().mitre_set_split_tables(split) == 0
    assert lib().mitre_set_dedupe(dedupe) == 0
    try:
        for name in sorted(MG.SCORING_CASES):
            c = MG.scoring_inputs(**MG.SCORING_CASES[name])
            gold = np.load(os.path.join(GOLD, name + ".npz"))["loglik"]
            assert rel(gpu_score(c), gold) < RTOL, name
        for N, p, P in [(1, 1, 9), (15, 2, 100), (16, 3, 333), (17, 1, 64), (32, 4, 1000), (33, 5, 257),
                        (48, 2, 500), (49, 7, 300), (64, 9, 400), (35, 24, 50)]:
            c = MG.scoring_inputs(seed=500 + N + p, N=N, p=p, P=P)
            opts = (c["truth"] * c["mask"]).astype(np.float64)
            want = oracle.likelihoods_of_all_options(c["X"], c["omega"], (c["y"] - 0.5) / c["omega"], opts,
                                                     c["sigma2"])
            assert rel(gpu_score(c), want) < RTOL, (N, p, P)
        c = MG.scoring_inputs(seed=77, N=35, p=4, P=50000)
        packed = D.pack_rows(c["truth"])
        out, status, stats = D.score_all_stats(packed, 35, c["X"], c["omega"], c["y"] - 0.5, 100.0)
        st = stats.cpu().numpy()
        want = float(torch.logsumexp(out, 0))
        assert abs(st[0] + np.log(st[1]) - want) < 1e-11 * abs(want)
    finally:
        lib().mitre_set_large_table_min_rows(4 << 20)
        lib().mitre_set_split_tables(0)
        lib().mitre_set_dedupe(1)


def _duplicated_sorted_rows(rng, N, n_distinct, P):
    """bool[P, N] lexsorted rows with long runs of equal rows (what a real lexsorted candidate table
    looks like: ~4 of 5 rows equal their predecessor at the S5 shape)."""
    distinct = rng.rand(n_distinct, N) > 0.5
    rows = distinct[rng.randint(n_distinct, size=P)]
    return rows[np.lexsort(np.rot90(rows))]


@pytest.mark.parametrize("N,p,P,n_distinct", [(35, 4, 40000, 3000), (35, 4, 5003, 40), (35, 2, 1025, 1),
                                              (64, 3, 30011, 20000), (17, 5, 7777, 500), (35, 7, 130000, 9000)])
def test_dedupe_kernel_runs_of_equal_rows(N, p, P, n_distinct):
    """score_w1d_kernel on lexsorted tables with runs of equal rows (every run scored once and copied to
    its members), ragged last tile, a mask, a log-prior with -inf entries: log-likelihoods against the C
    oracle, the fused (max, sum-exp), and the draw from the per-tile pairs the kernel leaves behind."""
    import torch
    from mitre_b200 import device as D
    from mitre_b200._lib import lib
    rng = np.random.RandomState(N * 1000 + p)
    c = MG.scoring_inputs(seed=4000 + N + p, N=N, p=p, P=8)
    truth = _duplicated_sorted_rows(rng, N, n_distinct, P)
    assert lib().mitre_set_large_table_min_rows(0) == 0
    try:
        packed = D.pack_rows(truth)
        mask = torch.as_tensor(D.pack_mask(c["mask"]), device=D.dev())
        prior = rng.normal(-12, 2, size=P)
        prior[::11] = -np.inf
        dprior = torch.as_tensor(prior, device=D.dev())
        draw = D.DrawBuffers(P)
        out, status, stats = D.score_all_stats(packed, N, c["X"], c["omega"], c["y"] - 0.5, 100.0, mask=mask,
                                               log_prior=dprior, draw_buffers=draw)
        got = out.cpu().numpy()
        D.raise_on_status(status)
        opts = (truth * c["mask"]).astype(np.float64)
        want = oracle.likelihoods_of_all_options(c["X"], c["omega"], (c["y"] - 0.5) / c["omega"], opts, 100.0)
        assert rel(got, want) < RTOL
        plain, _ = D.score_all(packed, N, c["X"], c["omega"], c["y"] - 0.5, 100.0, mask=mask)
        assert np.array_equal(plain.cpu().numpy(), got)
        w = got + prior
        st = stats.cpu().numpy()
        assert st[0] == w.max()
        assert abs(st[0] + np.log(st[1]) - float(torch.logsumexp(torch.as_tensor(w), 0))) < 1e-11 * abs(st[0])
        e = np.exp(w - w.max())
        cum = np.cumsum(e)
        for u in [0.0, 0.11, 0.5, 0.77, 0.999]:
            want_k = int(np.sum(cum < u * cum[-1]))
            k = int(D.draw_index(out, dprior, u, stabilise=True, buffers=draw, reuse_tiles=True).item())
            assert k == want_k, (u, k, want_k)
            k2 = int(D.draw_index(out, dprior, u, stabilise=True).item())
            assert k2 == want_k
    finally:
        lib().mitre_set_large_table_min_rows(4 << 20)


@pytest.mark.parametrize("tables", [1, 0])
def test_wide_path_parity(tables):
    """N > 64 (several 64-subject words per row): byte-table kernel (default) and the per-set-bit
    kernel against the C oracle, odd subject counts, masks, several candidates per warp and one."""
    from mitre_b200._lib import lib
    assert lib().mitre_set_wide_tables(tables) == 0
    try:
        for N, p, P in [(65, 1, 40), (72, 3, 333), (100, 4, 257), (129, 2, 100), (200, 5, 64), (257, 7, 129),
                        (600, 4, 50), (2000, 4, 12)]:
            c = MG.scoring_inputs(seed=900 + N + p, N=N, p=p, P=P)
            opts = (c["truth"] * c["mask"]).astype(np.float64)
            want = oracle.likelihoods_of_all_options(c["X"], c["omega"], (c["y"] - 0.5) / c["omega"], opts,
                                                     c["sigma2"])
            assert rel(gpu_score(c), want) < RTOL, (N, p, P)
    finally:
        lib().mitre_set_wide_tables(1)
