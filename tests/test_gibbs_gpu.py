"""The omega / beta Gibbs loop on the device: Polya-Gamma moments, the beta | omega posterior,
and a chain that uses it."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_device_polya_gamma_moments():
    import torch
    from mitre_b200 import device as D
    n = 400000
    for z in [0.0, 0.7, 2.5, 8.0, -3.0]:
        out = D.pg_draw(torch.full((n,), z, dtype=torch.float64, device=D.dev()), seed=17).cpu().numpy()
        za = abs(z)
        mean = 0.25 if za == 0 else np.tanh(za / 2) / (2 * za)
        var = 1 / 24.0 if za == 0 else (np.sinh(za) - za) / (4 * za ** 3 * np.cosh(za / 2) ** 2)
        assert np.all(out > 0)
        assert abs(out.mean() - mean) < 5 * np.sqrt(var / n)
        assert abs(out.var() - var) < 0.03 * var
    a = D.pg_draw(torch.linspace(0, 5, 1000, dtype=torch.float64, device=D.dev()), seed=3).cpu().numpy()
    b = D.pg_draw(torch.linspace(0, 5, 1000, dtype=torch.float64, device=D.dev()), seed=3).cpu().numpy()
    c = D.pg_draw(torch.linspace(0, 5, 1000, dtype=torch.float64, device=D.dev()), seed=4).cpu().numpy()
    assert np.array_equal(a, b) and not np.array_equal(a, c)


def test_device_beta_posterior_given_omega():
    from mitre_b200 import device as D
    rng = np.random.RandomState(0)
    N, p, v = 35, 4, 100.0
    X = (rng.rand(N, p) > 0.5).astype(float)
    X[:, -1] = 1.0
    omega = rng.gamma(2.0, 0.125, size=N) + 0.05
    kappa = (rng.rand(N) > 0.5) - 0.5
    V = np.linalg.inv(np.eye(p) / v + X.T.dot(omega[:, None] * X))
    mean = V.dot(X.T.dot(kappa))
    n = 20000
    _om, _b, betas, status = D.gibbs_omega_beta(X, kappa, v, np.zeros(p), n, seed=5, omega=omega)
    D.raise_on_status(status)
    B = betas.cpu().numpy()
    se = np.sqrt(np.diag(V) / n)
    assert np.all(np.abs(B.mean(axis=0) - mean) < 5 * se)
    assert np.max(np.abs(np.cov(B.T) - V)) < 0.08 * np.max(np.abs(V))


def test_full_sweep_matches_host_gibbs_in_distribution():
    """10 sweeps from the same start, many seeds: the marginal of the final beta agrees with the host
    sampler (mitre_b200/polyagamma.py + numpy) in mean."""
    from mitre_b200 import device as D, polyagamma
    rng = np.random.RandomState(1)
    N, p, v = 30, 3, 100.0
    X = (rng.rand(N, p) > 0.5).astype(float)
    X[:, -1] = 1.0
    y = rng.rand(N) > 0.4
    kappa = y - 0.5
    dev_final = []
    for s in range(300):
        _om, b, _bs, status = D.gibbs_omega_beta(X, kappa, v, np.zeros(p), 10, seed=1000 + s)
        dev_final.append(b.cpu().numpy())
    dev_final = np.array(dev_final)
    pg = polyagamma.PyPolyaGamma(2)
    host_final = []
    for s in range(300):
        beta = np.zeros(p)
        for _ in range(10):
            omega = np.array([pg.pgdraw(1, z) for z in X.dot(beta)])
            V = np.linalg.inv(np.eye(p) / v + X.T.dot(omega[:, None] * X))
            beta = rng.multivariate_normal(V.dot(X.T.dot(kappa)), V)
        host_final.append(beta)
    host_final = np.array(host_final)
    se = np.sqrt(dev_final.var(axis=0) / 300 + host_final.var(axis=0) / 300)
    assert np.all(np.abs(dev_final.mean(axis=0) - host_final.mean(axis=0)) < 5 * se)


def test_chain_with_device_gibbs_runs_and_is_seed_deterministic():
    from mitre_b200 import chains
    data, model = chains.build_model("S2", seed=3, variables=6)
    r0 = chains.initial_rule_list(model)
    a = chains.run_chain(model, 5, 25, r0, device_gibbs=True)
    b = chains.run_chain(model, 5, 25, r0, device_gibbs=True)
    assert a.comments == b.comments
    assert np.all(np.isfinite(a.likelihoods)) and len(a.states) == 25
    assert np.all(a.current_state.omega > 0)


def test_device_gibbs_chains_mix_with_host_gibbs_chains():
    """Chain-level check of the device omega/beta sweeps (pypolyagamma is absent and unpinned, so the device
    Philox / Devroye stream cannot be compared draw for draw): four chains with the host Polya-Gamma sampler
    and four with the device kernel, same model, pooled into one split R-hat (the reference's own
    convergence tool, mcmc_diagnostics.py:57-84).  If the device sweeps sampled a different conditional the
    two families would not mix and R-hat of the pooled runs would sit well above that of either family."""
    from mitre_b200 import chains
    data, model = chains.build_model("S2", seed=11, variables=5)
    r0 = chains.initial_rule_list(model)
    steps = 240
    host = [chains.scalar_traces(chains.run_chain(model, 100 + i, steps, r0, device_gibbs=False))[40:] for i in range(4)]
    dev = [chains.scalar_traces(chains.run_chain(model, 200 + i, steps, r0, device_gibbs=True))[40:] for i in range(4)]
    cols = [2, 3]                                        # log-likelihood, log-prior: continuous, every step
    r_host = chains.split_r_hat([t[:, cols] for t in host])
    r_dev = chains.split_r_hat([t[:, cols] for t in dev])
    r_all = chains.split_r_hat([t[:, cols] for t in host + dev])
    assert np.all(np.isfinite(r_all))
    assert np.all(r_all < 1.3), (r_host, r_dev, r_all)
    assert np.all(r_all < np.maximum(r_host, r_dev) + 0.15), (r_host, r_dev, r_all)
