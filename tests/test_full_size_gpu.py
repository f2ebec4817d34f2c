"""BASELINE.json's bench configuration at FULL size (S5: 35 subjects, 10 000 variables, 24 intervals,
mean + slope, full threshold grid -> 2.9e8 candidates on one GPU), checked through the
size-independent properties SURVEY 8c lists -- the oracle cannot enumerate this many candidates:

  * rows lexsorted (rules.py:450), stable in enumeration order;
  * every 'below' row is the complement of its 'above' row (rules.py:494-497);
  * K_g <= N - 1 and P = 2 sum K_g;
  * all-False mask => every candidate's likelihood is the base likelihood (the appended column is 0);
  * likelihoods_of_all_options[k] == likelihood_z_given_X([X | v_k]) (logit_rules.py:944-961) on
    sampled k, against the dense kernel AND the C oracle;
  * the reference Cython kernel itself on a lexsorted slice (when oracle/_ref is built).
"""
import os

import numpy as np
import pytest

import oracle
from mitre_b200 import synthetic

pytestmark = pytest.mark.gpu

RTOL = 1e-9      # BASELINE.json: <= 1e-9 relative on log-likelihoods


@pytest.fixture(scope="module")
def s5():
    import torch
    from mitre_b200 import rules as R
    if torch.cuda.mem_get_info()[1] < 60 * 2 ** 30:
        pytest.skip("needs a 60 GB+ GPU")
    d = synthetic.make_shape("S5", seed=0)
    V = d["X"][0].shape[0]
    data = R.Dataset(d["X"], d["T"], d["y"], ["v%d" % i for i in range(V)], d["variable_weights"],
                     d["experiment_start"], d["experiment_end"])
    pop = R.DiscreteRulePopulation(None, d["tmin"], d["N_intervals"], tmax=d["tmax"], data=data)
    yield data, pop
    del pop, data
    torch.cuda.empty_cache()


def test_full_size_table_structure(s5):
    import torch
    data, pop = s5
    N = data.n_subjects
    P = len(pop)
    assert N == 35 and data.n_variables == 10000 and P > 2.5e8
    K = pop._K_host
    assert int(K.max()) <= N - 1 and P == 2 * int(K.astype(np.int64).sum())
    rows = pop.packed_truth[:, 0]
    key = rows ^ torch.iinfo(torch.int64).min          # unsigned order of the packed words
    assert bool((key[1:] >= key[:-1]).all()), "rows must be lexsorted"
    perm = pop._perm
    same = key[1:] == key[:-1]
    assert bool((perm[1:][same] > perm[:-1][same]).all()), "equal rows keep enumeration order"
    del key, same
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(P, device=perm.device)
    valid = torch.tensor(-(1 << (64 - N)), dtype=torch.int64, device=rows.device)   # top N bits set
    above, below = rows[inv[0::2]], rows[inv[1::2]]
    assert bool((below == (~above & valid)).all()), "'below' row must be the complement of 'above'"
    assert bool(((above & ~valid) == 0).all())
    del inv, above, below
    torch.cuda.empty_cache()


def test_full_size_sweep_identities(s5):
    import torch
    from mitre_b200 import device as D
    data, pop = s5
    N, P = data.n_subjects, len(pop)
    rng = np.random.RandomState(3)
    state = synthetic.make_sweep_state(N, 3, np.stack([pop._truth_row(int(rng.randint(P))) for _ in range(64)]),
                                       seed=4)
    X, omega, kappa = state["X"], state["omega"], state["y"] - 0.5
    sigma2 = 100.0
    buffers = D.ScoreBuffers(P, N, X.shape[1])
    # (1) all-False mask: the appended column is zero for every candidate
    zero_mask = torch.as_tensor(D.pack_mask(np.zeros(N, dtype=bool)), device=D.dev())
    out, status = D.score_all(pop.packed_truth, N, X, omega, kappa, sigma2, mask=zero_mask, buffers=buffers)
    D.raise_on_status(status)
    Xz = np.concatenate([X, np.zeros((N, 1))], axis=1)
    base, st = D.likelihood_z_given_X_batch(Xz[None], omega, kappa, sigma2)
    base = float(base.item())
    assert float((out - base).abs().max()) <= RTOL * abs(base)
    assert abs(base - oracle.likelihood_z_given_X(omega, Xz, state["y"], sigma2)) <= RTOL * abs(base)
    # (2) sampled candidates under a real mask: sweep == dense kernel == C oracle
    mask = state["mask"]
    d_mask = torch.as_tensor(D.pack_mask(mask), device=D.dev())
    out, status = D.score_all(pop.packed_truth, N, X, omega, kappa, sigma2, mask=d_mask, buffers=buffers)
    D.raise_on_status(status)
    assert bool(torch.isfinite(out).all())
    ks = np.unique(np.concatenate([[0, 1, P // 2, P - 2, P - 1], rng.randint(0, P, size=59)]))
    got = out[torch.as_tensor(ks, device=out.device)].cpu().numpy()
    cols = np.stack([pop._truth_row(int(k)) & mask for k in ks]).astype(np.float64)
    Xs = np.stack([np.concatenate([X, c[:, None]], axis=1) for c in cols])
    dense, st = D.likelihood_z_given_X_batch(Xs, omega, kappa, sigma2)
    dense = dense.cpu().numpy()
    assert np.max(np.abs(got - dense) / np.abs(dense)) <= RTOL
    want = np.array([oracle.likelihood_z_given_X(omega, x, state["y"], sigma2) for x in Xs])
    assert np.max(np.abs(got - want) / np.abs(want)) <= RTOL
    # (3) the reference's own Cython kernel on a lexsorted slice of the table
    from oracle import build_ref
    try:
        ref = build_ref.load()
    except Exception:
        ref = None
    if ref is not None:
        a = int(rng.randint(0, P - 20000))
        sl = pop.packed_truth[a:a + 20000].cpu().numpy().view(np.uint64)
        from oracle import detectors_oracle as DO
        truth = DO.unpack_rows(sl, N)
        opts = (truth & mask).astype(np.float64)
        want = ref.likelihoods_of_all_options(X.astype(np.float64), omega, kappa / omega, opts, sigma2)
        got = out[a:a + 20000].cpu().numpy()
        assert np.max(np.abs(got - want) / np.abs(want)) <= RTOL
    # (4) the (max, sum-exp) fused into the launch agrees with a reduction of the written vector
    out2, status, stats = D.score_all_stats(pop.packed_truth, N, X, omega, kappa, sigma2, mask=d_mask,
                                            buffers=buffers)
    st = stats.cpu().numpy()
    lse = float(torch.logsumexp(out2, 0))
    assert abs(st[0] + np.log(st[1]) - lse) <= 1e-11 * abs(lse)
    del buffers, out, out2
    torch.cuda.empty_cache()
