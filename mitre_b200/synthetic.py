"""Synthetic MITRE-shaped inputs (host side, numpy only).

The reference's example datasets are absent from its tree (.MISSING_LARGE_BLOBS) and its
simulator needs them as templates (simulated_data.py:14-23, 391-426), so every measurement
and parity case is generated here with the *shape* simulated_data.py produces:

  * irregular per-subject sample times: a regular grid + jitter, first/last sample not moved
    outward                                              (simulated_data.py:497-517)
  * outcomes alternate case / control                    (simulated_data.py:663-669)
  * log-normal baseline x Gaussian random walk, clipped, renormalised per time point to
    relative abundance                                   (simulated_data.py:216-259, 345-347)
  * a fraction of (subject, variable) series all zero (zero inflation => exact ties)
  * one planted clade x window perturbation in the cases (simulated_data.py:606-675)
  * variable_weights ~ LogNormal(0, 1)

Shapes S1..S5 are the stand-ins for BASELINE.json's configs.
"""
import numpy as np

SHAPES = {
    # name: N subjects, V variables, (Tmin, Tmax) samples/subject, (t_start, t_end),
    #       N_intervals, tmin, tmax
    "S1": dict(N=40, V=60, T=(4, 10), span=(140.0, 210.0), N_intervals=1, tmin=1.0, tmax=100.0),
    "S2": dict(N=35, V=200, T=(8, 25), span=(0.0, 375.0), N_intervals=12, tmin=1.0, tmax=180.0),
    "S3": dict(N=2000, V=5000, T=(50, 50), span=(0.0, 375.0), N_intervals=12, tmin=1.0,
               tmax=180.0),
    "S5": dict(N=35, V=10000, T=(8, 25), span=(0.0, 375.0), N_intervals=24, tmin=1.0,
               tmax=None),
}


def sample_times(rng, n, t_start, t_end):
    """n irregular, strictly increasing sample times in [t_start, t_end]."""
    delta = (t_end - t_start) / float(n)
    base = np.linspace(t_start, t_end, n, endpoint=False) + 0.5 * delta
    jitter = np.clip(rng.normal(0.0, 0.1 * delta, size=n), -0.249 * delta, 0.249 * delta)
    jitter[0] = max(jitter[0], 0.0)
    jitter[-1] = min(jitter[-1], 0.0)
    return base + jitter


def make_series(N, V, T=(8, 25), span=(0.0, 375.0), seed=0, zero_fraction=0.1,
                perturb=True, dtype=np.float64, allow_constant=False, **_ignored):
    """Returns dict(X=list of [V, T_s] arrays, T=list of [T_s] arrays, y=bool[N],
    variable_weights=float64[V], experiment_start, experiment_end)."""
    rng = np.random.RandomState(seed)
    t_start, t_end = span
    y = (np.arange(N) % 2 == 0)
    X, Ts = [], []
    log_baseline_mean = rng.normal(-4.0, 1.5, size=V)
    pert_vars = rng.choice(V, size=max(1, V // 50), replace=False)
    pert_window = (t_start + 0.3 * (t_end - t_start), t_start + 0.6 * (t_end - t_start))
    for s in range(N):
        n_t = rng.randint(T[0], T[1] + 1)
        t = sample_times(rng, n_t, t_start, t_end)
        base = np.exp(log_baseline_mean + rng.normal(0.0, 0.5, size=V))
        walk = np.cumsum(rng.normal(0.0, 0.15, size=(V, n_t)), axis=1)
        x = base[:, None] * np.exp(walk)
        if perturb and y[s]:
            inside = (t >= pert_window[0]) & (t <= pert_window[1])
            x[np.ix_(pert_vars, inside)] *= 8.0
        x = np.clip(x, 0.0, 1.0)
        zero = rng.rand(V) < zero_fraction
        if not allow_constant and V - zero.sum() < 2:
            # allow_constant=False is for the parity fixtures on tiny V that compare WHOLE tables with the
            # reference: a subject with a single non-zero variable has the CONSTANT relative abundance 1.0,
            # whose least-squares slope is LAPACK rounding noise in the reference; keep two variables alive.
            # The named shapes (make_shape: bench.py, tools) are generated unsanitised; the product flags
            # such groups (slope guard band, DiscreteRulePopulation.slope_groups_flagged) and bench.py
            # reports the count.
            zero[rng.choice(V, size=min(2, V), replace=False)] = False
        x[zero, :] = 0.0
        tot = x.sum(axis=0)
        tot[tot == 0.0] = 1.0
        x = x / tot
        X.append(np.ascontiguousarray(x, dtype=dtype))
        Ts.append(t)
    weights = rng.lognormal(0.0, 1.0, size=V)
    return dict(X=X, T=Ts, y=y, variable_weights=weights,
                experiment_start=float(t_start), experiment_end=float(t_end))


def make_shape(name, seed=0, V=None, N=None):
    """Series for one of the named shapes (optionally with fewer variables/subjects)."""
    cfg = dict(SHAPES[name])
    if V is not None:
        cfg["V"] = V
    if N is not None:
        cfg["N"] = N
    data = make_series(seed=seed, allow_constant=True, **cfg)
    data["N_intervals"] = cfg["N_intervals"]
    data["tmin"] = cfg["tmin"]
    data["tmax"] = cfg["tmax"]
    return data


def make_sweep_state(N, n_rules, truth_rows_bool, seed=0):
    """A plausible mid-chain scoring state: omega ~ Gamma(2, 1/8) + 0.05 (PG(1,psi)-like),
    y alternating, a base design matrix of `n_rules` rule columns drawn from the candidate
    rows plus the constant column (rules.py:1173-1175), and a rule mask."""
    rng = np.random.RandomState(seed)
    omega = rng.gamma(2.0, 1.0 / 8.0, size=N) + 0.05
    y = (np.arange(N) % 2 == 0)
    P = truth_rows_bool.shape[0]
    cols = [truth_rows_bool[rng.randint(P)].astype(np.float64) for _ in range(n_rules)]
    X = np.stack(cols + [np.ones(N)], axis=1)
    mask = truth_rows_bool[rng.randint(P)].copy() if P else np.ones(N, dtype=bool)
    return dict(X=X, omega=omega, y=y, mask=mask)
