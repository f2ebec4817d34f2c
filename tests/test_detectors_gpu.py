"""GPU parity of detector evaluation against the golden vectors generated from the reference's
rules.py and against the numpy oracle.  Truth tables, sort order and mean values/thresholds are
compared BIT-EXACTLY; slope values within 1e-9 relative (closed form vs LAPACK gelsd)."""
import os

import numpy as np
import pytest

from mitre_b200 import synthetic
from oracle import detectors_oracle as DO
from tests.golden import make_golden as MG

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def build_gpu_population(d, pop_cfg, fused=True):
    """fused=True: the default device path (N <= 63): values kernel + staged sort/emit kernel;
    "tile" / "flat": other values kernels; "direct": sort/emit without shared-memory staging;
    "single": the one-kernel variant; False: phase-by-phase kernels (any N)."""
    from mitre_b200 import rules as R
    from mitre_b200._lib import lib
    lib().mitre_set_detector_split({"single": 0, "flat": 2, "tile": 3, "direct": 4}.get(fused, 1))
    lib().mitre_set_sort_emit_rolled(1 if fused == "rolled" else 0)
    walk = fused is True or fused == "rolled"   # the default values kernel; "resident" etc.: mitre_detectors_fused
    fused = bool(fused)
    V = d["X"][0].shape[0]
    data = R.Dataset(d["X"], d["T"], d["y"], ["v%d" % i for i in range(V)], d["variable_weights"],
                     d["experiment_start"], d["experiment_end"])
    old, old_walk = R.DiscreteRulePopulation.use_fused, R.DiscreteRulePopulation.use_walk
    R.DiscreteRulePopulation.use_fused = fused
    R.DiscreteRulePopulation.use_walk = walk
    try:
        pop = R.DiscreteRulePopulation(None, pop_cfg["tmin"], pop_cfg["N_intervals"], tmax=pop_cfg["tmax"],
                                       data=data)
    finally:
        R.DiscreteRulePopulation.use_fused = old
        R.DiscreteRulePopulation.use_walk = old_walk
        lib().mitre_set_detector_split(1)
        lib().mitre_set_sort_emit_rolled(0)
    return data, pop


def check_population(pop, gold_windows, gold_ok, gold_groups, gold_values, gold_order, gold_packed,
                     gold_rule, N):
    assert np.array_equal(np.array(pop.windows, dtype=np.float64).reshape(-1, 2), gold_windows)
    assert np.array_equal(np.array(pop.windows_ok_for_slope, dtype=bool), gold_ok)
    got_groups = np.stack([pop._group_window, pop._group_variable, pop._group_type], 1)
    assert np.array_equal(got_groups, gold_groups)
    vals = pop._values.cpu().numpy()
    is_mean = gold_groups[:, 2] == 1
    assert np.array_equal(vals[is_mean], gold_values[is_mean]), "window means must be bit-exact"
    sl, gs = vals[~is_mean], gold_values[~is_mean]
    if sl.size:
        scale = np.maximum(np.abs(gs).max(axis=1, keepdims=True), 1e-300)
        assert np.max(np.abs(sl - gs) / scale) < 1e-9
    assert pop.n_primitives == gold_packed.shape[0]
    got_packed = pop.packed_truth.cpu().numpy().view(np.uint64)
    assert np.array_equal(got_packed, gold_packed), "sorted truth table must be bit-exact"
    assert np.array_equal(pop._perm_host(), gold_order), "lexsort permutation must match np.lexsort"
    # rule descriptors in sorted order
    var, win, typ, dirn, thr = gold_rule
    P = len(var)
    idx = np.unique(np.r_[0:min(P, 50), np.linspace(0, P - 1, 200).astype(int)])
    for i in idx:
        t = pop.flat_rules[int(i)]
        assert t[0] == var[i] and tuple(t[1]) == tuple(gold_windows[win[i]])
        assert t[2] == ("slope", "average")[typ[i]] and t[3] == ("above", "below")[dirn[i]]
        if typ[i] == 1:
            assert t[4] == thr[i], "mean thresholds must be bit-exact"
        else:
            assert abs(t[4] - thr[i]) <= 1e-9 * max(abs(thr[i]), 1e-300) + 1e-300
        assert pop.primitive_index[t] == i
        assert np.array_equal(pop.truth_table[t], DO.unpack_rows(gold_packed[i:i + 1], N)[0])


@pytest.mark.parametrize("fused", [True, "rolled", "resident", "tile", "flat", "direct", "single", False])
@pytest.mark.parametrize("name", sorted(MG.DETECTOR_CASES))
def test_population_vs_reference_golden(name, fused):
    cfg = MG.DETECTOR_CASES[name]
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    d = synthetic.make_series(**cfg["series"])
    data, pop = build_gpu_population(d, cfg["pop"], fused)
    N = int(gold["n_subjects"])
    check_population(pop, gold["windows"], gold["ok_slope"], gold["groups"], gold["values"], gold["order"],
                     gold["sorted_truth_packed"],
                     (gold["rule_variable"], gold["rule_window"], gold["rule_type"], gold["rule_direction"],
                      gold["rule_threshold"]), N)
    assert np.array_equal(pop.flat_durations, gold["flat_durations"])
    assert np.array_equal(pop.flat_variable_weights, gold["flat_variable_weights"])
    assert np.array_equal(pop.flat_truth, DO.unpack_rows(gold["sorted_truth_packed"], N))


@pytest.mark.parametrize("fused", [True, "rolled", "resident", "tile", "flat", "direct", "single", False])
@pytest.mark.parametrize("N,V,T,nint", [(2, 3, (2, 3), 1), (3, 2, (2, 4), 2), (8, 40, (3, 6), 3), (9, 33, (3, 6), 3),
                                        (16, 3, (2, 5), 2), (24, 2, (3, 5), 2), (33, 5, (3, 20), 6),
                                        (40, 70, (3, 7), 2), (41, 2, (3, 7), 2), (48, 3, (3, 7), 2),
                                        (56, 2, (3, 7), 2), (63, 3, (3, 7), 2), (64, 3, (3, 7), 2),
                                        (100, 3, (5, 9), 4), (257, 2, (4, 6), 2)])
def test_population_vs_oracle(N, V, T, nint, fused):
    if not fused and N < 64 and N not in (3, 33, 63):
        pytest.skip("phase-by-phase path covered at a subset of sizes")
    d = synthetic.make_series(N=N, V=V, T=T, span=(0.0, 90.0), seed=N + V, zero_fraction=0.2)
    ref = DO.build_population(d["X"], d["T"], 0.0, 90.0, nint, 1.0, None)
    data, pop = build_gpu_population(d, dict(tmin=1.0, N_intervals=nint, tmax=None), fused)
    o = ref["order"]
    check_population(pop, ref["windows"], ref["ok_slope"],
                     np.stack([ref["group_window"], ref["group_variable"], ref["group_type"]], 1),
                     ref["values"], o, DO.pack_rows(ref["truth"][o]),
                     (ref["variable"][o], ref["window"][o], ref["type"][o], ref["direction"][o],
                      ref["threshold"][o]), N)


def test_structural_invariants_at_scale():
    """S2-shaped problem (200 variables): below == NOT above, rows sorted, K <= N-1,
    counts consistent -- the size-independent properties of SURVEY 8c."""
    d = synthetic.make_shape("S2", seed=0)
    data, pop = build_gpu_population(d, dict(tmin=d["tmin"], N_intervals=d["N_intervals"], tmax=d["tmax"]))
    N = 35
    rows = pop.packed_truth.cpu().numpy().view(np.uint64)[:, 0]
    assert np.all(rows[1:] >= rows[:-1])                                  # lexsorted
    K = pop._K_host
    assert K.max() <= N - 1 and pop.n_primitives == 2 * int(K.sum())
    valid = np.uint64(((1 << N) - 1) << (64 - N))
    perm = pop._perm_host()
    inv = np.empty_like(perm)
    inv[perm] = np.arange(len(perm))
    above = rows[inv[0::2]]
    below = rows[inv[1::2]]
    assert np.array_equal(below, ~above & valid)
    # stability: equal rows keep enumeration order
    same = rows[1:] == rows[:-1]
    assert np.all(perm[1:][same] > perm[:-1][same])


def test_no_windows_and_single_subject_edge_cases():
    d = synthetic.make_series(N=4, V=2, T=(3, 5), span=(0.0, 10.0), seed=1)
    data, pop = build_gpu_population(d, dict(tmin=100.0, N_intervals=3, tmax=None))   # no window long enough
    assert len(pop) == 0 and pop.windows == [] and pop.flat_truth.shape == (0, 4)
    d1 = synthetic.make_series(N=1, V=2, T=(6, 6), span=(0.0, 10.0), seed=2)
    data1, pop1 = build_gpu_population(d1, dict(tmin=1.0, N_intervals=2, tmax=None))
    assert len(pop1) == 0 and len(pop1.windows) > 0          # one subject => no thresholds


def test_update_truths_rebuilds_identically():
    """rules.py:300-323 regenerates the table WITHOUT sorting it again: enumeration order (same rows as before,
    permuted back by the lexsort's permutation); resort_after_update = True gives the sorted table again."""
    d = synthetic.make_series(N=20, V=3, T=(5, 9), span=(0.0, 40.0), seed=5)
    data, pop = build_gpu_population(d, dict(tmin=1.0, N_intervals=4, tmax=None))
    before = pop.packed_truth.cpu().numpy().copy()
    perm = pop._perm_host().copy()
    rules_before = list(pop.flat_rules)
    pop._update_truths()
    enum = np.empty_like(before)
    enum[perm] = before
    assert np.array_equal(pop.packed_truth.cpu().numpy(), enum)
    assert np.array_equal(pop._perm_host(), np.arange(len(perm)))
    assert [pop.flat_rules[int(perm[i])] for i in (0, 5, len(perm) - 1)] == [rules_before[i] for i in (0, 5, len(perm) - 1)]
    assert all(pop.primitive_index[pop.flat_rules[k]] == k for k in (0, 7, len(perm) - 1))
    pop.resort_after_update = True
    pop._update_truths()
    assert np.array_equal(pop.packed_truth.cpu().numpy(), before)
    assert np.array_equal(pop._perm_host(), perm)


@pytest.mark.parametrize("held_out", [0, 7, 19])
def test_update_truths_on_a_cross_validation_fold(held_out):
    """SURVEY 8f-4 (main.py:1192-1194, rules.py:300-323): leave-one-out keeps the population's windows
    (and their slope flags) and re-evaluates everything else on the training subjects; like the reference, the
    re-evaluated table stays in enumeration order."""
    from mitre_b200 import device as D, rules as R
    d = synthetic.make_series(N=20, V=3, T=(5, 9), span=(0.0, 40.0), seed=5)
    data, pop = build_gpu_population(d, dict(tmin=1.0, N_intervals=4, tmax=None))
    keep = [s for s in range(20) if s != held_out]
    X = [d["X"][s] for s in keep]
    T = [d["T"][s] for s in keep]
    train = R.Dataset(X, T, d["y"][keep], ["v%d" % i for i in range(3)], d["variable_weights"],
                      d["experiment_start"], d["experiment_end"])
    windows, ok = list(pop.windows), list(pop.windows_ok_for_slope)
    pop.data = train
    pop._update_truths()
    assert list(pop.windows) == windows and list(pop.windows_ok_for_slope) == ok
    groups = DO.calculate_values(X, T, windows, ok)
    ref = DO.enumerate_primitives(groups, len(keep))
    assert np.array_equal(pop.packed_truth.cpu().numpy().view(np.uint64), DO.pack_rows(ref["truth"]))
    assert pop.flat_truth.shape == (ref["truth"].shape[0], len(keep))
    assert train._primitive_result_cache is pop.truth_table
    # scoring does not need sorted rows: the unsorted fold table against the oracle
    c = synthetic.make_sweep_state(len(keep), 2, ref["truth"], seed=3)
    out, status = D.score_all(pop.packed_truth, len(keep), c["X"], c["omega"], c["y"] - 0.5, 100.0)
    D.raise_on_status(status)
    import oracle
    sel = np.r_[0:40, ref["truth"].shape[0] - 40:ref["truth"].shape[0]]
    want = oracle.likelihoods_of_all_options(c["X"], c["omega"], (c["y"] - 0.5) / c["omega"],
                                             ref["truth"][sel].astype(np.float64), 100.0)
    got = out.cpu().numpy()[sel]
    assert np.max(np.abs(got - want) / np.abs(want)) < 1e-9
    pop.resort_after_update = True
    pop._update_truths()
    order = DO.lexsort_rows(ref["truth"])
    assert np.array_equal(pop.packed_truth.cpu().numpy().view(np.uint64), DO.pack_rows(ref["truth"][order]))


def test_constant_series_slope_divergence_is_confined_to_slope_rows():
    """DESIGN.md "slope parity": a constant non-zero series has slope exactly 0 in closed form and
    ~1e-20 LAPACK noise in the reference, which changes the tie structure of that slope group only.
    Every 'average' primitive of the reference must still be present with an identical truth row."""
    cfg = MG.DIVERGENT_CASES["det_constslope"]
    gold = np.load(os.path.join(GOLD, "det_constslope.npz"))
    d = synthetic.make_series(**cfg["series"])
    data, pop = build_gpu_population(d, cfg["pop"])
    N = int(gold["n_subjects"])
    truth = DO.unpack_rows(gold["sorted_truth_packed"], N)
    n_avg = 0
    for i in np.nonzero(gold["rule_type"] == 1)[0]:
        key = (int(gold["rule_variable"][i]), tuple(gold["windows"][gold["rule_window"][i]]), "average",
               ("above", "below")[gold["rule_direction"][i]], gold["rule_threshold"][i])
        assert key in pop.truth_table
        assert np.array_equal(pop.truth_table[key], truth[i])
        n_avg += 1
    assert n_avg > 0
    # and the divergence really is there: some slope row differs
    got = pop.packed_truth.cpu().numpy().view(np.uint64)
    assert got.shape != gold["sorted_truth_packed"].shape or not np.array_equal(got, gold["sorted_truth_packed"])


def test_all_tied_group_and_sentinel_rows():
    """A variable that is zero for EVERY subject gives one threshold (0) whose 'below' row is all
    True -- the row that collides with the fused path's sentinel unless the padded sort uses N+1 bits."""
    for N in (8, 35, 40):
        d = synthetic.make_series(N=N, V=3, T=(4, 6), span=(0.0, 20.0), seed=N, zero_fraction=0.0)
        for x in d["X"]:
            x[1, :] = 0.0
        ref = DO.build_population(d["X"], d["T"], 0.0, 20.0, 2, 1.0, None)
        for fused in (True, False):
            data, pop = build_gpu_population(d, dict(tmin=1.0, N_intervals=2, tmax=None), fused)
            o = ref["order"]
            assert np.array_equal(pop.packed_truth.cpu().numpy().view(np.uint64), DO.pack_rows(ref["truth"][o]))
            assert np.array_equal(pop._perm_host(), o)


def test_long_windows_exceed_the_fused_staging_buffer():
    """Subjects with more in-window samples than one staging chunk holds take the direct-load path."""
    d = synthetic.make_series(N=6, V=34, T=(150, 200), span=(0.0, 50.0), seed=77, zero_fraction=0.0)
    ref = DO.build_population(d["X"], d["T"], 0.0, 50.0, 1, 1.0, None)
    data, pop = build_gpu_population(d, dict(tmin=1.0, N_intervals=1, tmax=None), True)
    o = ref["order"]
    assert np.array_equal(pop.packed_truth.cpu().numpy().view(np.uint64), DO.pack_rows(ref["truth"][o]))
    is_mean = ref["group_type"] == 1
    assert np.array_equal(pop._values.cpu().numpy()[is_mean], ref["values"][is_mean])


def test_exact_division_selftest():
    """Window means divide a pairwise sum by a small integer count with a 3-instruction exact
    sequence instead of the IEEE divide; it must be bit-identical (np.mean parity depends on it)."""
    from mitre_b200 import device as D
    rng = np.random.RandomState(0)
    n = rng.randint(1, 400, size=2000000).astype(np.int32)
    a = rng.rand(2000000) * 10.0 ** rng.uniform(-12, 3, size=2000000)
    a[:1000] = 0.0
    a[1000:2000] = n[1000:2000] * rng.rand(1000)          # near-exact quotients
    a[2000:3000] = np.nextafter(n[2000:3000].astype(np.float64), 0) * 3.0
    a[3000:3100] = 1e-300
    a[3100:3200] = 1e300
    a[3200:4200] = -a[3200:4200]
    fast, exact = D.selftest_divide(a, n)
    fast, exact = fast.cpu().numpy(), exact.cpu().numpy()
    assert np.array_equal(fast, exact)
    assert np.array_equal(exact, a / n)


@pytest.mark.parametrize("fused", [True, "rolled", "resident", "tile", "flat", "direct", "single", False])
def test_near_ties_below_key_resolution_take_the_exact_path(fused):
    """Values that differ only in their lowest mantissa bits collide in the 32-bit sort key (float32
    image of the value with the subject id in the low 6 bits); the kernel must detect the mis-ordered
    neighbours and repair them on the exact values.  Two-sample series have exact means."""
    from mitre_b200 import rules as R
    rng = np.random.RandomState(5)
    N, V = 20, 34
    base = 0.3
    X, T = [], []
    for s in range(N):
        t = np.array([1.0, 2.0])
        x = rng.rand(V, 2)
        delta = 2.0 ** -10 * (s + 1)          # distinct, well separated slopes; exact two-sample means
        # variable 0: the mean of subject s is base advanced by (N - s) ulps -> DEcreasing with the id,
        # so ordering by (truncated key, id) is exactly wrong; variable 1: a mix of ties and near-ties
        v0 = base
        for _ in range(N - s):
            v0 = np.nextafter(v0, 1.0)
        x[0, :] = [v0 - delta, v0 + delta]
        v1 = np.nextafter(base, 1.0) if s % 3 == 0 else (base if s % 3 == 1 else np.nextafter(base, 0.0))
        x[1, :] = [v1 - delta, v1 + delta]
        assert (x[0, 0] + x[0, 1]) / 2 == v0 and (x[1, 0] + x[1, 1]) / 2 == v1
        X.append(x)
        T.append(t)
    y = np.arange(N) % 2 == 0
    ref = DO.build_population(X, T, 0.0, 5.0, 1, 1.0, None)
    o = ref["order"]
    d = dict(X=X, T=T, y=y, variable_weights=np.ones(V), experiment_start=0.0, experiment_end=5.0)
    data, pop = build_gpu_population(d, dict(tmin=1.0, N_intervals=1, tmax=None), fused)
    is_mean = ref["group_type"] == 1
    assert np.array_equal(pop._values.cpu().numpy()[is_mean], ref["values"][is_mean])
    assert np.array_equal(pop._K_host[is_mean], ref["K"][is_mean]), (pop._K_host[is_mean][:4], ref["K"][is_mean][:4])
    assert np.array_equal(pop.packed_truth.cpu().numpy().view(np.uint64), DO.pack_rows(ref["truth"][o]))
    assert np.array_equal(pop._perm_host(), o)


# ---- slope guard band: groups whose rows are not guaranteed to be the reference's are FLAGGED ----------

def _group_rows(v):
    """The (above, below) truth rows of one group from its N values, as the reference builds them
    (rules.py:490-497, 536-539): sorted midpoints, unique, v > theta / v <= theta."""
    sv = np.sort(v)
    thr = np.unique(0.5 * (sv[:-1] + sv[1:]))
    return tuple((tuple(v > t), tuple(v <= t)) for t in thr)


def _criterion(values, group_type):
    """numpy restatement of the flag criterion (include/mitre_b200.h, mitre_slope_guard)."""
    flags = np.zeros(values.shape[0], dtype=bool)
    for g in np.nonzero(group_type == 0)[0]:
        v = values[g]
        sv = np.sort(v)
        scale = max(abs(sv[0]), abs(sv[-1]))
        d = np.diff(sv)
        near = np.any((d != 0) & (d <= 1e-10 * scale))
        const = np.any((v == 0) & (values[g + 1] != 0))
        flags[g] = near or const
    return flags


@pytest.mark.parametrize("fused", [True, "rolled", "resident", "direct", False])
def test_slope_guard_flags_the_constant_series_divergence(fused):
    """det_constslope (reference fixture): every group whose rows differ from the reference's is flagged,
    the flags are exactly the documented criterion, and nothing outside slope groups is flagged."""
    cfg = MG.DIVERGENT_CASES["det_constslope"]
    gold = np.load(os.path.join(GOLD, "det_constslope.npz"))
    d = synthetic.make_series(**cfg["series"])
    data, pop = build_gpu_population(d, cfg["pop"], fused)
    vals = pop._values.cpu().numpy()
    gtype = gold["groups"][:, 2]
    differ = np.array([_group_rows(vals[g]) != _group_rows(gold["values"][g]) for g in range(vals.shape[0])])
    flags = pop._slope_flags.cpu().numpy().astype(bool)
    assert differ.any(), "the fixture is meant to diverge"
    assert not differ[gtype == 1].any(), "average groups are bit-exact"
    assert np.all(flags[differ]), "a group whose rows differ from the reference must be flagged"
    assert np.array_equal(flags, _criterion(vals, gtype))
    assert pop.slope_groups_flagged == int(flags.sum()) > 0
    assert len(pop.flagged_groups()) == int(flags.sum())


@pytest.mark.parametrize("fused", [True, "rolled", False])
def test_slope_guard_flags_near_ties_and_nothing_else(fused):
    """Two subjects whose slopes agree to 1e-13: their order is not certain (closed form here, LAPACK gelsd
    in the reference), so the group is flagged; every unflagged group carries the oracle's rows."""
    rng = np.random.RandomState(11)
    N, V = 12, 6
    X, T = [], []
    for s in range(N):
        t = np.array([1.0 + 0.01 * s, 2.0, 3.5 + 0.02 * s])
        x = rng.rand(V, 3)
        X.append(x)
        T.append(t)
    # variable 2: subjects 3 and 7 get the same straight line up to 1e-13 (slope 0.25)
    for s, eps in ((3, 0.0), (7, 1e-13)):
        X[s][2, :] = 0.1 + 0.25 * (1.0 + eps) * (T[s] - T[s][0])
    y = np.arange(N) % 2 == 0
    d = dict(X=X, T=T, y=y, variable_weights=np.ones(V), experiment_start=0.0, experiment_end=5.0)
    ref = DO.build_population(X, T, 0.0, 5.0, 1, 1.0, None)
    data, pop = build_gpu_population(d, dict(tmin=1.0, N_intervals=1, tmax=None), fused)
    vals = pop._values.cpu().numpy()
    flags = pop._slope_flags.cpu().numpy().astype(bool)
    gtype = ref["group_type"]
    g_near = int(np.nonzero((pop._group_variable == 2) & (pop._group_type == 0))[0][0])
    assert flags[g_near] and flags.sum() == 1
    assert np.array_equal(flags, _criterion(vals, gtype))
    for g in np.nonzero(~flags)[0]:
        assert _group_rows(vals[g]) == _group_rows(ref["values"][g]), g


def test_s3_shape_population_vs_oracle():
    """BASELINE configs[2] shape (N = 2000 subjects, T = 50 samples, 12 intervals, t <= 180): the N > 64
    path end to end -- TMA-staged values kernel, block-wide threshold sort, rank-plane row kernel, 250-pass
    multi-word lexsort -- against the oracle for 4 variables and the 5-interval windows.  The permutation is
    checked without running np.lexsort over 2000 keys: the device order must (a) carry exactly the
    oracle's rows, (b) be lexicographically non-decreasing with subject 0 most significant and (c) keep
    enumeration order among equal rows; a stable sort is unique, so that IS np.lexsort's permutation."""
    shape = dict(synthetic.SHAPES["S3"], V=4)
    d = synthetic.make_series(seed=5, allow_constant=False, **shape)   # whole-table comparison: no constant series
    N = len(d["X"])
    assert N == 2000 and all(len(t) == 50 for t in d["T"])
    cfg = dict(tmin=140.0, N_intervals=12, tmax=180.0)
    data, pop = build_gpu_population(d, cfg, False)
    windows, ok_slope = DO.enumerate_windows(d["T"], d["experiment_start"], d["experiment_end"], 12, 140.0, 180.0)
    assert len(windows) == 8 and all(ok_slope)
    assert np.array_equal(np.array(pop.windows), np.array(windows))
    groups = DO.calculate_values(d["X"], d["T"], windows, ok_slope)
    gold_values = np.vstack([g[3] for g in groups])
    gtype = np.array([g[2] for g in groups])
    vals = pop._values.cpu().numpy()
    assert np.array_equal(vals[gtype == 1], gold_values[gtype == 1]), "window means must be bit-exact"
    scale = np.abs(gold_values[gtype == 0]).max(axis=1, keepdims=True)
    assert np.max(np.abs(vals[gtype == 0] - gold_values[gtype == 0]) / scale) < 1e-9
    prim = DO.enumerate_primitives(groups, N)
    flags = pop._slope_flags.cpu().numpy().astype(bool)
    assert np.array_equal(flags, _criterion(vals, gtype))
    assert not flags.any(), "this seed has no near-tie: the whole table must be the oracle's"
    # thresholds and K per group (slope thresholds to 1e-9: different slope arithmetic, same order)
    assert np.array_equal(pop._K_host, prim["K"])
    thr = pop._thresholds.cpu().numpy()
    off = 0
    for g, k in enumerate(prim["K"]):
        want = prim["thresholds"][off:off + k]
        if gtype[g] == 1:
            assert np.array_equal(thr[g, :k], want)
        else:
            assert np.max(np.abs(thr[g, :k] - want)) <= 1e-9 * np.abs(want).max()
        off += k
    rows_enum = DO.pack_rows(prim["truth"])
    P = rows_enum.shape[0]
    assert len(pop) == P
    perm = pop._perm_host()
    got = pop.packed_truth.cpu().numpy().view(np.uint64)
    assert np.array_equal(np.sort(perm), np.arange(P))
    assert np.array_equal(got, rows_enum[perm]), "sorted table must carry the oracle's rows, bit for bit"
    a, b = got[:-1], got[1:]
    diff = a != b
    differs = diff.any(axis=1)
    first = diff.argmax(axis=1)
    i = np.nonzero(differs)[0]
    assert np.all(b[i, first[i]] > a[i, first[i]]), "rows must be lexicographically non-decreasing"
    assert np.all(perm[1:][~differs] > perm[:-1][~differs]), "equal rows must keep enumeration order"


@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("name", sorted(MG.CAPPED_CASES))
def test_max_thresholds_population_vs_reference_golden(name, fused):
    """max_thresholds (the UPGMA threshold cap of the [model] section, rules.py:533-576; the reference's
    own worked example uses 10, logit_rule_test.py:18): windows, values, capped thresholds, truth rows,
    lexsort order and rule descriptors against fixtures generated by the reference's rules.py."""
    cfg = MG.CAPPED_CASES[name]
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    d = synthetic.make_series(**cfg["series"])
    from mitre_b200 import rules as R
    V = d["X"][0].shape[0]
    data = R.Dataset(d["X"], d["T"], d["y"], ["v%d" % i for i in range(V)], d["variable_weights"],
                     d["experiment_start"], d["experiment_end"])
    old = R.DiscreteRulePopulation.use_fused
    R.DiscreteRulePopulation.use_fused = fused
    try:
        pop = R.DiscreteRulePopulation(None, cfg["pop"]["tmin"], cfg["pop"]["N_intervals"], tmax=cfg["pop"]["tmax"],
                                       data=data, max_thresholds=cfg["pop"]["max_thresholds"])
    finally:
        R.DiscreteRulePopulation.use_fused = old
    N = int(gold["n_subjects"])
    assert int(pop._K_host.max()) <= cfg["pop"]["max_thresholds"]
    check_population(pop, gold["windows"], gold["ok_slope"], gold["groups"], gold["values"], gold["order"],
                     gold["sorted_truth_packed"],
                     (gold["rule_variable"], gold["rule_window"], gold["rule_type"], gold["rule_direction"],
                      gold["rule_threshold"]), N)
    assert np.array_equal(pop.flat_truth, DO.unpack_rows(gold["sorted_truth_packed"], N))


def test_model_builds_with_max_thresholds():
    """LogisticRuleModel(..., max_thresholds=10) as in the reference's worked example (logit_rule_test.py:18)."""
    from mitre_b200 import logit_rules as LR, rules as R
    cfg = MG.CAPPED_CASES["det_maxthr10"]
    d = synthetic.make_series(**cfg["series"])
    V = d["X"][0].shape[0]
    data = R.Dataset(d["X"], d["T"], d["y"], ["v%d" % i for i in range(V)], d["variable_weights"],
                     d["experiment_start"], d["experiment_end"])
    model = LR.LogisticRuleModel(data, tmin=1.0, tmax=None, N_intervals=4, max_thresholds=10)
    pop = model.rule_population
    assert 0 < len(pop) <= pop._G * 2 * 10
    out = model.likelihoods_of_all_options(R.RuleList([[R.PrimitiveRule(*pop.flat_rules[0])]]),
                                           np.full(35, 0.25), 0, 0)
    assert out.shape == (len(pop),) and np.isfinite(out).all()


def test_pipelined_detector_call_equals_back_to_back():
    """mitre_detectors_walk (values and sort/emit kernels pipelined over runs of start groups on two internal
    streams) must produce exactly what the two calls made back to back produce."""
    import torch
    from mitre_b200 import rules as R
    d = synthetic.make_series(N=35, V=70, T=(8, 25), span=(0.0, 375.0), seed=21)
    cfg = dict(tmin=1.0, N_intervals=12, tmax=None)
    data, pop = build_gpu_population(d, cfg, True)
    assert pop._walk_args is not None
    ref = [t.clone() for t in (pop._values, pop._thresholds, pop._K, pop._slope_flags)]
    old = R.DiscreteRulePopulation.walk_stages
    R.DiscreteRulePopulation.walk_stages = 5
    try:
        data2, pop2 = build_gpu_population(d, cfg, True)
    finally:
        R.DiscreteRulePopulation.walk_stages = old
    assert pop2._walk_args[15] is not None and len(pop2._walk_args[15][0]) > 2
    for a, b in zip(ref, (pop2._values, pop2._thresholds, pop2._K, pop2._slope_flags)):
        assert torch.equal(a, b)
    assert torch.equal(pop.packed_truth, pop2.packed_truth) and torch.equal(pop._perm, pop2._perm)
