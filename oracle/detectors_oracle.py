"""oracle/detectors_oracle.py -- CPU (numpy) restatement of the reference detector path.

TEST INFRASTRUCTURE ONLY: nothing under mitre_b200/ may import this module.  It is the
checker for the CUDA detector kernels (tests/, __graft_entry__.smoke(), bench.py cpu_baseline).

Follows /root/reference/mitre/rules.py:
  enumerate_windows        <- DiscreteRulePopulation.mine_rules step 1   rules.py:391-421
  evaluate                 <- PrimitiveRule.evaluate                      rules.py:80-141
  calculate_values         <- DiscreteRulePopulation.calculate_values     rules.py:578-600
  thresholds_from_values   <- DiscreteRulePopulation.thresholds_from_values rules.py:513-576
  enumerate_primitives     <- DiscreteRulePopulation.update_base_rules    rules.py:465-511
  lexsort_rows             <- DiscreteRulePopulation.sort_base_rules      rules.py:439-462
  pairwise_sum             <- numpy's float add-reduction order, which np.mean (rules.py:141)
                              inherits (numpy/_core/src/umath/loops_utils.h.src, pairwise sum,
                              block size 128, unroll 8); restated so the CUDA kernel and this
                              oracle can be checked against a numpy-independent definition.

The slope goes through np.linalg.lstsq exactly as the reference does (rules.py:110-114); LAPACK
gelsd is third-party arithmetic that ships in the numpy wheel of this image.

Parity pin: tests/test_oracle.py checks every function here against the reference's own
rules.py, imported through oracle/ref_shim.py in the build container, and against the golden
vectors in tests/golden/ generated from it (tests/golden/make_golden.py).
"""
import numpy as np

MIN_OBSERVATIONS = 1          # rules.py:19
MIN_OBSERVATIONS_SLOPE = 2    # rules.py:21
TYPES = ("slope", "average")  # enumeration order, rules.py:479, 588
DIRECTIONS = ("above", "below")  # rules.py:491


def enumerate_windows(T, experiment_start, experiment_end, N_intervals, tmin, tmax=None):
    """rules.py:391-421.  T: list of per-subject time vectors.  Returns (windows, ok_for_slope)."""
    if tmax is None:
        tmax = np.inf
    timepoints = np.linspace(experiment_start, experiment_end, N_intervals + 1)
    windows, ok_slope = [], []
    for i, t0 in enumerate(timepoints):
        for t1 in timepoints[i + 1:]:
            if t1 - t0 < tmin or t1 - t0 > tmax:
                continue
            long_enough, slope_ok = True, True
            for ts in T:
                n_in = np.sum((t0 <= ts) & (ts <= t1))
                if n_in < MIN_OBSERVATIONS_SLOPE:
                    slope_ok = False
                if n_in < MIN_OBSERVATIONS:
                    long_enough = False
                    break
            if long_enough:
                windows.append((t0, t1))
                ok_slope.append(slope_ok)
    return windows, ok_slope


def pairwise_sum(a):
    """numpy's pairwise float summation of a 1-D contiguous array (see module docstring)."""
    n = len(a)
    if n < 8:
        res = np.float64(0.0)
        for v in a:
            res = res + v
        return res
    if n <= 128:
        r = [np.float64(a[i]) for i in range(8)]
        i = 8
        while i < n - (n % 8):
            for j in range(8):
                r[j] = r[j] + a[i + j]
            i += 8
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
        while i < n:
            res = res + a[i]
            i += 1
        return res
    n2 = n // 2
    n2 -= n2 % 8
    return pairwise_sum(a[:n2]) + pairwise_sum(a[n2:])


def mean_restated(x):
    """np.mean of a contiguous float64 vector = pairwise_sum / n (true divide)."""
    return pairwise_sum(np.asarray(x, dtype=np.float64)) / np.float64(len(x))


def evaluate(x, t, window, type_):
    """rules.py:80-141: window mean, or least-squares slope of x on (t - midpoint)."""
    t0, t1 = window
    midpoint = (t0 + t1) * 0.5                          # rules.py:43
    relevant = (t >= t0) & (t <= t1)                    # rules.py:100
    tt = t[relevant] - midpoint
    xx = x[relevant]
    n_points = len(tt)
    if n_points < MIN_OBSERVATIONS:
        raise ValueError("Too few time points within rule time window.")
    if n_points < MIN_OBSERVATIONS_SLOPE and type_ == "slope":
        raise ValueError("Too few time points to calculate slope within rule time window.")
    if type_ == "slope":
        A = np.ones((n_points, 2))
        A[:, 0] = tt
        (slope, _intercept), _, _, _ = np.linalg.lstsq(A, xx, rcond=None)   # rules.py:113
        return slope
    return np.mean(xx)                                  # rules.py:141


def calculate_values(X, T, windows, ok_slope):
    """rules.py:578-600.  Returns a list of groups in enumeration order
    (window-major, variable, type in ('slope','average')):
        [(window_index, variable, type_index, values float64[N_subjects]), ...]"""
    n_variables = X[0].shape[0]
    groups = []
    for wi, (window, do_slope) in enumerate(zip(windows, ok_slope)):
        for variable in range(n_variables):
            for ti, type_ in enumerate(TYPES):
                if type_ == "slope" and not do_slope:
                    continue
                vals = np.array([evaluate(x[variable], t, window, type_) for x, t in zip(X, T)])
                groups.append((wi, variable, ti, vals))
    return groups


def thresholds_from_values(values, max_thresholds=np.inf):
    """rules.py:513-576 with delta=None (its only call site passes no delta, rules.py:489)."""
    values = np.sort(values)
    do_clustering = True
    if len(values) <= max_thresholds + 1:
        thresholds = 0.5 * (values[:-1] + values[1:])   # rules.py:539
        do_clustering = False
    if do_clustering:                                   # rules.py:547-574 (UPGMA cap)
        from scipy.cluster import hierarchy
        distances = []
        for i, v1 in enumerate(values):
            for v2 in values[i + 1:]:
                distances.append(v2 - v1)
        linkage = hierarchy.linkage(np.array(distances), method="average")
        n_clusters = min(max_thresholds + 1, len(values))
        clusters = hierarchy.fcluster(linkage, n_clusters, criterion="maxclust")
        boundaries = clusters[:-1] != clusters[1:]
        thresholds = 0.5 * (values[:-1][boundaries] + values[1:][boundaries])
    return np.unique(thresholds)                        # rules.py:576


def enumerate_primitives(groups, n_subjects, max_thresholds=np.inf):
    """rules.py:465-511.  Enumeration order: group, threshold ascending, direction (above, below).

    Returns dict of parallel arrays of length P (the reference's base_flat_* lists before the
    sort): group, window, variable, type, direction, threshold, truth bool[P, N]; plus
    per-group K (number of thresholds) and the concatenated thresholds."""
    g_idx, w_idx, var, typ, dirn, thr, truth = [], [], [], [], [], [], []
    K, all_thr = [], []
    for gi, (wi, variable, ti, values) in enumerate(groups):
        ths = thresholds_from_values(values, max_thresholds)
        K.append(len(ths))
        all_thr.append(ths)
        for th in ths:
            for di in range(2):
                g_idx.append(gi); w_idx.append(wi); var.append(variable)
                typ.append(ti); dirn.append(di); thr.append(th)
                truth.append(values > th if di == 0 else values <= th)   # rules.py:494-497
    P = len(thr)
    return dict(
        group=np.array(g_idx, dtype=np.int64), window=np.array(w_idx, dtype=np.int32),
        variable=np.array(var, dtype=np.int32), type=np.array(typ, dtype=np.int8),
        direction=np.array(dirn, dtype=np.int8), threshold=np.array(thr, dtype=np.float64),
        truth=(np.vstack(truth) if P else np.zeros((0, n_subjects), dtype=bool)),
        K=np.array(K, dtype=np.int32),
        thresholds=(np.concatenate(all_thr) if all_thr else np.zeros(0)),
    )


def lexsort_rows(truth):
    """rules.py:450: np.lexsort(np.rot90(truth)) -- stable; subject 0 is the primary key,
    False < True; ties keep enumeration order."""
    if truth.shape[0] == 0:
        return np.zeros(0, dtype=np.int64)
    return np.lexsort(np.rot90(truth))


def build_population(X, T, experiment_start, experiment_end, N_intervals, tmin, tmax=None,
                     max_thresholds=None):
    """The whole of DiscreteRulePopulation.mine_rules (rules.py:325-436) as arrays."""
    windows, ok_slope = enumerate_windows(T, experiment_start, experiment_end, N_intervals,
                                          tmin, tmax)
    groups = calculate_values(X, T, windows, ok_slope)
    mt = np.inf if max_thresholds is None else max_thresholds
    prim = enumerate_primitives(groups, len(X), mt)
    order = lexsort_rows(prim["truth"])
    return dict(windows=np.array(windows, dtype=np.float64).reshape(-1, 2),
                ok_slope=np.array(ok_slope, dtype=bool),
                group_window=np.array([g[0] for g in groups], dtype=np.int32),
                group_variable=np.array([g[1] for g in groups], dtype=np.int32),
                group_type=np.array([g[2] for g in groups], dtype=np.int8),
                values=(np.vstack([g[3] for g in groups]) if groups
                        else np.zeros((0, len(X)))),
                order=order, **prim)


def pack_rows(truth):
    """bool[P, N] -> uint64[P, ceil(N/64)], subject i at bit (63 - i%64) of word i//64, so that
    unsigned word order == the reference's lexicographic row order (subject 0 most significant,
    rules.py:450).  Test helper for comparing with the device layout (include/mitre_b200.h)."""
    P, N = truth.shape
    W = (N + 63) // 64
    padded = np.zeros((P, W * 64), dtype=bool)
    padded[:, :N] = truth
    # big-endian bits within bytes, big-endian bytes within words: subject 0 is the word's top bit
    return np.packbits(padded, axis=1).view('>u8').astype(np.uint64).reshape(P, W)


def unpack_rows(packed, N):
    P, W = packed.shape
    out = np.zeros((P, N), dtype=bool)
    for i in range(N):
        out[:, i] = ((packed[:, i // 64] >> np.uint64(63 - (i % 64))) & np.uint64(1)).astype(bool)
    return out
