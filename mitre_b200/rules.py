"""Host-side mirror of the reference's rule algebra (mitre/rules.py) for the hot path.

Same names, argument meaning and error behaviour as the reference classes
(PrimitiveRule rules.py:23-141, RuleList 143-243, DiscreteRulePopulation 245-810,
Dataset 945-1217, choose_from_relative_probabilities 1261-1264), but
DiscreteRulePopulation.mine_rules -- window admission, calculate_values, update_base_rules,
sort_base_rules -- runs as CUDA kernels through the C ABI (mitre_b200/device.py) and the
population lives on the GPU as packed bitmask rows in the reference's lexsorted order.
Python-level views (`flat_rules`, `flat_truth`, `truth_table`, `primitive_index`,
`primitive_values`) are materialised lazily from the device arrays.
"""
import copy
import logging

import numpy as np

logging.basicConfig(format='%(levelname)s:%(message)s')
logger = logging.getLogger(__name__)

min_observations = 1        # rules.py:19
min_observations_slope = 2  # rules.py:21

TYPES = ('slope', 'average')      # enumeration order, rules.py:479
DIRECTIONS = ('above', 'below')   # rules.py:491


class PrimitiveRule:
    """rules.py:23-141."""

    def __init__(self, variable, window, type_, direction, threshold):
        self.variable = variable
        self.window = window
        self.type_ = type_
        self.direction = direction
        self.threshold = threshold
        (t0, t1) = window
        self.midpoint = (t0 + t1) * 0.5

    def as_tuple(self):
        return (self.variable, self.window, self.type_, self.direction, self.threshold)

    def __repr__(self):
        return 'PrimitiveRule(%s)' % str(self.as_tuple())

    def apply(self, X, t):
        value = self.evaluate(X[self.variable, :], t)
        if self.direction == 'above':
            return value > self.threshold
        return value <= self.threshold

    def evaluate(self, x, t):
        """Single-subject evaluation on the host (used when a learned rule is applied to
        held-out data, rules.py:1187-1191).  Population mining never comes through here."""
        (t0, t1) = self.window
        relevant_times = (t >= t0) & (t <= t1)
        t = t[relevant_times] - self.midpoint
        x = x[relevant_times]
        n_points = len(t)
        if n_points < min_observations:
            raise ValueError('Too few time points within rule time window.')
        if (n_points < min_observations_slope) and self.type_ == 'slope':
            raise ValueError('Too few time points to calculate slope within rule time window.')
        if self.type_ == 'slope':
            A = np.ones((n_points, 2))
            A[:, 0] = t
            (slope, _intercept), _, _, _ = np.linalg.lstsq(A, x, rcond=None)
            return slope
        if self.type_ == 'average':
            return np.mean(x)


class RuleList:
    """rules.py:143-243."""

    def __init__(self, rules=[], model=None):
        self.model = model
        self.rules = []
        for input_rule in rules:
            if len(input_rule) < 1:
                raise ValueError('Each rule must contain one or more primitive rules.')
            translated_rule = []
            for primitive in input_rule:
                if isinstance(primitive, PrimitiveRule):
                    translated_rule.append(primitive)
                else:
                    translated_rule.append(PrimitiveRule(*primitive))
            self.rules.append(translated_rule)

    def as_tuple(self):
        return tuple([tuple([p.as_tuple() for p in subrule]) for subrule in self.rules])

    def __repr__(self):
        return 'RuleList(%s)' % str(self.rules)

    def __str__(self):
        overall_lines = ['Rule list with %d rules' % len(self)]
        for i, rule in enumerate(self.rules):
            rule_lines = []
            for p in rule:
                line = ('(%.3f,%.3f) variable %d %s %s %.4f' %
                        (p.window[0], p.window[1], p.variable, p.type_, p.direction, p.threshold))
                rule_lines.append('\t' + line)
            overall_lines.append('Rule %d:' % i)
            overall_lines = overall_lines + rule_lines
        return '\n'.join(overall_lines)

    def copy(self):
        return RuleList([[PrimitiveRule(*p.as_tuple()) for p in r] for r in self.rules], self.model)

    def __len__(self):
        return len(self.rules)

    def __getitem__(self, i):
        return self.rules[i]

    def __setitem__(self, i, item):
        self.rules[i] = item

    def __delitem__(self, i):
        del self.rules[i]

    def apply_rules(self, X, t):
        rule_truths = []
        for rule in self.rules:
            this_rule_result = True
            for primitive in rule:
                if not primitive.apply(X, t):
                    this_rule_result = False
                    break
            rule_truths.append(this_rule_result)
        return np.array(rule_truths)


# ---------------------------------------------------------------------------------------------
# lazy Python views over the device-resident population
# ---------------------------------------------------------------------------------------------

class _FlatRules:
    """Sequence view: flat_rules[i] -> (variable, window, type_, direction, threshold) in the
    reference's lexsorted order (rules.py:455-456), built on demand from the group tables."""

    def __init__(self, pop, index_map=None):
        self._pop = pop
        self._map = index_map   # positions into the base (unfiltered) order, or None

    def __len__(self):
        return len(self._map) if self._map is not None else self._pop._P

    def _one(self, i):
        n = len(self)
        if i < 0:
            i += n
        if not 0 <= i < n:
            raise IndexError('primitive index out of range')
        base_i = int(self._map[i]) if self._map is not None else i
        return self._pop._tuple_at(base_i)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self._one(j) for j in range(*i.indices(len(self)))]
        return self._one(int(i))

    def __iter__(self):
        for i in range(len(self)):
            yield self._one(i)


class _PrimitiveIndex:
    """Mapping view: tuple -> position in flat_rules (rules.py:458-460), by arithmetic on the
    group tables instead of a P-entry dict."""

    def __init__(self, pop, inverse_filter=None):
        self._pop = pop
        self._inv = inverse_filter   # base position -> filtered position (-1 = filtered out)

    def get(self, key, default=None):
        i = self._pop._index_of(key)
        if i is None:
            return default
        if self._inv is not None:
            i = int(self._inv[i])
            if i < 0:
                return default
        return i

    def __getitem__(self, key):
        i = self.get(key)
        if i is None:
            raise KeyError(key)
        return i

    def __contains__(self, key):
        return self.get(key) is not None

    def __len__(self):
        return self._pop.n_primitives


class _TruthTable:
    """Mapping view: tuple -> bool[N] truth vector (rules.py:498, 507); what
    Dataset._primitive_result_cache points at (logit_rules.py:113)."""

    def __init__(self, pop):
        self._pop = pop

    def __contains__(self, key):
        return self._pop._index_of(key) is not None

    def __getitem__(self, key):
        i = self._pop._index_of(key)
        if i is None:
            raise KeyError(key)
        return self._pop._truth_row(i)

    def get(self, key, default=None):
        i = self._pop._index_of(key)
        return default if i is None else self._pop._truth_row(i)

    def __len__(self):
        return self._pop._P


class _PrimitiveValues:
    """Mapping view: (window, variable, type_) -> float64[N] (rules.py:597-599)."""

    def __init__(self, pop):
        self._pop = pop

    def _group(self, key):
        window, variable, type_ = key
        w = self._pop._window_lookup.get((float(window[0]), float(window[1])))
        if w is None or type_ not in TYPES:
            return None
        ok = bool(self._pop.windows_ok_for_slope[w])
        if type_ == 'slope' and not ok:
            return None
        return int(self._pop._group_base[w] + variable * (2 if ok else 1) +
                   (TYPES.index(type_) if ok else 0))

    def __contains__(self, key):
        return self._group(key) is not None

    def __getitem__(self, key):
        g = self._group(key)
        if g is None:
            raise KeyError(key)
        return self._pop._values[g].cpu().numpy()

    def __len__(self):
        return self._pop._G


class DiscreteRulePopulation:
    """rules.py:245-810, mined on the GPU."""

    use_fused = True    # N <= 63: values + sort/emit kernels without a host round trip; False forces the phase-by-phase path
    resort_after_update = False   # _update_truths: True re-lexsorts the re-evaluated table (the reference does not)
    use_walk = True     # ... with the walk-forward values kernel (mitre_detector_values_walk); False: mitre_detectors_fused
    walk_stages = 1     # > 1: values / sort-emit kernels pipelined over that many runs of start groups on two streams
                        # (mitre_detectors_walk; measured SLOWER on B200 -- the kernels do not co-reside -- so off)

    def __init__(self, model, tmin, N_intervals, tmax=None, data=None, max_thresholds=None):
        self.model = model
        self.data = model.data if data is None else data
        self.tmin = tmin
        self.tmax = np.inf if tmax is None else tmax
        self.N_intervals = N_intervals
        # rules.py:287-290: None = the full grid (every midpoint); a number caps the thresholds per
        # (window, variable, type) group at that many by UPGMA clustering of the sorted values
        self.max_thresholds = np.inf if max_thresholds is None else max_thresholds
        self.mine_rules()
        self.applied_filters = []

    # -- pickling / copying: device tensors are a pure function of (data, windows) -------------
    def __getstate__(self):
        state = dict(self.__dict__)
        for k in list(state):
            if k.startswith('_d_') or k in ('_values', '_thresholds', '_K', '_row_offsets', '_rows',
                                            '_perm', '_sorted_rows', '_row_group', '_rows_padded', '_fused', '_fused_args', '_active_rows',
                                            '_slope_flags', '_slope_flagged', '_walk_args',
                                            '_d_windows'):
                state[k] = None
        state['_host_cache'] = {}
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        if getattr(self, 'windows', None) is not None:
            self._rebuild_device_state()

    def _rebuild_device_state(self):
        self.calculate_values()
        self.update_base_rules()
        self.sort_base_rules()

    # -- step 1: windows (rules.py:391-421) ---------------------------------------------------------
    def mine_rules(self):
        self._enumerate_windows()
        self.calculate_values()
        self.update_base_rules()
        self.sort_base_rules()

    def _enumerate_windows(self):
        """rules.py:391-421: admissible (t0, t1) windows and their slope flags."""
        from . import device as D
        import torch
        data = self.data
        timepoints = np.linspace(data.experiment_start, data.experiment_end, self.N_intervals + 1)
        self.timepoints = timepoints
        cand = []
        for i, t0 in enumerate(timepoints):
            for t1 in timepoints[i + 1:]:
                if t1 - t0 < self.tmin:
                    continue
                if t1 - t0 > self.tmax:
                    continue
                cand.append((t0, t1))
        ds = data.device_series()
        windows, ok_slope = [], []
        if cand:
            d_cand = torch.as_tensor(np.array(cand, dtype=np.float64).reshape(-1, 2), device=D.dev())
            _lo, cnt = D.window_ranges(ds['times'], ds['t_offsets'], d_cand)
            min_cnt = cnt.min(dim=1).values.cpu().numpy()
            for w, c in zip(cand, min_cnt):
                if c >= min_observations:
                    windows.append(w)
                    ok_slope.append(bool(c >= min_observations_slope))
        self.windows = windows
        self.windows_ok_for_slope = ok_slope

    # -- step 2 (rules.py:578-600) ----------------------------------------------------------------------
    def calculate_values(self):
        from . import device as D
        import torch
        ds = self.data.device_series()
        V = self.data.n_variables
        W = len(self.windows)
        ok = np.array(self.windows_ok_for_slope, dtype=np.uint8).reshape(W)
        per_window = V * (1 + ok.astype(np.int64))
        group_base = np.zeros(W + 1, dtype=np.int64)
        np.cumsum(per_window, out=group_base[1:])
        self._group_base = group_base
        self._G = int(group_base[-1])
        self._window_lookup = {(float(a), float(b)): i for i, (a, b) in enumerate(self.windows)}
        self._durations_by_window = np.array([w[1] - w[0] for w in self.windows], dtype=np.float64)
        # group -> (window, variable, type): per window either (v, slope), (v, average) pairs or
        # (v, average) alone -- two patterns, concatenated (no G-long integer divisions on the host)
        key = (W, V, ok.tobytes())
        cached = self.__dict__.get('_group_tables_cache')
        if cached is None or cached[0] != key:
            var2 = np.repeat(np.arange(V, dtype=np.int32), 2)
            var1 = np.arange(V, dtype=np.int32)
            typ2 = np.tile(np.array([0, 1], dtype=np.int8), V)
            typ1 = np.ones(V, dtype=np.int8)
            empty_i, empty_b = np.zeros(0, dtype=np.int32), np.zeros(0, dtype=np.int8)
            cached = (key,
                      np.repeat(np.arange(W, dtype=np.int32), per_window),
                      np.concatenate([var2 if o else var1 for o in ok] + [empty_i]),
                      np.concatenate([typ2 if o else typ1 for o in ok] + [empty_b]))
            self._group_tables_cache = cached
        _, self._group_window, self._group_variable, self._group_type = cached
        self._host_cache = {}
        self._slope_flags = self._slope_flagged = None
        if W == 0:
            self._values = torch.empty((0, self.data.n_subjects), dtype=torch.float64, device=D.dev())
            self._d_windows = None
            return
        self._d_windows = torch.as_tensor(np.array(self.windows, dtype=np.float64).reshape(-1, 2),
                                          device=D.dev())
        d_gb = torch.as_tensor(group_base[:-1].copy(), device=D.dev())
        d_ok = torch.as_tensor(ok, device=D.dev())
        N = self.data.n_subjects
        self._fused = None
        self._fused_args = None
        self._walk_args = None
        capped = N > self.max_thresholds + 1        # rules.py:535: groups with more values than the cap cluster
        if (self.use_fused and 2 <= N <= 63 and not capped and
                ds['series_t'].shape[0] * ds['series_t'].shape[1] < 2 ** 31 - 1):
            # values + thresholds + K + padded truth rows without a host round trip (fused_detectors.cuh)
            lists = D.window_lists(ds['times'], ds['t_offsets'], self._d_windows,
                                   ds['series_t'].shape[1])
            tw, tt = [], []
            for w in range(W):
                if ok[w]:
                    tw.append(w); tt.append(0)
                tw.append(w); tt.append(1)
            d_tw = torch.as_tensor(np.array(tw, dtype=np.int32), device=D.dev())
            d_tt = torch.as_tensor(np.array(tt, dtype=np.int8), device=D.dev())
            d_gt = torch.as_tensor(np.ascontiguousarray(self._group_type).view(np.uint8), device=D.dev())
            self._fused_args = (ds['series_t'], N, V, lists, d_gb, d_ok, d_tw, d_tt, self._G, d_gt)
            self._walk_args = None
            # start groups: runs of consecutive windows with the same start (rules.py:397-402 is start-major)
            t0s = np.array([w[0] for w in self.windows], dtype=np.float64)
            first = np.concatenate([[0], np.nonzero(t0s[1:] != t0s[:-1])[0] + 1, [W]]).astype(np.int32)
            plan = D.values_walk_plan(N, ds['S'], V, first) if self.use_walk else None
            if plan is not None:
                tbar = D.window_tbar(ds['times'], self._d_windows, lists['lo'], lists['cnt'])
                stages = (D.walk_stages(first, group_base, self._G, self.walk_stages)
                          if self.walk_stages > 1 else None)
                self._walk_args = (ds['series'], ds['times'], ds['t_offsets'], ds['S'], N, V, self._d_windows, lists,
                                   tbar, d_gb, d_ok, torch.as_tensor(first, device=D.dev()), self._G, d_gt, plan,
                                   stages)
                self._walk_first = first
                out = D.detectors_walk(*self._walk_args)
            else:
                out = D.detectors_fused(*self._fused_args)
            self._values, thr, K, rows_padded, self._slope_flags, self._slope_flagged = out
            self._fused = (thr, K, rows_padded)
        else:
            lo, cnt = D.window_ranges(ds['times'], ds['t_offsets'], self._d_windows)
            self._values = D.detector_values(ds['series'], ds['times'], ds['t_offsets'],
                                             self._d_windows, lo, cnt, d_gb, d_ok, self._G, ds['max_T'])
            d_gt = torch.as_tensor(np.ascontiguousarray(self._group_type).view(np.uint8), device=D.dev())
            self._slope_flags, self._slope_flagged = D.slope_guard(self._values, d_gt)
        self.primitive_values = _PrimitiveValues(self)

    @property
    def slope_groups_flagged(self):
        """Number of slope groups whose truth rows are not guaranteed to be the reference's: two sorted
        neighbours closer than 1e-10 of the group's largest |slope|, or a constant non-zero in-window series
        (rules.py:110-114 reaches the slope through LAPACK gelsd; see include/mitre_b200.h, mitre_slope_guard)."""
        f = getattr(self, '_slope_flagged', None)
        return 0 if f is None else int(f.item())

    def flagged_groups(self):
        """[(window index, variable)] of the flagged slope groups."""
        f = getattr(self, '_slope_flags', None)
        if f is None:
            return []
        idx = np.nonzero(f.cpu().numpy())[0]
        return [(int(self._group_window[g]), int(self._group_variable[g])) for g in idx]

    # -- step 3 (rules.py:465-511) ------------------------------------------------------------------------
    def update_base_rules(self):
        from . import device as D
        import torch
        N = self.data.n_subjects
        self._build_generation = getattr(self, '_build_generation', 0) + 1    # cache key of K / row offsets
        if self._G == 0:
            self._P = 0
            self._K_host = np.zeros(0, dtype=np.int32)
            self._row_offsets_host = np.zeros(1, dtype=np.int64)
            self._rows = torch.empty((0, D.words_for(N)), dtype=torch.int64, device=D.dev())
            self._rows_padded = None
            self._thresholds = torch.empty((0, max(N - 1, 0)), dtype=torch.float64, device=D.dev())
        elif getattr(self, '_fused', None) is not None:
            self._thresholds, self._K, self._rows_padded = self._fused
            self._fused = None
            self._row_offsets = D.detector_row_offsets(self._K)
            self._row_offsets_host = self._row_offsets.cpu().numpy()
            self._K_host = self._K.cpu().numpy()
            self._P = int(self._row_offsets_host[-1])
            self._rows = None
        else:
            self._rows_padded = None
            if N > self.max_thresholds + 1:
                self._thresholds, self._K = self._capped_thresholds()
            else:
                self._thresholds, self._K = D.detector_thresholds(self._values)
            self._row_offsets = D.detector_row_offsets(self._K)
            self._row_offsets_host = self._row_offsets.cpu().numpy()
            self._K_host = self._K.cpu().numpy()
            self._P = int(self._row_offsets_host[-1])
            if self._P > 0:
                self._rows = D.detector_rows(self._values, self._thresholds, self._K,
                                             self._row_offsets, self._P)
            else:
                self._rows = torch.empty((0, D.words_for(N)), dtype=torch.int64, device=D.dev())
        self.truth_table = _TruthTable(self)
        self._host_cache = {}

    def _capped_thresholds(self):
        """thresholds_from_values with max_thresholds set (rules.py:533-576): UPGMA (average-linkage)
        clustering of a group's sorted values, cut at max_thresholds + 1 clusters, thresholds midway
        between neighbouring clusters.  This is the reference's own host computation -- the same
        scipy.cluster.hierarchy calls on the same pairwise-distance vector -- applied to the values the
        device computed: the clustering walks a dendrogram whose merge order hangs on floating-point
        ties that only scipy's own Lance-Williams updates reproduce, and with the cap in force the table
        is small (at most 2 * max_thresholds rows per group), so the host pass is not the cost.  Values,
        truth rows, the lexsort and everything downstream stay on the device."""
        import torch
        from . import device as D
        thresholds, K = capped_thresholds_host(self._values.cpu().numpy(), self.max_thresholds)
        return torch.as_tensor(thresholds, device=D.dev()), torch.as_tensor(K, device=D.dev())

    # -- step 4 (rules.py:439-462) ------------------------------------------------------------------------
    def sort_base_rules(self):
        from . import device as D
        import torch
        N = self.data.n_subjects
        padded = getattr(self, '_rows_padded', None)
        if self._P == 0:
            self._perm = torch.empty(0, dtype=torch.int64, device=D.dev())
            self._sorted_rows = torch.empty((0, D.words_for(N)), dtype=torch.int64, device=D.dev())
        elif padded is not None:
            # padded table: sentinel slots sort last (N + 1 significant bits), valid rows keep the
            # reference's enumeration order among equals; slot indices -> dense enumeration indices
            # (one-sweep sort of (row | slot) words when they fit 64 bits: sentinels dropped in its first pass,
            # densified in its last)
            self._perm, self._sorted_rows = D.lexsort_padded(padded, N, self._P, self._row_offsets)
        else:
            self._perm, self._sorted_rows = D.lexsort_rows(self._rows, N)
        self._rows = None   # enumeration-order rows are no longer needed
        self._rows_padded = None
        self._host_cache = {}
        self.reset_filter()

    def _update_truths(self):
        """rules.py:300-323 (cross-validation): keep the windows, regenerate everything else.

        The reference does NOT sort again here: after update_base_rules its flat arrays are in enumeration
        order (rules.py:321-325 go straight to reset_filter), so a fold's primitive k is the k-th enumerated
        one.  That is the default (indices, draws and chains on a fold then match the reference's; the scoring
        kernels take rows in any order, a table without runs of equal rows just sweeps ~2.5x slower).
        `resort_after_update = True` re-lexsorts instead."""
        self.calculate_values()
        self.update_base_rules()
        old_applied_filters = self.applied_filters[:]
        if self.resort_after_update:
            self.sort_base_rules()
        else:
            self._keep_enumeration_order()
        for filter_method, args in old_applied_filters:
            getattr(self, filter_method)(*args)
        self.data._primitive_result_cache = self.truth_table

    def _keep_enumeration_order(self):
        """The population as update_base_rules leaves it (rules.py:465-511): identity permutation."""
        from . import device as D
        import torch
        N = self.data.n_subjects
        padded = getattr(self, '_rows_padded', None)
        if self._P == 0:
            rows = torch.empty((0, D.words_for(N)), dtype=torch.int64, device=D.dev())
        elif padded is not None:
            L = padded.shape[1]
            live = torch.arange(L, device=padded.device)[None, :] < (2 * self._K.to(torch.int64))[:, None]
            rows = padded[live].view(-1, 1)
        else:
            rows = self._rows
        self._perm = torch.arange(self._P, dtype=torch.int64, device=D.dev())
        self._sorted_rows = rows.contiguous()
        self._rows = None
        self._rows_padded = None
        self._host_cache = {}
        self.reset_filter()

    # -- host-side descriptor arithmetic ------------------------------------------------------------------
    @property
    def _reordering_cache(self):
        return self._perm_host()

    def _perm_host(self):
        c = self._host_cache
        if 'perm' not in c:
            c['perm'] = self._perm.cpu().numpy()
        return c['perm']

    def _inv_perm_host(self):
        c = self._host_cache
        if 'inv_perm' not in c:
            perm = self._perm_host()
            inv = np.empty_like(perm)
            inv[perm] = np.arange(len(perm), dtype=perm.dtype)
            c['inv_perm'] = inv
        return c['inv_perm']

    def _thresholds_host(self):
        c = self._host_cache
        if 'thr' not in c:
            c['thr'] = self._thresholds.cpu().numpy()
        return c['thr']

    def _row_group_host(self):
        """group id of every sorted row, int32[P]."""
        c = self._host_cache
        if 'row_group' not in c:
            perm = self._perm_host()
            c['row_group'] = (np.searchsorted(self._row_offsets_host, perm, side='right') - 1).astype(np.int32)
        return c['row_group']

    def _tuple_at(self, i):
        r = int(self._perm_host()[i])
        g = int(np.searchsorted(self._row_offsets_host, r, side='right') - 1)
        local = r - int(self._row_offsets_host[g])
        k, d = local >> 1, local & 1
        w = int(self._group_window[g])
        thr = self._thresholds_host()[g, k]
        return (int(self._group_variable[g]), self.windows[w], TYPES[int(self._group_type[g])],
                DIRECTIONS[d], thr)

    def _index_of(self, key):
        """Position in the base sorted order of a primitive tuple, or None."""
        try:
            variable, window, type_, direction, threshold = key
            w = self._window_lookup.get((float(window[0]), float(window[1])))
        except (TypeError, ValueError, IndexError):
            return None
        if w is None or type_ not in TYPES or direction not in DIRECTIONS:
            return None
        if not (0 <= variable < self.data.n_variables):
            return None
        ok = bool(self.windows_ok_for_slope[w])
        if type_ == 'slope' and not ok:
            return None
        g = int(self._group_base[w] + variable * (2 if ok else 1) + (TYPES.index(type_) if ok else 0))
        K = int(self._K_host[g])
        thr = self._thresholds_host()[g, :K]
        k = int(np.searchsorted(thr, threshold))
        if k >= K or thr[k] != threshold:
            return None
        r = int(self._row_offsets_host[g]) + 2 * k + DIRECTIONS.index(direction)
        return int(self._inv_perm_host()[r])

    def _truth_row(self, i):
        """bool[N] truth vector of base sorted row i (one small D2H copy, cached)."""
        c = self._host_cache.setdefault('rows', {})
        if i not in c:
            if len(c) > 4096:
                c.clear()
            N = self.data.n_subjects
            words = self._sorted_rows[i].cpu().numpy().view(np.uint64)
            c[i] = _unpack_words(words, N)
        return c[i]

    # -- filters (rules.py:602-629) -----------------------------------------------------------------------
    def reset_filter(self):
        self._filter_map = None          # positions (base order) of the active primitives
        self._filter_generation = getattr(self, '_filter_generation', 0) + 1   # cache key (id() can be reused)
        self._active_rows = self._sorted_rows
        self.flat_rules = _FlatRules(self)
        self.n_primitives = self._P
        self.primitive_index = _PrimitiveIndex(self)
        self.base_primitive_index = self.primitive_index
        self.base_flat_rules = self.flat_rules
        self.applied_filters = []
        self._host_cache.pop('flat_durations', None)
        self._host_cache.pop('flat_variable_weights', None)

    def _filter(self, boolean_index):
        import torch
        keep = np.asarray(boolean_index, dtype=bool)
        current = self._filter_map if self._filter_map is not None else np.arange(self._P)
        if keep.shape[0] != current.shape[0]:
            raise ValueError('filter index must match the active population')
        self._filter_map = current[keep]
        self._filter_generation = getattr(self, '_filter_generation', 0) + 1
        inv = -np.ones(self._P, dtype=np.int64)
        inv[self._filter_map] = np.arange(len(self._filter_map))
        idx = torch.as_tensor(self._filter_map, device=self._sorted_rows.device)
        self._active_rows = self._sorted_rows.index_select(0, idx)
        self.flat_rules = _FlatRules(self, self._filter_map)
        self.n_primitives = len(self._filter_map)
        self.primitive_index = _PrimitiveIndex(self, inv)
        self._host_cache.pop('flat_durations', None)
        self._host_cache.pop('flat_variable_weights', None)

    def filter_thresholds(self, *args, **kwargs):
        raise NotImplementedError('filter_thresholds (rules.py:631-791) is outside the device hot path')

    # -- device-side descriptors for the on-device prior / draw ------------------------------------------
    def row_group_device(self):
        """int32[P] CUDA tensor: group of every active lexsorted row."""
        from . import device as D
        import torch
        c = self._host_cache
        key = ('row_group_dev', self._filter_generation)
        if key not in c:
            if self._P == 0:
                base = torch.empty(0, dtype=torch.int32, device=D.dev())
            else:
                base = D.row_groups(self._perm, self._row_offsets)
            if self._filter_map is not None:
                base = base.index_select(0, torch.as_tensor(self._filter_map, device=base.device))
            c[key] = base
        return c[key]

    def group_tables_device(self):
        """(group_variable, group_window) int32[G] CUDA tensors."""
        from . import device as D
        import torch
        c = self._host_cache
        if 'group_tables_dev' not in c:
            c['group_tables_dev'] = (
                torch.as_tensor(np.ascontiguousarray(self._group_variable, dtype=np.int32), device=D.dev()),
                torch.as_tensor(np.ascontiguousarray(self._group_window, dtype=np.int32), device=D.dev()))
        return c['group_tables_dev']

    def group_counts(self):
        """int64[G]: number of active primitives in every group (2 K_g unless filtered)."""
        c = self._host_cache
        key = ('group_counts', self._filter_generation)
        if key not in c:
            if self._filter_map is None:
                c[key] = 2 * self._K_host.astype(np.int64)
            else:
                c[key] = np.bincount(self._flat_groups(), minlength=self._G).astype(np.int64)
        return c[key]

    # -- flat arrays the priors consume (rules.py:453-454, 606-608) ---------------------------------
    def _flat_groups(self):
        rg = self._row_group_host()
        return rg if self._filter_map is None else rg[self._filter_map]

    @property
    def flat_durations(self):
        c = self._host_cache
        if 'flat_durations' not in c:
            c['flat_durations'] = self._durations_by_window[self._group_window[self._flat_groups()]]
        return c['flat_durations']

    @property
    def flat_variable_weights(self):
        c = self._host_cache
        if 'flat_variable_weights' not in c:
            vw = np.asarray(self.data.variable_weights, dtype=np.float64)
            c['flat_variable_weights'] = vw[self._group_variable[self._flat_groups()]]
        return c['flat_variable_weights']

    base_flat_durations = flat_durations
    base_flat_variable_weights = flat_variable_weights

    @property
    def flat_truth(self):
        """bool[P, N] (rules.py:457, 608) -- unpacked on the device, copied to the host."""
        from . import device as D
        N = self.data.n_subjects
        if self.n_primitives == 0:
            return np.zeros((0, N), dtype=bool)
        return D.unpack_rows(self._active_rows, N).cpu().numpy().astype(bool)

    base_flat_truth = flat_truth

    @property
    def packed_truth(self):
        """Device tensor int64[P, W64]: the active population's packed rows, lexsorted."""
        return self._active_rows

    def get_primitive_index(self, primitive):
        return self.primitive_index.get(primitive.as_tuple())

    def sort_list_of_primitives(self, l):
        l.sort(key=self.get_primitive_index)

    def __len__(self):
        return self.n_primitives


def capped_thresholds_host(values, max_thresholds):
    """rules.py:533-576 for every row of values float64[G, N] (N > max_thresholds + 1): returns
    (thresholds float64[G, N-1] zero-padded, K int32[G])."""
    from scipy.cluster import hierarchy
    G, N = values.shape
    n_clusters = int(min(max_thresholds + 1, N))
    iu, ju = np.triu_indices(N, k=1)               # pair order of rules.py:548-551
    thresholds = np.zeros((G, max(N - 1, 0)), dtype=np.float64)
    K = np.zeros(G, dtype=np.int32)
    for g in range(G):
        v = np.sort(values[g])
        linkage = hierarchy.linkage(v[ju] - v[iu], method='average')
        clusters = hierarchy.fcluster(linkage, n_clusters, criterion='maxclust')
        boundaries = clusters[:-1] != clusters[1:]
        t = np.unique(0.5 * (v[:-1][boundaries] + v[1:][boundaries]))
        K[g] = len(t)
        thresholds[g, :len(t)] = t
    return thresholds, K


def _unpack_words(words, N):
    out = np.zeros(N, dtype=bool)
    for i in range(N):
        out[i] = bool((int(words[i >> 6]) >> (63 - (i & 63))) & 1)
    return out


# ---------------------------------------------------------------------------------------------

class RuleModel(object):
    """rules.py:814-922 (interface only)."""

    def __init__(self, data, lambda_rules, lambda_primitives, tmin, **kwargs):
        self.data = data
        self.lambda_rules = lambda_rules
        self.lambda_primitives = lambda_primitives
        self.tmin = tmin

    def prior(self, rule_list):
        raise NotImplementedError


class DiscreteRuleModel(RuleModel):
    def __init__(self, *args, **kwargs):
        raise NotImplementedError()


class Dataset:
    """rules.py:945-1217."""

    def __init__(self, X, T, y, variable_names, variable_weights, experiment_start, experiment_end,
                 subject_IDs=None, subject_data=None,
                 additional_subject_categorical_covariates=[],
                 additional_covariate_default_states=[],
                 additional_subject_continuous_covariates=[],
                 variable_annotations={}):
        self.X = X
        self.T = T
        self.y = np.array(y, dtype='bool')
        self.variable_names = variable_names
        self.variable_weights = variable_weights
        self.experiment_start = experiment_start
        self.experiment_end = experiment_end
        self.subject_data = subject_data
        self.subject_IDs = subject_IDs
        self.n_subjects = len(X)
        self.n_variables = len(variable_weights)
        self._primitive_result_cache = {}
        self._device_series = None
        for array in X:
            if len(array) == 0:
                continue
            this_subject_n_variables, _ = array.shape
            if this_subject_n_variables != self.n_variables:
                raise ValueError('Observation-prior dimension mismatch.')
        if len(self.variable_names) != self.n_variables:
            raise ValueError('Incorrect number of variable names.')
        self.additional_subject_categorical_covariates = additional_subject_categorical_covariates
        self.additional_covariate_default_states = additional_covariate_default_states
        self.additional_subject_continuous_covariates = additional_subject_continuous_covariates
        self.generate_additional_covariate_matrix()
        self.variable_annotations = variable_annotations

    def __getstate__(self):
        state = dict(self.__dict__)
        state['_device_series'] = None
        if not isinstance(state.get('_primitive_result_cache'), dict):
            state['_primitive_result_cache'] = {}
        return state

    def generate_additional_covariate_matrix(self):
        """rules.py:1075-1131."""
        columns = []
        explanations = []
        features_and_defaults = zip(self.additional_subject_categorical_covariates,
                                    self.additional_covariate_default_states)
        for feature, default_value in features_and_defaults:
            try:
                values = set(self.subject_data[feature])
            except KeyError:
                raise KeyError('Trying to control for covariate %s, but no "%s" '
                               'column in subject_data.' % (feature, feature))
            try:
                other_values = values.copy()
                other_values.remove(default_value)
            except KeyError:
                raise ValueError('Trying to control for covariate %s, but no '
                                 'subject has the default value "%s". To avoid '
                                 'identifiability problems, at least one subject '
                                 'must have the default value.' % (feature, default_value))
            for alternative in other_values:
                explanations.append((feature, alternative))
                columns.append((self.subject_data[feature] == alternative).values.astype('int64'))
        for feature in self.additional_subject_continuous_covariates:
            try:
                values = self.subject_data[feature].values
            except KeyError:
                raise KeyError('Trying to control for covariate %s, but no "%s" '
                               'column in subject_data.' % (feature, feature))
            columns.append(values - np.mean(values))
            explanations.append('continuous feature %s' % feature)
        self.additional_covariate_encoding = tuple(explanations)
        if columns:
            self.additional_covariate_matrix = np.vstack(columns).T
        else:
            self.additional_covariate_matrix = None
        self.n_fixed_covariates = 1 + len(columns)

    def copy(self):
        return copy.deepcopy(self)

    def device_series(self):
        """Stage the ragged series in HBM once: times f64[S_pad], t_offsets i32[N+1],
        series f64[V, S_pad] (variable-major; S_pad even so that TMA bulk copies stay aligned)."""
        if self._device_series is None:
            import torch
            from . import device as D
            lens = np.array([len(t) for t in self.T], dtype=np.int64)
            offsets = np.zeros(self.n_subjects + 1, dtype=np.int32)
            np.cumsum(lens, out=offsets[1:])
            S = int(offsets[-1])
            S_pad = (S + 2) & ~1
            times = np.zeros(S_pad, dtype=np.float64)
            series = np.zeros((self.n_variables, S_pad), dtype=np.float64)
            for s, (x, t) in enumerate(zip(self.X, self.T)):
                t = np.asarray(t, dtype=np.float64)
                if len(t) > 1 and np.any(np.diff(t) < 0):
                    raise ValueError('subject %d: observation times must be non-decreasing' % s)
                times[offsets[s]:offsets[s + 1]] = t
                if len(t):
                    series[:, offsets[s]:offsets[s + 1]] = x
            # sample-major transposed copy for the fused kernel: [S_pad, V rounded up to 32]
            Vp = (self.n_variables + 31) & ~31
            series_t = np.zeros((S_pad, Vp), dtype=np.float64)
            series_t[:, :self.n_variables] = series.T
            self._device_series = dict(
                times=torch.as_tensor(times, device=D.dev()),
                t_offsets=torch.as_tensor(offsets, device=D.dev()),
                series=torch.as_tensor(series, device=D.dev()),
                series_t=torch.as_tensor(series_t, device=D.dev()),
                max_T=int(lens.max()) if len(lens) else 1, S=S)
        return self._device_series

    def apply_rules(self, rule_list):
        """rules.py:1136-1150."""
        rule_results = []
        for rule in rule_list.rules:
            acc = None
            for p in rule:
                t = self.apply_primitive(p)
                acc = t if acc is None else np.logical_and(acc, t)
            rule_results.append(acc)
        if rule_results:
            return np.vstack(rule_results)
        return []

    def covariate_matrix(self, rule_list):
        """rules.py:1152-1179."""
        if len(rule_list) < 1:
            X = np.ones((self.n_subjects, 1), dtype=np.int64)
        else:
            X = np.hstack((1 * self.apply_rules(rule_list).T,
                           np.ones((self.n_subjects, 1), dtype=np.int64),))
        if self.additional_covariate_matrix is not None:
            X = np.hstack((X, self.additional_covariate_matrix))
        return X

    def apply_primitive(self, primitive):
        key = primitive.as_tuple()
        if key in self._primitive_result_cache:
            return self._primitive_result_cache[key]
        return self._apply_primitive(primitive)

    def _apply_primitive(self, primitive):
        values = []
        for subject_x, subject_t in zip(self.X, self.T):
            values.append(primitive.apply(subject_x, subject_t))
        return np.array(values)

    def stratify(self, rule_list):
        subjects_handled = np.zeros(self.n_subjects, dtype='bool')
        class_memberships = np.zeros((len(rule_list.rules) + 1, self.n_subjects), dtype='bool')
        i = -1
        for i, rule_results in enumerate(self.apply_rules(rule_list)):
            this_rule_subjects = rule_results & (~subjects_handled)
            class_memberships[i, :] = this_rule_subjects
            subjects_handled = subjects_handled | this_rule_subjects
        class_memberships[i + 1, :] = ~subjects_handled
        return class_memberships

    def y_by_class(self, rule_list):
        class_memberships = self.stratify(rule_list)
        return list(zip(np.sum(class_memberships * self.y, 1), np.sum(class_memberships, 1)))

    def __str__(self):
        template = ('<Dataset with %d subjects (%d observations of %d variables)>')
        n_obs = sum([len(timepoints) for timepoints in self.T])
        return template % (self.n_subjects, n_obs, self.n_variables)


class RuleListSampler:
    def sample(self, N):
        for i in range(N):
            logger.info('Sampling iteration %d' % i)
            self.step()


def choose_from_relative_probabilities(p):
    """rules.py:1261-1264 (host form, for the handful-of-options moves)."""
    p = np.cumsum(p)
    marker = np.random.rand() * p[-1]
    return np.sum(p < marker)
