"""Candidate sweeps over populations whose packed truth table does not fit in HBM
(configuration S3: 2 000 subjects x 5 000 variables x 50 timepoints, mean + slope,
full threshold grid -> 3.1e9 candidates, 0.8 TB packed; the reference would need 6 TB of bool
and a 50 TB float64 temporary, logit_rules.py:214-217).

The table is never materialised.  Variables are walked in tiles; per tile the detector kernels
rebuild values and thresholds (every group depends only on its own variable's series,
rules.py:585-599) and the tile's candidates are scored straight from them: within a group the 2K
candidates are nested sets along the sorted order of its values, so mitre_score_groups gets every
masked sum as a suffix sum -- O(N p) per group, no truth rows at all (group_scoring=False builds
the rows and uses the row kernels instead; the tests compare the two).  (max, sum exp) of
log-likelihood + log-prior is reduced per tile.  What survives a sweep is one 16-byte pair per tile: enough for the log-sum-exp the add/remove moves
need (logit_rules.py:959, 1064) and for a two-stage draw -- pick the tile that holds
u * total, rebuild that one tile, draw inside it -- exactly the exchange `parallel.two_stage_draw`
makes between ranks, here made between tiles in time.  With a process group the tiles are dealt
round-robin to the ranks (each GPU stages the series and walks its own tiles), the pairs are
all-gathered (16 bytes per tile) and the tile's owner makes the inner draw.

Candidates are indexed (tile, position in the tile's enumeration order).  That is a different
ORDER from the reference's global lexsort (which nothing that large can compute), over the same
candidate SET with the same weights, so the draw has the same distribution; chains are not
RNG-stream-aligned with a hypothetical resident run.  For populations that do fit,
tests/test_streaming_gpu.py checks tile by tile that every candidate carries the likelihood the
resident population gives it.
"""
import numpy as np

from .parallel import owner_of_marker, world
from .rules import DIRECTIONS, TYPES


def combine_tile_pairs(gathered):
    """gathered [world, n_tiles, 2]: every rank's (max, sum-exp) table in which only its own tiles
    (t = rank mod world) are filled -> [n_tiles, 2] with every tile's pair taken from its owner."""
    import torch
    ws, n_tiles = gathered.shape[0], gathered.shape[1]
    t = torch.arange(n_tiles, device=gathered.device)
    return gathered[t % ws, t]


class StreamingSweep:
    def __init__(self, data, tmin, N_intervals, tmax=None, tile_variables=64, group_scoring=True, group=None):
        import torch
        from . import device as D
        from . import rules as R
        self.data = data
        self.N = data.n_subjects
        self.V = data.n_variables
        self.tile_variables = int(tile_variables)
        self.group_scoring = bool(group_scoring)
        self.group = group
        self.ws, self.rank = world(group)
        pop = R.DiscreteRulePopulation.__new__(R.DiscreteRulePopulation)
        pop.model, pop.data, pop.tmin = None, data, tmin
        pop.tmax = np.inf if tmax is None else tmax
        pop.N_intervals = N_intervals
        pop._enumerate_windows()
        self.windows = list(pop.windows)
        self.ok_slope = np.array(pop.windows_ok_for_slope, dtype=np.uint8)
        self.durations = np.array([w[1] - w[0] for w in self.windows], dtype=np.float64)
        self._ds = data.device_series()
        W = len(self.windows)
        self._d_windows = torch.as_tensor(np.array(self.windows, dtype=np.float64).reshape(-1, 2), device=D.dev())
        self._d_ok = torch.as_tensor(self.ok_slope, device=D.dev())
        if W:
            self._lo, self._cnt = D.window_ranges(self._ds['times'], self._ds['t_offsets'], self._d_windows)
        self.tiles = [(v0, min(v0 + self.tile_variables, self.V)) for v0 in range(0, self.V, self.tile_variables)]
        self._buffers = None
        self._draw = None
        self._tile_static = {}
        self.tile_stats = None

    # -- one tile: detector kernels -> rows in enumeration order (window, variable, type, threshold, direction)
    def _group_base(self, n_vars):
        per_window = n_vars * (1 + self.ok_slope.astype(np.int64))
        gb = np.zeros(len(self.windows) + 1, dtype=np.int64)
        np.cumsum(per_window, out=gb[1:])
        return gb

    def build_tile(self, t, rows=False):
        """-> dict(values, K, thresholds, row_offsets (device), group_base (host), v0, v1, P, G) and, on
        request, rows int64[P_t, W64] in enumeration order (scoring does not need them)."""
        import torch
        from . import device as D
        v0, v1 = self.tiles[t]
        gb = self._group_base(v1 - v0)
        G = int(gb[-1])
        ds = self._ds
        d_gb = torch.as_tensor(gb[:-1].copy(), device=D.dev())
        values = D.detector_values(ds['series'][v0:v1], ds['times'], ds['t_offsets'], self._d_windows,
                                   self._lo, self._cnt, d_gb, self._d_ok, G, ds['max_T'])
        thresholds, K = D.detector_thresholds(values)
        row_offsets = D.detector_row_offsets(K)
        P = int(row_offsets[-1].item())
        tile = dict(values=values, K=K, thresholds=thresholds, row_offsets=row_offsets, group_base=gb, v0=v0,
                    v1=v1, P=P, G=G)
        if rows:
            tile['rows'] = D.detector_rows(values, thresholds, K, row_offsets, P) if P else \
                torch.empty((0, D.words_for(self.N)), dtype=torch.int64, device=D.dev())
        return tile

    def truth_row(self, tile, k):
        """bool[N] truth vector of candidate k of a built tile (what the sampler appends to X)."""
        ro = tile['row_offsets'].cpu().numpy()
        g = int(np.searchsorted(ro, k, side='right') - 1)
        slot = int(k - ro[g])
        v = tile['values'][g].cpu().numpy()
        thr = float(tile['thresholds'][g, slot >> 1].item())
        return (v > thr) if (slot & 1) == 0 else (v <= thr)          # rules.py:494-497

    def locate(self, primitive, tile=None):
        """(tile index, position in the tile's enumeration order) of a primitive tuple
        (variable, (t0, t1), type, direction, threshold), or None if it is not a candidate.  Builds the
        tile unless the caller passes the built one."""
        try:
            variable, window, type_, direction, threshold = primitive
            w = self.windows.index((window[0], window[1]))
        except (TypeError, ValueError, IndexError):
            return None
        if not (0 <= variable < self.V) or type_ not in TYPES or direction not in DIRECTIONS:
            return None
        ok = bool(self.ok_slope[w])
        if type_ == 'slope' and not ok:
            return None
        t = variable // self.tile_variables
        if tile is None:
            tile = self.build_tile(t)
        g = int(tile['group_base'][w]) + (variable - tile['v0']) * (2 if ok else 1) + (TYPES.index(type_) if ok else 0)
        K = int(tile['K'][g].item())
        thr = tile['thresholds'][g, :K].cpu().numpy()
        k = int(np.searchsorted(thr, threshold))
        if k >= K or thr[k] != threshold:
            return None
        return t, int(tile['row_offsets'][g].item()) + 2 * k + DIRECTIONS.index(direction)

    def _tile_prior(self, tile, prior_tables, deltas=None):
        """float64[P_t] log-prior of the tile's rows from the factored tables
        (by_variable[V], by_window[W], normalisation) of LogisticRuleModel._prior_tables, plus the
        sparse corrections `deltas` {(tile, position): additive term} the sampler applies to the
        primitives already in the rule list (logit_rules.py:861-862, 1058-1066)."""
        import torch
        from . import device as D
        if tile['P'] == 0:
            return None
        mine = [(k, dv) for (t, k), dv in (deltas or {}).items() if self.tiles[t][0] == tile['v0']]
        if prior_tables is None:
            if not mine:
                return None
            prior = torch.zeros(tile['P'], dtype=torch.float64, device=D.dev())
        else:
            prior = self._tile_prior_tables(tile, prior_tables)
        if mine:
            idx = torch.as_tensor([k for k, _ in mine], dtype=torch.int64, device=D.dev())
            prior[idx] += torch.as_tensor([dv for _, dv in mine], dtype=torch.float64, device=D.dev())
        return prior

    def _tile_prior_tables(self, tile, prior_tables):
        import torch
        from . import device as D
        bv, bw, norm = prior_tables
        static = self._tile_static_tables(tile)
        ident = torch.arange(tile['P'], dtype=torch.int64, device=D.dev())
        rg = D.row_groups(ident, tile['row_offsets'])
        return D.row_prior(rg, static[0], static[1],
                           torch.as_tensor(np.asarray(bv, dtype=np.float64), device=D.dev()),
                           torch.as_tensor(np.asarray(bw, dtype=np.float64), device=D.dev()), float(norm))

    def _tile_static_tables(self, tile):
        """(variable, window) of every group of the tile as device index tensors: geometry only."""
        import torch
        from . import device as D
        static = self._tile_static.get(tile['v0'])
        if static is None:
            gb = tile['group_base']
            nt = 1 + self.ok_slope.astype(np.int64)
            w_of = np.repeat(np.arange(len(self.windows), dtype=np.int32), np.diff(gb))
            within = np.arange(tile['G'], dtype=np.int64) - gb[w_of]
            gvar = (tile['v0'] + within // nt[w_of]).astype(np.int32)
            static = (torch.as_tensor(gvar, device=D.dev()), torch.as_tensor(w_of, device=D.dev()))
            self._tile_static[tile['v0']] = static
        return static

    def _weights_tile(self, tile, X, omega, kappa, sigma2, mask, prior_tables, deltas=None):
        """log-weights (likelihood + prior + sparse corrections) of a built tile and their
        (max, sum-exp): what sweep / draw / log_weight consume.  With group scoring the prior enters
        per GROUP inside the scoring launch (it depends only on window and variable), so no
        per-candidate prior vector exists."""
        import torch
        from . import device as D
        if not self.group_scoring:
            out, status, stats, prior = self._score_tile(tile, X, omega, kappa, sigma2, mask, prior_tables, deltas)
            return (out if prior is None else out + prior), status, stats
        P = tile['P']
        p = int(np.shape(X)[1])
        if self._buffers is None or self._buffers.out.shape[0] < P or self._buffers.p != p:
            self._buffers = D.ScoreBuffers(max(P, 1), self.N, p)
            self._buffers.p = p
            self._draw = D.DrawBuffers(max(P, 1))
        gp = None
        if prior_tables is not None:
            bv, bw, norm = prior_tables
            gvar, w_of = self._tile_static_tables(tile)
            gp = (torch.as_tensor(np.asarray(bv, dtype=np.float64), device=D.dev())[gvar.long()] +
                  torch.as_tensor(np.asarray(bw, dtype=np.float64), device=D.dev())[w_of.long()] - float(norm))
        d_mask = None if mask is None else torch.as_tensor(D.pack_mask(np.asarray(mask, dtype=bool)), device=D.dev())
        out, status = D.score_groups(tile['values'], tile['thresholds'], tile['K'], tile['row_offsets'], P, X, omega,
                                     kappa, sigma2, mask=d_mask, buffers=self._buffers, out=self._buffers.out[:P],
                                     group_prior=gp)
        mine = [(k, dv) for (t, k), dv in (deltas or {}).items() if self.tiles[t][0] == tile['v0']]
        if mine:
            idx = torch.as_tensor([k for k, _ in mine], dtype=torch.int64, device=D.dev())
            out[idx] += torch.as_tensor([dv for _, dv in mine], dtype=torch.float64, device=D.dev())
        stats = D.weight_stats(out, None, buffers=self._draw)
        return out, status, stats

    def _score_tile(self, tile, X, omega, kappa, sigma2, mask, prior_tables, deltas=None):
        import torch
        from . import device as D
        P = tile['P']
        p = int(np.shape(X)[1])
        if self._buffers is None or self._buffers.out.shape[0] < P or self._buffers.p != p:
            self._buffers = D.ScoreBuffers(max(P, 1), self.N, p)
            self._buffers.p = p
            self._draw = D.DrawBuffers(max(P, 1))
        prior = self._tile_prior(tile, prior_tables, deltas)
        d_mask = None if mask is None else torch.as_tensor(D.pack_mask(np.asarray(mask, dtype=bool)), device=D.dev())
        if self.group_scoring:
            # suffix sums along each group's sorted order: no truth rows (mitre_score_groups)
            out, status = D.score_groups(tile['values'], tile['thresholds'], tile['K'], tile['row_offsets'], P,
                                         X, omega, kappa, sigma2, mask=d_mask, buffers=self._buffers,
                                         out=self._buffers.out[:P])
            stats = D.weight_stats(out, prior, buffers=self._draw)
        else:
            if 'rows' not in tile:
                tile['rows'] = D.detector_rows(tile['values'], tile['thresholds'], tile['K'], tile['row_offsets'], P)
            out, status, stats = D.score_all_stats(tile['rows'], self.N, X, omega, kappa, sigma2, mask=d_mask,
                                                   buffers=self._buffers, out=self._buffers.out[:P],
                                                   log_prior=prior, draw_buffers=self._draw)
        return out, status, stats, prior

    def sweep(self, X, omega, kappa, sigma2, mask=None, prior_tables=None, deltas=None):
        """Scores every candidate once.  Returns (log_total, n_candidates) and keeps one
        (max, sum-exp) pair per tile for `draw`.  deltas: {(tile, position): term added to that
        candidate's log-weight} (see `locate`)."""
        import torch
        from . import device as D
        self._sweep_args = (X, omega, kappa, sigma2, mask, prior_tables, deltas)
        import torch.distributed as dist
        n_tiles = len(self.tiles)
        pairs = torch.zeros((n_tiles, 2), dtype=torch.float64, device=D.dev())
        pairs[:, 0] = -np.inf
        counts = torch.zeros(n_tiles, dtype=torch.int64, device=D.dev())
        # one status word for the whole sweep (the per-tile buffers may be reallocated on the way), raised
        # on EVERY rank AFTER the collectives: a rank that raised before them would leave the others blocked
        status_all = torch.zeros(1, dtype=torch.int32, device=D.dev())
        for t in range(self.rank, n_tiles, self.ws):          # tiles are dealt round-robin to the ranks
            tile = self.build_tile(t)
            counts[t] = tile['P']
            if tile['P'] == 0:
                continue
            out, status, stats = self._weights_tile(tile, X, omega, kappa, sigma2, mask, prior_tables, deltas)
            pairs[t] = stats
            status_all |= status
            status.zero_()
            del tile, out
        if self.ws > 1:
            # the sweep's only exchange: 16 bytes per tile (every rank ends up with every tile's pair)
            gathered = torch.empty((self.ws, n_tiles, 2), dtype=torch.float64, device=D.dev())
            dist.all_gather_into_tensor(gathered, pairs, group=self.group)
            pairs = combine_tile_pairs(gathered)
            dist.all_reduce(counts, group=self.group)
            dist.all_reduce(status_all, op=dist.ReduceOp.MAX, group=self.group)
        D.raise_on_status(status_all)
        self.tile_stats = pairs.cpu().numpy()
        self.tile_counts = counts.cpu().numpy()
        _, _, log_total = owner_of_marker(self.tile_stats, 0.5)
        return log_total, int(self.tile_counts.sum())

    def draw(self, u):
        """Two-stage draw from the weights of the last sweep: (tile, index within the tile's
        enumeration order, primitive tuple)."""
        from . import device as D
        if self.tile_stats is None:
            raise RuntimeError("draw() needs a sweep() first")
        t, u_local, _ = owner_of_marker(self.tile_stats, u)
        X, omega, kappa, sigma2, mask, prior_tables, deltas = self._sweep_args
        result = [None]
        if t % self.ws == self.rank:                           # the tile's owner redraws inside it
            tile = self.build_tile(t)
            w, status, stats = self._weights_tile(tile, X, omega, kappa, sigma2, mask, prior_tables, deltas)
            k = int(D.draw_index(w, None, u_local, stabilise=True, buffers=self._draw, reuse_stats=True).item())
            k = min(k, tile['P'] - 1)
            result = [(t, k, self.describe(tile, k))]
        if self.ws > 1:
            import torch.distributed as dist
            src = t % self.ws
            dist.broadcast_object_list(result, src=dist.get_global_rank(self.group, src) if self.group is not None
                                       else src, group=self.group)
        return result[0]

    def log_weight(self, t, k):
        """log-weight (likelihood + prior + correction) the last sweep gave candidate (t, k), and
        log-weight minus the sweep's log-total = the log-probability `draw` picks it (what the
        reverse-move probabilities of add/remove need, logit_rules.py:1168-1170)."""
        X, omega, kappa, sigma2, mask, prior_tables, deltas = self._sweep_args
        tile = self.build_tile(t)
        wts, status, stats = self._weights_tile(tile, X, omega, kappa, sigma2, mask, prior_tables, deltas)
        w = float(wts[k].item())
        _, _, log_total = owner_of_marker(self.tile_stats, 0.5)
        return w, w - log_total

    def describe(self, tile, k):
        """(variable, window, type, direction, threshold) of row k of a built tile (rules.py:23-44)."""
        ro = tile['row_offsets'].cpu().numpy()
        g = int(np.searchsorted(ro, k, side='right') - 1)
        slot = int(k - ro[g])
        gb = tile['group_base']
        w = int(np.searchsorted(gb, g, side='right') - 1)
        nt = 2 if self.ok_slope[w] else 1
        within = g - int(gb[w])
        v = tile['v0'] + within // nt
        ty = within % nt if nt == 2 else 1
        thr = float(tile['thresholds'][g, slot >> 1].item())
        return (int(v), self.windows[w], TYPES[ty], DIRECTIONS[slot & 1], thr)


class StreamingDetectorGibbs:
    """The sampler's detector update (LogisticRuleSampler.update_primitives, logit_rules.py:794-881)
    and its omega / beta Gibbs sweeps (logit_rules.py:567-573) over a StreamingSweep: the part of a
    chain that touches all P candidates, for tables that do not fit in HBM.  The rule list's
    STRUCTURE (number of rules, detectors per rule) is held fixed: the add / remove / move moves
    and the hyperparameter updates of the full sampler are not restated here.

    For detector j of rule i the conditional draws a replacement with log-weight
        log-likelihood([X_without_i | rule i with the candidate in place of j])
      + log-prior(candidate)                          (factored tables, LogisticRuleModel._prior_tables)
      + sum over candidates already in rule i:  -gammaln(n + 2) + log(1 + n)      (logit_rules.py:856-862)
    exactly the weights of the resident sampler (`_device_gibbs_draw`); only the ORDER in which the
    candidates are laid out differs (tile, enumeration) -- the same distribution over the same set.

    rule_list: list of rules, each a list of primitive tuples (variable, window, type, direction, threshold).
    """

    def __init__(self, sweep, y, rule_list, prior_coefficient_variance=100.0, prior_tables=None, seed=0):
        self.sweep = sweep
        self.y = np.asarray(y, dtype=bool)
        self.kappa = self.y.astype(np.float64) - 0.5
        self.sigma2 = float(prior_coefficient_variance)
        self.prior_tables = prior_tables
        self.rng = np.random.RandomState(seed)
        self.rules = [list(rule) for rule in rule_list]
        self.truth = [[self._truth_of(p) for p in rule] for rule in self.rules]
        N = sweep.N
        self.omega = np.full(N, 0.25)
        self.beta = np.zeros(len(self.rules) + 1)
        self.log = []

    def _truth_of(self, primitive):
        loc = self.sweep.locate(primitive)
        if loc is None:
            raise ValueError("not a candidate of this population: %r" % (primitive,))
        tile = self.sweep.build_tile(loc[0])
        return self.sweep.truth_row(tile, loc[1])

    def design_matrix(self, skip_rule=None):
        """[rule columns | 1] (rules.py:1173-1175); a rule's column is the AND of its detectors."""
        cols = [np.logical_and.reduce(t).astype(np.float64) for i, t in enumerate(self.truth) if i != skip_rule]
        return np.stack(cols + [np.ones(self.sweep.N)], axis=1)

    def update_omega_beta(self, n_iters=10):
        from . import device as D
        X = self.design_matrix()
        beta0 = self.beta if len(self.beta) == X.shape[1] else np.zeros(X.shape[1])
        omega, beta, betas, status = D.gibbs_omega_beta(X, self.kappa, self.sigma2, beta0, n_iters,
                                                        int(self.rng.randint(0, 2 ** 62)))
        D.raise_on_status(status)
        self.omega = omega.cpu().numpy()
        self.beta = beta.cpu().numpy()

    def conditional(self, i, j):
        """Runs the sweep for position (i, j); returns (log_total, n_candidates)."""
        from scipy.special import gammaln
        others = [t for jj, t in enumerate(self.truth[i]) if jj != j]
        mask = np.logical_and.reduce(others) if others else np.ones(self.sweep.N, dtype=bool)
        X = self.design_matrix(skip_rule=i)
        counts = {}
        for jj, p in enumerate(self.rules[i]):
            if jj != j:
                counts[p] = counts.get(p, 0) + 1
        deltas = {}
        for p, n in counts.items():
            loc = self.sweep.locate(p)
            if loc is not None:
                deltas[loc] = -1.0 * gammaln(n + 1 + 1) + np.log(1.0 + n)
        return self.sweep.sweep(X, self.omega, self.kappa, self.sigma2, mask=mask, prior_tables=self.prior_tables,
                                deltas=deltas)

    def update_detector(self, i=None, j=None):
        """One Gibbs update of a detector (uniformly chosen unless given)."""
        positions = [(a, b) for a, rule in enumerate(self.rules) for b in range(len(rule))]
        if i is None:
            i, j = positions[self.rng.randint(len(positions))]
        log_total, n = self.conditional(i, j)
        t, k, primitive = self.sweep.draw(self.rng.rand())
        tile = self.sweep.build_tile(t)
        old = self.rules[i][j]
        self.rules[i][j] = primitive
        self.truth[i][j] = self.sweep.truth_row(tile, k)
        self.log.append((i, j, old, primitive, log_total))
        return primitive

    def step(self, n_gibbs=10):
        """update_detector + the omega / beta sweeps: the P-dependent part of
        LogisticRuleSampler.step (logit_rules.py:469-504) at fixed structure."""
        self.update_detector()
        self.update_omega_beta(n_gibbs)


# ---- the full sampler step over a streamed population ------------------------------------------------------
#
# LogisticRuleSampler.step (logit_rules.py:567-596) = omega / beta sweeps + one of {update_primitives, add /
# remove, move_primitive} + the hyperparameter updates.  Everything in it that touches the candidate table does so
# through four things: (a) the log-sum-exp of likelihood + prior (+ sparse terms) over all candidates under a trial
# mask, (b) a draw from those weights, (c) the truth vector and (d) the prior of the handful of primitives that sit
# in the rule list.  StreamingSweep provides (a) and (b) tile by tile, its tiles rebuilt on demand provide (c), the
# factored prior tables (d) -- so the reference's step runs on a table that is never materialised (cfg3: 0.8 TB).

class _StreamedTruth:
    """Mapping view primitive tuple -> bool[N] (what Dataset._primitive_result_cache points at): the truth
    vector comes from rebuilding the primitive's tile; the few primitives a chain holds are memoised."""

    def __init__(self, pop):
        self._pop = pop
        self._rows = {}

    def get(self, key, default=None):
        row = self._rows.get(key)
        if row is None:
            loc = self._pop._locate(key)
            if loc is None:
                return default
            sweep = self._pop.sweep
            row = sweep.truth_row(sweep.build_tile(loc[0]), loc[1])
            if len(self._rows) > 4096:
                self._rows.clear()
            self._rows[key] = row
        return row

    def __contains__(self, key):
        return self.get(key) is not None

    def __getitem__(self, key):
        row = self.get(key)
        if row is None:
            raise KeyError(key)
        return row


class _StreamedPopulation:
    """The slice of DiscreteRulePopulation's surface that LogisticRuleModel / LogisticRuleSampler use, over a
    StreamingSweep.  A primitive's index is its (tile, position in the tile's enumeration order)."""

    def __init__(self, sweep):
        self.sweep = sweep
        self.windows = sweep.windows
        self.windows_ok_for_slope = [bool(x) for x in sweep.ok_slope]
        self._window_lookup = {(float(a), float(b)): i for i, (a, b) in enumerate(self.windows)}
        self._durations_by_window = sweep.durations
        self.truth_table = _StreamedTruth(self)
        self._filter_generation = 0
        self._build_generation = 0
        self._locs = {}
        self._counts = None

    def _locate(self, key):
        if key not in self._locs:
            if len(self._locs) > 4096:
                self._locs.clear()
            self._locs[key] = self.sweep.locate(key)
        return self._locs[key]

    def get_primitive_index(self, primitive):
        return self._locate(primitive.as_tuple())

    def sort_list_of_primitives(self, l):
        l.sort(key=lambda p: self.get_primitive_index(p) or (-1, -1))

    def count_matrix(self):
        """float64[W, V]: candidates per (window, variable), both detector types and directions -- one detector
        pass over all tiles (each rank its own tiles, then an all-reduce), kept for the life of the population."""
        if self._counts is None:
            import torch
            from . import device as D
            sw = self.sweep
            C = torch.zeros((len(self.windows), sw.V), dtype=torch.float64, device=D.dev())
            for t in range(sw.rank, len(sw.tiles), sw.ws):
                tile = sw.build_tile(t)
                gvar, w_of = sw._tile_static_tables(tile)
                C.index_put_((w_of.long(), gvar.long()), 2.0 * tile['K'].to(torch.float64), accumulate=True)
            if sw.ws > 1:
                import torch.distributed as dist
                dist.all_reduce(C, group=sw.group)
            self._counts = C.cpu().numpy()
        return self._counts

    @property
    def n_primitives(self):
        return int(self.count_matrix().sum())

    def __len__(self):
        return self.n_primitives


def make_streaming_model(data, tmin, tile_variables=64, group=None, **model_kwargs):
    """LogisticRuleModel (same constructor arguments) over a streamed candidate population."""
    from . import logit_rules as LR

    class StreamingModel(LR.LogisticRuleModel):
        """The model of logit_rules.py:77-457 with the candidate table behind a StreamingSweep.  The prior is
        always the factored form (_prior_tables); the P-long host / device vectors of the resident model
        (flat_prior, likelihoods_of_all_options) do not exist."""
        fast_prior = True

        def _build_population(self, N_intervals, kwargs):
            self.sweep = StreamingSweep(self.data, self.tmin, N_intervals, tmax=kwargs.get('tmax', None),
                                        tile_variables=tile_variables, group=group)
            return _StreamedPopulation(self.sweep)

        def _prior_count_matrix(self):
            return self.rule_population.count_matrix()

        def _prior_tables(self, state):
            try:
                return LR.LogisticRuleModel._prior_tables(self, state)
            except AttributeError:
                # degenerate hyperparameters (the resident model falls back to a per-group log-sum-exp): the
                # same sum over the (window, variable) count matrix
                from . import fastdist
                pop = self.rule_population
                log_vw = np.log(np.asarray(self.data.variable_weights, dtype=np.float64))
                bv = fastdist.norm_logpdf(log_vw, state.phylogeny_mean, state.phylogeny_std)
                total_duration = float(self.data.experiment_end - self.data.experiment_start)
                durations = pop._durations_by_window / ((1.0 + self.window_duration_epsilon) * total_duration)
                bw = fastdist.beta_logpdf(durations, state.window_concentration * state.window_typical_fraction,
                                          state.window_concentration * (1.0 - state.window_typical_fraction))
                C = pop.count_matrix()
                terms = bw[:, None] + bv[None, :]
                live = C > 0
                norm = fastdist.logsumexp(terms[live], b=C[live]) if live.any() else 0.0
                return bv, bw, float(norm)

        def _no_table(self, *a, **k):
            raise NotImplementedError("a streamed population has no P-long vectors; use StreamingSampler")
        flat_prior = flat_prior_device = likelihoods_of_all_options = likelihoods_of_all_options_device = \
            likelihoods_with_stats_device = _no_table

    return StreamingModel(data, tmin, **model_kwargs)


def make_streaming_sampler(model, r0, local_updates_per_structure_update=10, device_gibbs=True):
    """LogisticRuleSampler over a make_streaming_model(...) model: the reference's full step -- detector
    updates, add / remove, move, hyperparameter updates, omega / beta sweeps -- with np.random consumed in the
    resident sampler's order."""
    from scipy.special import gammaln
    from . import logit_rules as LR
    from .rules import PrimitiveRule

    class StreamingSampler(LR.LogisticRuleSampler):
        def _placeholder(self):
            return PrimitiveRule(0, self.model.rule_population.windows[0], 'average', 'above', 0.0)

        def _stream_sweep(self, rl, i, j, deltas):
            """log-sum-exp of likelihood + prior + sparse terms over every candidate for position (i, j)."""
            m = self.model
            X, mask = m._sweep_inputs(rl, i, j)
            log_total, _ = m.sweep.sweep(np.asarray(X, dtype=np.float64), self.current_state.omega,
                                         m.data.y - 0.5, m.prior_coefficient_variance, mask=mask,
                                         prior_tables=m._prior_tables(self.current_state), deltas=deltas or None)
            return log_total

        def _primitive_log_prior(self, primitive):
            bv, bw, norm = self.model._prior_tables(self.current_state)
            w = self.model.rule_population._window_lookup[(float(primitive.window[0]), float(primitive.window[1]))]
            return bv[primitive.variable] + bw[w] - norm

        def _draw_candidate(self, candidate_distribution):
            t, k, primitive = self.model.sweep.draw(np.random.rand())
            return (t, k), PrimitiveRule(*primitive)

        # logit_rules.py:794-881
        def update_primitives(self, N_updates=None):
            pop = self.model.rule_population
            rule_list = self.current_state.rl
            positions = [(i, j) for i, rule in enumerate(rule_list) for j, _ in enumerate(rule)]
            if len(positions) == 0:
                self.comments.append('Null update')
                return
            for _ in range(1 if N_updates is None else N_updates):
                i, j = positions[np.random.choice(len(positions))]
                k0 = pop.get_primitive_index(rule_list[i][j])
                counts = {}
                for jj, p in enumerate(rule_list[i]):
                    loc = pop.get_primitive_index(p)
                    if jj != j and loc is not None:
                        counts[loc] = counts.get(loc, 0) + 1
                deltas = {loc: -1.0 * gammaln(n + 1 + 1) + np.log(1.0 + n) for loc, n in counts.items()}
                self._stream_sweep(rule_list, i, j, deltas)
                k, new_primitive = self._draw_candidate(None)
                rule_list[i][j] = new_primitive
                self.comments.append('At (%d,%d) replacing %s with %s' % (i, j, k0, k))
            self.ensure_current_rl_sorted()
            self.current_X = self.model.data.covariate_matrix(rule_list)

        # logit_rules.py:894-970
        def consider_addition_to_new_rule(self, base_rule, position_to_add_at):
            m = self.model
            new_rl = base_rule.copy()
            new_rl.rules = (new_rl.rules[:position_to_add_at] + [[self._placeholder()]] +
                            new_rl.rules[position_to_add_at:])
            s = [m.structure_prior(rl, empty_probability=self.current_state.empty_probability)
                 for rl in [base_rule, new_rl]]
            log_relative_prior = s[1] - s[0]
            old_likelihood = m.likelihood_z_given_X(self.current_state.omega, m.data.covariate_matrix(base_rule))
            log_marginalized_likelihood = self._stream_sweep(new_rl, position_to_add_at, 0, None)
            log_likelihood_ratio = log_marginalized_likelihood - old_likelihood
            probability_remove_proposal = 1.0 / sum(map(len, new_rl))
            return (new_rl, log_relative_prior, log_likelihood_ratio, ('stream',), probability_remove_proposal)

        # logit_rules.py:972-1095
        def consider_addition_to_existing_rule(self, base_rule, position_to_add_to):
            m = self.model
            pop = m.rule_population
            base_subrule = base_rule[position_to_add_to]
            base_n_istar = len(base_subrule)
            counts, prior_of = {}, {}
            for primitive in base_subrule:
                loc = pop.get_primitive_index(primitive)
                if loc is not None:
                    counts[loc] = counts.get(loc, 0) + 1
                    prior_of[loc] = self._primitive_log_prior(primitive)
            c0 = np.log(base_n_istar + 1)
            deltas = {loc: -1.0 * np.log(n + 1) for loc, n in counts.items()}
            new_rl = base_rule.copy()
            new_rl[position_to_add_to].append(self._placeholder())
            s = [m.structure_prior(rl, empty_probability=self.current_state.empty_probability)
                 for rl in [base_rule, new_rl]]
            log_relative_structure_prior = s[1] - s[0]
            old_likelihood = m.likelihood_z_given_X(self.current_state.omega, m.data.covariate_matrix(base_rule))
            # A = lse(prior + deltas): the prior is normalised over the candidates, only the sparse terms move it
            lse_A = np.log1p(sum(np.exp(prior_of[loc]) * np.expm1(d) for loc, d in deltas.items()))
            # B = lse(prior + deltas + likelihood): the sweep (and what an accepted proposal draws from)
            lse_B = self._stream_sweep(new_rl, position_to_add_to, base_n_istar, deltas)
            # C = lse(prior + likelihood): B with the sparse terms taken back out of the few candidates that carry them
            lse_C = lse_B + np.log1p(sum(np.expm1(-d) * np.exp(m.sweep.log_weight(*loc)[0] - lse_B)
                                         for loc, d in deltas.items()))
            log_relative_prior_primitive_piece = c0 + lse_A
            log_relative_prior = log_relative_prior_primitive_piece + log_relative_structure_prior
            log_marginalized_likelihood = (c0 + lse_B) - log_relative_prior_primitive_piece
            log_likelihood_ratio = log_marginalized_likelihood - old_likelihood
            new_total_n_primitives = sum(map(len, new_rl))
            T_sstar_prefactor = (base_n_istar + 1.0) / new_total_n_primitives
            T_sstar_other_factor = np.exp(lse_C - (c0 + lse_B))
            return (new_rl, log_relative_prior, log_likelihood_ratio, ('stream',),
                    T_sstar_prefactor * T_sstar_other_factor)

    return StreamingSampler(model, r0, local_updates_per_structure_update, device_draw=True,
                            device_gibbs=device_gibbs)
