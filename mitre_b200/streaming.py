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
