"""The reference's full sampler step (logit_rules.py:567-596) on ONE problem whose candidate table is sharded
over the GPUs of a box (parallel.ShardedPopulation + ShardedSweep).

The sampler itself is replicated: every rank runs the same host code on the same np.random stream, so every
rank proposes the same move; whatever touches the P candidates goes through the sharded step -- each rank
scores its slice, the (max, sum-exp) pairs and the drawn index are exchanged inside the step (parallel.py) --
and the few per-primitive facts a move needs (a truth vector, a threshold, a weight) are broadcast by the rank
that owns them.  Candidates keep the reference's global lexsorted order, so with the same seed a sharded chain
makes the moves of the resident chain (tests/test_sharded_sampler_gpu.py).

    model = make_sharded_model(data, tmin, N_intervals=..., tmax=..., group=None)
    sampler = make_sharded_sampler(model, r0)
    sampler.sample(n)
"""
import numpy as np

from .parallel import ShardedPopulation, ShardedSweep, describe_group, world
from .rules import DIRECTIONS, TYPES


class _ShardedTruth:
    """Mapping view primitive tuple -> bool[N] (Dataset._primitive_result_cache): the owner of the primitive's
    slice position unpacks the row and broadcasts it (8 W64 bytes); rows of the rule list's primitives are
    memoised."""

    def __init__(self, pop):
        self._pop = pop
        self._rows = {}

    def get(self, key, default=None):
        row = self._rows.get(key)
        if row is None:
            k = self._pop._locate(key)
            if k is None:
                return default
            if len(self._rows) > 4096:
                self._rows.clear()
            row = self._rows[key] = self._pop.truth_row(k)
        return row

    def __contains__(self, key):
        return self.get(key) is not None

    def __getitem__(self, key):
        row = self.get(key)
        if row is None:
            raise KeyError(key)
        return row


class _ShardedPopulationView:
    """The slice of DiscreteRulePopulation's surface that LogisticRuleModel / LogisticRuleSampler use, over a
    ShardedPopulation.  A primitive's index is its position in the reference's global sorted order."""

    def __init__(self, spop):
        self.spop = spop
        self.windows = spop.windows
        self.windows_ok_for_slope = [bool(x) for x in spop.ok_slope]
        self._window_lookup = {(float(a), float(b)): i for i, (a, b) in enumerate(self.windows)}
        self._durations_by_window = np.array([w[1] - w[0] for w in self.windows], dtype=np.float64)
        self.truth_table = _ShardedTruth(self)
        self._filter_generation = 0
        self._build_generation = 0
        self._locs = {}
        self._counts = None
        self.n_primitives = spop.P

    def __len__(self):
        return self.spop.P

    # -- primitive tuple -> global sorted index -------------------------------------------------------------
    def _locate(self, key):
        if key in self._locs:
            return self._locs[key]
        if len(self._locs) > 4096:
            self._locs.clear()
        self._locs[key] = k = self._locate_uncached(key)
        return k

    def _locate_uncached(self, key):
        import torch
        sp = self.spop
        try:
            variable, window, type_, direction, threshold = key
            w = self._window_lookup[(float(window[0]), float(window[1]))]
        except (TypeError, ValueError, KeyError):
            return None
        if not (0 <= variable < sp.V) or type_ not in TYPES or direction not in DIRECTIONS:
            return None
        ok = bool(sp.ok_slope[w])
        if type_ == 'slope' and not ok:
            return None
        nt = 2 if ok else 1
        t = TYPES.index(type_) if ok else 0
        g_global = int(sp.GB[w]) + variable * nt + t
        # the threshold's slot: known to the rank that evaluated the variable
        vowner = sp.owner_of_variable(variable)
        slot = -1
        if sp.rank == vowner:
            pop = sp.local_population
            g_loc = int(pop._group_base[w]) + (variable - sp.v0) * nt + t
            K = int(pop._K_host[g_loc])
            thr = pop._thresholds[g_loc, :K].cpu().numpy()
            i = int(np.searchsorted(thr, threshold))
            if i < K and thr[i] == threshold:
                slot = i
        slot = int(sp._bcast(slot, vowner, torch.int64))
        if slot < 0:
            return None
        e = int(sp._row_offsets_host[g_global]) + 2 * slot + DIRECTIONS.index(direction)
        # sorted position of enumeration index e: on the rank whose slice holds it
        hit = torch.nonzero(sp.local_perm == e).ravel()
        k = torch.full((1,), -1, dtype=torch.int64, device=sp.local_perm.device)
        if hit.numel():
            k[0] = sp.start + hit[0]
        if sp.ws > 1:
            import torch.distributed as dist
            dist.all_reduce(k, op=dist.ReduceOp.MAX, group=sp.group)
        k = int(k.item())
        return k if k >= 0 else None

    def get_primitive_index(self, primitive):
        return self._locate(primitive.as_tuple())

    def sort_list_of_primitives(self, l):
        l.sort(key=lambda p: -1 if self.get_primitive_index(p) is None else self.get_primitive_index(p))

    def truth_row(self, k):
        import torch
        from . import device as D
        sp = self.spop
        owner = sp.owner_of_index(k)
        row = torch.zeros((1, D.words_for(sp.N)), dtype=torch.int64, device=sp.local_rows.device)
        if sp.rank == owner:
            row.copy_(sp.local_rows[k - sp.start:k - sp.start + 1])
        if sp.ws > 1:
            import torch.distributed as dist
            dist.broadcast(row, src=dist.get_global_rank(sp.group, owner) if sp.group is not None else owner,
                           group=sp.group)
        return D.unpack_rows(row, sp.N)[0].cpu().numpy().astype(bool)

    def count_matrix(self):
        """float64[W, V]: candidates per (window, variable) from the global row offsets (every rank has them)."""
        if self._counts is None:
            sp = self.spop
            counts = np.diff(sp._row_offsets_host).astype(np.float64)          # 2 K per global group
            nt = 1 + np.asarray(sp.ok_slope, dtype=np.int64)
            w_of = np.repeat(np.arange(len(self.windows), dtype=np.int64), sp.V * nt)
            within = np.arange(sp.G, dtype=np.int64) - sp.GB[w_of]
            C = np.zeros((len(self.windows), sp.V), dtype=np.float64)
            np.add.at(C, (w_of, within // nt[w_of]), counts)
            self._counts = C
        return self._counts

    def group_tables(self):
        """(variable, window) of every GLOBAL group as int32 device tensors."""
        import torch
        sp = self.spop
        nt = 1 + np.asarray(sp.ok_slope, dtype=np.int64)
        w_of = np.repeat(np.arange(len(self.windows), dtype=np.int64), sp.V * nt)
        within = np.arange(sp.G, dtype=np.int64) - sp.GB[w_of]
        dev = sp.local_rows.device
        return (torch.as_tensor((within // nt[w_of]).astype(np.int32), device=dev),
                torch.as_tensor(w_of.astype(np.int32), device=dev))


def make_sharded_model(data, tmin, group=None, exchange='p2p', **model_kwargs):
    """LogisticRuleModel (same constructor arguments) over a candidate table sharded across the ranks of
    `group` (default: the world).  Every rank passes the same full Dataset."""
    import torch
    from . import device as D, logit_rules as LR
    from ._lib import MAX_P

    class ShardedModel(LR.LogisticRuleModel):
        fast_prior = True

        def _build_population(self, N_intervals, kwargs):
            sp = ShardedPopulation(self.data, self.tmin, N_intervals, tmax=kwargs.get('tmax', None), group=group)
            self.spop = sp
            view = _ShardedPopulationView(sp)
            dev = sp.local_rows.device
            n = int(sp.local_rows.shape[0])
            self._d_prior = torch.zeros(max(n, 1), dtype=torch.float64, device=dev)[:n]
            self._row_group = D.row_groups(sp.local_perm, sp.row_offsets) if n else \
                torch.empty(0, dtype=torch.int32, device=dev)
            self._gvar, self._gwin = view.group_tables()
            self._prior_key = None
            self.sweep = ShardedSweep(sp.local_rows, self._d_prior, sp.P, sp.N, sp.start, group=group,
                                      exchange=exchange, max_p=MAX_P)
            return view

        def _prior_count_matrix(self):
            return self.rule_population.count_matrix()

        def refresh_prior(self, state):
            """This rank's slice of flat_prior (logit_rules.py:424-457) for `state`, in place in the buffer the
            sharded step reads."""
            key = (state.phylogeny_mean, state.phylogeny_std, state.window_concentration,
                   state.window_typical_fraction)
            if key != self._prior_key:
                bv, bw, norm = self._prior_tables(state)
                if self._d_prior.shape[0]:
                    dev = self._d_prior.device
                    D.row_prior(self._row_group, self._gvar, self._gwin, torch.as_tensor(bv, device=dev),
                                torch.as_tensor(bw, device=dev), norm, out=self._d_prior)
                self._prior_key = key

        def _no_table(self, *a, **k):
            raise NotImplementedError("a sharded population has no P-long vectors on one rank; use the sharded sampler")
        flat_prior = flat_prior_device = likelihoods_of_all_options = likelihoods_of_all_options_device = \
            likelihoods_with_stats_device = _no_table

    return ShardedModel(data, tmin, **model_kwargs)


def make_sharded_sampler(model, r0, local_updates_per_structure_update=10, device_gibbs=True):
    """LogisticRuleSampler over a make_sharded_model(...) model; run it on every rank with the same seed."""
    from scipy.special import gammaln
    from . import logit_rules as LR
    from .rules import PrimitiveRule

    class ShardedSampler(LR.LogisticRuleSampler):
        def _placeholder(self):
            return PrimitiveRule(0, self.model.rule_population.windows[0], 'average', 'above', 0.0)

        def _sharded_step(self, rl, i, j, deltas, u):
            """One sharded sweep for position (i, j): (drawn global index, log-sum-exp of likelihood + prior +
            sparse terms over every candidate).  deltas: {global index: additive term}."""
            import torch
            m = self.model
            sp, sw = m.spop, m.sweep
            m.refresh_prior(self.current_state)
            X, mask = m._sweep_inputs(rl, i, j)
            mine = [(k - sp.start, d) for k, d in (deltas or {}).items() if sp.start <= k < sp.stop]
            saved = None
            if mine:
                idx = torch.as_tensor([k for k, _ in mine], dtype=torch.int64, device=m._d_prior.device)
                saved = m._d_prior[idx].clone()
                m._d_prior[idx] = saved + torch.as_tensor([d for _, d in mine], dtype=torch.float64,
                                                          device=m._d_prior.device)
            try:
                return sw.step(np.asarray(X, dtype=np.float64), self.current_state.omega, m.data.y - 0.5, mask,
                               float(u), m.prior_coefficient_variance)
            finally:
                if saved is not None:
                    m._d_prior[idx] = saved

        def _weight_of(self, k):
            """log-weight (likelihood + prior as the last step saw it) of global candidate k, on every rank."""
            import torch
            m = self.model
            sp, sw = m.spop, m.sweep
            owner = sp.owner_of_index(k)
            w = 0.0
            if sp.rank == owner:
                w = float(sw.buffers.out[k - sp.start] + m._d_prior[k - sp.start])
            return float(sp._bcast(w, owner, torch.float64))

        def _primitive_log_prior(self, primitive):
            bv, bw, norm = self.model._prior_tables(self.current_state)
            w = self.model.rule_population._window_lookup[(float(primitive.window[0]), float(primitive.window[1]))]
            return bv[primitive.variable] + bw[w] - norm

        def _draw_candidate(self, candidate_distribution):
            """An accepted add: the step is repeated with the uniform the reference draws AFTER its acceptance
            test (the step makes its draw inside the sweep, and np.random must be consumed in the same order)."""
            _tag, rl, i, j, deltas = candidate_distribution
            k, _ = self._sharded_step(rl, i, j, deltas, np.random.rand())
            return k, self._primitive_at(k)

        def _primitive_at(self, k):
            """PrimitiveRule of global sorted index k; its index is remembered (looking a tuple's position up
            again would scan the slices' permutations)."""
            tup = self.model.spop.describe(k)
            self.model.rule_population._locs[tup] = k
            return PrimitiveRule(*tup)

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
                    k = pop.get_primitive_index(p)
                    if jj != j and k is not None:
                        counts[k] = counts.get(k, 0) + 1
                deltas = {k: -1.0 * gammaln(n + 1 + 1) + np.log(1.0 + n) for k, n in counts.items()}
                k, _ = self._sharded_step(rule_list, i, j, deltas, np.random.rand())
                rule_list[i][j] = self._primitive_at(k)
                self.comments.append('At (%d,%d) replacing %d with %d' % (i, j, -1 if k0 is None else k0, k))
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
            _, log_marginalized_likelihood = self._sharded_step(new_rl, position_to_add_at, 0, None, 0.5)
            log_likelihood_ratio = log_marginalized_likelihood - old_likelihood
            probability_remove_proposal = 1.0 / sum(map(len, new_rl))
            return (new_rl, log_relative_prior, log_likelihood_ratio,
                    ('sharded', new_rl, position_to_add_at, 0, None), probability_remove_proposal)

        # logit_rules.py:972-1095 (the three log-sum-exps as in streaming.make_streaming_sampler)
        def consider_addition_to_existing_rule(self, base_rule, position_to_add_to):
            m = self.model
            pop = m.rule_population
            base_subrule = base_rule[position_to_add_to]
            base_n_istar = len(base_subrule)
            counts, prior_of = {}, {}
            for primitive in base_subrule:
                k = pop.get_primitive_index(primitive)
                if k is not None:
                    counts[k] = counts.get(k, 0) + 1
                    prior_of[k] = self._primitive_log_prior(primitive)
            c0 = np.log(base_n_istar + 1)
            deltas = {k: -1.0 * np.log(n + 1) for k, n in counts.items()}
            new_rl = base_rule.copy()
            new_rl[position_to_add_to].append(self._placeholder())
            s = [m.structure_prior(rl, empty_probability=self.current_state.empty_probability)
                 for rl in [base_rule, new_rl]]
            log_relative_structure_prior = s[1] - s[0]
            old_likelihood = m.likelihood_z_given_X(self.current_state.omega, m.data.covariate_matrix(base_rule))
            lse_A = np.log1p(sum(np.exp(prior_of[k]) * np.expm1(d) for k, d in deltas.items()))
            _, lse_B = self._sharded_step(new_rl, position_to_add_to, base_n_istar, deltas, 0.5)
            # the step's prior buffer is restored by now: the weight the step saw = out + prior + delta
            lse_C = lse_B + np.log1p(sum(np.expm1(-d) * np.exp(self._weight_of(k) + d - lse_B)
                                         for k, d in deltas.items()))
            log_relative_prior_primitive_piece = c0 + lse_A
            log_relative_prior = log_relative_prior_primitive_piece + log_relative_structure_prior
            log_marginalized_likelihood = (c0 + lse_B) - log_relative_prior_primitive_piece
            log_likelihood_ratio = log_marginalized_likelihood - old_likelihood
            new_total_n_primitives = sum(map(len, new_rl))
            T_sstar_prefactor = (base_n_istar + 1.0) / new_total_n_primitives
            T_sstar_other_factor = np.exp(lse_C - (c0 + lse_B))
            return (new_rl, log_relative_prior, log_likelihood_ratio,
                    ('sharded', new_rl, position_to_add_to, base_n_istar, deltas),
                    T_sstar_prefactor * T_sstar_other_factor)

    return ShardedSampler(model, r0, local_updates_per_structure_update, device_draw=True,
                          device_gibbs=device_gibbs)
