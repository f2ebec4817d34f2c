"""Host carrier of the reference's model / sampler surface (mitre/logit_rules.py) with the
likelihood hot path on the GPU.

Public names and call signatures follow the reference -- LogisticRuleModel
(logit_rules.py:77-457), LogisticRuleState (479-501), LogisticRuleSampler (503-1420) -- so that
a chain can be driven exactly like `main.sample` does (main.py:929-980).  What changes:

  * likelihoods_of_all_options (186-229): no bool->float64 [P, N] temporary; the mask, base
    design matrix and omega go to the device (~KBs) and mitre_score_all runs over the
    device-resident packed table.
  * likelihood_z_given_X (171-177): never forms the N x N covariance; move_primitive's
    2(2R'+1) evaluations (1375-1401) are one batched launch.
  * flat_prior (424-457): evaluated per variable and per window (V + W scipy evaluations) and
    gathered to the P primitives -- elementwise identical to the reference's P-long evaluation.

Everything else (priors, MH bookkeeping, the order in which np.random is consumed) is host
logic kept step-for-step compatible with the reference so that, given the same seed and the same
Polya-Gamma sampler, both samplers walk the same chain (tests/golden/make_golden_chain.py).
"""
import time

import numpy as np
import scipy.stats
from scipy.special import gammaln, logsumexp
from scipy.stats import nbinom  # noqa: F401  (kept for pickles of older models)

from . import fastdist

from . import polyagamma
from .rules import (PrimitiveRule, RuleList, DiscreteRulePopulation, Dataset, RuleListSampler,
                    choose_from_relative_probabilities, logger)

default_n_workers = 6

ppg = polyagamma.PyPolyaGamma(0)   # logit_rules.py:62


def log_multinomial_coefficient(list_of_primitive_rules):
    """logit_rules.py:64-75."""
    counts = {}
    for p in list_of_primitive_rules:
        t = p.as_tuple()
        counts[t] = counts.get(t, 0) + 1
    n = len(list_of_primitive_rules)
    v = np.array(list(counts.values()))
    return gammaln(n + 1) - np.sum(gammaln(v + 1))


class LogisticRuleModel(object):
    """logit_rules.py:77-457."""

    def __init__(self, data, tmin,
                 prior_coefficient_variance=100.,
                 hyperparameter_alpha_m=0.5,
                 hyperparameter_beta_m=2.0,
                 hyperparameter_alpha_primitives=2.0,
                 hyperparameter_beta_primitives=4.0,
                 window_concentration_typical=5.0,
                 window_concentration_update_ratio=0.2,
                 window_fraction_proposal_std=0.2,
                 n_workers=default_n_workers,
                 N_intervals=10,
                 hyperparameter_a_empty=0.5,
                 hyperparameter_b_empty=0.5,
                 min_phylogeny_delta_l=1e-3,
                 max_rules=10,
                 max_primitives=10,
                 delta_l_scale_mean=50.,
                 delta_l_scale_sigma=25.,
                 lambda_l_offset=0.,
                 **kwargs):
        self.data = data
        self.tmin = tmin
        self.rule_population = self._build_population(N_intervals, kwargs)
        self.data._primitive_result_cache = self.rule_population.truth_table

        self.prior_coefficient_variance = prior_coefficient_variance
        self.hyperparameter_alpha_m = hyperparameter_alpha_m
        self.hyperparameter_beta_m = hyperparameter_beta_m
        self.hyperparameter_alpha_primitives = hyperparameter_alpha_primitives
        self.hyperparameter_beta_primitives = hyperparameter_beta_primitives
        self.n_workers = n_workers
        self.max_rules = max_rules
        self.max_primitives = max_primitives
        self.hyperparameter_a_empty = hyperparameter_a_empty
        self.hyperparameter_b_empty = hyperparameter_b_empty

        log_w = np.log(self.data.variable_weights)
        phylogeny_low, phylogeny_high = np.percentile(log_w, [2.5, 100.0])
        self.phylogeny_delta_l = max(phylogeny_high - phylogeny_low, min_phylogeny_delta_l)
        self.phylogeny_lambda_l = np.median(log_w) + lambda_l_offset
        self.phylogeny_mean_hyperprior_variance = delta_l_scale_mean * self.phylogeny_delta_l ** 2
        self.phylogeny_std_upper_bound = delta_l_scale_sigma * self.phylogeny_delta_l
        self.window_concentration_typical = window_concentration_typical
        self.window_concentration_update_ratio = window_concentration_update_ratio
        self.window_fraction_proposal_std = window_fraction_proposal_std
        self.window_duration_epsilon = 0.01
        self._device_state = None

    def _build_population(self, N_intervals, kwargs):
        """The resident candidate table (streaming.StreamingModel overrides this with a view of a
        population that is never materialised)."""
        return DiscreteRulePopulation(self, self.tmin, N_intervals, max_thresholds=kwargs.get('max_thresholds'),
                                      tmax=kwargs.get('tmax', None))

    # ---- device-side per-model buffers (never pickled) ---------------------------------------------
    def __getstate__(self):
        state = dict(self.__dict__)
        state['_device_state'] = None
        state.pop('_flat_prior_cache', None)
        state.pop('_prior_table_cache', None)
        state.pop('_structure_terms', None)
        state.pop('_prior_device_cache', None)
        state.pop('_prior_count_cache', None)
        state.pop('_step_memo', None)
        state.pop('_prior_cx_cache', None)
        return state

    def _dev(self):
        """Pinned host staging + device buffers reused by every sweep."""
        if self._device_state is None:
            import torch
            from . import device as D
            from ._lib import MAX_P
            N = self.data.n_subjects
            W64 = D.words_for(N)
            P = len(self.rule_population)
            n_small = N * MAX_P + 2 * N + W64
            st = dict(
                N=N, W64=W64, MAX_P=MAX_P,
                h_small=torch.empty(n_small, dtype=torch.float64).pin_memory(),
                d_small=torch.empty(n_small, dtype=torch.float64, device=D.dev()),
                buffers=D.ScoreBuffers(P, N),
                h_out=torch.empty(max(P, 1), dtype=torch.float64).pin_memory(),
            )
            self._device_state = st
        return self._device_state

    def _stage_sweep_inputs(self, X, omega, mask_bool):
        """One pinned -> device copy of (X, omega, kappa, packed mask); returns device views."""
        import torch
        from . import device as D
        st = self._dev()
        N, W64 = st['N'], st['W64']
        p = X.shape[1]
        if p > st['MAX_P']:
            raise ValueError('design matrix has %d columns; the kernels support %d' % (p, st['MAX_P']))
        h = st['h_small'].numpy()
        h[:N * p] = np.asarray(X, dtype=np.float64).reshape(-1)
        o = N * st['MAX_P']
        h[o:o + N] = omega
        h[o + N:o + 2 * N] = self.data.y - 0.5
        h[o + 2 * N:o + 2 * N + W64] = D.pack_mask(mask_bool).view(np.float64)
        st['d_small'].copy_(st['h_small'], non_blocking=True)
        d = st['d_small']
        return (d[:N * p].view(N, p), d[o:o + N], d[o + N:o + 2 * N],
                d[o + 2 * N:o + 2 * N + W64].view(dtype=torch.int64))

    # ---- likelihoods --------------------------------------------------------------------------------
    def likelihood_z_given_rl(self, omega, rl):
        return self.likelihood_z_given_X(omega, self.data.covariate_matrix(rl))

    def likelihood_z_given_X(self, omega, X):
        """logit_rules.py:171-177."""
        return float(self.likelihoods_z_given_Xs(omega, [X])[0])

    def likelihoods_z_given_Xs(self, omega, Xs):
        """Batched form: one launch for a list of design matrices with the same omega."""
        from . import device as D
        N = self.data.n_subjects
        pmax = max(x.shape[1] for x in Xs)
        batch = np.zeros((len(Xs), N, pmax), dtype=np.float64)
        for b, x in enumerate(Xs):
            batch[b, :, :x.shape[1]] = x
        out, status = D.likelihood_z_given_X_batch(batch, omega, self.data.y - 0.5,
                                                   self.prior_coefficient_variance,
                                                   p_b=[x.shape[1] for x in Xs])
        return D.read_with_status(out, status)

    def likelihood_y_given_X_beta(self, X, beta):
        psi = np.dot(X, beta)
        probabilities = 1.0 / (1.0 + np.exp(-1.0 * psi))
        return np.sum(np.log(np.where(self.data.y, probabilities, 1.0 - probabilities)))

    def _sweep_inputs(self, rl, i, j):
        if (i >= len(rl.rules)) or (j >= len(rl.rules[i])):
            raise ValueError('No position (%d,%d) in rule list.' % (i, j))
        n_subjects = self.data.n_subjects
        mask = np.ones(n_subjects, dtype='bool')
        for p in rl.rules[i][:j] + rl.rules[i][j + 1:]:
            mask = np.logical_and(mask, self.data.apply_primitive(p))
        base_rl = rl.copy()
        del base_rl[i]
        X = self.data.covariate_matrix(base_rl)
        return X, mask

    def likelihoods_of_all_options_device(self, rl, current_omega, i, j):
        """Device-resident result (float64[P] CUDA tensor, status): for the fused on-device draw."""
        from . import device as D
        X, mask = self._sweep_inputs(rl, i, j)
        dX, dom, dka, dmask = self._stage_sweep_inputs(X, current_omega, mask)
        st = self._dev()
        pop = self.rule_population
        return D.score_all(pop.packed_truth, st['N'], dX, dom, dka, self.prior_coefficient_variance,
                           mask=dmask, buffers=st['buffers'])

    def likelihoods_of_all_options(self, rl, current_omega, i, j, N_workers=None):
        """logit_rules.py:186-229: host ndarray[P] out.  NOTE: the array is a VIEW of a reused pinned
        staging buffer -- the next call overwrites it (the reference returns a fresh array; copying
        2.35 GB per sweep at S5 would cost more than the sweep).  Callers that keep it must .copy()."""
        import torch
        from . import device as D
        out, status = self.likelihoods_of_all_options_device(rl, current_omega, i, j)
        st = self._dev()
        P = out.shape[0]
        st['h_out'][:P].copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        D.raise_on_status(status)
        return st['h_out'][:P].numpy()

    # ---- priors (logit_rules.py:232-457) ----------------------------------------------------------
    def prior(self, state):
        memo = self.__dict__.get('_step_memo')
        if memo is not None:
            return self._prior_with_fixed_terms(state, memo)
        terms = [self.prior_contribution_hyperpriors(state), self.rule_list_prior(state),
                 self.prior_contribution_coefficients(state)]
        return np.sum(terms)

    def _prior_with_fixed_terms(self, state, memo):
        """prior(state) while the sampler runs a step's four hyperparameter updates: the states that come
        through differ in the phylogeny / window hyperparameters only, so the terms of the rule list's
        structure, its multinomial coefficients, the empty probability and the coefficients are evaluated
        once (the sampler opens and closes the memo, LogisticRuleSampler.step).  Same values summed in the
        same order as prior()."""
        fixed = memo.get('fixed')
        if fixed is None:
            fixed = memo['fixed'] = (
                self.prior_contribution_empty_probability(state),
                self.structure_prior(state.rl, empty_probability=state.empty_probability),
                self.multinomial_prior(state.rl),
                self.prior_contribution_coefficients(state))
        empty_piece, structure_piece, multinomial_piece, coefficient_piece = fixed
        hyperpriors = np.sum([self.prior_contribution_phylogeny_parameters(state),
                              self.prior_contribution_window_parameters(state), empty_piece])
        rule_list_piece = self.prior_contribution_from_primitives(state) + structure_piece + multinomial_piece
        return np.sum([hyperpriors, rule_list_piece, coefficient_piece])

    def prior_contribution_coefficients(self, state):
        dimensions = len(state.beta)
        normalization = -0.5 * dimensions * (np.log(2.0 * np.pi * self.prior_coefficient_variance))
        exponent = (-0.5 * np.dot(state.beta, state.beta) / (self.prior_coefficient_variance))
        return normalization + exponent

    def prior_contribution_hyperpriors(self, state):
        terms = [self.prior_contribution_phylogeny_parameters(state),
                 self.prior_contribution_window_parameters(state),
                 self.prior_contribution_empty_probability(state)]
        return np.sum(terms)

    def prior_contribution_empty_probability(self, state):
        return fastdist.beta_logpdf(state.empty_probability, self.hyperparameter_a_empty,
                                    self.hyperparameter_b_empty)

    def prior_contribution_window_parameters(self, state):
        return fastdist.expon_logpdf(state.window_concentration, self.window_concentration_typical)

    def prior_contribution_phylogeny_parameters(self, state):
        mean_prior = fastdist.norm_logpdf(state.phylogeny_mean, self.phylogeny_lambda_l,
                                          np.sqrt(self.phylogeny_mean_hyperprior_variance))
        if (0. <= state.phylogeny_std and state.phylogeny_std <= self.phylogeny_std_upper_bound):
            std_prior = -1.0 * np.log(self.phylogeny_std_upper_bound)
        else:
            std_prior = -np.inf
        return mean_prior + std_prior

    def rule_list_prior(self, state):
        rule_list = state.rl
        primitive_piece = self.prior_contribution_from_primitives(state)
        structure_piece = self.structure_prior(rule_list, empty_probability=state.empty_probability)
        multinomial_piece = self.multinomial_prior(rule_list)
        return primitive_piece + structure_piece + multinomial_piece

    def structure_prior(self, rule_list, empty_probability):
        """logit_rules.py:337-402."""
        frozen = self.__dict__.get('_structure_terms')
        if frozen is None:      # pure functions of the model constants: freeze once
            length_distribution = fastdist.NegativeBinomial(
                self.hyperparameter_alpha_m, self.hyperparameter_beta_m / (1.0 + self.hyperparameter_beta_m))
            sublength_distribution = fastdist.NegativeBinomial(
                self.hyperparameter_alpha_primitives,
                self.hyperparameter_beta_primitives / (1.0 + self.hyperparameter_beta_primitives))
            # log-pmf tables over the admissible lengths: the same values logpmf returns, looked up
            frozen = (length_distribution.logpmf(np.arange(self.max_rules + 1)),
                      length_distribution.logcdf(self.max_rules - 1),
                      sublength_distribution.logpmf(np.arange(self.max_primitives + 1)),
                      sublength_distribution.logcdf(self.max_primitives - 1))
            self._structure_terms = frozen
        length_logpmf, log_rule_normalization, sublength_logpmf, log_primitive_normalization = frozen
        rules = rule_list.rules
        if len(rules) == 0:
            return np.log(empty_probability)
        if len(rules) > self.max_rules:
            return -np.inf
        sublengths = np.array([len(rule) for rule in rules])
        if np.max(sublengths) > self.max_primitives:
            return -np.inf
        if np.min(sublengths) < 1:      # logpmf(-1) = -inf (never produced by the sampler's moves)
            return -np.inf
        l0 = np.log(1 - empty_probability)
        l1 = (length_logpmf[len(rules) - 1] - log_rule_normalization)
        l2 = (np.sum(sublength_logpmf[sublengths - 1]) -
              len(rules) * log_primitive_normalization)
        return l0 + l1 + l2

    def prior_contribution_from_primitives(self, state):
        if getattr(self, 'fast_prior', False):
            # factored form: no P-long vector (see _prior_tables)
            bv, bw, norm = self._prior_tables(state)
            lookup = self.rule_population._window_lookup
            return sum([(bv[p.variable] + bw[lookup[(float(p.window[0]), float(p.window[1]))]]) - norm
                        for rule in state.rl for p in rule])
        flat_prior = self.flat_prior(state)
        return sum([flat_prior[self.rule_population.get_primitive_index(p)]
                    for rule in state.rl for p in rule])

    def _prior_tables(self, state):
        """flat_prior in factored form: (by_variable[V], by_window[W], normalisation).  A primitive's
        log-prior depends only on its (variable, window); the normalisation is
        log sum_w e^{bw[w]} sum_v C[w, v] e^{bv[v]} with C[w, v] the number of active primitives of
        (window, variable) -- a W x V matrix-vector product instead of a log-sum-exp over the P
        primitives (identical up to summation order)."""
        pop = self.rule_population
        key = (state.phylogeny_mean, state.phylogeny_std, state.window_concentration,
               state.window_typical_fraction, pop._filter_generation, pop.n_primitives, pop._build_generation)
        cache = self.__dict__.setdefault('_prior_table_cache', [])
        for k, v in cache:
            if k == key:
                return v
        log_vw = np.log(np.asarray(self.data.variable_weights, dtype=np.float64))
        by_variable = fastdist.norm_logpdf(log_vw, state.phylogeny_mean, state.phylogeny_std)
        total_duration = float(self.data.experiment_end - self.data.experiment_start)
        durations = pop._durations_by_window / ((1.0 + self.window_duration_epsilon) * total_duration)
        window_a = state.window_concentration * state.window_typical_fraction
        window_b = state.window_concentration * (1.0 - state.window_typical_fraction)
        by_window = fastdist.beta_logpdf(durations, window_a, window_b)
        C = self._prior_count_matrix()
        m_v = float(np.max(by_variable)) if by_variable.size else np.nan
        m_w = float(np.max(by_window)) if by_window.size else np.nan
        if np.isfinite(m_v) and np.isfinite(m_w) and not (np.isnan(by_variable).any() or np.isnan(by_window).any()):
            # C e^{bv - m_v} depends on the phylogeny parameters only: a step's two window proposals reuse it
            # (a [W, V] matrix-vector product, 20 MB at cfg5's 256 x 10 000 -- 0.6 ms on one host thread)
            cx_key = (state.phylogeny_mean, state.phylogeny_std) + key[4:]      # + the population's generations
            cx_cache = self.__dict__.setdefault('_prior_cx_cache', [])
            Cx = next((v for k, v in cx_cache if k == cx_key), None)
            if Cx is None:
                Cx = C.dot(np.exp(by_variable - m_v))
                cx_cache.append((cx_key, Cx))
                del cx_cache[:-4]
            total = float(np.dot(np.exp(by_window - m_w), Cx))
            normalization = m_v + m_w + np.log(total) if total > 0 else 0.0
        else:     # degenerate hyperparameters (e.g. a non-positive scale proposed by the MH update)
            counts = pop.group_counts()
            terms = by_variable[pop._group_variable] + by_window[pop._group_window]
            live = counts > 0
            normalization = fastdist.logsumexp(terms[live], b=counts[live]) if live.any() else 0.0
        result = (by_variable, by_window, float(normalization))
        cache.append((key, result))
        del cache[:-6]
        return result

    def _prior_count_matrix(self):
        """float64[W, V]: active primitives per (window, variable), both detector types together."""
        pop = self.rule_population
        key = (pop._filter_generation, pop.n_primitives, pop._build_generation)
        cached = self.__dict__.get('_prior_count_cache')
        if cached is not None and cached[0] == key:
            return cached[1]
        C = np.zeros((len(pop.windows), self.data.n_variables), dtype=np.float64)
        np.add.at(C, (pop._group_window, pop._group_variable), pop.group_counts().astype(np.float64))
        self._prior_count_cache = (key, C)
        return C

    def flat_prior_device(self, state):
        """flat_prior as a float64[P] CUDA tensor (mitre_row_prior), cached per hyperparameter state."""
        import torch
        from . import device as D
        pop = self.rule_population
        key = (state.phylogeny_mean, state.phylogeny_std, state.window_concentration,
               state.window_typical_fraction, pop._filter_generation, pop.n_primitives, pop._build_generation)
        cached = self.__dict__.get('_prior_device_cache')
        if cached is not None and cached[0] == key:
            return cached[1]
        bv, bw, norm = self._prior_tables(state)
        gvar, gwin = pop.group_tables_device()
        out = cached[1] if cached is not None and cached[1].shape[0] == pop.n_primitives else None
        d = D.row_prior(pop.row_group_device(), gvar, gwin, torch.as_tensor(bv, device=D.dev()),
                        torch.as_tensor(bw, device=D.dev()), norm, out=out)
        self._prior_device_cache = (key, d)
        return d

    def likelihoods_with_stats_device(self, rl, current_omega, i, j, log_prior):
        """One launch: log-likelihoods of all options (device) + (max, sum-exp) of lik + log_prior."""
        from . import device as D
        X, mask = self._sweep_inputs(rl, i, j)
        dX, dom, dka, dmask = self._stage_sweep_inputs(X, current_omega, mask)
        st = self._dev()
        if 'draw' not in st:
            st['draw'] = D.DrawBuffers(max(len(self.rule_population), 1))
        return D.score_all_stats(self.rule_population.packed_truth, st['N'], dX, dom, dka,
                                 self.prior_coefficient_variance, mask=dmask, buffers=st['buffers'],
                                 log_prior=log_prior, draw_buffers=st['draw'])

    def multinomial_prior(self, rule_list):
        return sum([log_multinomial_coefficient(subrule) for subrule in rule_list.rules])

    def flat_prior(self, state):
        """logit_rules.py:424-457.  norm.logpdf / beta.logpdf are elementwise, so evaluating them
        on the V distinct variable weights and W distinct durations and gathering to the P
        primitives gives bit-identical entries; the normalisation is the same scipy logsumexp
        over the same P-vector."""
        pop = self.rule_population
        key = (state.phylogeny_mean, state.phylogeny_std, state.window_concentration,
               state.window_typical_fraction, pop._filter_generation, pop.n_primitives, pop._build_generation)
        cache = self.__dict__.setdefault('_flat_prior_cache', [])
        for k, v in cache:
            if k == key:
                return v
        log_vw = np.log(np.asarray(self.data.variable_weights, dtype=np.float64))
        by_variable = scipy.stats.norm.logpdf(log_vw, loc=state.phylogeny_mean,
                                              scale=state.phylogeny_std)
        total_duration = float(self.data.experiment_end - self.data.experiment_start)
        durations = pop._durations_by_window / ((1.0 + self.window_duration_epsilon) * total_duration)
        window_a = state.window_concentration * state.window_typical_fraction
        window_b = state.window_concentration * (1.0 - state.window_typical_fraction)
        by_window = scipy.stats.beta.logpdf(durations, window_a, window_b)
        groups = pop._flat_groups()
        weights = by_variable[pop._group_variable[groups]] + by_window[pop._group_window[groups]]
        normalization = logsumexp(weights)
        result = weights - normalization
        result.setflags(write=False)
        cache.append((key, result))
        del cache[:-4]
        return result


class LogisticRuleModelPriorOnly(LogisticRuleModel):
    """logit_rules.py:459-477: likelihood == 0, so a chain samples the prior."""

    def likelihood_z_given_X(self, omega, X):
        return 0.

    def likelihoods_z_given_Xs(self, omega, Xs):
        return np.zeros(len(Xs))

    def likelihood_y_given_X_beta(self, X, beta):
        return 0.

    def likelihoods_of_all_options(self, rl, current_omega, i, j, N_workers=default_n_workers):
        return np.zeros(len(self.rule_population))


class LogisticRuleState:
    """logit_rules.py:479-501."""

    def __init__(self, rl, omega, beta, phylogeny_mean, phylogeny_std,
                 window_typical_fraction=0.5, window_concentration=2.0, empty_probability=0.5):
        self.rl = rl
        self.omega = omega
        self.beta = beta
        self.phylogeny_mean = phylogeny_mean
        self.phylogeny_std = phylogeny_std
        self.window_typical_fraction = window_typical_fraction
        self.window_concentration = window_concentration
        self.empty_probability = empty_probability

    def copy(self):
        return LogisticRuleState(self.rl.copy(), self.omega.copy(), self.beta.copy(),
                                 self.phylogeny_mean, self.phylogeny_std,
                                 self.window_typical_fraction, self.window_concentration,
                                 self.empty_probability)


class LogisticRuleSampler(RuleListSampler):
    """logit_rules.py:503-1420."""

    def __init__(self, model, r0, local_updates_per_structure_update=10, device_draw=False,
                 device_gibbs=False):
        """device_draw=False walks the reference's arithmetic exactly (log-likelihoods copied to the
        host, numpy cumsum draw).  device_draw=True keeps the P-long vectors on the GPU: factored
        prior (model._prior_tables / flat_prior_device), log-sum-exp reduced in the scoring kernel,
        proposal drawn by mitre_draw_index -- same np.random consumption, same chain except where a
        uniform draw lands within rounding of a partial sum."""
        self.model = model
        self.device_draw = device_draw
        # device_gibbs=True: the 10 x (omega, beta) Gibbs updates of a step are one kernel launch
        # (mitre_gibbs_omega_beta: Polya-Gamma + posterior draw on the device, Philox stream keyed
        # from np.random) -- same conditionals, different random stream than the host samplers.
        self.device_gibbs = device_gibbs
        if device_draw:
            model.fast_prior = True
        r0 = r0.copy()
        for i in range(len(r0)):
            model.rule_population.sort_list_of_primitives(r0[i])
        omega = np.ones(self.model.data.y.shape)
        phylogeny_mean = self.model.phylogeny_lambda_l
        phylogeny_std = min(self.model.phylogeny_std_upper_bound, 5.0 * self.model.phylogeny_delta_l)
        self.current_state = LogisticRuleState(
            r0, omega, np.zeros(len(r0) + model.data.n_fixed_covariates), phylogeny_mean, phylogeny_std)
        self.current_X = self.model.data.covariate_matrix(r0)
        self.initial_state = self.current_state.copy()
        self.states = []
        self.likelihoods = []
        self.priors = []
        self.comments = []
        self.attempts = {}
        self.successes = {}
        self.local_updates_per_structure_update = local_updates_per_structure_update
        self.additional_beta_values = []
        self.phylogeny_mean_proposal_std = 0.5 * self.model.phylogeny_delta_l
        self.phylogeny_std_proposal_std = 0.5 * self.model.phylogeny_delta_l
        self.window_fraction_proposal_std = self.model.window_fraction_proposal_std
        self.window_concentration_proposal_std = (self.model.window_concentration_typical *
                                                  self.model.window_concentration_update_ratio)

    def __getstate__(self):
        state = dict(self.__dict__)
        state.pop('_gibbs_ws', None)          # device / pinned staging: rebuilt on first use
        state['_current_prior_value'] = None
        return state

    def ensure_current_rl_sorted(self):
        rl = self.current_state.rl
        for i in range(len(rl)):
            self.model.rule_population.sort_list_of_primitives(rl[i])

    def step(self):
        """logit_rules.py:567-596."""
        if self.device_gibbs:
            beta_values = self._device_gibbs_sweeps(self.local_updates_per_structure_update)
        else:
            beta_values = []
            for _ in range(self.local_updates_per_structure_update):
                self.update_omega()
                self.update_beta()
                beta_values.append(self.current_state.beta)
        self.additional_beta_values.append(beta_values)
        move_type_indicator = np.random.rand()
        if move_type_indicator < 0.45:
            self.update_primitives()
        elif move_type_indicator < 0.9:
            self.add_remove()
        else:
            self.move_primitive()
        self.update_beta()
        self.update_empty_probability()
        self._in_step = True
        self.model._step_memo = {}      # the four updates below change hyperparameters only (model.prior)
        try:
            self.update_window_parameters()
            prior = self.update_phylogeny_parameters()
        finally:
            self._in_step = False
            self._current_prior_value = None
            self.model._step_memo = None
        self.states.append(self.current_state.copy())
        self.likelihoods.append(
            self.model.likelihood_y_given_X_beta(self.current_X, self.current_state.beta))
        self.priors.append(prior)

    def update_empty_probability(self):
        if len(self.current_state.rl) > 0:
            conditional_a = self.model.hyperparameter_a_empty
            conditional_b = self.model.hyperparameter_b_empty + 1
        else:
            conditional_a = self.model.hyperparameter_a_empty + 1
            conditional_b = self.model.hyperparameter_b_empty
        self.current_state.empty_probability = scipy.stats.beta.rvs(conditional_a, conditional_b)

    def parameter_mh_update(self, proposed_state):
        """logit_rules.py:615-649."""
        proposed_prior = self.model.prior(proposed_state)
        # the four hyperparameter updates of a step run back to back on a state that only they change, so the
        # prior of the current state is the value the previous update ended with (the reference recomputes it:
        # same function of the same state, same number); update_window_parameters starts the run
        current_prior = self.__dict__.get('_current_prior_value')
        if current_prior is None:
            current_prior = self.model.prior(self.current_state)
        acceptance_probability = np.exp(proposed_prior - current_prior)
        if np.random.rand() < acceptance_probability:
            self.current_state = proposed_state
            self._current_prior_value = proposed_prior
            return proposed_prior
        self._current_prior_value = current_prior
        return current_prior

    def update_window_parameters(self):
        self._current_prior_value = None      # the structure move / beta / empty-probability updates came before
        delta_f = self.window_fraction_proposal_std * np.random.randn()
        current_f = self.current_state.window_typical_fraction
        proposed_state = self.current_state.copy()
        proposed_state.window_typical_fraction = np.abs(((delta_f + current_f + 1.0) % 2.0) - 1.0)
        self.parameter_mh_update(proposed_state)
        delta_cw = self.window_concentration_proposal_std * np.random.randn()
        current_cw = self.current_state.window_concentration
        proposed_state = self.current_state.copy()
        proposed_state.window_concentration = np.abs(delta_cw + current_cw)
        self.parameter_mh_update(proposed_state)

    def update_phylogeny_parameters(self):
        if not self.__dict__.get('_in_step'):
            self._current_prior_value = None  # called on its own: nothing is known about the current state
        delta_mean = self.phylogeny_mean_proposal_std * np.random.randn()
        current_mean = self.current_state.phylogeny_mean
        proposed_state = self.current_state.copy()
        proposed_state.phylogeny_mean = (delta_mean + current_mean)
        self.parameter_mh_update(proposed_state)
        delta_std = self.phylogeny_std_proposal_std * np.random.randn()
        current_std = self.current_state.phylogeny_std
        proposed_std = delta_std + current_std
        upper = self.model.phylogeny_std_upper_bound
        while (proposed_std < 0. or proposed_std > upper):
            proposed_std = np.abs(proposed_std)
            if proposed_std > upper:
                proposed_std = (2.0 * upper - proposed_std)
        proposed_state = self.current_state.copy()
        proposed_state.phylogeny_std = proposed_std
        return self.parameter_mh_update(proposed_state)

    def sample(self, N):
        for i in range(N):
            logger.info('Step %d/%d' % (i, N))
            self.step()

    def sample_for(self, time_in_seconds):
        tstop = time.time() + time_in_seconds
        i = 0
        while time.time() <= tstop:
            logger.info('Step %d' % i)
            self.step()
            i += 1

    def _device_gibbs_sweeps(self, n):
        """n x (update_omega, update_beta) (logit_rules.py:567-573) in one launch."""
        from . import device as D
        state = self.current_state
        p = self.current_X.shape[1]
        beta0 = state.beta if len(state.beta) == p else np.zeros(p)
        seed = int(np.random.randint(0, 2 ** 62))
        ws = self.__dict__.setdefault('_gibbs_ws', {})
        key = (self.current_X.shape[0], p, n)
        if key not in ws:
            ws[key] = D.GibbsWorkspace(*key)
        state.omega, state.beta, betas_h = ws[key].run(self.current_X, self.model.data.y - 0.5,
                                                       self.model.prior_coefficient_variance, beta0, seed)
        return [betas_h[i].copy() for i in range(n)]

    def update_omega(self):
        """logit_rules.py:753-760."""
        state = self.current_state
        psi = np.dot(self.current_X, state.beta)
        ni = np.ones(psi.shape)
        new_omega = np.zeros(state.omega.shape)
        ppg.pgdrawv(ni, psi, new_omega)
        state.omega = new_omega

    def update_beta(self):
        """logit_rules.py:762-787."""
        state = self.current_state
        _, p = self.current_X.shape
        v = self.model.prior_coefficient_variance
        kappa = self.model.data.y - 0.5
        posterior_variance = np.linalg.inv(
            np.eye(p) / v + np.dot(self.current_X.T, (np.dot(np.diag(state.omega), self.current_X))))
        posterior_mean = np.dot(posterior_variance, np.dot(self.current_X.T, kappa))
        state.beta = np.random.multivariate_normal(posterior_mean, posterior_variance)

    # ---- structure moves --------------------------------------------------------------------------
    def update_primitives(self, N_updates=None):
        """logit_rules.py:794-881."""
        population = self.model.rule_population
        primitives = population.flat_rules
        primitive_priors = None if self.device_draw else self.model.flat_prior(self.current_state)
        rule_list = self.current_state.rl
        positions = []
        counts = []
        for i, rule in enumerate(rule_list):
            this_rule_counts = {}
            for j, primitive in enumerate(rule):
                positions.append((i, j))
                k = population.get_primitive_index(primitive)
                this_rule_counts[k] = this_rule_counts.get(k, 0) + 1
            counts.append(this_rule_counts)
        if N_updates is None:
            N_updates = 1
        if len(positions) == 0:
            logger.info('No detectors to update')
            self.comments.append('Null update')
            return
        for _ in range(N_updates):
            choice_index = np.random.choice(len(positions))
            i, j = positions[choice_index]
            primitive = rule_list[i][j]
            k0 = population.get_primitive_index(primitive)
            if not self.device_draw:
                multinomial_priors = np.zeros(len(primitives))
                relative_rates = np.ones(len(primitives))
            base_counts = counts[i].copy()
            if k0 is not None:
                base_counts[k0] -= 1
            else:
                k0 = -1
                logger.warning('Not decrementing base counts of atypical detector')
            for k, n in base_counts.items():
                if self.device_draw:
                    break
                multinomial_priors[k] = -1.0 * gammaln(n + 1 + 1)
                relative_rates[k] += n
            if self.device_draw:
                k = self._device_gibbs_draw(rule_list, i, j, base_counts)
            else:
                likelihoods = self.model.likelihoods_of_all_options(rule_list, self.current_state.omega, i, j)
                posteriors = likelihoods + primitive_priors + multinomial_priors
                weights = posteriors + np.log(relative_rates)
                k = choose_from_relative_probabilities(np.exp(weights))
            rule_list[i][j] = PrimitiveRule(*primitives[k])
            logger.info('At (%d,%d) replacing %d with %d' % (i, j, k0, k))
            self.comments.append('At (%d,%d) replacing %d with %d' % (i, j, k0, k))
            counts[i][k] = counts[i].get(k, 0) + 1
            old_k0_counts = counts[i][k0]
            if old_k0_counts > 1:
                counts[i][k0] -= 1
            else:
                counts[i].pop(k0)
        self.ensure_current_rl_sorted()
        self.current_X = self.model.data.covariate_matrix(rule_list)

    # ---- device-resident weight assembly + draw (device_draw=True) ------------------------------------
    def _with_sparse(self, d_vec, deltas):
        """Context helper: add sparse {index: delta} to a device vector, return the restore info."""
        import torch
        idx = [k for k in deltas if k is not None and k >= 0]
        if not idx:
            return None
        t_idx = torch.as_tensor(idx, device=d_vec.device)
        saved = d_vec[t_idx].clone()
        d_vec[t_idx] = saved + torch.as_tensor([deltas[k] for k in idx], dtype=d_vec.dtype,
                                               device=d_vec.device)
        return (t_idx, saved)

    @staticmethod
    def _restore(d_vec, token):
        if token is not None:
            d_vec[token[0]] = token[1]

    def _device_draw(self, d_a, d_b, deltas=None, status=None):
        """index = #{cumsum(exp(a + b + sparse)) < u * total} with u = np.random.rand()
        (choose_from_relative_probabilities, rules.py:1261-1264) on the device."""
        from . import device as D
        st = self.model._dev()
        if 'draw' not in st:
            st['draw'] = D.DrawBuffers(max(len(self.model.rule_population), 1))
        token = self._with_sparse(d_b, deltas) if deltas else None
        try:
            u = np.random.rand()
            idx = D.draw_index(d_a, d_b, u, stabilise=False, buffers=st['draw'])
            k = int(D.read_with_status(idx, status)[0]) if status is not None else int(idx.item())
        finally:
            self._restore(d_b, token)     # also when the status word raises LinAlgError
        return min(k, d_a.shape[0] - 1)

    def _device_gibbs_draw(self, rule_list, i, j, base_counts):
        """update_primitives' weights = lik + prior + multinomial + log(relative_rates)
        (logit_rules.py:861-862) assembled and drawn on the device."""
        from . import device as D
        d_prior = self.model.flat_prior_device(self.current_state)
        deltas = {k: -1.0 * gammaln(n + 1 + 1) + np.log(1.0 + n) for k, n in base_counts.items()
                  if k is not None and k >= 0}
        # the sparse terms go into the prior BEFORE the sweep: the scoring launch then leaves the per-tile
        # (max, sum-exp) pairs of exactly the weights the draw needs, and the draw makes no pass over them
        token = self._with_sparse(d_prior, deltas) if deltas else None
        try:
            out, status, _stats = self.model.likelihoods_with_stats_device(
                rule_list, self.current_state.omega, i, j, d_prior)
            u = np.random.rand()
            st = self.model._dev()
            idx = D.draw_index(out, d_prior, u, stabilise=False, buffers=st['draw'], reuse_tiles=True)
            k = int(D.read_with_status(idx, status)[0])                   # index and status in one read
        finally:
            self._restore(d_prior, token)
        return min(k, out.shape[0] - 1)

    def add_remove(self):
        if np.random.rand() < 0.5:
            self.add()
        else:
            self.remove()
        self.update_beta()

    def _placeholder(self):
        return PrimitiveRule(*self.model.rule_population.flat_rules[0])

    def consider_addition_to_new_rule(self, base_rule, position_to_add_at):
        """logit_rules.py:894-970."""
        new_rl = base_rule.copy()
        new_rl.rules = (new_rl.rules[:position_to_add_at] + [[self._placeholder()]] +
                        new_rl.rules[position_to_add_at:])
        structure_contributions = [
            self.model.structure_prior(rl, empty_probability=self.current_state.empty_probability)
            for rl in [base_rule, new_rl]]
        log_relative_prior = (structure_contributions[1] - structure_contributions[0])
        old_X = self.model.data.covariate_matrix(base_rule)
        old_likelihood = self.model.likelihood_z_given_X(self.current_state.omega, old_X)
        if self.device_draw:
            from . import device as D
            d_prior = self.model.flat_prior_device(self.current_state)
            out, status, stats = self.model.likelihoods_with_stats_device(
                new_rl, self.current_state.omega, position_to_add_at, 0, d_prior)
            st = D.read_with_status(stats, status)
            log_marginalized_likelihood = st[0] + np.log(st[1])
            primitive_distribution = ('device', out, d_prior, None)
        else:
            priors = self.model.flat_prior(self.current_state)
            likelihoods = self.model.likelihoods_of_all_options(new_rl, self.current_state.omega,
                                                                position_to_add_at, 0)
            primitive_distribution = priors + likelihoods
            log_marginalized_likelihood = logsumexp(primitive_distribution)
        log_likelihood_ratio = (log_marginalized_likelihood - old_likelihood)
        new_total_n_primitives = sum(map(len, new_rl))
        probability_remove_proposal = 1.0 / new_total_n_primitives
        return (new_rl, log_relative_prior, log_likelihood_ratio, primitive_distribution,
                probability_remove_proposal)

    def consider_addition_to_existing_rule(self, base_rule, position_to_add_to):
        """logit_rules.py:972-1095."""
        population = self.model.rule_population
        base_subrule = base_rule[position_to_add_to]
        base_n_istar = len(base_subrule)
        base_rule_counts = {}
        for primitive in base_subrule:
            k = population.get_primitive_index(primitive)
            base_rule_counts[k] = base_rule_counts.get(k, 0) + 1
        if self.device_draw:
            return self._consider_addition_to_existing_rule_device(base_rule, position_to_add_to,
                                                                   base_n_istar, base_rule_counts)
        log_relative_multinomial_coefficient = (np.log(base_n_istar + 1) +
                                                np.zeros(len(population.flat_rules)))
        for k, n in base_rule_counts.items():
            log_relative_multinomial_coefficient[k] += (-1.0 * np.log(n + 1))
        new_rl = base_rule.copy()
        new_rl[position_to_add_to].append(self._placeholder())
        structure_contributions = [
            self.model.structure_prior(rl, empty_probability=self.current_state.empty_probability)
            for rl in [base_rule, new_rl]]
        log_relative_structure_prior = (structure_contributions[1] - structure_contributions[0])
        primitive_prior = self.model.flat_prior(self.current_state)
        log_relative_prior_primitive_piece = logsumexp(log_relative_multinomial_coefficient +
                                                       primitive_prior)
        log_relative_prior = (log_relative_prior_primitive_piece + log_relative_structure_prior)
        old_X = self.model.data.covariate_matrix(base_rule)
        old_likelihood = self.model.likelihood_z_given_X(self.current_state.omega, old_X)
        priors = primitive_prior
        likelihoods = self.model.likelihoods_of_all_options(new_rl, self.current_state.omega,
                                                            position_to_add_to, base_n_istar)
        primitive_distribution = (priors + log_relative_multinomial_coefficient + likelihoods)
        log_marginalized_likelihood = (logsumexp(primitive_distribution) -
                                       log_relative_prior_primitive_piece)
        log_likelihood_ratio = (log_marginalized_likelihood - old_likelihood)
        new_total_n_primitives = sum(map(len, new_rl))
        T_sstar_prefactor = (base_n_istar + 1.0) / (new_total_n_primitives)
        T_sstar_other_factor = np.exp(
            logsumexp(priors + likelihoods) -
            logsumexp(log_relative_multinomial_coefficient + priors + likelihoods))
        probability_remove_proposal = (T_sstar_prefactor * T_sstar_other_factor)
        return (new_rl, log_relative_prior, log_likelihood_ratio, primitive_distribution,
                probability_remove_proposal)

    def _consider_addition_to_existing_rule_device(self, base_rule, position_to_add_to, base_n_istar,
                                                   base_rule_counts):
        """consider_addition_to_existing_rule with every P-long quantity reduced on the device.
        log_relative_multinomial_coefficient = c0 + sparse deltas (logit_rules.py:1005-1013)."""
        from . import device as D
        c0 = np.log(base_n_istar + 1)
        deltas = {k: -1.0 * np.log(n + 1) for k, n in base_rule_counts.items() if k is not None}
        new_rl = base_rule.copy()
        new_rl[position_to_add_to].append(self._placeholder())
        structure_contributions = [
            self.model.structure_prior(rl, empty_probability=self.current_state.empty_probability)
            for rl in [base_rule, new_rl]]
        log_relative_structure_prior = (structure_contributions[1] - structure_contributions[0])
        d_prior = self.model.flat_prior_device(self.current_state)
        st = self.model._dev()
        if 'draw' not in st:
            st['draw'] = D.DrawBuffers(max(len(self.model.rule_population), 1))
        old_X = self.model.data.covariate_matrix(base_rule)
        old_likelihood = self.model.likelihood_z_given_X(self.current_state.omega, old_X)
        # C = logsumexp(prior + lik) with the unmodified prior, in the scoring launch
        out, status, stats = self.model.likelihoods_with_stats_device(
            new_rl, self.current_state.omega, position_to_add_to, base_n_istar, d_prior)
        import torch
        sC_d = stats.clone()
        token = self._with_sparse(d_prior, deltas)
        sA_d = D.weight_stats(d_prior, None, buffers=st['draw']).clone()     # prior + deltas
        sB_d = D.weight_stats(out, d_prior, buffers=st['draw']).clone()      # + lik
        self._restore(d_prior, token)
        three = D.read_with_status(torch.cat([sC_d, sA_d, sB_d]), status)     # one read for all three pairs
        sC, sA, sB = three[0:2], three[2:4], three[4:6]
        lse = lambda s: s[0] + np.log(s[1])
        log_relative_prior_primitive_piece = c0 + lse(sA)
        log_relative_prior = (log_relative_prior_primitive_piece + log_relative_structure_prior)
        log_marginalized_likelihood = (c0 + lse(sB)) - log_relative_prior_primitive_piece
        log_likelihood_ratio = (log_marginalized_likelihood - old_likelihood)
        new_total_n_primitives = sum(map(len, new_rl))
        T_sstar_prefactor = (base_n_istar + 1.0) / (new_total_n_primitives)
        T_sstar_other_factor = np.exp(lse(sC) - (c0 + lse(sB)))
        probability_remove_proposal = (T_sstar_prefactor * T_sstar_other_factor)
        return (new_rl, log_relative_prior, log_likelihood_ratio, ('device', out, d_prior, deltas),
                probability_remove_proposal)

    def _draw_candidate(self, candidate_distribution):
        """(index, PrimitiveRule) drawn from the distribution a consider_* function returned."""
        if isinstance(candidate_distribution, tuple):
            _tag, d_out, d_prior, deltas = candidate_distribution
            k = self._device_draw(d_out, d_prior, deltas)
        else:
            k = choose_from_relative_probabilities(np.exp(candidate_distribution))
        return k, PrimitiveRule(*self.model.rule_population.flat_rules[k])

    def add(self, force_acceptance=False):
        """logit_rules.py:1097-1192."""
        current_rl = self.current_state.rl
        starting_n_rules = len(current_rl)
        option = np.random.randint(2 * starting_n_rules + 1)
        comment = 'add detector to '
        if option < starting_n_rules:
            rule_index = option
            primitive_index = len(current_rl[rule_index])
            comment += 'rule %d' % rule_index
            consider_function = self.consider_addition_to_existing_rule
        else:
            rule_index = option - starting_n_rules
            primitive_index = 0
            comment += 'new rule at position %d' % rule_index
            consider_function = self.consider_addition_to_new_rule
        (new_rl, relative_prior, relative_likelihood, candidate_distribution,
         probability_remove_proposal) = consider_function(current_rl, rule_index)
        probability_this_option = 1. / (2. * len(current_rl) + 1.)
        probability_reverse_option = probability_remove_proposal
        acceptance_ratio = (np.exp(relative_prior + relative_likelihood) *
                            (probability_reverse_option / probability_this_option))
        self.attempts['add'] = self.attempts.get('add', 0) + 1
        if (np.random.rand() < acceptance_ratio) or force_acceptance:
            comment += ': accepted (ratio %.3g)' % acceptance_ratio
            self.successes['add'] = self.successes.get('add', 0) + 1
            k, new_primitive = self._draw_candidate(candidate_distribution)
            comment += ': adds detector %s' % (k,)
            new_rl[rule_index][primitive_index] = new_primitive
            self.current_state.rl = new_rl
            self.ensure_current_rl_sorted()
            self.current_X = self.model.data.covariate_matrix(self.current_state.rl)
        else:
            comment += ': rejected (ratio %.3g)' % acceptance_ratio
        logger.info(comment)
        self.comments.append(comment)

    def remove(self, dry_run=False):
        """logit_rules.py:1194-1300."""
        population = self.model.rule_population
        current_rl = self.current_state.rl
        if len(current_rl) == 0:
            logger.info('Remove block: Nothing to remove.')
            self.comments.append('Nothing to remove')
            return
        positions = [(i, j) for i, subrule in enumerate(current_rl.rules) for j, _ in enumerate(subrule)]
        choice_index = np.random.choice(len(positions))
        rule_index, primitive_index = positions[choice_index]
        comment = 'remove detector %d from rule %d' % (primitive_index, rule_index)
        shorter_rule = current_rl.copy()
        if len(current_rl[rule_index]) == 1:
            consider_function = self.consider_addition_to_new_rule
            del shorter_rule[rule_index]
            copies_removed_primitive_in_rule = 1.0
        else:
            consider_function = self.consider_addition_to_existing_rule
            copies_removed_primitive_in_rule = 0.0
            removed_primitive_as_int = population.get_primitive_index(
                current_rl[rule_index][primitive_index])
            for primitive in current_rl[rule_index]:
                if population.get_primitive_index(primitive) == removed_primitive_as_int:
                    copies_removed_primitive_in_rule += 1.0
            del shorter_rule[rule_index][primitive_index]
        (_, relative_prior, relative_likelihood, _, _) = consider_function(shorter_rule, rule_index)
        relative_prior *= -1.0
        relative_likelihood *= -1.0
        probability_reverse_option = 1. / (2. * len(shorter_rule) + 1.)
        probability_this_option = (copies_removed_primitive_in_rule / float(len(positions)))
        acceptance_ratio = (np.exp(relative_prior + relative_likelihood) *
                            (probability_reverse_option / probability_this_option))
        if dry_run:
            return
        self.attempts['remove'] = self.attempts.get('remove', 0) + 1
        if np.random.rand() < acceptance_ratio:
            comment += ': accepted (ratio %.3g)' % acceptance_ratio
            self.successes['remove'] = self.successes.get('add', 0) + 1   # sic, logit_rules.py:1288
            self.current_state.rl = shorter_rule
            self.current_X = self.model.data.covariate_matrix(shorter_rule)
        else:
            comment += ': rejected (ratio %.3g)' % acceptance_ratio
        logger.info(comment)
        self.comments.append(comment)

    def move_primitive(self):
        """logit_rules.py:1302-1420; the candidate likelihoods are one batched launch."""
        base_rl = self.current_state.rl.copy()
        if len(base_rl) == 0 or (len(base_rl) == 1 and len(base_rl[0]) == 1):
            self.comments.append('Null detector move')
            logger.info('Null detector move')
            return
        primitives = []
        for i, rule in enumerate(base_rl.rules):
            for j, primitive in enumerate(rule):
                primitives.append((i, j, primitive))
        index = np.random.choice(len(primitives))
        i, j, primitive = primitives[index]
        comment_prefix = 'detector (%d,%d) moves to ' % (i, j)
        inverted_primitive = PrimitiveRule(*primitive.as_tuple())
        inverted_primitive.direction = 'below' if primitive.direction == 'above' else 'above'
        del base_rl[i][j]
        if len(base_rl[i]) == 0:
            del base_rl[i]
        comments = []
        rls = []
        for i, rule in enumerate(base_rl.rules):
            for p, comment in [(primitive, 'existing rule %d'),
                               (inverted_primitive, 'existing rule %d (inverted)')]:
                rl = base_rl.copy()
                rl[i].append(p)
                comments.append(comment % i)
                rls.append(rl)
        for i in range(len(base_rl) + 1):
            for p, comment in [(primitive, 'new rule at position %d'),
                               (inverted_primitive, 'new rule at position %d (inverted)')]:
                rl = base_rl.copy()
                rl.rules = rl.rules[:i] + [[p]] + rl.rules[i:]
                comments.append(comment % i)
                rls.append(rl)
        priors = np.zeros(len(rls))
        Xs = []
        for k, rl in enumerate(rls):
            candidate_state = self.current_state.copy()
            candidate_state.rl = rl
            priors[k] = self.model.rule_list_prior(candidate_state)
            Xs.append(self.model.data.covariate_matrix(rl))
        likelihoods = self.model.likelihoods_z_given_Xs(self.current_state.omega, Xs)
        k = choose_from_relative_probabilities(np.exp(priors + likelihoods))
        self.current_state.rl = rls[k]
        self.ensure_current_rl_sorted()
        self.current_X = self.model.data.covariate_matrix(self.current_state.rl)
        self.comments.append(comment_prefix + comments[k])
        logger.info(comment_prefix + comments[k])


def probabilities(rule_list, beta, test_data):
    X = test_data.covariate_matrix(rule_list)
    psi = np.dot(X, beta)
    return 1.0 / (1.0 + np.exp(-1.0 * psi))


def prediction(rule_list, beta, test_data):
    return probabilities(rule_list, beta, test_data) >= 0.5
