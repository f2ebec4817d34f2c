"""The slice of the reference's PosteriorSummary (mitre/posterior.py) that defines the final
parity observable: burn-in selection (posterior.py:60-71), the rule-list length tabulation
(posterior.py:163-275, lengths only) and `point_summarize` (posterior.py:492-526).  Reports,
Bayes factors, clustering and plots are outside the hot path and not rebuilt."""
import numpy as np


class PosteriorSummary:
    def __init__(self, sampler, burnin_fraction=0.1, tag=None, primitive_prior=None):
        self.sampler = sampler
        self.model = sampler.model
        self.tag = '' if tag is None else tag + '_'
        self.primitive_prior = primitive_prior
        self.burnin = int(burnin_fraction * len(sampler.states))
        # off-by-one between states and the beta samples recorded before each structure update
        # (posterior.py:62-67)
        self.state_ensemble = sampler.states[int(self.burnin):-1]
        self.ensemble_beta_samples = sampler.additional_beta_values[int(self.burnin) + 1:]
        self.ensemble_likelihoods = np.array(sampler.likelihoods[int(self.burnin):-1])
        self.ensemble_priors = np.array(sampler.priors[int(self.burnin):-1])
        self.tabulate()
        self.reports = {}

    def tabulate(self):
        self.n_states = len(self.state_ensemble)
        self.rl_lengths = np.array([len(s.rl) for s in self.state_ensemble])
        self.rl_total_primitives = np.array([sum(map(len, s.rl)) for s in self.state_ensemble])
        n = float(self.n_states)
        self.rl_length_distribution = {
            d: int(np.sum(self.rl_lengths == d)) for d in range(int(max(self.rl_lengths)) + 1)}
        self.rl_total_primitive_distribution = {
            d: int(np.sum(self.rl_total_primitives == d))
            for d in range(int(max(self.rl_total_primitives)) + 1)}
        self.normalized_rl_total_primitive_distribution = {
            d: c / n for d, c in self.rl_total_primitive_distribution.items()}
        self.rl_length_posterior = sorted(
            [(v / n, k) for k, v in self.rl_length_distribution.items()], reverse=True)
        self.rl_total_primitive_posterior = sorted(
            [(v / n, k) for k, v in self.rl_total_primitive_distribution.items()], reverse=True)
        self.modal_rl_length = self.rl_length_posterior[0][1]
        self.modal_rl_total_primitives = self.rl_total_primitive_posterior[0][1]
        if self.modal_rl_length == 0 and len(self.rl_length_posterior) > 1:
            self.modal_nonzero_rl_length = self.rl_length_posterior[1][1]
        else:
            self.modal_nonzero_rl_length = self.modal_rl_length
        if self.modal_rl_total_primitives == 0 and len(self.rl_total_primitive_posterior) > 1:
            self.modal_nonzero_rl_total_primitives = self.rl_total_primitive_posterior[1][1]
        else:
            self.modal_nonzero_rl_total_primitives = self.modal_rl_total_primitives

    def point_summarize(self):
        """posterior.py:492-526: the sampled state of highest prior+likelihood among those with
        the modal total number of primitives (modal non-zero if P(empty) < 0.5)."""
        if self.normalized_rl_total_primitive_distribution[0] < 0.5:
            target_length = self.modal_nonzero_rl_total_primitives
        else:
            target_length = self.modal_rl_total_primitives
        states, posteriors, betas = [], [], []
        for state, prior, likelihood, beta in zip(self.state_ensemble, self.ensemble_priors,
                                                  self.ensemble_likelihoods,
                                                  self.ensemble_beta_samples):
            if sum(map(len, state.rl)) == target_length:
                states.append(state)
                posteriors.append(prior + likelihood)
                betas.append(beta)
        best_index = np.argmax(posteriors)
        self.point = states[best_index]
        self.point_betas = np.vstack(betas[best_index])
