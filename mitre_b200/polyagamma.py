"""Polya-Gamma PG(b, z) sampler with the pypolyagamma call surface the reference uses
(logit_rules.py:13, 62, 759: `PyPolyaGamma(seed)`, `pgdrawv(n, z, out)`).

pypolyagamma is a third-party dependency of the reference that is neither vendored nor pinned
(setup.py:57) and is absent from this image, so whole-chain RNG parity is defined with THIS
sampler plugged into both the reference sampler and the mitre_b200 carrier
(tests/golden/make_golden_chain.py).  omega is an *input* of the likelihood hot path.

Algorithm: Devroye's exact alternating-series sampler for J*(1, z/2) (Polson, Scott & Windle
2013, "Bayesian inference for logistic models using Polya-Gamma latent variables", section 4),
PG(1, z) = J*(1, z/2) / 4; PG(b, z) for integer b is the sum of b independent PG(1, z) draws.
"""
import math

import numpy as np

_TRUNC = 0.64
_SQRT2 = math.sqrt(2.0)


def _pnorm(x):
    return 0.5 * math.erfc(-x / _SQRT2)


def _a_coef(n, x):
    k = (n + 0.5) * math.pi
    if x > _TRUNC:
        return k * math.exp(-0.5 * k * k * x)
    if x > 0.0:
        expnt = -1.5 * (math.log(0.5 * math.pi) + math.log(x)) + math.log(k) - 2.0 * (n + 0.5) ** 2 / x
        return math.exp(expnt)
    return 0.0


class PyPolyaGamma:
    def __init__(self, seed=0):
        self.rng = np.random.RandomState(seed)

    def _rtigauss(self, z):
        """Inverse Gaussian(1/z, 1) truncated to (0, TRUNC)."""
        rng = self.rng
        t = _TRUNC
        if z < 1.0 / t:     # mean above the truncation point: chi-square tail proposal
            while True:
                while True:
                    e1 = rng.exponential()
                    e2 = rng.exponential()
                    if e1 * e1 <= 2.0 * e2 / t:
                        break
                x = t / (1.0 + t * e1) ** 2
                if rng.uniform() <= math.exp(-0.5 * z * z * x):
                    return x
        mu = 1.0 / z
        while True:
            y = rng.normal()
            y *= y
            half_mu = 0.5 * mu
            mu_y = mu * y
            x = mu + half_mu * mu_y - half_mu * math.sqrt(4.0 * mu_y + mu_y * mu_y)
            if rng.uniform() > mu / (mu + x):
                x = mu * mu / x
            if x <= t:
                return x

    def _mass_texpon(self, z):
        """Probability of proposing from the exponential (right of TRUNC) branch."""
        t = _TRUNC
        K = 0.125 * math.pi * math.pi + 0.5 * z * z
        w = math.sqrt(1.0 / t)
        x0 = math.log(K) + K * t
        pb = _pnorm(w * (t * z - 1.0))
        pa = _pnorm(-w * (t * z + 1.0))
        xb = x0 - z + (math.log(pb) if pb > 0.0 else -math.inf)
        xa = x0 + z + (math.log(pa) if pa > 0.0 else -math.inf)
        # exp(xb) overflows a double from |psi| ~ 98 on (xb ~ 0.32 z^2 - z); BayesLogit / pypolyagamma
        # get inf there and a ratio of 0, so do we
        m = max(xb, xa)
        if m > 700.0:
            return 0.0
        qdivp = 4.0 / math.pi * (math.exp(xb) + math.exp(xa))
        return 1.0 / (1.0 + qdivp)

    def _draw_one(self, z):
        rng = self.rng
        z = abs(z) * 0.5
        K = 0.125 * math.pi * math.pi + 0.5 * z * z
        ratio = self._mass_texpon(z)
        while True:
            if rng.uniform() < ratio:
                x = _TRUNC + rng.exponential() / K
            else:
                x = self._rtigauss(z)
            s = _a_coef(0, x)
            y = rng.uniform() * s
            n = 0
            while True:
                n += 1
                if n % 2 == 1:
                    s -= _a_coef(n, x)
                    if y <= s:
                        return 0.25 * x
                else:
                    s += _a_coef(n, x)
                    if y > s:
                        break

    def pgdraw(self, n, z):
        total = 0.0
        for _ in range(int(n)):
            total += self._draw_one(float(z))
        return total

    def pgdrawv(self, n, z, out):
        """out[i] ~ PG(n[i], z[i]) (in place, like pypolyagamma's vector API)."""
        for i in range(len(out)):
            out[i] = self.pgdraw(n[i], z[i])
