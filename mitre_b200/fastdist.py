"""The handful of scipy.stats / scipy.special evaluations the sampler makes per step
(logit_rules.py:232-457 of the reference), restated without scipy's generic argument machinery.

Every MCMC step evaluates the model prior ~8 times in its hyperparameter updates
(logit_rules.py:476-528), each through `scipy.stats.<dist>.logpdf` on scalars or short vectors;
at ~70 us of argument checking per call that was three quarters of a step once the P-long work
had moved to the device.  The functions below evaluate the SAME formulas scipy 1.x uses, in the
same order of operations, so the results are bit-identical (tests/test_host_cpu.py checks that
against the installed scipy) and the chain's accept/reject sequence does not move.
"""
import math

import numpy as np
from scipy import special as sc

_NORM_PDF_LOGC = np.log(np.sqrt(2 * np.pi))      # scipy.stats._continuous_distns._norm_pdf_logC


def norm_logpdf(x, loc, scale):
    """scipy.stats.norm.logpdf(x, loc=loc, scale=scale)."""
    if not scale > 0:
        return np.full(np.shape(x), np.nan) if np.ndim(x) else np.nan
    scalar = np.ndim(x) == 0
    y = (np.atleast_1d(np.asarray(x, dtype=np.float64)) - loc) / scale   # arrays: y**2 is a multiply,
    out = (-y ** 2 / 2.0 - _NORM_PDF_LOGC) - np.log(scale)                # numpy scalars would call pow()
    return out[0] if scalar else out


def expon_logpdf(x, scale):
    """scipy.stats.expon.logpdf(x, scale=scale) for scalar x."""
    if not scale > 0 or x != x:
        return np.nan
    y = x / scale
    if y < 0:
        return -np.inf
    return -y - np.log(scale)


def beta_logpdf(x, a, b):
    """scipy.stats.beta.logpdf(x, a, b) (loc 0, scale 1)."""
    x = np.asarray(x, dtype=np.float64)
    if not (a > 0 and b > 0):
        out = np.full(x.shape, np.nan)
        return out if out.ndim else out[()]
    inside = (x >= 0) & (x <= 1)
    with np.errstate(divide='ignore', invalid='ignore'):
        lpx = sc.xlog1py(b - 1.0, -x) + sc.xlogy(a - 1.0, x)
        lpx = lpx - sc.betaln(a, b)
    out = np.where(inside, lpx, -np.inf)
    out = np.where(np.isnan(x), np.nan, out)
    return out if out.ndim else out[()]


class NegativeBinomial:
    """Frozen scipy.stats.nbinom(n, p): logpmf on non-negative integers, and the log-cdf constant
    the structure prior needs (computed once through scipy itself)."""

    def __init__(self, n, p):
        from scipy.stats import nbinom
        self.n, self.p = float(n), float(p)
        self._frozen = nbinom(n, p)
        self._n_log_p = self.n * np.log(self.p)

    def logcdf(self, k):
        return self._frozen.logcdf(k)

    def logpmf(self, k):
        k = np.asarray(k, dtype=np.float64)
        coeff = sc.gammaln(self.n + k) - sc.gammaln(k + 1) - sc.gammaln(self.n)
        out = coeff + self._n_log_p + sc.xlog1py(k, -self.p)
        out = np.where((k >= 0) & (k == np.floor(k)), out, -np.inf)
        return out if out.ndim else out[()]


def logsumexp(a, b=None):
    """scipy.special.logsumexp(a, b=b) for real 1-D input: the maximal terms are taken out of the
    sum, the rest enters through log1p (scipy/special/_logsumexp.py:_logsumexp)."""
    a = np.array(a, dtype=np.float64, copy=True).ravel()
    if a.size == 0:
        return -np.inf
    if b is not None:
        b = np.asarray(b, dtype=np.float64).ravel()
    with np.errstate(divide='ignore', invalid='ignore', over='ignore'):
        direct = np.log(np.sum(np.exp(a) if b is None else b * np.exp(a)))
        if b is not None:
            a[b == 0] = -np.inf
        a_max = np.max(a)
        i_max = a == a_max
        a[i_max] = -np.inf
        m = np.sum(i_max.astype(np.float64) if b is None else b * i_max.astype(np.float64))
        e = np.exp(a - a_max)
        if b is not None:
            e = b * e
        s = np.sum(e)
        if s != 0:
            s = s / m
        sgn = np.sign(s + 1) * np.sign(m)
        if s < -1:
            s = -s - 2
        m = abs(m)
        out = np.log1p(s) + np.log(m) + a_max
        if sgn < 0:
            out = np.nan
    return out if math.isfinite(out) else direct
