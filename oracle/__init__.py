"""oracle/ -- CPU restatement of the reference's hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package; the product (mitre_b200/) never does and fails loudly without its CUDA
library.

  scoring_oracle.c      C restatement of efficient_likelihoods.pyx / mvnpdf_cholesky.py
  detectors_oracle.py   numpy restatement of rules.py's detector evaluation
  build_ref.py          builds the UNMODIFIED reference .pyx into oracle/_ref/ (git-ignored)
  ref_shim.py           imports the reference's rules.py under py3 (build container only)

Parity pin: pinned (tests/test_oracle.py) against oracle/_ref (the reference's own Cython
kernel compiled here), the reference's rules.py run through ref_shim, the internal identity
likelihoods_of_all_options[k] == likelihood_z_given_X([X | v_k]) the reference relies on
(logit_rules.py:944-961), and the golden vectors under tests/golden/.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "scoring_oracle.c")
_LIB = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force=False):
    """gcc the C restatement into oracle/_build/liboracle.so."""
    if (not force and os.path.exists(_LIB)
            and (not os.path.exists(_SRC) or os.path.getmtime(_LIB) >= os.path.getmtime(_SRC))):
        return _LIB
    os.makedirs(os.path.dirname(_LIB), exist_ok=True)
    subprocess.check_call(["gcc", "-O2", "-std=c99", "-D_GNU_SOURCE", "-ffp-contract=off",
                           "-shared", "-fPIC", _SRC, "-o", _LIB, "-lm"])
    return _LIB


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        dp = ctypes.POINTER(ctypes.c_double)
        L.oracle_likelihoods_of_all_options.argtypes = [dp, dp, dp, dp, ctypes.c_double,
                                                        ctypes.c_int, ctypes.c_int,
                                                        ctypes.c_int64, dp]
        L.oracle_likelihoods_of_all_options.restype = ctypes.c_int
        L.oracle_mc_logpdf.argtypes = [dp, dp, ctypes.c_int, dp]
        L.oracle_mc_logpdf.restype = ctypes.c_int
        L.oracle_likelihood_z_given_X.argtypes = [dp, dp, dp, ctypes.c_double, ctypes.c_int,
                                                  ctypes.c_int, dp]
        L.oracle_likelihood_z_given_X.restype = ctypes.c_int
        L.oracle_choose.argtypes = [dp, ctypes.c_int64, ctypes.c_double]
        L.oracle_choose.restype = ctypes.c_int64
        _lib = L
    return _lib


def _c(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def likelihoods_of_all_options(X, omega, response, list_of_truth_vectors,
                               prior_coefficient_variance):
    """Same signature as the reference extension (efficient_likelihoods.pyx:15-21)."""
    X, pX = _c(X)
    omega, pw = _c(omega)
    response, pz = _c(response)
    tv, pt = _c(list_of_truth_vectors)
    n_opt, k = tv.shape
    out = np.zeros(n_opt)
    info = lib().oracle_likelihoods_of_all_options(
        pX, pw, pz, pt, float(prior_coefficient_variance), k, X.shape[1], n_opt,
        out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
    if info:
        raise np.linalg.LinAlgError("Matrix is not positive definite")
    return out


def mc_logpdf_np(x, cov):
    """mvnpdf_cholesky.py:53-74."""
    x, px = _c(x)
    cov, pc = _c(cov)
    out = ctypes.c_double(0.0)
    info = lib().oracle_mc_logpdf(px, pc, len(x), ctypes.byref(out))
    if info:
        raise np.linalg.LinAlgError("Matrix is not positive definite")
    return out.value


def likelihood_z_given_X(omega, X, y, prior_coefficient_variance):
    """logit_rules.py:171-177 (y: boolean outcomes, kappa = y - 0.5)."""
    X, pX = _c(X)
    omega, pw = _c(omega)
    kappa, pk = _c(np.asarray(y, dtype=np.float64) - 0.5)
    out = ctypes.c_double(0.0)
    info = lib().oracle_likelihood_z_given_X(pX, pw, pk, float(prior_coefficient_variance),
                                             X.shape[0], X.shape[1], ctypes.byref(out))
    if info:
        raise np.linalg.LinAlgError("Matrix is not positive definite")
    return out.value


def choose_from_relative_probabilities(p, u):
    """rules.py:1261-1264 with the np.random.rand() draw passed in as u."""
    p, pp = _c(p)
    return int(lib().oracle_choose(pp, len(p), float(u)))


def mc_logpdf_np_numpy(x, cov):
    """numpy/scipy restatement of mvnpdf_cholesky.py:65-74 (same library calls as the
    reference: np.linalg.cholesky + scipy solve_triangular)."""
    import scipy.linalg
    k = cov.shape[0]
    R = np.linalg.cholesky(cov)
    log_det_cov = np.sum(np.log(R.diagonal())) * 2.0
    y = scipy.linalg.solve_triangular(R, x, lower=True, check_finite=False)
    exponent = -0.5 * np.sum(np.square(y))
    return exponent - 0.5 * (k * np.log(2.0 * np.pi) + log_det_cov)
