"""Drop-in replacement for the reference extension module `efficient_likelihoods`
(mitre/efficient_likelihoods.pyx, imported at logit_rules.py:17).

    likelihoods_of_all_options(X, omega, response, list_of_truth_vectors,
                               prior_coefficient_variance) -> float64[P]

Same signature and return value as efficient_likelihoods.pyx:15-21,169; host ndarrays in, host
ndarray out.  The work is done by mitre_score_all_host in libmitre_b200.so (rows packed to bits on
the host, 8 bytes per candidate over PCIe, scored on the GPU).  A failed Cholesky raises numpy.linalg.LinAlgError like np.linalg.cholesky does at
pyx:42.  The model-level path (mitre_b200.logit_rules) never builds the float64[P, N] array this
signature requires; it scores the device-resident packed table directly.
"""
import ctypes

import numpy as np

from ._lib import MITRE_ERR_NOT_PD, check, lib


def _dp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def likelihoods_of_all_options(X, omega, response, list_of_truth_vectors,
                               prior_coefficient_variance):
    X = np.ascontiguousarray(X, dtype=np.float64)
    omega = np.ascontiguousarray(omega, dtype=np.float64)
    response = np.ascontiguousarray(response, dtype=np.float64)
    tv = np.ascontiguousarray(list_of_truth_vectors, dtype=np.float64)
    if X.ndim != 2 or tv.ndim != 2 or omega.ndim != 1 or response.ndim != 1:
        raise ValueError('Buffer has wrong number of dimensions')
    N, p = X.shape
    P = tv.shape[0]
    if tv.shape[1] != N or omega.shape[0] != N or response.shape[0] != N:
        raise ValueError('inconsistent shapes')
    out = np.empty(P, dtype=np.float64)
    rc = lib().mitre_score_all_host(_dp(X), _dp(omega), _dp(response), _dp(tv),
                                    float(prior_coefficient_variance), N, p, P, _dp(out))
    if rc == MITRE_ERR_NOT_PD:
        raise np.linalg.LinAlgError('Matrix is not positive definite')
    check(rc, 'mitre_score_all_host')
    return out
