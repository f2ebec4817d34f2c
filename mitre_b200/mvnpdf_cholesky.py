"""Drop-in replacement for the reference module `mvnpdf_cholesky` (only `mc_logpdf_np` is used
by the reference: logit_rules.py:12, 177).

    mc_logpdf_np(x, cov) -> float     log N(x; 0, cov) through a Cholesky factorisation

mvnpdf_cholesky.py:53-74.  Runs mitre_mc_logpdf (single-CTA device Cholesky).  The model-level
dense likelihood does not come through here: it never forms the N x N covariance
(mitre_likelihood_z_given_X_batch).
"""
import numpy as np


def mc_logpdf_np(x, cov):
    from . import device as D
    out, status = D.mc_logpdf(np.asarray(x, dtype=np.float64), np.asarray(cov, dtype=np.float64))
    val = float(out.item())
    D.raise_on_status(status)
    return val
