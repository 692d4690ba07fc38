"""
Priors with a device log-pdf.  Same constructors and semantics as smcpy/priors.py
(ImproperUniform :6-35, ImproperCov :93-147); scipy.stats.uniform frozen objects are
accepted as they are (examples use them).  `logpdf` runs on the GPU through
smcb_mh_init's prior table; `rvs` stays on the host (one-time, needs scipy's stream
for replay parity) -- the native sampler draws uniform priors on the device instead.
"""

import math

import numpy as np

from . import _lib


class ImproperUniform:
    """smcpy/priors.py:6-35: log-pdf 0 inside the inclusive bounds, -inf outside."""

    def __init__(self, lower_bound=None, upper_bound=None):
        self._lower = -np.inf if lower_bound is None else lower_bound
        self._upper = np.inf if upper_bound is None else upper_bound

    def logpdf(self, x):
        array_x = np.array(x).squeeze()
        if array_x.ndim > 1:
            raise ValueError("Input array must be 1D or must squeeze to 1D")
        from .vector_mcmc import device_log_priors
        return device_log_priors([self], np.atleast_1d(array_x).reshape(-1, 1))[:, 0]


class ImproperCov:
    """smcpy/priors.py:93-147: improper uniform over positive-definite covariance matrices;
    rvs draws inverse-Wishart samples (host)."""

    def __init__(self, num_cov_cols, dof=None, S=None):
        self._ncols = num_cov_cols
        self._dim = int(num_cov_cols * (num_cov_cols + 1) / 2)
        self._iw_dof = dof if dof is not None else num_cov_cols
        self._iw_scale = S if S is not None else np.eye(self._ncols)

    @property
    def dim(self):
        return self._dim

    def logpdf(self, samples):
        samples = np.asarray(samples, dtype=float)
        if len(samples.shape) != 2 or samples.shape[1] != self._dim:
            raise ValueError("Covariance samples must be a 2D array (num samples, %d packed "
                             "upper-triangular terms)" % self._dim)
        from .vector_mcmc import device_log_priors
        return device_log_priors([self], samples)

    def rvs(self, num_samples, random_state=None):
        from scipy.stats import invwishart
        cov = invwishart(self._iw_dof, self._iw_scale).rvs(num_samples, random_state=random_state)
        idx1, idx2 = np.triu_indices(self._ncols)
        return cov[:, idx1, idx2]


def prior_dim(p):
    """smcpy/mcmc/vector_mcmc.py:113-114."""
    return p.dim if hasattr(p, "dim") else 1


def describe_prior(p):
    """-> (kind, ncols, a, b, logp_in) for the device prior table, or raises TypeError for
    priors without a device log-pdf (InvWishart, constrained callables: use the host
    reference classes with the torch-callable route)."""
    name = type(p).__name__
    dist = getattr(p, "dist", None)
    if dist is not None and getattr(dist, "name", None) == "uniform":        # scipy frozen uniform
        args = list(getattr(p, "args", ()))
        kw = dict(getattr(p, "kwds", {}))
        loc = float(args[0]) if len(args) > 0 else float(kw.get("loc", 0.0))
        scale = float(args[1]) if len(args) > 1 else float(kw.get("scale", 1.0))
        if not scale > 0:
            raise ValueError("uniform prior needs scale > 0")
        return (_lib.PRIOR_UNIFORM, 1, loc, scale, -math.log(scale))
    if name == "ImproperUniform":
        return (_lib.PRIOR_IMPROPER_UNIFORM, 1, float(p._lower), float(p._upper), 0.0)
    if name == "ImproperCov":
        if p._ncols > 3:
            raise NotImplementedError("ImproperCov on the device supports up to 3x3 matrices")
        return (_lib.PRIOR_IMPROPER_COV, int(p._ncols), 0.0, 0.0, 0.0)
    raise TypeError("prior %r has no device log-pdf (supported: scipy.stats.uniform, "
                    "ImproperUniform, ImproperCov)" % (p,))
