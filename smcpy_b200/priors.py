"""
Priors.  Same constructors and semantics as smcpy/priors.py.
  * device log-pdf (prior table of the fused kernels): scipy.stats.uniform frozen objects (accepted as
    they are; examples use them), ImproperUniform (:6-35), ImproperCov (:93-147);
  * host log-pdf (SMCB_PRIOR_EXTERNAL: evaluated by the object itself on the host and added to the
    log-prior column of the un-fused Metropolis route): InvWishart (:38-90), ImproperConstrainedUniform
    (:150-247) and any other object with `logpdf(x (N, dim))` / optional `dim`, `rvs`.
`rvs` stays on the host (one-time; replay parity needs scipy's stream); the native sampler draws
uniform and ImproperCov priors on the device (smcb_sample_priors).
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


class InvWishart:
    """Inverse-Wishart prior over a covariance matrix given as packed upper-triangular row-major columns
    (smcpy/priors.py:38-90).  Host-evaluated through scipy.stats.invwishart, as in the reference."""

    def __init__(self, dof, scale):
        from scipy.stats import invwishart
        scale = np.asarray(scale, dtype=np.float64)
        self._p = int(scale.shape[0])
        self._dim = self._p * (self._p + 1) // 2
        self._dist = invwishart(dof, scale)

    @property
    def dim(self):
        return self._dim

    def _matrices(self, x):
        """(N, dim) packed rows -> (p, p, N) symmetric matrices (scipy's batch layout)."""
        x = np.asarray(x, dtype=np.float64)
        r, c = np.triu_indices(self._p)
        m = np.zeros((self._p, self._p, x.shape[0]))
        m[r, c, :] = x.T
        m[c, r, :] = x.T
        return m

    def rvs(self, num_samples, random_state=None):
        cov = self._dist.rvs(num_samples, random_state=random_state)
        cov = cov.reshape(-1, self._p, self._p)                # scipy drops the batch axis for one sample
        r, c = np.triu_indices(self._p)
        return cov[:, r, c]

    def pdf(self, x):
        try:
            return np.asarray(self._dist.pdf(self._matrices(x))).reshape(-1, 1)
        except np.linalg.LinAlgError:                          # any matrix of the batch not PD
            return np.zeros((np.asarray(x).shape[0], 1))

    def logpdf(self, x):
        try:
            return np.asarray(self._dist.logpdf(self._matrices(x))).reshape(-1, 1)
        except np.linalg.LinAlgError:
            return np.full((np.asarray(x).shape[0], 1), -np.inf)


class ImproperConstrainedUniform:
    """Improper uniform prior over `dim` parameters restricted by a user constraint and optional bounds
    (smcpy/priors.py:150-247): log-pdf 0 where constraint_function(rows) is true and the row lies inside the
    inclusive bounds, -inf elsewhere; rvs by rejection inside finite bounds."""

    def __init__(self, constraint_function, dim=None, bounds=None, max_rvs_tries=1000):
        if dim is None and bounds is None:
            raise ValueError('Either "bounds" or "dim" must be provided.')
        self._constraint = constraint_function
        self._bounds = None if bounds is None else np.asarray(bounds, dtype=np.float64)
        self._dim = dim
        self._max_tries = max_rvs_tries

    @property
    def dim(self):
        return self._dim if self._dim else self._bounds.shape[1]

    def logpdf(self, samples):
        samples = np.asarray(samples, dtype=np.float64)
        ok = np.asarray(self._constraint(samples)).astype(bool).reshape(-1)
        if self._bounds is not None:
            ok &= (self._bounds[0] <= samples).all(axis=1) & (self._bounds[1] >= samples).all(axis=1)
        return np.where(ok, 0.0, -np.inf).reshape(-1, 1)

    def rvs(self, num_samples, random_state=None):
        if random_state is not None and not isinstance(random_state, np.random.Generator):
            raise TypeError("Random number generator must be a numpy generator.")
        if self._bounds is None or not np.isfinite(self._bounds).all():
            raise NotImplementedError("rvs cannot be used without finite bounds.")
        rng = np.random.default_rng() if random_state is None else random_state
        kept, tries = [], 0
        while len(kept) < num_samples and tries < self._max_tries:
            cand = rng.uniform(*self._bounds, (num_samples, self.dim))
            kept += list(cand[self._constraint(cand)])
            tries += 1
        if len(kept) < num_samples:
            raise ValueError(f"Max rvs tries exceeded; {len(kept)}/{num_samples} generated")
        return np.array(kept)[:num_samples]


def prior_dim(p):
    """smcpy/mcmc/vector_mcmc.py:113-114."""
    return p.dim if hasattr(p, "dim") else 1


def describe_prior(p):
    """-> (kind, ncols, a, b, logp_in) for the device prior table.  Objects without a device log-pdf
    (InvWishart, ImproperConstrainedUniform, the reference's own classes, user classes) become
    SMCB_PRIOR_EXTERNAL columns: flat on the device, evaluated by the object on the host."""
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
        return (_lib.PRIOR_IMPROPER_COV, int(p._ncols), float(p._iw_dof), 0.0, 0.0)
    if callable(getattr(p, "logpdf", None)):
        return (_lib.PRIOR_EXTERNAL, int(prior_dim(p)), 0.0, 0.0, 0.0)
    raise TypeError("prior %r has neither a device log-pdf (scipy.stats.uniform, ImproperUniform, ImproperCov) "
                    "nor a logpdf method" % (p,))
