"""Host helper identical in interface to smcpy/proposals.py:4-34 (initial-particle proposal
used with GeometricPath(proposal=...)); sampling and its log-pdf stay on the host, the
resulting (N,1) log q column is uploaded once per SMC step and consumed by smcb_reweight."""

import numpy as np


class MultivarIndependent:
    def __init__(self, *args):
        self._dist_list = args
        self._dims = [d.rvs(1).size for d in self._dist_list]

    def rvs(self, num_samples, random_state=None):
        return np.hstack([np.c_[d.rvs(num_samples, random_state=random_state)] for d in self._dist_list])

    def logpdf(self, inputs):
        parts = np.split(inputs, np.cumsum(self._dims)[:-1], axis=1)
        cols = np.hstack([np.asarray(d.logpdf(x)).reshape(inputs.shape[0], -1)
                          for d, x in zip(self._dist_list, parts)])
        return np.sum(cols, axis=1, keepdims=True)


def describe_proposal(proposal):
    """Per-column (kind, loc, scale) table for the device log q kernel, or None when the
    proposal is not a MultivarIndependent of scalar scipy.stats norm / uniform components
    (then only the reweighting can use it, through its host logpdf)."""
    from . import _lib
    dists = getattr(proposal, "_dist_list", None)
    if dists is None:
        return None
    cols = []
    for dist in dists:
        name = getattr(getattr(dist, "dist", None), "name", None)
        args = list(getattr(dist, "args", ()))
        kw = dict(getattr(dist, "kwds", {}))
        loc = float(args[0]) if len(args) > 0 else float(kw.get("loc", 0.0))
        scale = float(args[1]) if len(args) > 1 else float(kw.get("scale", 1.0))
        if name == "norm":
            cols.append((_lib.PROPOSAL_NORMAL, loc, scale))
        elif name == "uniform":
            cols.append((_lib.PROPOSAL_UNIFORM, loc, scale))
        else:
            return None
    return cols
