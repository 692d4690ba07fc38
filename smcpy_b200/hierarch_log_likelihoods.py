"""
Approximate hierarchical likelihood (smcpy/hierarch_log_likelihoods.py:6-100) as torch callables on device
tensors: the likelihood of the overall-effect parameters z given posterior samples x_i of R random effects,
    log L(z) = sum_i [ mll_i - log(n_i) + logsumexp_s( log p(x_is | z) - log prior_i(x_is) ) ].
`model` is a CLASS: model(inputs) binds the proposed parameter rows and the resulting object maps one random
effect's samples (n_s, p) to the conditional log-probabilities (N, n_s).  MVNHierarchModel is the built-in
conditional: x | z ~ N(mean = z[:p], cov packed in z[p:]).

These objects plug into B200VectorMCMC through the un-fused route (callable_route.py): everything below is
batched torch on the device (torch.linalg for the p x p factorisations) -- coverage of the reference's
likelihood family, not one of the measured kernels.
"""

import math

import numpy as np
import torch

from .log_likelihoods import BaseLogLike


class ApproxHierarch(BaseLogLike):
    kind = "torch"

    def __init__(self, model, data, args):
        """data: (R, n_s, p) posterior samples of the R random effects; args[0]: their marginal log-likelihoods
        (R,); args[1]: log prior probabilities of those samples (R, n_s)."""
        super().__init__(model, data, args)
        self._dev = None

    def _tensors(self, device):
        if self._dev is None or self._dev[0] != device:
            data = [torch.as_tensor(np.asarray(d, dtype=np.float64), device=device) for d in self._data]
            mll = [float(m) for m in self._args[0]]
            lpr = [torch.as_tensor(np.asarray(q, dtype=np.float64), device=device) for q in self._args[1]]
            self._dev = (device, data, mll, lpr)
        return self._dev[1:]

    def torch_loglike(self, params):
        """(N, d) float64 device tensor -> (N,) log-likelihoods."""
        data, mll, lpr = self._tensors(params.device)
        cond = self._model(params)
        total = torch.zeros(params.shape[0], dtype=torch.float64, device=params.device)
        for d, m, q in zip(data, mll, lpr):
            total += m - math.log(d.shape[0]) + torch.logsumexp(cond(d) - q, dim=1)
        return total

    def __call__(self, inputs):
        """(N, 1) host array, evaluated on the device (the reference's calling convention)."""
        from . import _lib
        _lib.require_cuda()
        t = torch.as_tensor(np.asarray(inputs, dtype=np.float64), device="cuda")
        return self.torch_loglike(t).cpu().numpy().reshape(-1, 1)


class MVNHierarchModel:
    """p(x | z): multivariate normal with mean z[:, :p] and covariance from the packed upper-triangular
    z[:, p:] (smcpy/hierarch_log_likelihoods.py:58-100).  d = p + p (p + 1) / 2."""

    def __init__(self, inputs):
        inputs = torch.as_tensor(inputs, dtype=torch.float64)
        d = inputs.shape[1]
        p = int(-3 / 2 + math.sqrt(9 / 4 + 2 * d))
        if p * (p + 1) // 2 != d - p:
            raise IndexError("Provided numbers of parameters and covariance terms are invalid.")
        self._mean = inputs[:, :p]
        r, c = torch.triu_indices(p, p, device=inputs.device)
        cov = torch.zeros((inputs.shape[0], p, p), dtype=torch.float64, device=inputs.device)
        cov[:, r, c] = inputs[:, p:]
        cov[:, c, r] = inputs[:, p:]
        self._p = p
        # the reference takes slogdet / inv of whatever symmetric matrix it is handed; LU does the same here
        self._logdet = torch.linalg.slogdet(cov)[1]
        self._vi = torch.linalg.inv(cov)

    def __call__(self, data):
        """data (n_s, p) -> (N, n_s) conditional log-probabilities."""
        delta = data.unsqueeze(0) - self._mean.unsqueeze(1)                    # (N, n_s, p)
        maha = torch.einsum("inj,ijk,ink->in", delta, self._vi, delta)
        return -0.5 * (self._p * math.log(2 * math.pi) + self._logdet.unsqueeze(1) + maha)
