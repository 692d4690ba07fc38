"""
Tempering path: same host-side bookkeeping as smcpy/paths.py (phi stack :9-61, geometric
path with proposal / required phi :64-112).  The arithmetic (logpdf, inc_log_weights) runs
on the device: smcb_reweight for the particle weights, the fused Metropolis kernel for the
acceptance ratio; the methods here are thin host-array wrappers over the same kernels.
"""

import numpy as np


class PathBase:
    def __init__(self, proposal):
        self._phi_list = [0]
        self._proposal = proposal

    @property
    def phi(self):
        return self._phi_list[-1]

    @phi.setter
    def phi(self, phi):
        if phi <= self._phi_list[-1]:
            raise ValueError("phi updates must be monotonic; " f"tried {self.phi} -> {phi}")
        self._phi_list.append(phi)

    @property
    def previous_phi(self):
        return self._phi_list[-2] if len(self._phi_list) > 1 else None

    @property
    def delta_phi(self):
        return self._phi_list[-1] - self._phi_list[-2] if len(self._phi_list) > 1 else None

    def undo_phi_set(self):
        self._phi_list = self._phi_list[:-1]

    @property
    def proposal(self):
        return self._proposal


class GeometricPath(PathBase):
    def __init__(self, proposal=None, required_phi=1):
        super().__init__(proposal)
        self._lambda = None
        self.required_phi_list = required_phi

    @property
    def required_phi_list(self):
        return self._required_phi_list.copy()

    @required_phi_list.setter
    def required_phi_list(self, phi):
        if isinstance(phi, (float, int)):
            phi = [phi]
        self._required_phi_list = sorted([p for p in phi if p < 1])
        self._lambda = min(self._required_phi_list + [1])

    @property
    def lam(self):
        return self._lambda

    def _proposal_logpdf(self, inputs, log_prior):
        if self._proposal:
            return np.asarray(self._proposal.logpdf(inputs), dtype=np.float64).reshape(-1, 1)
        return None

    def logpdf(self, inputs, log_like, log_prior):
        """smcpy/paths.py:81-84, (N,1).  log_prior may be (N,1) or (N,n_priors)."""
        from .engine import host_path_logpdf
        lp = np.asarray(log_prior, dtype=np.float64)
        lp = lp.reshape(lp.shape[0], -1).sum(axis=1, keepdims=True)
        return host_path_logpdf(np.asarray(log_like, dtype=np.float64), lp,
                                self._proposal_logpdf(inputs, log_prior), self.phi, self._lambda)

    def inc_log_weights(self, inputs, log_like, log_prior):
        """smcpy/paths.py:86-91, (N,1)."""
        from .engine import host_inc_log_weights
        lp = np.asarray(log_prior, dtype=np.float64).reshape(-1, 1)
        return host_inc_log_weights(np.asarray(log_like, dtype=np.float64), lp,
                                    self._proposal_logpdf(inputs, log_prior), self.phi,
                                    self.previous_phi, self._lambda)
