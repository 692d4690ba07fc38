"""
Kernel adapter with the interface of smcpy/mcmc/kernel_base.py:10-66 and
vector_mcmc_kernel.py:7-39: dict <-> array conversion, path ownership, one shared generator.
`rng` may be a numpy Generator (host draws in the reference's order: replay / parity mode) or
a PhiloxRNG / int seed / None (native device generator).
"""

import numpy as np

from . import engine
from .paths import GeometricPath


class KernelBase:
    def __init__(self, mcmc_object, param_order, path, rng=None):
        self._mcmc = mcmc_object
        self._param_order = param_order
        self.path = GeometricPath() if path is None else path
        if rng is None or isinstance(rng, (int, np.integer)):
            rng = engine.PhiloxRNG(rng)
        self.rng = rng

    def has_proposal(self):
        return self.path.proposal

    @property
    def rng(self):
        return self._rng

    @rng.setter
    def rng(self, rng):
        self.set_mcmc_rng(rng)
        self._rng = rng

    @property
    def param_order(self):
        return self._param_order

    @param_order.setter
    def param_order(self, value):
        self._param_order = value

    def _conv_param_array_to_dict(self, param_array):
        return dict(zip(self.param_order, np.c_[param_array].T))

    def _conv_param_dict_to_array(self, param_dict):
        first = param_dict[self.param_order[0]]
        dim0 = 1 if isinstance(first, (int, float)) else len(first)
        out = np.zeros((dim0, len(self.param_order)))
        for i, k in enumerate(self.param_order):
            out[:, i] = param_dict[k]
        return out


class VectorMCMCKernel(KernelBase):
    def __init__(self, vector_mcmc_object, param_order, path=None, rng=None):
        super().__init__(vector_mcmc_object, param_order, path, rng)
        self._mcmc.evaluate_log_posterior = self.path.logpdf
        if hasattr(self._mcmc, "bind_path"):
            self._mcmc.bind_path(self.path)
        self.param_order = tuple(str(p) for p in param_order)

    def mutate_particles(self, param_dict, num_samples, cov):
        arr = self._conv_param_dict_to_array(param_dict)
        arr, log_likes = self._mcmc.smc_metropolis(arr, num_samples, cov)
        return self._conv_param_array_to_dict(arr), log_likes

    def sample_from_prior(self, num_samples):
        return self._conv_param_array_to_dict(self._mcmc.sample_from_priors(num_samples))

    def sample_from_proposal(self, num_samples):
        rng = self.rng
        if isinstance(rng, engine.PhiloxRNG):
            # the proposal object draws on the host (scipy): give it a numpy Generator keyed by the native
            # generator's seed and next stream -- the same on every rank
            rng = np.random.default_rng([rng.seed & 0xFFFFFFFF, rng.seed >> 32, rng.next_stream()])
        arr = self.path.proposal.rvs(num_samples, random_state=rng)
        return self._conv_param_array_to_dict(arr)

    def get_log_likelihoods(self, param_dict):
        return self._mcmc.evaluate_log_likelihood(self._conv_param_dict_to_array(param_dict))

    def get_log_priors(self, param_dict):
        lp = self._mcmc.evaluate_log_priors(self._conv_param_dict_to_array(param_dict))
        return np.sum(lp, axis=1).reshape(-1, 1)

    def set_mcmc_rng(self, rng):
        self._mcmc.rng = rng

    # -- device-resident extras used by smcpy_b200.samplers -------------------------------------
    @property
    def spec(self):
        return self._mcmc.spec

    def device_path(self):
        p = self.path
        prev = p.previous_phi if p.previous_phi is not None else p.phi
        return engine.make_path(p.phi, prev, p.lam)
