"""Particle mutation: weighted covariance -> K Metropolis steps  (smcpy/smc/mutator.py:39-81)."""

from . import engine
from .kernel import KernelBase
from .particles import Particles


class Mutator:
    def __init__(self, mcmc_kernel):
        self.mcmc_kernel = mcmc_kernel

    @property
    def mcmc_kernel(self):
        return self._mcmc_kernel

    @mcmc_kernel.setter
    def mcmc_kernel(self, mcmc_kernel):
        if not isinstance(mcmc_kernel, KernelBase):
            raise TypeError
        self._mcmc_kernel = mcmc_kernel

    def mutate_state(self, st, num_samples):
        """In place on a device working set; returns the mutation ratio (mutator.py:69-71)."""
        kernel = self._mcmc_kernel
        if kernel.has_proposal() and not (kernel.spec.has_device_proposal and st.with_lq):
            raise NotImplementedError("mutation under GeometricPath(proposal=...) needs a MultivarIndependent "
                                      "of scipy norm / uniform components (device log q)")
        import numpy as np
        native = not isinstance(kernel.rng, np.random.Generator)
        engine.covariance(st)
        engine.cholesky(st, lazy=native)
        p = kernel.path
        path = engine.make_path(p.phi, p.phi, p.lam)
        return engine.mutate(st, kernel.spec, num_samples, path, kernel.rng, lazy_chol=native)

    def mutate(self, particles, num_samples):
        st = particles.to_state()
        ratio = self.mutate_state(st, num_samples)
        attrs = {"total_unnorm_log_weight": particles.total_unnorm_log_weight,
                 "phi": self._mcmc_kernel.path.phi, "mutation_ratio": ratio}
        return Particles.from_state(st, particles.param_names, attrs, clone=False)
