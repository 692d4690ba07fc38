"""Initial particles: prior (or proposal) draw, one log-likelihood evaluation, uniform
log-weights  (smcpy/smc/initializer.py:42-78)."""

import numpy as np

from . import engine
from ._lib import call, ptr, stream_ptr
from .kernel import KernelBase
from .particles import Particles

import ctypes as C


class Initializer:
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

    def initialize_state(self, num_particles):
        """Device working set for `num_particles` global particles (this rank's shard)."""
        kernel = self.mcmc_kernel
        spec = kernel.spec
        rank, world = engine.dist_info()
        lo, hi = engine.shard_bounds(num_particles, rank, world)
        st = engine.DeviceState(hi - lo, num_particles, spec.d, id_offset=lo, with_lq=spec.has_device_proposal,
                                symmetric=(world > 1))
        host_rng = isinstance(kernel.rng, np.random.Generator)
        if kernel.has_proposal() or host_rng:
            if kernel.has_proposal():
                params = kernel.sample_from_proposal(num_particles)
            else:
                params = kernel.sample_from_prior(num_particles)
            arr = np.column_stack([np.asarray(params[k], dtype=np.float64) for k in kernel.param_order])
            st.set_params_host(arr[lo:hi])
        else:
            rng = kernel.rng
            stream_id = rng.next_stream()
            try:
                call("smcb_sample_priors", C.byref(spec.c), ptr(st.cols), st.n, st.n, rng.seed, stream_id,
                     st.id_offset, stream_ptr())
            except NotImplementedError:
                # priors the device cannot draw (host-evaluated ones, non-integer inverse-Wishart dof): their own
                # rvs on the host, from a numpy Generator keyed by the Philox seed -- the same on every rank
                gen = np.random.default_rng([rng.seed & 0xFFFFFFFF, rng.seed >> 32, stream_id])
                cols = [np.asarray(p.rvs(num_particles, random_state=gen), dtype=np.float64).reshape(num_particles, -1)
                        for p in spec.priors]
                st.set_params_host(np.hstack(cols)[lo:hi])
        if not kernel.has_proposal():
            st.lp_const = spec.lp_const      # drawn from the (flat) priors: every particle is inside the support
        engine.eval_incoming(st, spec)
        # the reference's Initializer does not test the priors (initializer.py:66-73): particles
        # drawn from a proposal may lie outside the prior support and simply get zero weight
        st.check_flags("initialize_particles", allow_oob=True)
        st.set_uniform_weights()
        lw0 = np.log(1 / num_particles)
        st.total_unnorm_log_weight = float(lw0 + np.log(1 + (num_particles - 1) * 1.0))   # _logsum of N equal terms
        return st

    def initialize_particles(self, num_particles):
        st = self.initialize_state(num_particles)
        return self.particles_from_state(st)

    def particles_from_state(self, st):
        attrs = {"total_unnorm_log_weight": st.total_unnorm_log_weight, "phi": 0, "mutation_ratio": 0}
        return Particles.from_state(st, self.mcmc_kernel.param_order, attrs)
