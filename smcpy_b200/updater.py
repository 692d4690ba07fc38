"""Weight update, ESS test and resampling  (smcpy/smc/updater.py:40-165) on the device."""

import numpy as np

from . import engine
from .particles import Particles
from .resampler_rngs import standard


class Updater:
    def __init__(self, ess_threshold, mcmc_kernel, resample_rng=standard, particles_warn_threshold=0.01):
        if not callable(resample_rng):
            raise TypeError
        self.ess_threshold = ess_threshold
        self._ess = np.nan
        self._resampled = False
        self._mcmc_kernel = mcmc_kernel
        self.resample_rng = resample_rng
        self.particles_warn_threshold = particles_warn_threshold

    @property
    def ess_threshold(self):
        return self._ess_threshold

    @ess_threshold.setter
    def ess_threshold(self, ess_threshold):
        if ess_threshold < 0.0 or ess_threshold > 1.0:
            raise ValueError("ESS threshold must be between 0 and 1.")
        self._ess_threshold = ess_threshold

    @property
    def particles_warn_threshold(self):
        return self._particles_warn_threshold

    @particles_warn_threshold.setter
    def particles_warn_threshold(self, v):
        if v < 0.0 or v > 1.0:
            raise ValueError("Particle warn threshold must be between 0 and 1")
        self._particles_warn_threshold = v

    @property
    def ess(self):
        return self._ess

    @property
    def resampled(self):
        return self._resampled

    def update_state(self, st, defer=False):
        """In-place update of a device working set: reweight with the kernel path's
        (phi_prev -> phi), then resample if ESS < threshold * N.  The reweight launch applies the same
        test on the device and leaves the weights of a resampling step unmaterialised (the CDF scan
        recomputes them); the replay / parity mode always writes them.  defer=True (samplers): the
        collapse warning is raised at the end of the step without a host sync of its own."""
        kernel = self._mcmc_kernel
        if kernel.has_proposal() and not st.with_lq:   # no device log q: evaluate it on the host
            st.lq = engine.dev(np.asarray(kernel.path.proposal.logpdf(st.params_host()),
                                          dtype=np.float64).reshape(-1))
        native = not isinstance(kernel.rng, np.random.Generator)
        ess, _ = engine.reweight(st, kernel.device_path(), self.ess_threshold if native else -1.0)
        self._ess = ess
        self._resampled = False
        self.last_indices = None
        if ess < self.ess_threshold * st.n_global:
            self._resampled = True
            self.last_indices = engine.resample(st, kernel, self.resample_rng, self.particles_warn_threshold,
                                                defer=defer)
        else:
            engine.materialize_weights(st)             # (never taken: host and device apply the same test)
        return st

    def update(self, particles):
        st = particles.to_state()
        if not particles._lp_valid:
            from .callable_route import log_priors_into         # device prior table + host-evaluated priors
            log_priors_into(st, self._mcmc_kernel.spec, st.params, st.lp)
        self.update_state(st)
        attrs = {"total_unnorm_log_weight": st.total_unnorm_log_weight}
        return Particles.from_state(st, particles.param_names, attrs, clone=False)

    def resample_if_needed(self, particles):
        st = particles.to_state()
        ess = particles.compute_ess()
        self._ess = ess
        self._resampled = False
        if ess < self.ess_threshold * particles.num_particles:
            self._resampled = True
            engine.resample(st, self._mcmc_kernel, self.resample_rng, self.particles_warn_threshold)
            attrs = {"total_unnorm_log_weight": particles.total_unnorm_log_weight}
            return Particles.from_state(st, particles.param_names, attrs, clone=False)
        return particles
