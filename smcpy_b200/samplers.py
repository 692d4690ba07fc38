"""
SMC samplers with the interface of smcpy/smc/samplers.py:50-301 (SamplerBase,
FixedPhiSampler, AdaptiveSampler), re-hosted around the device kernels: the phi control loop
stays on the host, every per-particle computation is a C-ABI call on a device-resident
working set, and the adaptive phi search (scipy bisect in the reference) runs on the device
without a host round-trip per iteration.  FixedTimeSampler / MaxStepSampler (samplers.py:304-454)
are host phi-policies on top of AdaptiveSampler.optimize_step and are provided unchanged in
behaviour.
"""

import time

import numpy as np

from . import engine
from .initializer import Initializer
from .mutator import Mutator
from .particles import Particles
from .resampler_rngs import standard
from .storage import ContextManager, InMemoryStorage
from .updater import Updater


def rank_zero_value(value):
    """Rank 0's copy of a host scalar on every rank (the reference decorates optimize_step with
    rank_zero_output_only, smcpy/utils/mpi_utils.py:18-28, so that every MPI rank follows rank 0's decision);
    the identity without a process group."""
    if engine.dist_info()[1] == 1:
        return value
    import torch.distributed as dist
    box = [value]
    dist.broadcast_object_list(box, src=0)
    return box[0]


class SamplerBase:
    def __init__(self, mcmc_kernel, show_progress_bar=True):
        self._mcmc_kernel = mcmc_kernel
        self._initializer = Initializer(self._mcmc_kernel)
        self._mutator = Mutator(self._mcmc_kernel)
        self._updater = None
        self._phi_sequence = []
        self._show_progress_bar = show_progress_bar
        self._st = None
        self._step = None
        try:
            self._result = ContextManager.get_context()
        except TypeError:
            self._result = InMemoryStorage()

    @property
    def phi_sequence(self):
        return np.array(self._phi_sequence)

    @property
    def step(self):
        return self._step

    @step.setter
    def step(self, step):
        if step is not None:
            self._save_step(step)
            self._step = step

    def _save_step(self, step):
        self._result.save_step(step)          # every rank keeps (or writes) its own shard

    def _finish(self):
        """End of sample(): file-backed stores stream steps out asynchronously -- wait for the last write."""
        flush = getattr(self._result, "flush", None)
        if flush is not None:
            flush()
        if self._step is not None:
            self._step.detach()
        return self._result, self._result.estimate_marginal_log_likelihoods()

    def _initialize(self, num_particles):
        """samplers.py:87-92.  With a file-backed result store opened on an existing file the run
        resumes from the last stored step: its particles go back to the device (log-priors are
        re-evaluated there), the stored phi sequence is adopted and the path is advanced to the
        stored phi.  (The reference leaves the path at phi = 0 after a restart, so its first
        resumed reweight spans 0 -> phi_next; that is not reproduced.)"""
        if self._result is not None and getattr(self._result, "is_restart", False) and len(self._result) > 0:
            if engine.dist_info()[1] > 1:
                raise NotImplementedError("restart from a file-backed store is single-GPU")
            last = self._result[-1]
            kernel = self._mcmc_kernel
            spec = kernel.spec
            # fresh working set with the layout this problem needs (a proposal log-pdf column when the path has
            # one); log-priors, log-likelihoods and log q are re-evaluated from the stored parameters
            st = engine.DeviceState(last.num_particles, last.num_particles, spec.d, with_lq=spec.has_device_proposal)
            st.cols[:spec.d].copy_(last._cols[:spec.d])
            st.log_w.copy_(last._log_w)
            st.w.copy_(last._w)
            st.total_unnorm_log_weight = last.attrs.get("total_unnorm_log_weight")
            engine.eval_incoming(st, spec)
            st.check_flags("restart", allow_oob=True)
            if not kernel.has_proposal():
                st.lp_const = spec.lp_const
            if isinstance(kernel.rng, engine.PhiloxRNG) and "rng_stream" in last.attrs:
                # continue the counter-based generator where the stored run left it: resumed steps must not
                # reuse the (seed, stream, particle) tuples of the original run
                kernel.rng._stream = max(kernel.rng._stream, int(last.attrs["rng_stream"]))
            self._st = st
            self._step = last                      # already stored: not saved again
            self._restart_phi = list(self._result.phi_sequence)
            phi = float(self._restart_phi[-1])
            if phi > self._mcmc_kernel.path.phi:
                self._mcmc_kernel.path.phi = phi
            return None
        self._restart_phi = None
        self._st = self._initializer.initialize_state(num_particles)
        return self._initializer.particles_from_state(self._st)

    def _do_smc_step(self, phi, num_mcmc_samples):
        """samplers.py:94-98 on the live device working set."""
        self._mcmc_kernel.path.phi = phi
        st = self._updater.update_state(self._st, defer=True)
        ratio = self._mutator.mutate_state(st, num_mcmc_samples)      # reads the deferred checks with its one sync
        attrs = {"total_unnorm_log_weight": st.total_unnorm_log_weight, "phi": self._mcmc_kernel.path.phi,
                 "mutation_ratio": ratio}
        if isinstance(self._mcmc_kernel.rng, engine.PhiloxRNG):
            attrs["rng_stream"] = self._mcmc_kernel.rng._stream
        # a store that only ever shows the latest step gets a view of the working set (no per-step snapshot copy);
        # the view of the final step is turned into an owning snapshot by _finish()
        snap = getattr(self._result, "wants_snapshots", True)
        self.step = Particles.from_state(st, self._mcmc_kernel.param_order, attrs, clone=snap)

    # progress bar (samplers.py:104-124), rank 0 only
    def _pbar_open(self, initial):
        if not self._show_progress_bar or engine.dist_info()[0] != 0:
            return None
        from tqdm import tqdm
        fmt = "{desc}: {percentage:.2f}%|{bar}| phi: {n:.5f}/{total_fmt} [{elapsed}<{remaining}"
        bar = tqdm(total=1.0, bar_format=fmt, initial=initial)
        bar.set_description(f"[ mutation ratio: {self.step.attrs['mutation_ratio']}")
        return bar

    def _pbar_update(self, bar):
        if bar is not None:
            bar.set_description(f"[ mutation ratio: {self.step.attrs['mutation_ratio']}")
            bar.update(float(self._phi_sequence[-1] - self._phi_sequence[-2]))

    @staticmethod
    def _pbar_close(bar):
        if bar is not None:
            bar.close()


class FixedPhiSampler(SamplerBase):
    """SMC sampler using a fixed phi sequence (samplers.py:127-178)."""

    def __init__(self, mcmc_kernel, show_progress_bar=True):
        super().__init__(mcmc_kernel, show_progress_bar)

    def sample(self, num_particles, num_mcmc_samples, phi_sequence, ess_threshold, resample_rng=standard,
               particles_warn_threshold=0.01):
        self._updater = Updater(ess_threshold, self._mcmc_kernel, resample_rng=resample_rng,
                                particles_warn_threshold=particles_warn_threshold)
        self._phi_sequence = phi_sequence
        self.step = self._initialize(num_particles)
        done = float(self._restart_phi[-1]) if self._restart_phi else None
        bar = self._pbar_open(float(self._phi_sequence[0]))
        for i, phi in enumerate(self._phi_sequence[1:]):
            if done is not None and phi <= done:      # resumed run: these steps are already stored
                continue
            self._do_smc_step(phi, num_mcmc_samples)
            if bar is not None:
                bar.set_description(f"[ mutation ratio: {self.step.attrs['mutation_ratio']}")
                bar.update(float(self._phi_sequence[i + 1] - self._phi_sequence[i]))
        self._pbar_close(bar)
        return self._finish()


class AdaptiveSampler(SamplerBase):
    """SMC sampler using an adaptive phi sequence (samplers.py:181-301)."""

    def __init__(self, mcmc_kernel, show_progress_bar=True):
        self.req_phi_index = None
        super().__init__(mcmc_kernel, show_progress_bar)

    def sample(self, num_particles, num_mcmc_samples, target_ess=0.8, min_dphi=None, resample_rng=standard,
               particles_warn_threshold=0.01):
        if target_ess <= 0.0 or target_ess >= 1.0:
            raise ValueError
        self._updater = Updater(ess_threshold=1, mcmc_kernel=self._mcmc_kernel, resample_rng=resample_rng,
                                particles_warn_threshold=particles_warn_threshold)
        self._phi_sequence = [0]
        self.step = self._initialize(num_particles)
        if self._restart_phi:
            self._phi_sequence = list(self._restart_phi)
        bar = self._pbar_open(float(self._phi_sequence[-1]))
        while self._phi_sequence[-1] < 1:
            proposed_phi = self.optimize_step(self.step, self._phi_sequence[-1], target_ess)
            proposed_phi = self._verify_min_dphi(proposed_phi, min_dphi)
            self._do_smc_step(proposed_phi, num_mcmc_samples)
            self._phi_sequence.append(proposed_phi)
            self._pbar_update(bar)
        self._pbar_close(bar)
        self.req_phi_index = [i for i, phi in enumerate(self.phi_sequence)
                              if phi in self._mcmc_kernel.path.required_phi_list]
        return self._finish()

    def _state_of(self, particles):
        if particles is self._step and self._st is not None:
            return self._st
        return particles.to_state()

    def _prepare_lq(self, st):
        kernel = self._mcmc_kernel
        if kernel.has_proposal() and not st.with_lq:
            st.lq = engine.dev(np.asarray(kernel.path.proposal.logpdf(st.params_host()),
                                          dtype=np.float64).reshape(-1))

    def optimize_step(self, particles, phi_old, target_ess=1):
        """samplers.py:250-262: next phi = root of ESS(phi) - target*N on (phi_old, 1] (device
        bisection with scipy's decision sequence), merged with the path's required phis."""
        st = self._state_of(particles)
        self._prepare_lq(st)
        kernel = self._mcmc_kernel
        phi = engine.find_phi(st, kernel.spec, phi_old, target_ess, kernel.path.lam)
        proposed = kernel.path.required_phi_list
        proposed.append(phi)
        return self._select_phi(proposed, phi_old)

    def predict_ess_margin(self, phi_new, phi_old, particles, target_ess):
        """samplers.py:264-276 for one candidate phi (device reduction)."""
        st = self._state_of(particles)
        n = st.n_global
        if not phi_new - phi_old > 0:
            return n - n * target_ess
        self._prepare_lq(st)
        return engine.ess_margin(st, self._mcmc_kernel.spec, phi_new, phi_old, target_ess,
                                 self._mcmc_kernel.path.lam)

    @staticmethod
    def _select_phi(proposed_phi_list, phi_old):
        return min([p for p in proposed_phi_list if p > phi_old])

    def _verify_min_dphi(self, phi, min_dphi):
        dphi = phi - self._phi_sequence[-1]
        if min_dphi and dphi < min_dphi:
            return self._phi_sequence[-1] + min_dphi
        return phi

    def _full_step_meets_target(self, phi_old, particles, target_ess):
        return self.predict_ess_margin(1, phi_old, particles, target_ess) > 0


class FixedTimeSampler(AdaptiveSampler):
    """Host phi-policy of samplers.py:304-379 (wall-clock budget)."""

    def __init__(self, mcmc_kernel, time, show_progress_bar=True, norm_time_threshold=0.8,
                 time_knockdown_factor=0.95):
        self.buffer_time = time * norm_time_threshold
        self.final_time = time * time_knockdown_factor
        self._time_history = [0]
        self._buffer_phi = None
        self._start_time = None
        super().__init__(mcmc_kernel, show_progress_bar)

    @property
    def prev_time_step(self):
        return self._time_history[-1] - self._time_history[-2]

    def sample(self, *args, **kwargs):
        self._start_time = time.time()
        return super().sample(*args, **kwargs)

    def _do_smc_step(self, phi, num_mcmc_samples):
        super()._do_smc_step(phi=phi, num_mcmc_samples=num_mcmc_samples)
        # one clock for the whole job: every phi decision below derives from this history, so the ranks of a
        # multi-GPU run cannot disagree (and then issue mismatched collectives)
        self._time_history.append(rank_zero_value(time.time() - self._start_time))
        estimated_future_time = self.prev_time_step + self._time_history[-1]
        if self._buffer_phi is None and estimated_future_time >= self.buffer_time:
            self._buffer_phi = phi

    def optimize_step(self, particles, phi_old, target_ess=1):
        if len(self._time_history) >= 2:
            if 2 * self.prev_time_step + self._time_history[-1] >= self.final_time:
                return 1
        phi = super().optimize_step(particles=particles, phi_old=phi_old, target_ess=target_ess)
        if self._buffer_phi:
            t = self.prev_time_step + self._time_history[-1]
            return 10 ** max(np.log10(phi), np.interp(t, [self.buffer_time, self.final_time],
                                                      [np.log10(self._buffer_phi), 0]))
        return phi


class MaxStepSampler(AdaptiveSampler):
    """Host phi-policy of samplers.py:382-454 (bounded number of SMC steps)."""

    def __init__(self, mcmc_kernel, max_steps, show_progress_bar=True, norm_step_threshold=0.8):
        self.buffer_step = int(max_steps * norm_step_threshold) - 1
        self.final_step = int(max_steps) - 1
        self._step_count = 0
        self._buffer_phi = None
        super().__init__(mcmc_kernel, show_progress_bar)

    def sample(self, *args, **kwargs):
        self._step_count = 0
        return super().sample(*args, **kwargs)

    def _do_smc_step(self, phi, num_mcmc_samples):
        super()._do_smc_step(phi=phi, num_mcmc_samples=num_mcmc_samples)
        self._step_count += 1
        if self._buffer_phi is None and self._step_count > self.buffer_step:
            self._buffer_phi = phi

    def optimize_step(self, particles, phi_old, target_ess=1):
        if self._step_count >= self.final_step:
            return 1.0
        phi = super().optimize_step(particles=particles, phi_old=phi_old, target_ess=target_ess)
        if self._buffer_phi:
            return 10 ** max(np.log10(phi), np.interp(self._step_count, [self.buffer_step, self.final_step],
                                                      [np.log10(self._buffer_phi), 0]))
        return phi
