"""
B200VectorMCMC -- drop-in sibling of smcpy.VectorMCMC / ParallelVectorMCMC
(smcpy/mcmc/vector_mcmc.py:10-236, parallel_vector_mcmc.py:7-47) whose smc_metropolis,
log-prior and log-likelihood evaluations run as sm_100a kernels.

Same constructor and method names, numpy in / numpy out, so it can sit inside the reference's
own VectorMCMCKernel and samplers (parity path: arrays cross the bus every SMC step and the
host numpy Generator supplies the normals and uniforms in the reference's order).  The
device-resident samplers in smcpy_b200.samplers bypass the host arrays altogether.
"""

import ctypes as C

import numpy as np
import torch

from . import _lib, engine
from ._lib import call, ptr, stream_ptr
from .log_likelihoods import Normal, loglike_kind
from .priors import prior_dim


def _spec_for_priors(priors):
    """Minimal problem carrying only a prior table (for stand-alone prior log-pdfs)."""
    from .priors import describe_prior
    c = _lib.Problem()
    d = 0
    for i, p in enumerate(priors):
        k, ncols, a, b, lpin = describe_prior(p)
        c.priors[i] = _lib.Prior(k, ncols, a, b, lpin)
        d += ncols * (ncols + 1) // 2 if k == _lib.PRIOR_IMPROPER_COV else (ncols if k == _lib.PRIOR_EXTERNAL else 1)
    c.n_priors = len(priors)
    c.n_params = d
    return c, d


def host_prior_columns(priors, x):
    """{prior index: (N,) log-pdf} of the host-evaluated (SMCB_PRIOR_EXTERNAL) priors on host rows x (N, d):
    each object sees its own `dim` columns as a 2-D block, as vector_mcmc.py:105-109 hands them over."""
    from .priors import describe_prior, prior_dim
    out, c0 = {}, 0
    for i, p in enumerate(priors):
        k, ncols, _, _, _ = describe_prior(p)
        dim = ncols * (ncols + 1) // 2 if k == _lib.PRIOR_IMPROPER_COV else (ncols if k == _lib.PRIOR_EXTERNAL else 1)
        if k == _lib.PRIOR_EXTERNAL:
            out[i] = np.asarray(p.logpdf(x[:, c0:c0 + dim]), dtype=np.float64).reshape(-1)
        c0 += dim
    return out


def device_log_priors(priors, x, problem_c=None):
    """(N, n_priors) un-summed log-priors of host rows x (N, d)  (vector_mcmc.py:98-111): device prior table,
    plus the host-evaluated columns."""
    x = np.asarray(x, dtype=np.float64)
    if problem_c is None:
        problem_c, d = _spec_for_priors(priors)
    else:
        d = int(problem_c.n_params)
    if x.ndim != 2 or x.shape[1] != d:
        raise ValueError("Num prior distributions != num input params")
    n = x.shape[0]
    rt = engine.Runtime.get()
    soa = torch.empty((d, n), dtype=torch.float64, device=rt.device)
    x_dev = engine.dev(x)
    call("smcb_aos_to_soa", ptr(x_dev), ptr(soa), n, d, n, stream_ptr())
    lp = torch.empty(n, dtype=torch.float64, device=rt.device)
    cols = torch.empty((int(problem_c.n_priors), n), dtype=torch.float64, device=rt.device)
    call("smcb_log_priors", C.byref(problem_c), ptr(soa), n, n, ptr(lp), ptr(cols), stream_ptr())
    out = cols.t().contiguous().cpu().numpy()
    for i, v in host_prior_columns(priors, x).items():
        out[:, i] = v
    return out


def device_log_likelihood(loglike, inputs):
    """(N,1) log-likelihoods for a log-likelihood descriptor whose priors are irrelevant:
    evaluated with wide improper-uniform priors so every row is evaluated."""
    from .priors import ImproperUniform
    inputs = np.asarray(inputs, dtype=np.float64)
    priors = [ImproperUniform() for _ in range(inputs.shape[1])]
    mc = B200VectorMCMC(loglike._model, loglike._data, priors, loglike._args, type(loglike))
    return mc.evaluate_log_likelihood(inputs)


class B200VectorMCMC:
    def __init__(self, model, data, priors, log_like_args=None, log_like_func=Normal):
        """
        :param model: a smcpy_b200.models.DeviceModel (fused kernels), a models.NvrtcModel, or a torch
            callable mapping a device tensor (N, d_model) to (N, M) (un-fused route); for
            ApproxHierarch the conditional-probability model class (hierarch_log_likelihoods.py)
        :param data: 1-D data array (Normal) or (snapshots, features) array (MVNormal)
        :param priors: scipy.stats.uniform frozen objects, ImproperUniform, ImproperCov (device
            log-pdf); InvWishart, ImproperConstrainedUniform or any object with logpdf (host-evaluated)
        :param log_like_args: likelihood hyper-parameters as in the reference
        :param log_like_func: Normal, MultiSourceNormal, MVNormal (these or the reference's classes),
            ApproxHierarch, or a class with a torch_loglike(params) method
        """
        self._eval_model = model
        self._data = data
        self._priors = priors
        self._log_like_args = log_like_args
        self._kind = loglike_kind(log_like_func)
        self._log_like_func = log_like_func(model, data, log_like_args)
        self._rng = np.random.default_rng()
        self._spec = None
        self._proposal = None

    # -- reference surface ---------------------------------------------------------------------
    @property
    def rng(self):
        return self._rng

    @rng.setter
    def rng(self, rng):
        if isinstance(rng, (np.random.Generator, engine.PhiloxRNG)):
            self._rng = rng
        else:
            raise TypeError("Random number generator must be a numpy generator.")

    @property
    def spec(self):
        if self._spec is None:
            self._spec = engine.ProblemSpec(self._eval_model, self._data, self._priors, self._log_like_args,
                                            self._kind, loglike=self._log_like_func)
            self._spec.set_proposal(self._proposal)
        return self._spec

    def bind_path(self, path):
        """Tells the MCMC object about the kernel's path: a describable proposal
        (MultivarIndependent of scipy norm / uniform) gets a device log q column."""
        self._proposal = getattr(path, "proposal", None)
        if self._spec is not None:
            self._spec.set_proposal(self._proposal)

    @staticmethod
    def evaluate_log_posterior(inputs, log_likelihood, log_priors):
        """vector_mcmc.py:120-122; VectorMCMCKernel overwrites this attribute with path.logpdf."""
        return np.sum(np.hstack((log_likelihood, log_priors)), axis=1)

    def evaluate_model(self, inputs):
        return self._eval_model(inputs)

    def sample_from_priors(self, num_samples):
        """vector_mcmc.py:91-96 (host draw from the shared generator, reference order)."""
        if not isinstance(self._rng, np.random.Generator):
            raise TypeError("host prior sampling needs a numpy Generator; the native sampler draws "
                            "uniform priors on the device")
        samples = [p.rvs(num_samples, random_state=self._rng).T for p in self._priors]
        return np.vstack(samples).T

    def _get_prior_dims(self):
        return [prior_dim(p) for p in self._priors]

    def evaluate_log_priors(self, inputs):
        inputs = np.asarray(inputs, dtype=np.float64)
        if inputs.shape[1] != sum(self._get_prior_dims()):
            raise ValueError("Num prior distributions != num input params")
        return device_log_priors(self._priors, inputs, self.spec.c)

    def _load(self, inputs):
        inputs = np.ascontiguousarray(np.asarray(inputs, dtype=np.float64))
        if inputs.ndim != 2 or inputs.shape[1] != self.spec.d:
            raise ValueError("Num prior distributions != num input params")
        st = engine.DeviceState(inputs.shape[0], inputs.shape[0], self.spec.d,
                                with_lq=self.spec.has_device_proposal)
        st.set_params_host(inputs)
        return st

    def evaluate_log_likelihood(self, inputs):
        """(N,1)  (vector_mcmc.py:116-118); every row is evaluated, whatever its prior."""
        st = self._load(inputs)
        engine.eval_incoming(st, self.spec)
        f = int(st.flags[0].item())
        if f & _lib.FLAG_MODEL_NAN:
            raise ValueError("nan in model output.")
        return st.ll.cpu().numpy().reshape(-1, 1)

    def _path_from_posterior(self):
        """The tempered target reaches VectorMCMC only through the bound method the kernel
        assigns (vector_mcmc_kernel.py:10): recover the path object from it."""
        owner = getattr(self.evaluate_log_posterior, "__self__", None)
        if owner is not None and hasattr(owner, "phi") and hasattr(owner, "_lambda"):
            prop = getattr(owner, "proposal", None)
            if prop is not self._proposal:
                self.bind_path(owner)
            if prop and not self.spec.has_device_proposal:
                raise NotImplementedError("smc_metropolis with a path proposal needs a MultivarIndependent of "
                                          "scipy norm / uniform components (device log q)")
            return engine.make_path(owner.phi, owner.phi, owner._lambda)
        return engine.make_path(1.0, 1.0, 1.0)       # plain posterior: ll + sum(log priors)

    def smc_metropolis(self, inputs, num_samples, cov):
        """vector_mcmc.py:46-62 on the device; returns (inputs (N,d), log_like (N,1))."""
        path = self._path_from_posterior()             # binds a path proposal before the state is laid out
        st = self._load(inputs)
        engine.eval_incoming(st, self.spec)
        st.check_flags("smc_metropolis")
        cov = np.asarray(cov, dtype=np.float64).reshape(self.spec.d, self.spec.d)
        st.cov.copy_(engine.dev(cov))
        engine.cholesky(st)
        engine.mutate(st, self.spec, num_samples, path, self._rng)
        return st.params_host(), st.ll.cpu().numpy().reshape(-1, 1)

    def metropolis(self, inputs, num_samples, cov, adapt_interval=None, adapt_delay=0, progress_bar=False,
                   **kwargs):
        """Stand-alone long chain (vector_mcmc.py:64-86): every step is the fused device kernel,
        the chain (N, d, num_samples + 1) is recorded on the device and copied back once; the
        windowed covariance adaptation (:153-160,185-197) is host logic on the recorded window."""
        path = self._path_from_posterior()
        st = self._load(inputs)
        engine.eval_incoming(st, self.spec)
        st.check_flags("metropolis")
        n, d = st.n, st.d
        cov = np.asarray(cov, dtype=np.float64).reshape(d, d)
        st.cov.copy_(engine.dev(cov))
        engine.cholesky(st)
        chain = torch.empty((num_samples + 1, d, n), dtype=torch.float64, device=st.cols.device)
        chain[0].copy_(st.params)
        it = range(1, num_samples + 1)
        if progress_bar:
            from tqdm import tqdm
            it = tqdm(it)
        for i in it:
            engine.mutate(st, self.spec, 1, path, self._rng)        # one step, no covariance rescaling
            chain[i].copy_(st.params)
            if self._is_adapt_iteration(adapt_interval, i, adapt_delay):
                start = self._get_window_start(i, adapt_delay, adapt_interval)
                win = chain[start:i + 1].permute(1, 2, 0).reshape(d, -1).cpu().numpy()   # (d, N * window)
                st.cov.copy_(engine.dev(np.cov(win).reshape(d, d)))
                engine.cholesky(st)
        return chain.permute(2, 1, 0).contiguous().cpu().numpy()

    @staticmethod
    def _is_adapt_iteration(adapt_interval, idx, adapt_delay):
        if adapt_interval is None:
            return False
        return idx >= adapt_delay and (idx - adapt_delay) % adapt_interval == 0

    @staticmethod
    def _get_window_start(idx, adapt_delay, adapt_interval):
        if idx >= adapt_delay + adapt_interval:
            return adapt_delay + 1
        return max(adapt_delay - adapt_interval + 1, 1)
