"""
Device-resident particle container with the interface of smcpy/smc/particles.py:58-205.
Arrays live on the GPU as structure-of-arrays columns; the numpy views the reference exposes
(`params` (N,d), `log_likes` (N,1), `log_weights` (N,1), `weights`) are materialised lazily.
Weight normalisation / total_unnorm_log_weight / covariance run through the same C-ABI kernels
as the samplers (smcb_reweight, smcb_weighted_cov).
"""

import copy
import ctypes as C

import numpy as np
import torch

from . import engine
from ._lib import call, ptr, stream_ptr


class Particles:
    def __init__(self, params, log_likes, log_weights):
        """Same constructor as the reference: dict of 1-D arrays, log-likelihoods, log-weights;
        the weights are normalised on construction (particles.py:63-81)."""
        if not isinstance(params, dict):
            raise TypeError('"params" must be dict of array-like objects.')
        names = tuple(params.keys())
        try:
            arr = np.vstack([np.asarray(v, dtype=np.float64) for v in params.values()]).T
        except ValueError as e:
            raise ValueError("parameter arrays must have equal length") from e
        n, d = arr.shape
        ll = np.array(log_likes, dtype=np.float64).reshape(-1)
        lw = np.array(log_weights, dtype=np.float64).reshape(-1)
        if ll.shape[0] != n:
            raise ValueError(f"log_likes.shape[0] != number particles: {ll.shape[0]} != {n}")
        if lw.shape[0] != n:
            raise ValueError(f"log_weights.shape[0] != number particles: {lw.shape[0]} != {n}")
        st = engine.DeviceState(n, n, d)
        st.set_params_host(arr)
        st.ll.copy_(engine.dev(ll))
        st.lp.fill_(float("nan"))                    # unknown until a kernel evaluates the priors
        # normalise: reweight with a zero increment (phi == phi_prev)
        zeros = torch.zeros(n, dtype=torch.float64, device=st.cols.device)
        st.log_w.copy_(engine.dev(lw))
        path = engine.make_path(0.0, 0.0, 1.0)
        call("smcb_reweight", n, ptr(zeros), ptr(zeros), None, ptr(st.log_w), n, C.byref(path),
             ptr(st.log_w_alt), ptr(st.w_alt), ptr(st.scalars), ptr(st.ws), stream_ptr())
        sc = st.scalars[:6].cpu().numpy()
        self._adopt(st.cols, st.log_w_alt, st.w_alt, d, names, n, 0, lp_valid=False)
        self.attrs = {"total_unnorm_log_weight": float(sc[3])}

    # -- internal constructors -------------------------------------------------------------
    def _adopt(self, cols, log_w, w, d, names, n_global, id_offset, lp_valid=True):
        self._cols, self._log_w, self._w = cols, log_w, w
        self._d = d
        self._param_names = tuple(names)
        self._num_particles = int(cols.shape[1])
        self._n_global = int(n_global)
        self._id_offset = int(id_offset)
        self._lp_valid = lp_valid
        self._cache = {}

    @classmethod
    def from_state(cls, st, names, attrs=None, clone=True):
        """Snapshot of a DeviceState (device-to-device copies unless clone=False)."""
        self = cls.__new__(cls)
        f = (lambda t: t.clone()) if clone else (lambda t: t)
        self._adopt(f(st.cols), f(st.log_w), f(st.w), st.d, names, st.n_global, st.id_offset)
        self._is_view = not clone
        self.attrs = dict(attrs or {})
        return self

    def detach(self):
        """Turns a zero-copy view of a live working set (from_state(clone=False)) into an owning snapshot."""
        if getattr(self, "_is_view", False):
            self._cols, self._log_w, self._w = self._cols.clone(), self._log_w.clone(), self._w.clone()
            self._is_view = False
        return self

    def to_state(self):
        """Fresh DeviceState holding copies of this step."""
        st = engine.DeviceState(self._num_particles, self._n_global, self._d, self._id_offset,
                                with_lq=(self._cols.shape[0] == self._d + 3))
        st.cols.copy_(self._cols)
        st.log_w.copy_(self._log_w)
        st.w.copy_(self._w)
        st.total_unnorm_log_weight = self.attrs.get("total_unnorm_log_weight")
        return st

    # -- reference interface ------------------------------------------------------------------
    @property
    def params(self):
        if "params" not in self._cache:
            n, d = self._num_particles, self._d
            out = torch.empty((n, d), dtype=torch.float64, device=self._cols.device)
            call("smcb_soa_to_aos", ptr(self._cols), ptr(out), n, d, n, stream_ptr())
            self._cache["params"] = engine.host(out)
        return self._cache["params"]

    @property
    def param_dict(self):
        return dict(zip(self._param_names, self.params.T))

    @property
    def param_names(self):
        return self._param_names

    @property
    def num_particles(self):
        return self._num_particles

    @property
    def log_likes(self):
        if "ll" not in self._cache:
            self._cache["ll"] = engine.host(self._cols[self._d]).reshape(-1, 1)
        return self._cache["ll"]

    @property
    def log_weights(self):
        if "lw" not in self._cache:
            self._cache["lw"] = engine.host(self._log_w).reshape(-1, 1)
        return self._cache["lw"]

    @property
    def weights(self):
        if "w" not in self._cache:
            self._cache["w"] = engine.host(self._w).reshape(-1, 1)
        return self._cache["w"]

    @property
    def total_unnorm_log_weight(self):
        return self.attrs["total_unnorm_log_weight"]

    def copy(self):
        new = Particles.__new__(Particles)
        new._adopt(self._cols.clone(), self._log_w.clone(), self._w.clone(), self._d, self._param_names,
                   self._n_global, self._id_offset, self._lp_valid)
        new.attrs = copy.deepcopy(self.attrs)
        return new

    def compute_ess(self):
        """particles.py:158-162."""
        s2 = torch.sum(self._w * self._w)
        _, world = engine.dist_info()
        if world > 1:
            engine._all_reduce_sum(s2)
        return float(1.0 / s2.item())

    def _wsum_cols(self, vals):
        s = torch.sum(vals * self._w[None, :], dim=1)
        _, world = engine.dist_info()
        if world > 1:
            engine._all_reduce_sum(s)
        return s

    def compute_mean(self, package=True):
        """particles.py:164-169."""
        m = self._wsum_cols(self._cols[:self._d]).cpu().numpy()
        return dict(zip(self._param_names, m)) if package else m

    def compute_variance(self, package=True):
        """particles.py:171-180 (weighted sample formula)."""
        p = self._cols[:self._d]
        mean = self._wsum_cols(p)
        s2 = torch.sum(self._w * self._w)
        _, world = engine.dist_info()
        if world > 1:
            engine._all_reduce_sum(s2)
        var = (self._wsum_cols((p - mean[:, None]) ** 2) / (1 - s2)).cpu().numpy()
        return dict(zip(self._param_names, var)) if package else var

    def compute_std_dev(self, package=True):
        sd = np.sqrt(self.compute_variance(package=False))
        return dict(zip(self._param_names, sd)) if package else sd

    def compute_covariance(self):
        """particles.py:189-198: np.cov(params.T, ddof=0, aweights=w) on the device."""
        st = engine.DeviceState.__new__(engine.DeviceState)
        rt = engine.Runtime.get()
        d, n = self._d, self._num_particles
        st.n, st.n_global, st.d = n, self._n_global, d
        st.cols, st.w = self._cols, self._w
        st.mean = torch.zeros(d, dtype=torch.float64, device=rt.device)
        st.cov = torch.zeros((d, d), dtype=torch.float64, device=rt.device)
        st.sums = torch.zeros(1 + d + d * d, dtype=torch.float64, device=rt.device)
        st.ws = rt.workspace(n, d)
        return engine.covariance(st).cpu().numpy().reshape(d, d)
