"""
Built-in device models.  In SMCPy the model is a user callable (N, d) -> (N, M) evaluated by
numpy (smcpy/mcmc/vector_mcmc.py:88-89); here the built-ins are evaluated inside the fused
Metropolis kernel (smcpy_b200/csrc/metropolis.cu), so outputs never touch HBM.  Arbitrary user
models plug in as torch callables on device tensors (see vector_mcmc.B200VectorMCMC).
"""

import numpy as np

from . import _lib


class DeviceModel:
    model_id = 0
    n_model_params = 0

    def fill(self, prob, keep):
        raise NotImplementedError

    def __call__(self, inputs):
        """Model outputs for host inputs, computed on the device with torch ops
        (diagnostics only -- the hot path never materialises them)."""
        import torch
        _lib.require_cuda()
        t = torch.as_tensor(np.asarray(inputs, dtype=np.float64), device="cuda")
        return self.torch_eval(t).cpu().numpy()


class LinearModel(DeviceModel):
    """y_j = a * x_j + b  (README.md:64-79)."""
    model_id = _lib.MODEL_LINEAR
    n_model_params = 2

    def __init__(self, x):
        self.x = np.ascontiguousarray(np.asarray(x, dtype=np.float64).reshape(-1))
        self.n_out = self.x.shape[0]

    def torch_eval(self, t):
        import torch
        x = torch.as_tensor(self.x, device=t.device)
        return t[:, 0:1] * x[None, :] + t[:, 1:2]


class SpringMassModel(DeviceModel):
    """x'' = -K x + g from (x0, v0), classical RK4 with `nsub` sub-steps per output interval,
    outputs x(j * dt), j = 0..n_out-1 (examples/spring_mass/spring_mass_model.py:8-52 with the
    LSODA integrator replaced by a fixed-step device RK4)."""
    model_id = _lib.MODEL_SPRING_MASS
    n_model_params = 2

    def __init__(self, dt, n_out, nsub=4, x0=0.0, v0=0.0):
        self.dt, self.n_out, self.nsub, self.x0, self.v0 = float(dt), int(n_out), int(nsub), float(x0), float(v0)

    def torch_eval(self, t):
        import torch
        K, g = t[:, 0], t[:, 1]
        x = torch.full_like(K, self.x0)
        v = torch.full_like(K, self.v0)
        h = self.dt / self.nsub
        out = [x.clone()]
        for _ in range(1, self.n_out):
            for _ in range(self.nsub):
                k1x, k1v = v, g - K * x
                k2x, k2v = v + 0.5 * h * k1v, g - K * (x + 0.5 * h * k1x)
                k3x, k3v = v + 0.5 * h * k2v, g - K * (x + 0.5 * h * k2x)
                k4x, k4v = v + h * k3v, g - K * (x + h * k3x)
                x = x + (h / 6.0) * (k1x + 2.0 * k2x + 2.0 * k3x + k4x)
                v = v + (h / 6.0) * (k1v + 2.0 * k2v + 2.0 * k3v + k4v)
            out.append(x.clone())
        return torch.stack(out, dim=1)


class IdentityModel(DeviceModel):
    """y_j = theta_j (Gaussian-target config C5)."""
    model_id = _lib.MODEL_IDENTITY

    def __init__(self, n):
        self.n_out = int(n)
        self.n_model_params = int(n)

    def torch_eval(self, t):
        return t.clone()


class NvrtcModel:
    """User model given as CUDA source, registered through the C ABI (smcb_model_register_nvrtc)
    -- the device-side counterpart of the reference's user callable `model(inputs)`
    (smcpy/mcmc/vector_mcmc.py:88-89).  `source` defines

        __device__ double smcb_user_model(const double* theta, int j, double x, double* state)

    = output j of the model for parameters theta[0..n_params) (x = x[j] if a grid is given;
    state = 8 doubles per particle, zero at j = 0, persisting over j).  Used by the un-fused
    Metropolis route with the Normal likelihood; model outputs stay in registers."""

    def __init__(self, source, n_params, n_out, x=None):
        self.source = str(source)
        self.n_model_params = int(n_params)
        self.n_out = int(n_out)
        self.x = None if x is None else np.ascontiguousarray(np.asarray(x, dtype=np.float64).reshape(-1))
        if self.x is not None and self.x.shape[0] != self.n_out:
            raise ValueError("grid length and n_out differ")
        self._handle = None
        self._x_dev = None

    @property
    def handle(self):
        if self._handle is None:
            import ctypes as C
            _lib.require_cuda()
            h = C.c_int32(0)
            _lib.call("smcb_model_register_nvrtc", self.source.encode(), self.n_model_params, C.byref(h), launches=0)
            self._handle = int(h.value)
        return self._handle

    def x_dev(self):
        if self.x is not None and self._x_dev is None:
            import torch
            self._x_dev = torch.as_tensor(self.x, device="cuda")
        return self._x_dev

    def __call__(self, inputs):
        raise TypeError("NvrtcModel is evaluated inside the log-likelihood kernel; it has no host form")
