"""
Log-likelihood descriptors with the constructor signature of smcpy/log_likelihoods.py
((model, data, args)).  They do not compute on the host: B200VectorMCMC turns them into the
likelihood id / arguments of the fused device kernels.
"""

import numpy as np


class BaseLogLike:
    kind = None

    def __init__(self, model, data, args):
        self._model = model
        self._data = data
        self._args = args
        self._model_wrapper = None

    def set_model_wrapper(self, wrapper):
        """smcpy/log_likelihoods.py:18-19.  Accepted for interface parity; the device path
        shards particles instead of wrapping the model (parallel_vector_mcmc.py:36-47)."""
        self._model_wrapper = wrapper

    def __call__(self, inputs):
        """(N,) log-likelihoods of host inputs, evaluated on the device."""
        from .vector_mcmc import device_log_likelihood
        return device_log_likelihood(self, np.asarray(inputs, dtype=np.float64)).reshape(-1)


class Normal(BaseLogLike):
    """iid Gaussian errors with std-dev = args, or the last parameter column when args is
    None (smcpy/log_likelihoods.py:22-49)."""
    kind = "normal"


class MultiSourceNormal(Normal):
    """Several data sources with their own std-dev: args = [segment lengths, std-devs], None
    std-devs are sampled as the trailing parameter columns (smcpy/log_likelihoods.py:52-118)."""
    kind = "multisource"

    def __init__(self, model, data, args):
        super().__init__(model, data, args)
        self._num_nones = list(self._args[1]).count(None)
        if sum(args[0]) != np.asarray(data).shape[0]:
            raise ValueError("data segments in args[0] must sum to dim of data")


class MVNormal(BaseLogLike):
    """Multivariate normal errors; args = packed row-major upper-triangular covariance terms,
    None entries are sampled as the trailing parameter columns
    (smcpy/log_likelihoods.py:121-192).  data is (snapshots, features)."""
    kind = "mvnormal"

    def __init__(self, model, data, args):
        super().__init__(model, data, args)
        self._num_nones = list(args).count(None)


def loglike_kind(obj_or_cls):
    """Accepts these classes, or the reference's own Normal / MVNormal by class name."""
    name = obj_or_cls.__name__ if isinstance(obj_or_cls, type) else type(obj_or_cls).__name__
    if name == "Normal":
        return "normal"
    if name == "MVNormal":
        return "mvnormal"
    if name == "MultiSourceNormal":
        return "multisource"
    if callable(getattr(obj_or_cls, "torch_loglike", None)):
        return "torch"                  # user likelihood evaluated as a torch callable on device tensors
    raise NotImplementedError("log-likelihood %s has neither a device kernel (Normal, MultiSourceNormal, MVNormal) "
                              "nor a torch_loglike(params) method" % name)
