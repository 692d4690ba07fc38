"""
Resampling uniforms, same call signature as smcpy/resampler_rngs.py:4-9
(`resample_rng(mcmc_kernel, N) -> u[N]`).  With a numpy Generator on the kernel they draw on
the host exactly as the reference does (replay parity); with the native Philox generator the
sampler recognises `.kind` and generates them on the device (smcb_resample_uniforms).
`systematic` is an extension with no reference counterpart.
"""

import numpy as np

from . import _lib


def _host_only(mcmc_kernel):
    if not isinstance(mcmc_kernel.rng, np.random.Generator):
        raise TypeError("host resampling uniforms need a numpy Generator on the kernel")


def standard(mcmc_kernel, N):
    _host_only(mcmc_kernel)
    return mcmc_kernel.rng.uniform(0, 1, N)


def stratified(mcmc_kernel, N):
    _host_only(mcmc_kernel)
    return 1 / N * (np.arange(N) + mcmc_kernel.rng.uniform(0, 1, N))


def systematic(mcmc_kernel, N):
    _host_only(mcmc_kernel)
    return 1 / N * (np.arange(N) + mcmc_kernel.rng.uniform(0, 1))


standard.kind = _lib.RESAMPLE_STANDARD
stratified.kind = _lib.RESAMPLE_STRATIFIED
systematic.kind = _lib.RESAMPLE_SYSTEMATIC
