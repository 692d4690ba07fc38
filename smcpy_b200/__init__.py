"""
smcpy_b200 -- B200-native (sm_100a) implementation of nasa/SMCPy's data-parallel particle
hot path behind SMCPy's own API.  See DESIGN.md; the C ABI is in include/smcb200.h.

Importing the package does not touch the GPU; any computation needs the built
libsmcb200.so and a CUDA device -- there is no CPU fallback.
"""

__version__ = "0.1.0"

from .engine import PhiloxRNG                                               # noqa: F401
from .kernel import KernelBase, VectorMCMCKernel                            # noqa: F401
from .hierarch_log_likelihoods import ApproxHierarch, MVNHierarchModel      # noqa: F401
from .log_likelihoods import MultiSourceNormal, MVNormal, Normal            # noqa: F401
from .models import IdentityModel, LinearModel, NvrtcModel, SpringMassModel             # noqa: F401
from .particles import Particles                                            # noqa: F401
from .paths import GeometricPath                                            # noqa: F401
from .priors import ImproperConstrainedUniform, ImproperCov, ImproperUniform, InvWishart  # noqa: F401
from .proposals import MultivarIndependent                                  # noqa: F401
from .resampler_rngs import standard, stratified, systematic                # noqa: F401
from .samplers import AdaptiveSampler, FixedPhiSampler, FixedTimeSampler, MaxStepSampler  # noqa: F401
from .storage import HDF5Storage, InMemoryStorage, PickleStorage                                        # noqa: F401
from .vector_mcmc import B200VectorMCMC                                     # noqa: F401

VectorMCMC = B200VectorMCMC
