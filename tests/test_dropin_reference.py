"""
Drop-in test with the UNMODIFIED reference (optional): when the reference is installed under
baseline/_ref (pip --target, git-ignored; it travels to the GPU box with the snapshot), its own
VectorMCMCKernel + AdaptiveSampler / FixedPhiSampler are run with smcpy_b200.B200VectorMCMC in
place of smcpy.VectorMCMC and must reproduce the recorded reference run (tests/golden).
Skipped when baseline/_ref is absent; nothing here reads /root/reference.
"""

import os
import sys
import types
import warnings

import numpy as np
import pytest

from tests.golden import configs
from tests.helpers import golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    if not os.path.isdir(os.path.join(REF, "smcpy")):
        pytest.skip("reference not installed under baseline/_ref")
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))     # only HDF5Storage needs it
    sys.path.insert(0, REF)
    import smcpy
    return smcpy


@pytest.mark.parametrize("name", ["c1_adaptive", "c2_fixed_unknown_std", "c3_spring", "c5_gauss8", "c1_proposal_reqphi"])
def test_reference_samplers_with_b200_mcmc(ref, name):
    from scipy import stats
    import smcpy_b200 as sb
    from smcpy.paths import GeometricPath
    from smcpy.proposals import MultivarIndependent
    from smcpy.resampler_rngs import standard, stratified
    g = golden(name)
    build, kind, kw, seed = configs.CASES[name]
    problem = build()
    mkind, arg = problem["model"]
    model = {"linear": lambda: sb.LinearModel(arg), "identity": lambda: sb.IdentityModel(arg),
             "spring_mass": lambda: sb.SpringMassModel(arg["dt"], arg["n_out"], arg["nsub"], arg["x0"], arg["v0"])
             }[mkind]()
    priors = [stats.uniform(p[1], p[2]) for p in problem["priors"]]
    mcmc = sb.B200VectorMCMC(model, problem["data"], priors, problem["loglike"][1], sb.Normal)
    path = None
    if problem.get("proposal"):
        dists = [stats.norm(loc, sc) for _, loc, sc in problem["proposal"]]
        path = GeometricPath(proposal=MultivarIndependent(*dists), required_phi=problem["required_phi"])
    rng = np.random.default_rng(seed)
    kernel = ref.VectorMCMCKernel(mcmc, param_order=problem["names"], path=path, rng=rng)   # the reference's kernel
    res = {"standard": standard, "stratified": stratified}[kw["resampler"]]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if kind == "adaptive":
            smc = ref.AdaptiveSampler(kernel, show_progress_bar=False)                     # the reference's sampler
            steps, mll = smc.sample(kw["n"], kw["K"], target_ess=kw["target_ess"], resample_rng=res)
        else:
            smc = ref.FixedPhiSampler(kernel, show_progress_bar=False)
            steps, mll = smc.sample(kw["n"], kw["K"], kw["phi"], kw["ess_threshold"], resample_rng=res)
    steps = list(steps)
    np.testing.assert_allclose(smc.phi_sequence, g["phi"], rtol=1e-10)
    np.testing.assert_allclose(mll, g["mll"], rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(steps[-1].params, g["params_last"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(steps[-1].log_likes, g["log_likes_last"], rtol=1e-10)
    np.testing.assert_array_equal(rng.uniform(0, 1, 4), g["rng_check"])
