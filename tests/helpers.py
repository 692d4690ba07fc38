"""Shared test helpers (CPU): fixture loading and oracle drivers."""

import os

import numpy as np

from oracle import smc_oracle as orc
from tests.golden import configs

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_CACHE = {}


def golden(name):
    if name not in _CACHE:
        with np.load(os.path.join(GOLDEN_DIR, name + ".npz")) as z:
            _CACHE[name] = {k: z[k] for k in z.files}
    return _CACHE[name]


def run_oracle_case(name, g=None):
    """Runs the oracle sampler for a tests/golden/configs.py case with the same
    generator seed the reference fixture was made with."""
    build, kind, kw, seed = configs.CASES[name]
    problem = build()
    rng = np.random.default_rng(seed)
    init = g["init_params"] if (g is not None and "init_params" in g) else None
    if kind == "adaptive":
        steps, mll, phis = orc.adaptive_sample(problem, kw["n"], kw["K"], kw["target_ess"], rng,
                                               kw["resampler"], init_params=init)
    else:
        steps, mll = orc.fixed_phi_sample(problem, kw["n"], kw["K"], kw["phi"], kw["ess_threshold"],
                                          rng, kw["resampler"], init_params=init)
        phis = np.asarray(kw["phi"], float)
    return steps, mll, phis, rng


# ---- product-side builders (need the built library; GPU only when something is evaluated) ----
from tests.golden.b200_build import make_b200          # noqa: E402,F401  (re-exported)


def run_b200_case(name, g=None, rng=None):
    """Runs the product sampler for a configs.py case (replay mode when rng is a numpy
    Generator seeded like the fixture)."""
    import smcpy_b200 as sb
    build, kind, kw, seed = configs.CASES[name]
    problem = build()
    rng = np.random.default_rng(seed) if rng is None else rng
    mcmc, kernel = make_b200(problem, rng=rng)
    if g is not None and "init_params" in g:
        init = g["init_params"]
        kernel.sample_from_prior = lambda n: dict(zip(kernel.param_order, init.T))
    res = {"standard": sb.standard, "stratified": sb.stratified}[kw["resampler"]]
    if kind == "adaptive":
        smc = sb.AdaptiveSampler(kernel, show_progress_bar=False)
        steps, mll = smc.sample(kw["n"], kw["K"], target_ess=kw["target_ess"], resample_rng=res)
    else:
        smc = sb.FixedPhiSampler(kernel, show_progress_bar=False)
        steps, mll = smc.sample(kw["n"], kw["K"], kw["phi"], kw["ess_threshold"], resample_rng=res)
    return smc, list(steps), mll, rng
