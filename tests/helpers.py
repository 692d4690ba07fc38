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
def make_b200(problem, rng=None, path=None):
    """B200VectorMCMC + VectorMCMCKernel for an oracle-style problem dict."""
    from scipy import stats
    import smcpy_b200 as sb
    kind, arg = problem["model"]
    if kind in ("linear", "linear_features"):
        model = sb.LinearModel(arg)
    elif kind == "spring_mass":
        model = sb.SpringMassModel(arg["dt"], arg["n_out"], arg["nsub"], arg["x0"], arg["v0"])
    elif kind == "identity":
        model = sb.IdentityModel(arg)
    else:
        raise ValueError(kind)
    priors = []
    for p in problem["priors"]:
        if p[0] == "uniform":
            priors.append(stats.uniform(p[1], p[2]))
        elif p[0] == "improper_uniform":
            priors.append(sb.ImproperUniform(p[1], p[2]))
        elif p[0] == "improper_cov":
            priors.append(sb.ImproperCov(int(p[1]), dof=5, S=np.eye(int(p[1]))))
    lk, largs = problem["loglike"]
    cls = {"normal": sb.Normal, "mvnormal": sb.MVNormal, "multisource": sb.MultiSourceNormal}[lk]
    mcmc = sb.B200VectorMCMC(model, problem["data"], priors, largs, cls)
    if path is None and problem.get("proposal"):
        dists = [stats.norm(loc, sc) if k == "normal" else stats.uniform(loc, sc)
                 for k, loc, sc in problem["proposal"]]
        path = sb.GeometricPath(proposal=sb.MultivarIndependent(*dists),
                                required_phi=problem.get("required_phi", 1))
    kernel = sb.VectorMCMCKernel(mcmc, param_order=problem["names"], rng=rng, path=path)
    return mcmc, kernel


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
