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
