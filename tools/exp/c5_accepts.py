"""Experiment: per-step global accept fractions of the C5 adaptive run (are the covariance-scale decisions predictable?)."""
import os, sys, warnings
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import smcpy_b200 as sb
from smcpy_b200 import engine
from tests.golden.b200_build import make_b200
from tests.golden import configs
log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 23
problem = configs.gauss_problem(32)
mcmc, kernel = make_b200(problem, rng=2)
smc = sb.AdaptiveSampler(kernel, show_progress_bar=False)
smc._result = sb.InMemoryStorage(retain=1)
rows = []
orig = engine.mutate
def wrapped(st, spec, K, *a, **k):
    r = orig(st, spec, K, *a, **k)
    rows.append((st.accept_counts[:K].cpu().numpy() / st.n_global))
    return r
engine.mutate = wrapped
import smcpy_b200.mutator as mu
for m in (mu, ):
    if hasattr(m, "engine"): m.engine.mutate = wrapped
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    steps, mll = smc.sample(1 << log2n, 10, target_ess=0.8, resample_rng=sb.standard)
for i, r in enumerate(rows):
    print(i, " ".join("%.3f" % v for v in r))
