"""Experiment: per-step accept fractions of the C2 bench workload (how far from the 0.3 / 0.7 thresholds?)."""
import os, sys, warnings
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
import smcpy_b200 as sb
from smcpy_b200 import engine
x, data = bench.linear_data(bench.M_DATA)
rows = []
orig = engine.mutate
def wrapped(st, spec, K, *a, **k):
    r = orig(st, spec, K, *a, **k)
    rows.append((st.accept_counts[:K].cpu().numpy() / st.n_global))
    return r
engine.mutate = wrapped
smc, mcmc = bench.build_c2(sb, x, data, 5, None)
smc._result = sb.InMemoryStorage(retain=1)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    steps, mll = smc.sample(bench.N_PER_GPU, bench.K_MCMC, bench.PHI, bench.ESS_THRESHOLD, resample_rng=sb.standard)
for i, r in enumerate(rows):
    print(i, " ".join("%.3f" % v for v in r))
