"""Small driver for ncu: a few fused Metropolis steps on the C2 workload (2^20 particles)."""
import os, sys, warnings
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import smcpy_b200 as sb
from smcpy_b200 import engine
cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
from tests.golden import configs
if cfg == "c2":
    problem = configs.linear_problem(1000, sigma=0.5, data_seed=0); n = 1 << 20
else:
    problem = configs.gauss_problem(32); n = 1 << 23
from tests.golden.b200_build import make_b200
mcmc, kernel = make_b200(problem, rng=3)
from smcpy_b200.initializer import Initializer
st = Initializer(kernel).initialize_state(n)
engine.covariance(st); st.cov.mul_(0.01); engine.cholesky(st)
path = engine.make_path(0.5, 0.5, 1.0)
r = engine.mutate(st, mcmc.spec, 4, path, kernel.rng)
torch.cuda.synchronize()
print("ok", cfg, r)
