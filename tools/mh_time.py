"""Times the fused Metropolis step with CUDA events (tuning aid)."""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import smcpy_b200 as sb
from smcpy_b200 import engine
from smcpy_b200.initializer import Initializer
from tests.golden.b200_build import make_b200
from tests.golden import configs
cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
K = 20
if cfg == "c2":
    problem = configs.linear_problem(1000, sigma=0.5, data_seed=0); n = 1 << 20
elif cfg == "c5":
    problem = configs.gauss_problem(32); n = 1 << 23
elif cfg == "c3":
    problem = configs.spring_problem(500, 4); n = 1 << 20; K = 3
elif cfg == "c4":
    problem = configs.mvn_problem(50); n = 1 << 22
if len(sys.argv) > 2:
    n = 1 << int(sys.argv[2])
if os.environ.get("SMCB_N"):
    n = int(os.environ["SMCB_N"])
mcmc, kernel = make_b200(problem, rng=3)
if cfg == "c4":
    init = configs.mvn_init_params(4096, 1)
    reps = np.tile(init, (n // 4096, 1))
    kernel.rng = np.random.default_rng(0)
    kernel.sample_from_prior = lambda m: dict(zip(kernel.param_order, reps.T))
    st = Initializer(kernel).initialize_state(n)
    kernel.rng = sb.PhiloxRNG(3)
else:
    st = Initializer(kernel).initialize_state(n)
engine.covariance(st); st.cov.mul_(0.01); engine.cholesky(st)
path = engine.make_path(0.5, 0.5, 1.0)
engine.mutate(st, mcmc.spec, 3, path, kernel.rng)
engine.PROFILE = {}
r = engine.mutate(st, mcmc.spec, K, path, kernel.rng)
torch.cuda.synchronize()
t = np.array([a.elapsed_time(b) for a, b in engine.PROFILE["mh_step"]])
d = st.d
print("%s cfg=%s n=%d d=%d: mh_step avg %.4f ms (min %.4f) -> %.3e particle-steps/s, %.1f GB/s algorithmic, moved %.3f"
      % (cfg, os.environ.get("SMCB_LINEAR_CFG", "-"), n, d, t.mean(), t.min(), n / t.mean() * 1e3, (16 * d + 32) * n / t.mean() / 1e6, r))
