"""Phase timing (CUDA events) of a whole sampler run on one GPU. Usage: adaptive_time.py c5|c1big|c3|c4 [log2n] [K]"""
import os, sys, time, warnings
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import smcpy_b200 as sb
from smcpy_b200 import engine
from tests.golden.b200_build import make_b200
from tests.golden import configs
cfg = sys.argv[1]; log2n = int(sys.argv[2]) if len(sys.argv) > 2 else 23; K = int(sys.argv[3]) if len(sys.argv) > 3 else 10
problem = {"c5": lambda: configs.gauss_problem(32), "c1big": lambda: configs.linear_problem(100),
           "c3": lambda: configs.spring_problem(500, 4)}[cfg]()
n = 1 << log2n
def run(seed):
    mcmc, kernel = make_b200(problem, rng=seed)
    smc = sb.AdaptiveSampler(kernel, show_progress_bar=False)
    smc._result = sb.InMemoryStorage(retain=1)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        steps, mll = smc.sample(n, K, target_ess=0.8, resample_rng=sb.stratified)
    return smc, mll
run(1)
torch.cuda.synchronize(); t0 = time.perf_counter()
smc, mll = run(2)
torch.cuda.synchronize(); wall = time.perf_counter() - t0
engine.PROFILE = {}
smc, mll = run(3)
torch.cuda.synchronize()
ns = len(smc.phi_sequence) - 1
print("%s n=2^%d K=%d: %d SMC steps, wall %.1f ms (%.2f ms/SMC step), mll %.4f, %.3e particle-steps/s"
      % (cfg, log2n, K, ns, wall * 1e3, wall * 1e3 / ns, mll[-1], n * K * ns / wall))
for k, v in engine.PROFILE.items():
    t = np.array([a.elapsed_time(b) for a, b in v])
    print("  %-10s calls %4d  total %8.2f ms  avg %7.3f ms  (%.1f%% of wall)" % (k, len(t), t.sum(), t.mean(), 100 * t.sum() / (wall * 1e3)))
