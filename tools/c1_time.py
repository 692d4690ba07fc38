"""Wall time of config C1 (README example: N=500, K=10, M=100, Adaptive) through the public API, and where the host time goes."""
import os, sys, time, warnings, cProfile, pstats
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from scipy import stats
import bench
import smcpy_b200 as sb
from smcpy_b200 import _lib
x1, d1 = bench.linear_data(100)
def c1(seed):
    mcmc = sb.B200VectorMCMC(sb.LinearModel(x1), d1, [stats.uniform(-6.0, 12.0)] * 2, 0.5)
    kernel = sb.VectorMCMCKernel(mcmc, param_order=("a", "b"), rng=seed)
    smc = sb.AdaptiveSampler(kernel, show_progress_bar=False)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        steps, mll = smc.sample(500, 10, target_ess=0.8)
    p = steps[-1].params
    return time.perf_counter() - t0, len(smc.phi_sequence) - 1, float(mll[-1])
c1(1)
_lib.launch_count = 0
r = [c1(10 + i) for i in range(9)]
print("C1 wall ms:", sorted(round(1e3 * a[0], 2) for a in r), "steps", r[0][1], "mll", [round(a[2], 3) for a in r[:3]], "launches/run", _lib.launch_count / 9)
pr = cProfile.Profile(); pr.enable(); c1(30); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
