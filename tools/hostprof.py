import cProfile, pstats, sys, os, time, warnings
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
import smcpy_b200 as sb
problem = bench.c2_problem()
def one_pass(seed, spec=None):
    smc, mcmc = bench.build_sampler(sb, problem, bench.N_PER_GPU, seed, spec)
    smc._result = sb.InMemoryStorage(retain=1)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        steps, mll = smc.sample(bench.N_PER_GPU, bench.K_MCMC, bench.PHI, bench.ESS_THRESHOLD, resample_rng=sb.standard)
    return steps[-1], mll, mcmc
_,_,mc = one_pass(1)
spec = mc.spec
for i in range(3): one_pass(10+i, spec)
torch.cuda.synchronize(); t0=time.perf_counter()
for i in range(3): one_pass(20+i, spec)
torch.cuda.synchronize(); print("ms per pass", (time.perf_counter()-t0)/3*1e3)
pr = cProfile.Profile(); pr.enable()
one_pass(30, spec); torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
