import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np
from baseline import fork_comm as fc
assert fc.import_smcpy() is not None
from tests.golden import configs
p = configs.linear_problem(1000, sigma=0.5, data_seed=0)
x, data = p["model"][1], p["data"]
phi = np.linspace(0, 1, 20)
for ranks in (1, 2, 4, 8):
    dt, mll = fc.c2_fixed_run(x, data, 2048, 20, phi, 0.8, 1, ranks)
    print(ranks, "ranks: %.2f s  %.3e steps/s  mll %.3f" % (dt, 2048 * 20 * 19 / dt, mll), flush=True)
