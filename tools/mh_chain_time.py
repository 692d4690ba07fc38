"""Times a whole mutate() call (K chained Metropolis launches enqueued by smcb_mh_steps) with CUDA events: the
per-step figure includes the launch gaps that tools/mh_time.py's per-launch events cannot see (tuning aid;
SMCB_NO_PDL=1 switches the programmatic dependent launch chaining off)."""
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
elif cfg == "c1":
    problem = configs.linear_problem(100); n = 500; K = 10
elif cfg == "c5":
    problem = configs.gauss_problem(32); n = 1 << 23; K = 10
elif cfg == "c3":
    problem = configs.spring_problem(500, 4); n = 1 << 20; K = 3
if len(sys.argv) > 2:
    n = 1 << int(sys.argv[2])
mcmc, kernel = make_b200(problem, rng=3)
st = Initializer(kernel).initialize_state(n)
engine.covariance(st); st.cov.mul_(0.01); engine.cholesky(st)
path = engine.make_path(0.5, 0.5, 1.0)
for _ in range(3):
    engine.mutate(st, mcmc.spec, K, path, kernel.rng)
ts = []
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = engine.mutate(st, mcmc.spec, K, path, kernel.rng)
    e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
t = np.array(ts)
print("%s pdl=%s n=%d K=%d: mutate() median %.4f ms (min %.4f) = %.4f ms per step, moved %.3f, ll sum %.10e"
      % (cfg, "off" if os.environ.get("SMCB_NO_PDL") else "on", n, K, np.median(t), t.min(), np.median(t) / K, r,
         float(st.ll.sum().item())))
