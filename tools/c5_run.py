"""North-star target run: 32-dim Gaussian target, AdaptiveSampler, 2^23 particles per GPU (64M on 8 GPUs),
K=10, to phi=1.  Launch: torchrun --nproc-per-node G tools/c5_run.py [standard|stratified] [log2 n per gpu]"""
import os, sys, time, warnings
sys.path.insert(0, os.getcwd())
import numpy as np, torch, torch.distributed as dist
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import smcpy_b200 as sb
from smcpy_b200 import engine
from tests.golden.b200_build import make_b200
from tests.golden import configs
kind = sys.argv[1] if len(sys.argv) > 1 else "standard"
log2n = int(sys.argv[2]) if len(sys.argv) > 2 else 23
res = {"standard": sb.standard, "stratified": sb.stratified, "systematic": sb.systematic}[kind]
problem = configs.gauss_problem(32)
n_global = (1 << log2n) * world
K = 10
def run(seed, profile=False):
    mcmc, kernel = make_b200(problem, rng=seed)
    smc = sb.AdaptiveSampler(kernel, show_progress_bar=False)
    smc._result = sb.InMemoryStorage(retain=1)
    engine.PROFILE = {} if profile else None
    if world > 1: dist.barrier()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        steps, mll = smc.sample(n_global, K, target_ess=0.8, resample_rng=res)
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    return smc, steps, mll, time.perf_counter() - t0
run(1)
smc, steps, mll, wall = run(2)
ns = len(smc.phi_sequence) - 1
last = steps[-1]
mean_err = float(np.max(np.abs(last.compute_mean(package=False) - problem["data"])))
std = last.compute_std_dev(package=False)
smc2, _, _, wall_p = run(3, profile=True)
if rank == 0:
    print("C5 %s: %d GPUs x 2^%d = %d particles, d=32, K=%d: %d SMC steps to phi=1, wall %.3f s, mll %.4f (analytic %.4f), "
          "max |mean - truth| %.4f, std in [%.3f, %.3f], %.3e particle-steps/s"
          % (kind, world, log2n, n_global, K, ns, wall, mll[-1], -32 * np.log(20.0), mean_err, std.min(), std.max(), n_global * K * ns / wall))
    for k, v in engine.PROFILE.items():
        t = np.array([a.elapsed_time(b) for a, b in v])
        print("  %-10s calls %4d  avg %8.3f ms  total %8.1f ms (%.1f%% of wall)" % (k, len(t), t.mean(), t.sum(), 100 * t.sum() / (wall_p * 1e3)))
    mh = np.array([a.elapsed_time(b) for a, b in engine.PROFILE["mh_step"]])
    print("  mutation kernel: %.1f GB/s algorithmic per GPU (544 B/particle-step) = %.1f%% of 6551 GB/s"
          % (544.0 * (1 << log2n) / mh.mean() / 1e6, 100 * 544.0 * (1 << log2n) / mh.mean() / 1e6 / 6551.4))
if world > 1:
    dist.destroy_process_group()
