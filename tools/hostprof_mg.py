"""Host profile of one sharded C2 pass (rank 0) under torchrun: where the Python time of the multi-GPU control loop goes."""
import cProfile, pstats, sys, os, time, warnings
sys.path.insert(0, os.getcwd())
import numpy as np, torch, torch.distributed as dist
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import bench
import smcpy_b200 as sb
x, data = bench.linear_data(bench.M_DATA)
n_global = bench.N_PER_GPU * world
def one_pass(seed, spec=None):
    smc, mcmc = bench.build_c2(sb, x, data, seed, spec)
    smc._result = sb.InMemoryStorage(retain=1)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        steps, mll = smc.sample(n_global, bench.K_MCMC, bench.PHI, bench.ESS_THRESHOLD, resample_rng=sb.standard)
    return steps[-1], mll, mcmc
_, _, mc = one_pass(1)
spec = mc.spec
for i in range(3): one_pass(10 + i, spec)
ts = []
for i in range(6):
    if world > 1: dist.barrier()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    one_pass(20 + i, spec)
    torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
if rank == 0: print("world", world, "ms per pass", [round(t, 1) for t in ts])
pr = cProfile.Profile()
if world > 1: dist.barrier()
pr.enable(); one_pass(30, spec); torch.cuda.synchronize(); pr.disable()
if rank == 0:
    pstats.Stats(pr).sort_stats("tottime").print_stats(22)
    pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
if world > 1: dist.destroy_process_group()
