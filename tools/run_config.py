"""BASELINE.json configs[2] / configs[3] as configured, end to end, with anchors, clocks and kernel times.
  C3: spring-mass ODE (RK4 device kernel), AdaptiveSampler, 4M particles, 3 params (K, g, sigma), 500 observations, K=5
  C4: MVNormal with unknown covariance, AdaptiveSampler, 16M particles (sharded over the job's GPUs), 8 params, K=20
Launch: python tools/run_config.py c3        |  torchrun --nproc-per-node G tools/run_config.py c4 [log2 n_global]
Prints one JSON line (rank 0)."""
import json, os, sys, time, warnings
sys.path.insert(0, os.getcwd())
import numpy as np, torch, torch.distributed as dist
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import bench
import smcpy_b200 as sb
from smcpy_b200 import engine, _lib
from scipy import stats

cfg = sys.argv[1]
peaks, _ = bench.measured_peaks()


def noisy(clean, std, seed):
    return bench.noisy(clean, std, seed)[0]


if cfg == "c3":
    log2n = int(sys.argv[2]) if len(sys.argv) > 2 else 22
    K_mcmc, M, nsub = 5, 500, 4
    dt = 5.0 / (M - 1)
    t = np.arange(M) * dt
    Kt, gt = 1.67, 4.62
    data = noisy((gt / Kt) * (1.0 - np.cos(np.sqrt(Kt) * t)), 0.5, 3)
    names = ["K", "g", "std"]

    def build(seed):
        mcmc = sb.B200VectorMCMC(sb.SpringMassModel(dt, M, nsub, 0.0, 0.0), data, [stats.uniform(0.0, 10.0)] * 3, None)
        return mcmc, sb.VectorMCMCKernel(mcmc, param_order=names, rng=seed)
    d, flops_unit = 3, (18.0 * 2 - 0) * nsub * (M - 1) + 3.0 * M        # 18 FP64 instr (mostly FMA) per RK4 sub-step
    label = "C3: spring-mass ODE (RK4 x%d sub-steps), AdaptiveSampler, 2^%d particles, 3 params, %d observations, K=%d" % (nsub, log2n, M, K_mcmc)
else:
    log2n = int(sys.argv[2]) if len(sys.argv) > 2 else 24
    K_mcmc = 20
    X = np.arange(3, dtype=float)
    TRUE_COV = np.array([[0.5, 0.25, 0.005], [0.25, 0.25, 0.04], [0.005, 0.04, 1.0]])
    rs = np.random.RandomState(200)
    data = np.tile(2.0 * X + 3.5, (50, 1)) + rs.multivariate_normal([0, 0, 0], TRUE_COV, 50)
    names = ["a", "b"] + ["cov%d" % i for i in range(6)]

    def build(seed):
        priors = [stats.uniform(0.0, 6.0), stats.uniform(0.0, 6.0), sb.ImproperCov(3, dof=5, S=np.eye(3))]
        mcmc = sb.B200VectorMCMC(sb.LinearModel(X), data, priors, [None] * 6, sb.MVNormal)
        return mcmc, sb.VectorMCMCKernel(mcmc, param_order=names, rng=seed)
    d, flops_unit = 8, 0.0
    label = "C4: MVNormal likelihood, unknown 3x3 covariance (50 snapshots x 3 features), AdaptiveSampler, 2^%d particles over %d GPU(s), 8 params, K=%d" % (log2n, world, K_mcmc)
n_global = 1 << log2n
n_local = n_global // world


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def run(seed, profile=False, keep=False):
    mcmc, kernel = build(seed)
    smc = sb.AdaptiveSampler(kernel, show_progress_bar=False)
    smc._result = sb.InMemoryStorage(retain=1)
    engine.PROFILE = {} if profile else None
    engine.PROFILE_BYTES.clear()
    barrier()
    ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    ev[0].record()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        steps, mll = smc.sample(n_global, K_mcmc, target_ess=0.8, resample_rng=sb.standard)
    ev[1].record()
    barrier()
    tt = torch.tensor([ev[0].elapsed_time(ev[1]) * 1e-3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    last = steps[-1]
    rec = {"wall_s": float(tt.item()), "smc_steps": len(smc.phi_sequence) - 1, "mll": float(mll[-1]),
           "mean": last.compute_mean(package=False).tolist(), "std": last.compute_std_dev(package=False).tolist(),
           "mutation_ratio": float(last.attrs["mutation_ratio"]), "variant": _lib.last_mh_variant()}
    return rec


import gc
run(1); gc.collect()
clk = bench.ClockSampler(range(world) if world > 1 else [local], enabled=(rank == 0))
clk.start()
runs = []
for i in range(2):
    runs.append(run(2 + i)); gc.collect()
clocks = clk.stop()
prof_rec = run(2, profile=True)
prof = {k: [a.elapsed_time(b) for a, b in v] for k, v in engine.PROFILE.items()}
engine.PROFILE = None
if rank == 0:
    r0 = min(runs, key=lambda r: r["wall_s"])
    ns = r0["smc_steps"]
    mh = float(np.mean(prof["mh_step"]))
    out = {"workload": label, "particles": n_global, "n_gpus": world, "wall_s_to_phi1": r0["wall_s"],
           "wall_s_runs": [r["wall_s"] for r in runs], "smc_steps": ns, "steps_per_s": n_global * K_mcmc * ns / r0["wall_s"],
           "mll_runs": [r["mll"] for r in runs], "posterior_mean": dict(zip(names, r0["mean"])),
           "posterior_std": dict(zip(names, r0["std"])), "mutation_ratio_last": r0["mutation_ratio"],
           "kernel": r0["variant"], "clocks": clocks,
           "kernels_ms": {k: {"avg": float(np.mean(v)), "calls": len(v), "share_of_wall": float(np.sum(v)) / (prof_rec["wall_s"] * 1e3)}
                          for k, v in prof.items()},
           "mutation": {"avg_ms": mh, "alg_bytes_per_particle_step": 16 * d + 32,
                        "gbps": (16 * d + 32) * n_local / mh / 1e6, "frac_hbm": (16 * d + 32) * n_local / mh / 1e6 / peaks["hbm_gbs"]}}
    if cfg == "c3":
        probes = bench.fp64_probes(_lib, torch)
        if "rk4=propagator" in r0["variant"]:
            # RK4 propagator of the linear ODE: 4 FMAs per observation (the nsub sub-steps composed once per particle
            # and step) + the residual (sub, fma) -- the staged form runs 18 FP64 instructions per sub-step
            flops_unit = 11.0 * M + 40.0 * nsub
            out["mutation"]["rk4_form"] = "propagator"
        else:
            out["mutation"]["rk4_form"] = "stages"
        out["mutation"]["fp64_tflops"] = flops_unit * n_local / mh / 1e9
        out["mutation"]["fp64_peak_tflops"] = max(probes.values())
        out["mutation"]["fp64_frac"] = out["mutation"]["fp64_tflops"] / max(probes.values())
        out["mutation"]["flops_per_particle_step"] = flops_unit
        out["anchor"] = {"truth": {"K": Kt, "g": gt, "std": 0.5},
                         "note": "data = closed-form trajectory (g/K)(1-cos(sqrt(K) t)) + N(0, 0.5); posterior mean within a few "
                                 "posterior std of the truth; RK4 x4 vs closed form 1.2e-11 (SURVEY 8c)"}
    else:
        ybar = data.mean(axis=0)
        Wm = (data - ybar).T @ (data - ybar) / data.shape[0]
        out["anchor"] = {"truth": {"a": 2.0, "b": 3.5, "cov": TRUE_COV[np.triu_indices(3)].tolist()},
                         "sample_cov_of_residuals_mle": Wm[np.triu_indices(3)].tolist(),
                         "note": "flat prior on (a, b) and on PD covariances: the posterior of the covariance terms centres near "
                                 "the sample scatter (inverse-Wishart(S-ish) shaped), a and b near the GLS fit"}
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
