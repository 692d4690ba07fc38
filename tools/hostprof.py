"""Host-side profile of one C2 pass (cProfile) and the GPU idle gaps of the same pass (torch.profiler / CUPTI
timeline of every kernel in the process, ours included).  Usage: python tools/hostprof.py [gaps|cprofile]"""
import cProfile, pstats, sys, os, time, warnings, json, tempfile
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
import smcpy_b200 as sb
x, data = bench.linear_data(bench.M_DATA)
def one_pass(seed, spec=None):
    smc, mcmc = bench.build_c2(sb, x, data, seed, spec)
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
mode = sys.argv[1] if len(sys.argv) > 1 else "gaps"
if mode == "cprofile":
    pr = cProfile.Profile(); pr.enable()
    one_pass(30, spec); torch.cuda.synchronize()
    pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(25)
else:
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        one_pass(31, spec); torch.cuda.synchronize()
    f = os.path.join(tempfile.mkdtemp(), "t.json")
    prof.export_chrome_trace(f)
    ev = json.load(open(f))["traceEvents"]
    ks = sorted([e for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e], key=lambda e: e["ts"])
    busy = sum(e["dur"] for e in ks)
    span = ks[-1]["ts"] + ks[-1]["dur"] - ks[0]["ts"]
    print("GPU activities %d, span %.2f ms, busy %.2f ms, idle %.2f ms" % (len(ks), span / 1e3, busy / 1e3, (span - busy) / 1e3))
    gaps = {}
    end = ks[0]["ts"] + ks[0]["dur"]
    prev = ks[0]
    for e in ks[1:]:
        g = e["ts"] - end
        if g > 2:
            key = (prev["name"][:50], e["name"][:50])
            a = gaps.setdefault(key, [0, 0.0])
            a[0] += 1; a[1] += g
        if e["ts"] + e["dur"] > end:
            end = e["ts"] + e["dur"]; prev = e
    for k, v in sorted(gaps.items(), key=lambda kv: -kv[1][1])[:25]:
        print("%8.1f us total  %4d x %6.1f us   %s  ->  %s" % (v[1], v[0], v[1] / v[0], k[0], k[1]))
