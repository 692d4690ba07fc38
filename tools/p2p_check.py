"""2-rank check of the fused P2P gather against the NCCL path (same seed -> same particles)."""
import os, sys, warnings
sys.path.insert(0, os.getcwd())
import numpy as np, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import smcpy_b200 as sb
from smcpy_b200 import engine
from tests.golden.b200_build import make_b200
from tests.golden import configs
problem = configs.gauss_problem(8)
out = {}
for mode in ("p2p", "nccl"):
    if mode == "nccl":
        os.environ["SMCB_NO_P2P"] = "1"
    mcmc, kernel = make_b200(problem, rng=5)
    smc = sb.AdaptiveSampler(kernel, show_progress_bar=False)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        steps, mll = smc.sample(1 << 16, 4, target_ess=0.8, resample_rng=sb.stratified)
    st = smc._st
    out[mode] = (mll.copy(), steps[-1].params.copy(), st.peer_cols is not None, len(steps))
    if rank == 0:
        p = engine.SymmetricPool._cache
        print(mode, "steps", len(steps), "mll", mll[-1], "symmetric", st.peer_cols is not None,
              "offsets", [int(t.data_ptr()) - int(h.buffer_ptrs[rank]) for k in p for (t, _, h) in p[k]])
same = np.array_equal(out["p2p"][0], out["nccl"][0]) and np.array_equal(np.sort(out["p2p"][1], axis=0), np.sort(out["nccl"][1], axis=0))
print("rank", rank, "p2p == nccl:", same)
dist.destroy_process_group()
