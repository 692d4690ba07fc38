import os, sys, ctypes as C
sys.path.insert(0, os.getcwd())
import numpy as np, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from smcpy_b200 import engine, _lib
from smcpy_b200._lib import call, ptr, stream_ptr
rows, n = 5, 1000
pool = engine.SymmetricPool.get(rows, n)
(cols, pc, _), (alt, pa, _) = pool
cols.copy_(torch.arange(rows * n, dtype=torch.float64, device="cuda").reshape(rows, n) + 1e6 * rank)
alt.fill_(-1.0)
torch.cuda.synchronize(); dist.barrier()
# every rank serves n requests: first half from rank 0, second half from rank 1 (n/2 each)
half = n // 2
idx = torch.arange(n, dtype=torch.int64, device="cuda").flip(0).contiguous()      # reversed rows
block_off = engine.dev(np.array([0, half, n]), torch.int64)
slot_off = engine.dev(np.array([rank * half, rank * half]), torch.int64)          # my block lands at rank*half
dst = (C.c_void_p * world)(*pa)
call("smcb_gather_p2p", ptr(cols), n, ptr(idx), n, rows, world, ptr(block_off), ptr(slot_off), dst, n, stream_ptr())
torch.cuda.synchronize(); dist.barrier()
got = alt.cpu().numpy()
# expected: alt[c][q*half + j] on rank R = cols_of_rank_q[c][idx[R*half + j]]  (block R of server q goes to requester R)
exp = np.empty_like(got)
for q in range(world):
    src = (np.arange(rows * n, dtype=np.float64).reshape(rows, n) + 1e6 * q)
    ids = np.arange(n)[::-1][rank * half:(rank + 1) * half]
    exp[:, q * half:(q + 1) * half] = src[:, ids]
print("rank", rank, "ok" if np.array_equal(got, exp) else "MISMATCH", got[0, :3], exp[0, :3], got[0, half:half+3], exp[0, half:half+3], "ptrs", [hex(p) for p in pa], hex(alt.data_ptr()))
dist.destroy_process_group()
