"""Times the one-pass covariance (wmoments_dmma_kernel + reduce + finalise) at 2^23 x 32 and checks it against
torch.cov in fp64 (tuning aid)."""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from smcpy_b200 import engine
n, d = 1 << 23, 32
if len(sys.argv) > 1:
    n = 1 << int(sys.argv[1])
st = engine.DeviceState(n, n, d)
st.cols.normal_()
st.cols[:d] += torch.arange(d, device="cuda", dtype=torch.float64)[:, None] * 0.1
st.set_uniform_weights()
for _ in range(3):
    engine.covariance(st)
torch.cuda.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for _ in range(10):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); engine.covariance(st); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
ref = torch.cov(st.cols[:d, :n], correction=0)
err = float((st.cov - ref).abs().max() / ref.abs().max())
print("cov 2^%d x %d: median %.3f ms (min %.3f) = %.0f GB/s on 8d+8 B per particle; max rel err vs torch.cov %.2e"
      % (int(np.log2(n)), d, np.median(ts), min(ts), (8 * d + 8) * n / np.median(ts) / 1e6, err))
