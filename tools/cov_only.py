import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from smcpy_b200 import engine
n, d = 1 << 23, 32
st = engine.DeviceState(n, n, d)
st.cols.normal_()
st.set_uniform_weights()
for _ in range(3):
    engine.covariance(st)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); engine.covariance(st); b.record(); torch.cuda.synchronize()
print("cov 2^23 x 32: %.3f ms" % a.elapsed_time(b))
