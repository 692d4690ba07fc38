"""Times the cooperative reweight launch (smcb_reweight_fused) and the phi search (smcb_ess_search) with CUDA events.
Usage: python tools/rw_time.py [log2n ...]"""
import os, sys, ctypes as C
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from smcpy_b200 import engine, _lib
from smcpy_b200._lib import call, ptr, stream_ptr
sizes = [int(a) for a in sys.argv[1:]] or [20, 23, 26]
for log2n in sizes:
    n = 1 << log2n
    st = engine.DeviceState(n, n, 2)
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    st.cols[2].copy_(torch.randn(n, generator=g, device="cuda", dtype=torch.float64) * 30 - 700)
    st.cols[3].fill_(-4.97)
    st.set_uniform_weights()
    path = engine.make_path(0.05, 0.0, 1.0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for label, thr, lw_in in (("resampling step (nothing written)", 2.0, None), ("weights written", -1.0, None),
                              ("weights written, log_w read", -1.0, st.log_w)):
        ts = []
        for it in range(8):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            call("smcb_reweight_fused", n, ptr(st.ll), None, -4.97, None, ptr(lw_in), n, C.byref(path), thr,
                 ptr(st.log_w_alt), ptr(st.w_alt), ptr(st.scalars), ptr(st.ws), None, stream_ptr())
            b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        t = float(np.median(ts[2:]))
        byt = n * (8 + (8 if lw_in is not None else 0) + (16 if thr < 0 else 0))
        print("2^%d reweight %-34s %7.1f us  %6.0f GB/s (%.0f %% of 6551)" % (log2n, label, t * 1e3, byt / t / 1e6, byt / t / 1e6 / 65.514))
    ts = []
    for it in range(6):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        call("smcb_ess_search", n, ptr(st.ll), None, None, -4.97, 0.0, 1.0, 0.8, n, 0.0, ptr(st.bis), ptr(st.ws), None, stream_ptr(), launches=1)
        b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    bis = st.bis.cpu().numpy()
    print("2^%d search: %7.1f us, %d passes, phi %.6e" % (log2n, float(np.median(ts[2:])) * 1e3, int(bis[10]), bis[0]))
    del st
