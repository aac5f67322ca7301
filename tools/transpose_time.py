"""Device-resident transpose timing: achieved HBM GB/s (every bit read once + written once)."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import m4ri_b200  # noqa: E402

lib = m4ri_b200.load_library()
torch.cuda.init()
ts = torch.cuda.Stream()
torch.cuda.set_stream(ts)
sh = ctypes.c_void_p(ts.cuda_stream)
for m, n in [(16384, 16384), (65536, 65536), (32768, 131072)]:
    a = torch.randint(-2**62, 2**62, (m, n // 64), dtype=torch.int64, device="cuda")
    d = torch.zeros((n, m // 64), dtype=torch.int64, device="cuda")
    dA = lib.m4ri_b200_dmat_wrap(a.data_ptr(), n // 64, m, n)
    dD = lib.m4ri_b200_dmat_wrap(d.data_ptr(), m // 64, n, m)
    for _ in range(3):
        lib.m4ri_b200_dtranspose(dD, dA, sh)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        lib.m4ri_b200_dtranspose(dD, dA, sh)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    gbs = 2.0 * m * n / 8 / (ms * 1e-3) / 1e9
    print(f"transpose {m}x{n}: {ms:.3f} ms  {gbs:.0f} GB/s  ({gbs/6550.4:.2f} of measured HBM copy peak)", flush=True)
