"""Time the left triangular solves: device resident, host API, and the reference on the CPU."""
import ctypes
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import m4ri_b200  # noqa: E402
from tests import harness as H  # noqa: E402

lib = m4ri_b200.load_library()
torch.cuda.init()
ts = torch.cuda.Stream()
torch.cuda.set_stream(ts)
sh = ctypes.c_void_p(ts.cuda_stream)


def dev_random(rows, cols):
    pitch = cols // 64
    t = torch.randint(-2**62, 2**62, (rows, pitch), dtype=torch.int64, device="cuda")
    t ^= torch.randint(-2**62, 2**62, (rows, pitch), dtype=torch.int64, device="cuda") << 2
    return t, lib.m4ri_b200_dmat_wrap(t.data_ptr(), pitch, rows, cols)


for m, n in [(16384, 16384), (65536, 65536)]:
    tT, dT = dev_random(m, m)
    tB, dB = dev_random(m, n)
    for upper, left in ((0, 1), (1, 1), (0, 0), (1, 0)):
        for _ in range(2):
            lib.m4ri_b200_dtrsm(dT, dB, upper, left, 0, sh)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = lib.m4ri_b200_kernel_launches()
        e0.record()
        for _ in range(3):
            lib.m4ri_b200_dtrsm(dT, dB, upper, left, 0, sh)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(f"device trsm_{'upper' if upper else 'lower'}_{'left' if left else 'right'} m={m} n={n}: {ms:.2f} ms "
              f"({1.0*m*m*n/ms/1e9:.0f} T bit-ops/s nominal m^2 n) launches/call {(lib.m4ri_b200_kernel_launches()-l0)//3}", flush=True)
    del tT, tB

m = n = 16384
T, B = H.new(m, m), H.new(m, n)
rng = np.random.default_rng(1)
for M in (T, B):
    st = H.storage(M)
    st[:, :] = rng.integers(0, 2**64, size=st.shape, dtype=np.uint64)
Bc = H.clone(B)
for _ in range(2):
    lib.mzd_trsm_lower_left(T, B, 0)
t0 = time.perf_counter(); lib.mzd_trsm_lower_left(T, B, 0); t1 = time.perf_counter()
print(f"host API mzd_trsm_lower_left {m}x{n} (pageable): {(t1-t0)*1e3:.1f} ms")
R = H.ref()
if R is not None:
    # proper unit lower triangular input for the reference
    st = H.storage(T)
    for i in range(m):
        w, b = divmod(i, 64)
        st[i, w] = (int(st[i, w]) & ((1 << b) - 1)) | (1 << b)
        st[i, w + 1:] = 0
    t0 = time.perf_counter(); R.mzd_trsm_lower_left(T, Bc, 0); t1 = time.perf_counter()
    print(f"reference CPU mzd_trsm_lower_left {m}x{n} (1 thread): {(t1-t0)*1e3:.1f} ms")
