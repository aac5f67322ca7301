"""Time the device PLE (m4ri_b200_dple) on device-resident random matrices; `ref:N` times the reference's mzd_ple
on the host (oracle/_ref, test infrastructure, one core — the reference's PLE has no OpenMP path)."""
import ctypes
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import m4ri_b200  # noqa: E402

lib = m4ri_b200.load_library()
torch.cuda.init()


def run(n):
    g = torch.Generator(device="cuda").manual_seed(n)
    t = torch.randint(-2**62, 2**62, (n, n // 64), dtype=torch.int64, device="cuda", generator=g)
    t ^= torch.randint(-2**62, 2**62, (n, n // 64), dtype=torch.int64, device="cuda", generator=g) << 2
    keep = t.clone()
    d = lib.m4ri_b200_dmat_wrap(t.data_ptr(), n // 64, n, n)
    P, Q = (ctypes.c_int * n)(), (ctypes.c_int * n)()
    torch.cuda.synchronize()
    lib.m4ri_b200_dple(d, P, Q, None)                # warm-up (workspace, tensor maps)
    t.copy_(keep)
    torch.cuda.synchronize()
    l0 = lib.m4ri_b200_kernel_launches()
    t0 = time.perf_counter()
    r = lib.m4ri_b200_dple(d, P, Q, None)
    dt = time.perf_counter() - t0
    print(f"ple {n}x{n}: rank {r}  {dt*1e3:.1f} ms  {lib.m4ri_b200_kernel_launches() - l0} launches", flush=True)


for a in sys.argv[1:]:
    if a.startswith("ref:"):
        from tests import harness as H
        from tests.test_ple_reference_canonical import _ref
        n = int(a[4:])
        ref = _ref()
        H.libc.srandom(1)
        A = H.random_matrix(n, n)
        P, Q = ref.mzp_init(n), ref.mzp_init(n)
        t0 = time.perf_counter()
        r = ref.mzd_ple(A, P, Q, 0)
        print(f"reference mzd_ple {n}x{n}: rank {r}  {(time.perf_counter() - t0)*1e3:.1f} ms (1 core)", flush=True)
    else:
        run(int(a))
