"""Time the device RREF (m4ri_b200_dechelonize) on device-resident random matrices; optionally the reference's
mzd_echelonize_m4ri(A, 1, 0) on the host cores for one size (oracle/_ref, test infrastructure)."""
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
    torch.cuda.synchronize()
    lib.m4ri_b200_dechelonize(d, 1, None)            # warm-up (workspace, tensor maps)
    t.copy_(keep)
    torch.cuda.synchronize()
    l0 = lib.m4ri_b200_kernel_launches()
    t0 = time.perf_counter()
    r = lib.m4ri_b200_dechelonize(d, 1, None)
    dt = time.perf_counter() - t0
    print(f"rref {n}x{n}: rank {r}  {dt*1e3:.1f} ms  {lib.m4ri_b200_kernel_launches() - l0} launches", flush=True)


for a in sys.argv[1:]:
    if a.startswith("ref:"):
        from tests import harness as H
        n = int(a[4:])
        ref = H.ref(required=True)
        H.libc.srandom(1)
        A = H.random_matrix(n, n)
        t0 = time.perf_counter()
        r = ref.mzd_echelonize_m4ri(A, 1, 0)
        print(f"reference mzd_echelonize_m4ri {n}x{n}: rank {r}  {(time.perf_counter() - t0)*1e3:.1f} ms (1 core)", flush=True)
    else:
        run(int(a))
