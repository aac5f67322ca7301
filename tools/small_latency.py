"""Latency of small host-API products (the calls an L4 caller such as PLE/TRSM makes)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import m4ri_b200  # noqa: E402
from tests import harness as H  # noqa: E402

lib = m4ri_b200.load_library()
for n in (64, 256, 1024, 2048, 4096):
    H.libc.srandom(1)
    A, B, C = H.random_matrix(n, n), H.random_matrix(n, n), H.new(n, n)
    for fn in ("mzd_mul", "mzd_addmul"):
        f = getattr(lib, fn)
        for _ in range(5):
            f(C, A, B, 0)
        t0 = time.perf_counter()
        iters = 200 if n <= 1024 else 50
        for _ in range(iters):
            f(C, A, B, 0)
        dt = (time.perf_counter() - t0) / iters
        print(f"{fn} {n}^3: {dt*1e6:8.1f} us/call", flush=True)
    H.free(A, B, C)
