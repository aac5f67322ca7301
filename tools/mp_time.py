"""Time the in-process multi-GPU entry point mzd_mul_mp (host mzd_t in, host mzd_t out)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import m4ri_b200  # noqa: E402

lib = m4ri_b200.load_library()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
G = int(sys.argv[2]) if len(sys.argv) > 2 else lib.m4ri_b200_device_count()
rng = np.random.default_rng(0)


def host(r, c):
    M = m4ri_b200.mzd_init(r, c)
    w = m4ri_b200.words(M)
    step = max(1, (1 << 24) // w.shape[1])
    for i in range(0, r, step):
        w[i:i + step] = rng.integers(0, 2**64, size=w[i:i + step].shape, dtype=np.uint64)
    return M


A, B, C = host(n, n), host(n, n), m4ri_b200.mzd_init(n, n)
for g in sorted({1, G}):
    lib.m4ri_b200_set_num_devices(g)
    lib.mzd_mul_mp(C, A, B, 0)
    t0 = time.perf_counter()
    iters = 3
    for _ in range(iters):
        lib.mzd_mul_mp(C, A, B, 0)
    dt = (time.perf_counter() - t0) / iters
    print(f"mzd_mul_mp {n}^3 on {g} GPU(s), pageable host matrices: {dt*1e3:.1f} ms  {2.0*n**3/dt:.3e} bit-ops/s "
          f"path={lib.m4ri_b200_last_path().decode()}", flush=True)
