"""Run the M4RM leaf kernel a few times on device-resident random matrices (for ncu)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import m4ri_b200  # noqa: E402

lib = m4ri_b200.load_library()
m = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
l = int(sys.argv[2]) if len(sys.argv) > 2 else m
n = int(sys.argv[3]) if len(sys.argv) > 3 else m
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 3
cutoff = int(sys.argv[5]) if len(sys.argv) > 5 else -1   # -1: leaf only

rng = np.random.default_rng(0)


def host(r, c):
    M = m4ri_b200.mzd_init(r, c)
    w = m4ri_b200.words(M)
    w[:, :] = rng.integers(0, 2**64, size=w.shape, dtype=np.uint64)
    w[:, -1] &= np.uint64(M.contents.high_bitmask)
    return M


A, B = host(m, l), host(l, n)
dA, dB, dC = lib.m4ri_b200_dmat_alloc(m, l), lib.m4ri_b200_dmat_alloc(l, n), lib.m4ri_b200_dmat_alloc(m, n)
lib.m4ri_b200_upload(dA, A, None)
lib.m4ri_b200_upload(dB, B, None)
for _ in range(iters):
    if cutoff < 0:
        lib.m4ri_b200_dmul_m4rm(dC, dA, dB, 1, None)
    else:
        lib.m4ri_b200_dmul(dC, dA, dB, cutoff, 1, None)
lib.m4ri_b200_sync(None)
print("done", lib.m4ri_b200_last_path().decode(), lib.m4ri_b200_kernel_launches())
