"""Quick on-GPU diagnostic (not a test, not the bench): correctness probes with mismatch
statistics, then rough timings of the leaf and of Strassen via the device-resident API."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import m4ri_b200  # noqa: E402
from tests import harness as H  # noqa: E402

lib = m4ri_b200.load_library()
print("devices:", lib.m4ri_b200_device_count(), flush=True)


def fill(M, seed):
    rng = np.random.default_rng(seed)
    st = H.storage(M)
    st[:, :] = rng.integers(0, 2**64, size=st.shape, dtype=np.uint64)
    w = M.contents.width
    st[:, w - 1] &= np.uint64(M.contents.high_bitmask)
    st[:, w:] = 0


def check(m, l, n, fn="mzd_mul_m4rm", arg=0):
    A, B = H.new(m, l), H.new(l, n)
    fill(A, 1); fill(B, 2)
    C = H.new(m, n)
    getattr(lib, fn)(C, A, B, arg)
    want = H.oracle().orc_mul_m4rm(None, A, B, 0, 1)
    g, w = m4ri_b200.valid_words(C), m4ri_b200.valid_words(want)
    diff = g ^ w
    bad = np.argwhere(diff != 0)
    status = "OK" if len(bad) == 0 else f"MISMATCH words={len(bad)} rows={len(set(bad[:,0]))} first={bad[:5].tolist()}"
    print(f"{fn}({m}x{l}x{n}, {arg}) path={lib.m4ri_b200_last_path().decode()}: {status}", flush=True)
    if len(bad):
        r, c = bad[0]
        print(f"   got {int(g[r,c]):016x} want {int(w[r,c]):016x}", flush=True)
    H.free(A, B, C, want)
    return len(bad) == 0


ok = True
for shape in [(64, 64, 64), (16, 16, 128), (256, 128, 1024), (300, 300, 300), (1024, 1024, 1024), (1030, 200, 1100),
              (2048, 2048, 4096), (1500, 3000, 2500)]:
    ok &= check(*shape)
for shape, cut in [((1024, 1024, 1024), 256), ((2048, 2048, 4096), 1024), ((1710, 1290, 1000), 256)]:
    ok &= check(*shape, fn="mzd_mul", arg=cut)
print("ALL OK" if ok else "FAILURES", flush=True)


def time_dev(m, l, n, cutoff, iters=5, leaf_only=False):
    dA, dB, dC = lib.m4ri_b200_dmat_alloc(m, l), lib.m4ri_b200_dmat_alloc(l, n), lib.m4ri_b200_dmat_alloc(m, n)
    # random device contents: upload one random row-block repeatedly is slow; use a host matrix of full size
    A, B = H.new(m, l), H.new(l, n)
    fill(A, 5); fill(B, 6)
    lib.m4ri_b200_upload(dA, A, None); lib.m4ri_b200_upload(dB, B, None)
    H.free(A, B)
    def run():
        if leaf_only:
            lib.m4ri_b200_dmul_m4rm(dC, dA, dB, 1, None)
        else:
            lib.m4ri_b200_dmul(dC, dA, dB, cutoff, 1, None)
    run(); lib.m4ri_b200_sync(None)
    t0 = time.perf_counter()
    for _ in range(iters):
        run()
    lib.m4ri_b200_sync(None)
    dt = (time.perf_counter() - t0) / iters
    print(f"dev {m}x{l}x{n} cutoff={'leaf' if leaf_only else cutoff} path={lib.m4ri_b200_last_path().decode()}: "
          f"{dt*1e3:.3f} ms  {2.0*m*l*n/dt:.3e} bit-ops/s", flush=True)
    for d in (dA, dB, dC):
        lib.m4ri_b200_dmat_free(d)


if ok or os.environ.get("FORCE_TIMING"):
    time_dev(4096, 4096, 4096, 0, leaf_only=True)
    time_dev(8192, 8192, 8192, 0, leaf_only=True)
    time_dev(16384, 16384, 16384, 0, leaf_only=True)
    for cut in (4096, 8192):
        time_dev(16384, 16384, 16384, cut)
    if os.environ.get("BIG"):
        time_dev(32768, 32768, 32768, 0, iters=2, leaf_only=True)
        for cut in (8192, 16384):
            time_dev(32768, 32768, 32768, cut, iters=2)
