"""Small end-to-end cases for compute-sanitizer (memcheck / racecheck / synccheck): every kernel of the
library on shapes with ragged edges.  Checks results against the oracle as well."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import m4ri_b200  # noqa: E402
from tests import harness as H  # noqa: E402

lib = m4ri_b200.load_library()
O = H.oracle()
H.libc.srandom(3)
ok = True
for (m, l, n, cutoff) in [(70, 130, 200, 0), (300, 300, 300, 0), (1030, 200, 1100, 0), (1100, 1300, 1050, 256)]:
    A, B, C = H.random_matrix(m, l), H.random_matrix(l, n), H.random_matrix(m, n)
    want = O.orc_addmul(H.clone(C), A, B, 0)
    lib.mzd_addmul(C, A, B, cutoff)
    ok &= H.equal(C, want)
    P = lib.mzd_mul(None, A, B, cutoff)
    ok &= H.equal(P, O.orc_mul(None, A, B, 0))
    print("mul", m, l, n, cutoff, lib.m4ri_b200_last_path().decode(), "OK" if ok else "BAD", flush=True)
for (m, n) in [(200, 300), (700, 260)]:
    for side in ("lower", "upper"):
        for hand in ("left", "right"):
            t = m if hand == "left" else n
            T, B = H.random_matrix(t, t), H.random_matrix(m, n)
            want = H.clone(B)
            getattr(O, f"orc_trsm_{side}_{hand}")(T, want)
            getattr(lib, f"mzd_trsm_{side}_{hand}")(T, B, 0)
            ok &= H.equal(B, want)
    print("trsm", m, n, "OK" if ok else "BAD", flush=True)
for (m, n) in [(65, 63), (700, 1300)]:
    A = H.random_matrix(m, n)
    T = lib.m4ri_b200_transpose(None, A)
    ok &= H.equal(T, O.orc_transpose(None, A))
    print("transpose", m, n, "OK" if ok else "BAD", flush=True)
# round 2: the tensor-core leaf (bulk copies, mbarriers, TMEM, red.xor epilogue) alone and under one Strassen level
for (m, l, n, cutoff) in [(256, 1024, 256, 0), (384, 2048, 512, 0), (1024, 4096, 1024, 512)]:
    A, B = H.random_matrix(m, l), H.random_matrix(l, n)
    P = lib.mzd_mul(None, A, B, cutoff) if cutoff else lib.mzd_mul_m4rm(None, A, B, 0)
    ok &= H.equal(P, O.orc_mul(None, A, B, 0)) and lib.m4ri_b200_last_leaf_variant() == 3
    print("tensor leaf", m, l, n, cutoff, lib.m4ri_b200_last_path().decode(), "OK" if ok else "BAD", flush=True)
# round 2: the tall-tile leaf with the hybrid partition and store mode under a two-level Strassen node with the fused
# two-level additions (8192^3 at cutoff 2048 with the tall leaf forced: 49 products of 2048^3 in one launch), and the
# accumulating form; checked by Freivalds (the scalar oracle is too slow under the sanitizer at this size)
if os.environ.get("SANITIZE_BIG", "1") == "1":
    prev = lib.m4ri_b200_set_leaf_variant(2)
    n = 8192
    rng = np.random.default_rng(4)
    A, B, C = H.new(n, n), H.new(n, n), H.new(n, n)
    for M in (A, B):
        H.storage(M)[:, :] = rng.integers(0, 2**64, size=H.storage(M).shape, dtype=np.uint64)
    lib.mzd_mul(C, A, B, 2048)
    path = lib.m4ri_b200_last_path().decode()
    Aw, Bw, Cw = (m4ri_b200.valid_words(M) for M in (A, B, C))

    def matvec(M, x):                       # bit-packed M [rows, words] times packed vector x -> packed result
        anded = M & x[None, :]
        f = np.bitwise_xor.reduce(anded, axis=1)
        for sft in (32, 16, 8, 4, 2, 1):
            f ^= f >> np.uint64(sft)
        bits = (f & np.uint64(1)).astype(np.uint8)
        return np.packbits(bits.reshape(-1, 64)[:, ::-1], axis=1, bitorder="big").view(">u8").astype(np.uint64).ravel()

    for _ in range(8):
        x = rng.integers(0, 2**64, size=n // 64, dtype=np.uint64)
        ok &= bool(np.array_equal(matvec(Aw, matvec(Bw, x)), matvec(Cw, x)))
    lib.mzd_addmul(C, A, B, 2048)           # C ^= A*B -> zero
    ok &= not H.storage(C).any()
    lib.m4ri_b200_set_leaf_variant(prev)
    print("big", n, path, "OK" if ok else "BAD", flush=True)
    # PLE with the cluster strip kernel
    import ctypes
    from ctypes import POINTER, c_int
    from m4ri_b200 import MzpT
    E = H.random_matrix(1500, 1100)
    F = H.clone(E)
    pw, qw = (c_int * 1500)(), (c_int * 1100)()
    rw = O.orc_ple(F, pw, qw)
    pv, qv = (c_int * 1500)(), (c_int * 1100)()
    r = lib.mzd_ple(E, ctypes.byref(MzpT(ctypes.cast(pv, POINTER(c_int)), 1500)), ctypes.byref(MzpT(ctypes.cast(qv, POINTER(c_int)), 1100)), 0)
    ok &= r == rw and H.equal(E, F) and list(pv) == list(pw)
    print("ple", r, "OK" if ok else "BAD", flush=True)
print("ALL OK" if ok else "FAILURES")
sys.exit(0 if ok else 1)
