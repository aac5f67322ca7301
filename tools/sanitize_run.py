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
print("ALL OK" if ok else "FAILURES")
sys.exit(0 if ok else 1)
