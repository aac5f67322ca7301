"""CPU: the oracle's PLE restatement (oracle/m4rm_oracle.c: orc_ple, after m4ri/ple.c:222-272) against the
unmodified reference's mzd_ple / _mzd_ple_naive: factored matrix, P, rank and Q[0..rank) bit for bit.  The
reference's own tests/test_ple.c only checks A == P*L*E; the canonical form is pinned by
tests/test_ple_reference_canonical.py."""
import numpy as np
import pytest
from ctypes import POINTER, c_int

from tests import harness as H
from tests.test_ple_reference_canonical import _ref, _run


def structured(m, n, kind, seed):
    """inputs that exercise the pivot rule: random, low rank, zero column bands (pivot columns not contiguous:
    r1 < n1 with r2 > 0 in the recursive splits), and a matrix whose first pivot rows sit at the bottom"""
    H.libc.srandom(seed)
    if kind == "random":
        return H.random_matrix(m, n)
    if kind == "lowrank":
        r = max(1, min(m, n) // 7)
        X, Y = H.random_matrix(m, r), H.random_matrix(r, n)
        A = H.oracle().orc_mul(None, X, Y, 0)
        H.free(X, Y)
        return A
    A = H.random_matrix(m, n)
    st = H.storage(A)
    if kind == "zerobands":
        for w in range(0, A.contents.width, 3):
            st[:, w] &= np.uint64(0xFFFF00000000FFFF)
        st[: m // 2, 0] = 0
    elif kind == "bottom":
        st[: (3 * m) // 4, : max(1, A.contents.width // 2)] = 0
    return A


def oracle_ple(A0):
    A = H.clone(A0)
    m, n = A0.contents.nrows, A0.contents.ncols
    P, Q = (c_int * m)(), (c_int * n)()
    r = H.oracle().orc_ple(A, P, Q)
    out = (r, H.storage(A).copy(), np.array(P[:m]), np.array(Q[:n]))
    H.free(A)
    return out


def same_ple(a, b):
    r = a[0]
    return a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]) and np.array_equal(a[3][:r], b[3][:r])


@pytest.mark.parametrize("m,n", [(1, 1), (64, 64), (65, 129), (200, 300), (300, 200), (513, 511), (1000, 1000)])
@pytest.mark.parametrize("kind", ["random", "lowrank", "zerobands", "bottom"])
def test_orc_ple_matches_reference(m, n, kind):
    ref = _ref()
    A = structured(m, n, kind, 3 * m + n)
    want = _run(ref, "mzd_ple", A, 0)
    assert same_ple(want, _run(ref, "_mzd_ple_naive", A))
    assert same_ple(want, oracle_ple(A)), (m, n, kind)
    H.free(A)
