"""CPU, reference only: design input for a device PLE/PLUQ.

The reference's PLE / PLUQ results (the factored matrix, both permutations and the rank) do NOT depend on the
algorithm variant — naive, "russian" with any k, recursive with any cutoff all agree bit for bit — because every
variant honours the same pivot rule: leftmost column, first row in the CURRENT (already swapped) row order.
A device implementation therefore cannot pick "any row with the bit" (as the device RREF does, whose result is
unique anyway): it has to reproduce that rule to stay bit-exact.  This test pins the property on the reference."""
import numpy as np
import pytest
from ctypes import POINTER, Structure, c_int

from m4ri_b200 import MzdP
from tests import harness as H


class Mzp(Structure):                      # m4ri/mzp.h:37-49
    _fields_ = [("values", POINTER(c_int)), ("length", c_int)]


def _ref():
    ref = H.ref(required=True)
    ref.mzp_init.argtypes, ref.mzp_init.restype = [c_int], POINTER(Mzp)
    ref.mzp_free.argtypes = [POINTER(Mzp)]
    for name, extra in (("mzd_ple", [c_int]), ("_mzd_ple", [c_int]), ("_mzd_ple_naive", []), ("_mzd_ple_russian", [c_int]),
                        ("mzd_pluq", [c_int]), ("_mzd_pluq_naive", []), ("_mzd_pluq_russian", [c_int])):
        f = getattr(ref, name)
        f.argtypes, f.restype = [MzdP, POINTER(Mzp), POINTER(Mzp)] + extra, c_int
    return ref


def _run(ref, name, A0, *extra):
    A = H.clone(A0)
    m, n = A0.contents.nrows, A0.contents.ncols
    P, Q = ref.mzp_init(m), ref.mzp_init(n)
    r = getattr(ref, name)(A, P, Q, *extra)
    out = (r, H.storage(A).copy(), np.array(P.contents.values[:m]), np.array(Q.contents.values[:n]))
    ref.mzp_free(P)
    ref.mzp_free(Q)
    H.free(A)
    return out


def _same(a, b):
    return a[0] == b[0] and all(np.array_equal(x, y) for x, y in zip(a[1:], b[1:]))


@pytest.mark.parametrize("m,n,rank", [(64, 64, 0), (200, 300, 0), (300, 200, 0), (1000, 1000, 0), (700, 900, 100)])
def test_every_variant_gives_the_same_factorisation(m, n, rank):
    ref = _ref()
    H.libc.srandom(m + n + rank)
    if rank:
        X, Y = H.random_matrix(m, rank), H.random_matrix(rank, n)
        A = H.oracle().orc_mul(None, X, Y, 0)
        H.free(X, Y)
    else:
        A = H.random_matrix(m, n)
    ple = _run(ref, "_mzd_ple_naive", A)
    for name, extra in (("mzd_ple", (0,)), ("_mzd_ple", (64,)), ("_mzd_ple_russian", (0,)), ("_mzd_ple_russian", (4,))):
        assert _same(ple, _run(ref, name, A, *extra)), name
    pluq = _run(ref, "_mzd_pluq_naive", A)
    for name, extra in (("mzd_pluq", (0,)), ("_mzd_pluq_russian", (0,))):
        assert _same(pluq, _run(ref, name, A, *extra)), name
    H.free(A)
