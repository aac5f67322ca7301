"""GPU tests (-m gpu) of the device-resident PLE decomposition (csrc/ple.cu) through the reference-named entry
points mzd_ple / _mzd_ple: factored matrix, P, rank and Q[0..rank) bit for bit against the oracle (orc_ple, pinned
to the reference by tests/test_ple_oracle.py) and, at sizes the scalar oracle is slow for, against the compiled
reference itself.  Shapes and structure follow the reference's tests/test_ple.c / test_pluq.c (random, half-rank,
structured inputs) plus the cases that exercise the pivot rule across strips and recursion levels."""
import ctypes
from ctypes import POINTER, c_int

import numpy as np
import pytest

import m4ri_b200
from m4ri_b200 import MzpT
from tests import harness as H
from tests.test_ple_oracle import oracle_ple, same_ple, structured

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    L = m4ri_b200.load_library()
    assert L.m4ri_b200_device_count() > 0, "no CUDA device: the product has no CPU fallback"
    return L


def device_ple(lib, A0, fn="mzd_ple"):
    A = H.clone(A0)
    m, n = A0.contents.nrows, A0.contents.ncols
    pv, qv = (c_int * max(m, 1))(), (c_int * max(n, 1))()
    P, Q = MzpT(ctypes.cast(pv, POINTER(c_int)), m), MzpT(ctypes.cast(qv, POINTER(c_int)), n)
    launches0 = lib.m4ri_b200_kernel_launches()
    r = getattr(lib, fn)(A, ctypes.byref(P), ctypes.byref(Q), 0)
    assert lib.m4ri_b200_kernel_launches() > launches0 or m == 0 or n == 0
    out = (r, H.storage(A).copy(), np.array(pv[:m]), np.array(qv[:n]))
    H.free(A)
    return out


@pytest.mark.parametrize("m,n", [(1, 1), (64, 64), (65, 129), (128, 128), (200, 300), (300, 200), (513, 511), (1000, 1000),
                                 (1500, 257), (257, 1500), (2100, 2050)])
@pytest.mark.parametrize("kind", ["random", "lowrank", "zerobands", "bottom"])
def test_ple_matches_oracle(lib, m, n, kind):
    A = structured(m, n, kind, 5 * m + n)
    want = oracle_ple(A)
    got = device_ple(lib, A, "mzd_ple" if (m + n) % 2 else "_mzd_ple")
    assert got[0] == want[0], "rank"
    assert same_ple(want, got), (m, n, kind)
    H.free(A)


@pytest.mark.parametrize("m,n,kind", [(4096, 4096, "random"), (5000, 3000, "lowrank"), (3000, 6000, "zerobands"),
                                      (8192, 8192, "random"), (9000, 4100, "bottom")])
def test_ple_matches_compiled_reference(lib, m, n, kind):
    if H.ref() is None:
        pytest.skip("oracle/_ref/libm4ri_ref.so not present")
    from tests.test_ple_reference_canonical import _ref, _run
    A = structured(m, n, kind, m + 7 * n)
    want = _run(_ref(), "mzd_ple", A, 0)
    got = device_ple(lib, A)
    assert same_ple(want, got), (m, n, kind)
    H.free(A)


def test_ple_of_identity_zero_and_permutation(lib):
    for n in (1, 64, 130, 700):
        Z = H.new(n, n)
        r, st, P, Q = device_ple(lib, Z)
        assert r == 0 and not st.any() and np.array_equal(P, np.arange(n))
        I = H.new(n, n)
        for i in range(n):
            H.storage(I)[n - 1 - i, i // 64] |= np.uint64(1) << np.uint64(i % 64)    # anti-diagonal: every pivot needs a swap
        want = oracle_ple(I)
        assert same_ple(want, device_ple(lib, I))
        H.free(Z, I)
