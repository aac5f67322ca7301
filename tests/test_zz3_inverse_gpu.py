"""GPU parity of m4ri_b200_inv_m4ri (the reference's mzd_inv_m4ri, m4ri/brilliantrussian.c:971-997, on the device
RREF): bit-exact against the compiled reference where present, and A * A^-1 == I through the oracle.  Sizes: the
reference's tests/test_invert.c."""
import numpy as np
import pytest

import m4ri_b200
from tests import harness as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    L = m4ri_b200.load_library()
    assert L.m4ri_b200_device_count() > 0, "no CUDA device: the product has no CPU fallback"
    return L


def _invertible(n):
    """random invertible matrix: product of a unit lower and a unit upper triangular random matrix"""
    L, U = H.random_matrix(n, n), H.random_matrix(n, n)
    sl, su = H.storage(L), H.storage(U)
    for i in range(n):
        w, b = divmod(i, 64)
        sl[i, w] = (int(sl[i, w]) & ((1 << b) - 1)) | (1 << b)         # keep bits below the diagonal, set it
        sl[i, w + 1:] = 0
        su[i, :w] = 0
        su[i, w] = ((int(su[i, w]) >> b) << b) | (1 << b)               # keep bits from the diagonal on, set it
    su[:, L.contents.width - 1] &= np.uint64(L.contents.high_bitmask)
    A = H.oracle().orc_mul(None, L, U, 0)
    H.free(L, U)
    return A


@pytest.mark.parametrize("n", [1, 2, 3, 21, 64, 128, 193, 1000, 1024, 1025, 1290, 1710, 2048, 2065])
def test_inverse(lib, n):
    H.libc.srandom(100 + n)
    A = _invertible(n)
    B = lib.m4ri_b200_inv_m4ri(None, A)
    P = H.oracle().orc_mul(None, A, B, 0)
    sp = H.storage(P)
    for i in range(n):                                   # A * B == I
        w, b = divmod(i, 64)
        assert int(sp[i, w]) == 1 << b and not sp[i, :w].any() and not sp[i, w + 1:].any()
    ref = H.ref()
    if ref is not None:
        R = ref.mzd_inv_m4ri(None, A, 0)
        assert np.array_equal(H.storage(R)[:, :A.contents.width], H.storage(B)[:, :A.contents.width])
        ref.mzd_free(R)
    lib.m4ri_b200_mzd_free(B)
    H.free(A, P)


def test_singular_input_gives_the_same_block_as_the_reference(lib):
    ref = H.ref()
    if ref is None:
        pytest.skip("needs the compiled reference")
    H.libc.srandom(7)
    X, Y = H.random_matrix(300, 100), H.random_matrix(100, 300)
    A = H.oracle().orc_mul(None, X, Y, 0)                # rank <= 100
    B = lib.m4ri_b200_inv_m4ri(None, A)
    R = ref.mzd_inv_m4ri(None, A, 0)
    assert np.array_equal(H.storage(R)[:, :A.contents.width], H.storage(B)[:, :A.contents.width])
    ref.mzd_free(R)
    H.free(X, Y, A)
