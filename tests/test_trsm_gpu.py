"""GPU parity tests of the widened row "triangular solve, left variants" (m4ri/triangular.c:394-516):
the C-ABI entry points mzd_trsm_lower_left / mzd_trsm_upper_left on HOST mzd_t operands against the
oracle, on the reference's own shapes (tests/test_trsm.c: windows with offsets) and at sizes where
only a property check (T * X == B through independent numpy mat-vec products) is affordable."""
import numpy as np
import pytest

import m4ri_b200
from tests import harness as H
from tests.test_trsm_oracle import SHAPES, make_case
from tests.test_parity_gpu import _fill_fast, _gf2_matvec_rows, _pack

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    L = m4ri_b200.load_library()
    assert L.m4ri_b200_device_count() > 0
    return L


@pytest.mark.parametrize("shape", SHAPES + [(129, 200, 0, 0), (1000, 1000, 0, 0), (2500, 3000, 64, 64), (4096, 5000, 0, 0)])
@pytest.mark.parametrize("side", ["lower", "upper"])
@pytest.mark.parametrize("hand", ["left", "right"])
def test_trsm_matches_oracle(lib, shape, side, hand):
    m, n, off_t, off_b = shape
    if hand == "right" and n > 3000:
        n = 3000        # the oracle's right solve is O(m n^2 / 64) row operations
    Tbase, Bbase, T, B = make_case(m, n, off_t, off_b, 7 + m, hand == "right")
    want_base = H.clone(Bbase)
    H.storage(want_base)[:, :] = H.storage(Bbase)
    want = H.window(want_base, 1, off_b, 1 + m, off_b + n)
    launches = lib.m4ri_b200_kernel_launches()
    getattr(H.oracle(), f"orc_trsm_{side}_{hand}")(T, want)
    getattr(lib, f"mzd_trsm_{side}_{hand}")(T, B, 0)
    assert lib.m4ri_b200_last_path().decode() == f"trsm_{side}_{hand}"
    assert lib.m4ri_b200_kernel_launches() > launches
    assert np.array_equal(H.storage(Bbase), H.storage(want_base))   # incl. every bit outside the window
    H.free(T, B, want, Tbase, Bbase, want_base)


def _effective_triangle(Tw, m, side):
    """strict triangle of the bit-packed square matrix Tw plus the unit diagonal"""
    rows = np.arange(m)
    colword = (np.arange(Tw.shape[1], dtype=np.int64) * 64)[None, :]
    rel = rows[:, None] - colword                 # diagonal position relative to each word
    if side == "lower":   # keep bits j <= i
        full = rel >= 63
        part = (rel >= 0) & (rel < 63)
        mask = np.where(full, np.uint64(2**64 - 1), np.uint64(0))
        mask = np.where(part, (np.uint64(2) << np.clip(rel, 0, 62).astype(np.uint64)) - np.uint64(1), mask)
    else:                 # keep bits j >= i
        full = rel < 0
        part = (rel >= 0) & (rel <= 63)
        mask = np.where(full, np.uint64(2**64 - 1), np.uint64(0))
        mask = np.where(part, ~((np.uint64(1) << np.clip(rel, 0, 63).astype(np.uint64)) - np.uint64(1)), mask)
    Teff = Tw & mask
    Teff[rows, rows // 64] |= np.uint64(1) << (rows % 64).astype(np.uint64)
    return Teff


@pytest.mark.parametrize("m,n", [(16384, 16384), (20000, 4100)])
@pytest.mark.parametrize("side", ["lower", "upper"])
@pytest.mark.parametrize("hand", ["left", "right"])
def test_trsm_large_property(lib, m, n, side, hand):
    """left: T (X v) == B0 v;  right: X (T v) == B0 v, for 24 random vectors v (independent numpy mat-vecs)."""
    t = m if hand == "left" else n
    T, B = H.new(t, t), H.new(m, n)
    _fill_fast(T, 5); _fill_fast(B, 6)
    B0 = m4ri_b200.valid_words(B)
    getattr(lib, f"mzd_trsm_{side}_{hand}")(T, B, 0)
    Teff = _effective_triangle(m4ri_b200.valid_words(T), t, side)
    Xw = m4ri_b200.valid_words(B)
    rng = np.random.default_rng(9)
    for _ in range(24):
        v = rng.integers(0, 2, size=n).astype(bool)
        if hand == "left":
            lhs = _gf2_matvec_rows(Teff, _pack(_gf2_matvec_rows(Xw, _pack(v))))
        else:
            lhs = _gf2_matvec_rows(Xw, _pack(_gf2_matvec_rows(Teff, _pack(v))))
        assert np.array_equal(lhs, _gf2_matvec_rows(B0, _pack(v)))
    H.free(T, B)
