"""CPU tests: pin the oracle's triangular solves (left variants) against the compiled reference on the
reference's own shapes (tests/test_trsm.c) — windows with offsets, garbage below/above the triangle."""
import numpy as np
import pytest

from tests import harness as H

needs_ref = pytest.mark.skipif(H.ref() is None, reason="oracle/_ref not built")

# (m = order of the triangular matrix = rows of B, n = columns of B, offsetT, offsetB) — test_trsm.c
SHAPES = [(10, 20, 0, 0), (10, 80, 0, 0), (70, 20, 0, 0), (70, 80, 0, 0), (53, 53, 0, 0), (54, 54, 0, 0),
          (63, 63, 0, 0), (64, 64, 0, 0), (65, 65, 0, 0), (57, 150, 64, 0), (57, 80, 0, 64), (770, 1600, 64, 128),
          (1764, 1345, 128, 64), (1, 1, 0, 0), (2, 300, 0, 0)]


def make_case(m, n, off_t, off_b, seed, right=False):
    """left: T is m x m, B is m x n.  right: T is n x n, B is m x n."""
    H.libc.srandom(seed)
    t = n if right else m
    Tbase, Bbase = H.random_matrix(t + 3, t + off_t + 64), H.random_matrix(m + 3, n + off_b + 64)
    T = H.window(Tbase, 1, off_t, 1 + t, off_t + t)      # random bits everywhere: diagonal and the other
    B = H.window(Bbase, 1, off_b, 1 + m, off_b + n)      # triangle must be ignored by the solver
    return Tbase, Bbase, T, B


@needs_ref
@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("side", ["lower", "upper"])
@pytest.mark.parametrize("hand", ["left", "right"])
def test_oracle_trsm_matches_compiled_reference(shape, side, hand):
    m, n, off_t, off_b = shape
    right = hand == "right"
    O, R = H.oracle(), H.ref()
    Tbase, Bbase, T, B = make_case(m, n, off_t, off_b, 100 + m + n, right)
    m_t = n if right else m
    # the reference requires a proper triangular matrix with unit diagonal: give it one, but keep the
    # garbage for the oracle to prove the garbage is ignored
    Tclean_base = H.clone(Tbase)
    H.storage(Tclean_base)[:, :] = H.storage(Tbase)
    Tclean = H.window(Tclean_base, 1, off_t, 1 + m_t, off_t + m_t)
    tw = H.storage(Tclean_base)
    for i in range(m_t):
        for j in range(m_t):
            keep = (j < i) if side == "lower" else (j > i)
            if not keep:
                col = off_t + j
                w, b = col // 64, col % 64
                val = 1 if i == j else 0
                tw[1 + i, w] = (int(tw[1 + i, w]) & ~(1 << b)) | (val << b)
    Bref_base = H.clone(Bbase)
    H.storage(Bref_base)[:, :] = H.storage(Bbase)
    Bref = H.window(Bref_base, 1, off_b, 1 + m, off_b + n)
    getattr(O, f"orc_trsm_{side}_{hand}")(T, B)
    getattr(R, f"mzd_trsm_{side}_{hand}")(Tclean, Bref, 0)
    assert np.array_equal(H.storage(Bbase), H.storage(Bref_base))   # also: nothing outside the window moved
    H.free(T, B, Tclean, Bref, Tbase, Bbase, Tclean_base, Bref_base)
