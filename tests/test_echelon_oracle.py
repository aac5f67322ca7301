"""CPU: the oracle's echelon form (restating m4ri/mzd.c:208-233) pinned against the compiled reference —
its M4RI (k = auto and k = 8), naive and PLUQ variants all agree on the reduced row echelon form, which is
unique (the reference's own tests/test_elimination.c checks exactly that, on the shape list used here)."""
import numpy as np
import pytest

from tests import harness as H

# tests/test_elimination.c:96-125
SHAPES = [(4, 67), (17, 121), (65, 17), (128, 128), (1024, 1024), (2047, 2047), (65, 65), (100, 100), (21, 171),
          (31, 121), (193, 65), (1025, 1025), (2048, 2048), (64, 64), (1024, 1025), (1000, 1000), (1000, 10),
          (1710, 1290), (1290, 1710)]


def _low_rank(m, n, rank):
    """random m x n matrix of rank <= `rank`: product of random m x rank and rank x n factors"""
    X, Y = H.random_matrix(m, rank), H.random_matrix(rank, n)
    P = H.oracle().orc_mul(None, X, Y, 0)
    H.free(X, Y)
    return P


@pytest.mark.parametrize("m,n", SHAPES)
def test_reduced_form_matches_every_reference_variant(m, n):
    ref = H.ref(required=True)
    H.libc.srandom(17 + m + n)
    A = H.random_matrix(m, n)
    want = H.clone(A)
    r = H.oracle().orc_echelonize(want, 1)
    for fn, args in (("mzd_echelonize_m4ri", (1, 0)), ("mzd_echelonize_m4ri", (1, 8)), ("mzd_echelonize_naive", (1,)),
                     ("mzd_echelonize_pluq", (1,))):
        B = H.clone(A)
        assert getattr(ref, fn)(B, *args) == r, fn
        assert np.array_equal(H.storage(B), H.storage(want)), fn
        H.free(B)
    H.free(A, want)


@pytest.mark.parametrize("m,n,rank", [(300, 500, 40), (700, 200, 64), (257, 257, 129), (1000, 1000, 1)])
def test_rank_deficient_inputs(m, n, rank):
    ref = H.ref(required=True)
    H.libc.srandom(5 + rank)
    A = _low_rank(m, n, rank)
    want, B = H.clone(A), H.clone(A)
    r = H.oracle().orc_echelonize(want, 1)
    assert r <= rank and ref.mzd_echelonize_m4ri(B, 1, 0) == r
    assert np.array_equal(H.storage(B), H.storage(want))
    assert not H.storage(want)[r:].any()            # rows below the rank are zero
    H.free(A, want, B)


def test_upper_triangular_form_then_top_reduction_is_the_reduced_form():
    """full == 0 is not unique across algorithms (the reference compares it only after mzd_top_echelonize_m4ri);
    the oracle's naive full == 0 form must at least have the same rank and pivot columns."""
    H.libc.srandom(3)
    A = H.random_matrix(200, 300)
    U, Rr = H.clone(A), H.clone(A)
    r0, r1 = H.oracle().orc_echelonize(U, 0), H.oracle().orc_echelonize(Rr, 1)
    assert r0 == r1
    su, sr = H.storage(U), H.storage(Rr)
    for i in range(r0):      # same pivot column in every row
        first = lambda row: next(64 * w + int(np.log2(int(row[w]) & -int(row[w]))) for w in range(len(row)) if row[w])
        assert first(su[i]) == first(sr[i])
    H.free(A, U, Rr)
