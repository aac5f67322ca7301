"""CPU tests of the interchange functions (csrc/io.cu; no GPU needed: they are host code) against the unmodified
reference's mzd_from_str / mzd_from_jcf / mzd_fprint_row (m4ri/io.c:49-68, 297-357) where it is present."""
import ctypes
import os

import numpy as np
import pytest

import m4ri_b200
from tests import harness as H

libc = ctypes.CDLL(None)
libc.fopen.restype = ctypes.c_void_p
libc.fopen.argtypes = [ctypes.c_char_p, ctypes.c_char_p]
libc.fclose.argtypes = [ctypes.c_void_p]


@pytest.fixture(scope="module")
def lib():
    return m4ri_b200.load_library()


def _ref_io():
    ref = H.ref()
    if ref is None:
        pytest.skip("oracle/_ref/libm4ri_ref.so not present")
    ref.mzd_from_str.argtypes, ref.mzd_from_str.restype = [ctypes.c_int, ctypes.c_int, ctypes.c_char_p], m4ri_b200.MzdP
    ref.mzd_from_jcf.argtypes, ref.mzd_from_jcf.restype = [ctypes.c_char_p, ctypes.c_int], m4ri_b200.MzdP
    ref.mzd_fprint_row.argtypes = [ctypes.c_void_p, m4ri_b200.MzdP, ctypes.c_int]
    return ref


@pytest.mark.parametrize("m,n", [(1, 1), (3, 70), (65, 64), (20, 129)])
def test_from_str_matches_reference(lib, m, n):
    rng = np.random.default_rng(m * 131 + n)
    text = "".join("1" if b else "0" for b in rng.integers(0, 2, size=m * n)).encode()
    A = lib.m4ri_b200_from_str(m, n, text)
    ref = _ref_io()
    B = ref.mzd_from_str(m, n, text)
    assert np.array_equal(m4ri_b200.valid_words(A), m4ri_b200.valid_words(B))
    lib.m4ri_b200_mzd_free(A)
    ref.mzd_free(B)


def test_jcf_round_trip_and_reference_reader(lib, tmp_path):
    rng = np.random.default_rng(5)
    m, n = 37, 150
    A = m4ri_b200.mzd_init(m, n)
    w = m4ri_b200.words(A)
    w[:, :] = rng.integers(0, 2**64, size=w.shape, dtype=np.uint64) & rng.integers(0, 2**64, size=w.shape, dtype=np.uint64)
    w[:, -1] &= np.uint64(A.contents.high_bitmask)
    w[:, 0] |= np.uint64(1)                        # JCF cannot express an empty row
    fn = str(tmp_path / "a.jcf").encode()
    assert lib.m4ri_b200_to_jcf(A, fn) == 0
    B = lib.m4ri_b200_from_jcf(fn, 0)
    assert np.array_equal(m4ri_b200.valid_words(A), m4ri_b200.valid_words(B))
    ref = _ref_io()
    C = ref.mzd_from_jcf(fn, 0)
    assert np.array_equal(m4ri_b200.valid_words(A), m4ri_b200.valid_words(C))
    ref.mzd_free(C)
    Z = m4ri_b200.mzd_init(2, 5)                   # an empty row: the writer refuses
    assert lib.m4ri_b200_to_jcf(Z, str(tmp_path / "z.jcf").encode()) == 2
    assert not lib.m4ri_b200_from_jcf(str(tmp_path / "missing.jcf").encode(), 0)
    for M in (A, B, Z):
        lib.m4ri_b200_mzd_free(M)


def test_fprint_row_matches_reference(lib, tmp_path):
    ref = _ref_io()
    H.libc.srandom(9)
    A = H.random_matrix(4, 200)
    outs = []
    for name, fn in (("ours", lib.m4ri_b200_fprint_row), ("ref", ref.mzd_fprint_row)):
        path = str(tmp_path / name).encode()
        fh = libc.fopen(path, b"w")
        for i in range(4):
            fn(fh, A, i)
        libc.fclose(fh)
        outs.append(open(path, "rb").read())
    assert outs[0] == outs[1] and outs[0].count(b"\n") == 4
    H.free(A)


def test_pbm_round_trip(lib, tmp_path):
    H.libc.srandom(10)
    A = H.random_matrix(33, 77)
    fn = str(tmp_path / "a.pbm").encode()
    assert lib.m4ri_b200_to_pbm(A, fn) == 0
    raw = open(fn, "rb").read()
    assert raw.startswith(b"P4\n77 33\n") and len(raw) == len(b"P4\n77 33\n") + 33 * 10
    B = lib.m4ri_b200_from_pbm(fn)
    assert np.array_equal(m4ri_b200.valid_words(A), m4ri_b200.valid_words(B))
    lib.m4ri_b200_mzd_free(B)
    H.free(A)
