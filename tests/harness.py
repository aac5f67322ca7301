"""Test-only helpers: the oracle (oracle/liboracle_m4rm.so), the compiled reference
(oracle/_ref/libm4ri_ref.so, when present) and matrix plumbing shared by the test files.

Nothing here is imported by the product package."""
from __future__ import annotations

import ctypes
import hashlib
import os
import subprocess
from ctypes import POINTER, c_int, c_uint64

import numpy as np

import m4ri_b200
from m4ri_b200 import MzdP, MzdT

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle_m4rm.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libm4ri_ref.so")
REF_OMP_SO = os.path.join(ORACLE_DIR, "_ref", "libm4ri_ref_omp.so")
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

libc = ctypes.CDLL(None)
libc.srandom.argtypes = [ctypes.c_uint]

_oracle = None
_ref = None


def oracle():
    """The plain-C restatement, (re)built on demand with gcc."""
    global _oracle
    if _oracle is None:
        src = os.path.join(ORACLE_DIR, "m4rm_oracle.c")
        if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
            subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "oracle"])
        lib = ctypes.CDLL(ORACLE_SO, mode=ctypes.RTLD_LOCAL)
        lib.orc_init.argtypes, lib.orc_init.restype = [c_int, c_int], MzdP
        lib.orc_init_window.argtypes, lib.orc_init_window.restype = [MzdP, c_int, c_int, c_int, c_int], MzdP
        lib.orc_free.argtypes = [MzdP]
        lib.orc_random_word.restype = c_uint64
        lib.orc_randomize.argtypes = [MzdP]
        lib.orc_equal.argtypes, lib.orc_equal.restype = [MzdP, MzdP], c_int
        lib.orc_copy.argtypes = [MzdP, MzdP]
        lib.orc_add.argtypes = [MzdP, MzdP, MzdP]
        lib.orc_gray_code.argtypes, lib.orc_gray_code.restype = [c_int, c_int], c_int
        lib.orc_build_code.argtypes = [POINTER(c_int), POINTER(c_int), c_int]
        lib.orc_make_table.argtypes = [MzdP, c_int, c_int, MzdP, POINTER(c_int)]
        lib.orc_mul_naive.argtypes, lib.orc_mul_naive.restype = [MzdP, MzdP, MzdP, c_int], MzdP
        lib.orc_mul_m4rm.argtypes, lib.orc_mul_m4rm.restype = [MzdP, MzdP, MzdP, c_int, c_int], MzdP
        lib.orc_mul.argtypes, lib.orc_mul.restype = [MzdP, MzdP, MzdP, c_int], MzdP
        lib.orc_addmul.argtypes, lib.orc_addmul.restype = [MzdP, MzdP, MzdP, c_int], MzdP
        lib.orc_trsm_lower_left.argtypes = [MzdP, MzdP]
        lib.orc_trsm_upper_left.argtypes = [MzdP, MzdP]
        lib.orc_trsm_lower_right.argtypes = [MzdP, MzdP]
        lib.orc_trsm_upper_right.argtypes = [MzdP, MzdP]
        lib.orc_transpose.argtypes, lib.orc_transpose.restype = [MzdP, MzdP], MzdP
        lib.orc_echelonize.argtypes, lib.orc_echelonize.restype = [MzdP, c_int], c_int
        lib.orc_ple.argtypes, lib.orc_ple.restype = [MzdP, POINTER(c_int), POINTER(c_int)], c_int
        _oracle = lib
    return _oracle


def _declare_ref(lib):
    three = [MzdP, MzdP, MzdP, c_int]
    for name in ("mzd_mul", "mzd_addmul", "mzd_mul_m4rm", "mzd_addmul_m4rm"):
        fn = getattr(lib, name)
        fn.argtypes, fn.restype = three, MzdP
    for name in ("mzd_mul_naive", "mzd_addmul_naive"):
        fn = getattr(lib, name)
        fn.argtypes, fn.restype = [MzdP, MzdP, MzdP], MzdP
    lib._mzd_mul_m4rm.argtypes, lib._mzd_mul_m4rm.restype = [MzdP, MzdP, MzdP, c_int, c_int], MzdP
    lib.mzd_init.argtypes, lib.mzd_init.restype = [c_int, c_int], MzdP
    lib.mzd_init_window.argtypes, lib.mzd_init_window.restype = [MzdP, c_int, c_int, c_int, c_int], MzdP
    lib.mzd_free.argtypes = [MzdP]
    lib.mzd_randomize.argtypes = [MzdP]
    lib.mzd_equal.argtypes, lib.mzd_equal.restype = [MzdP, MzdP], c_int
    lib.mzd_copy.argtypes, lib.mzd_copy.restype = [MzdP, MzdP], MzdP
    lib._mzd_add.argtypes, lib._mzd_add.restype = [MzdP, MzdP, MzdP], MzdP
    lib.m4ri_gray_code.argtypes, lib.m4ri_gray_code.restype = [c_int, c_int], c_int
    lib.m4ri_build_code.argtypes = [POINTER(c_int), POINTER(c_int), c_int]
    lib.mzd_make_table.argtypes = [MzdP, c_int, c_int, c_int, MzdP, POINTER(c_int)]
    lib.m4ri_random_word.restype = c_uint64
    lib.mzd_transpose.argtypes, lib.mzd_transpose.restype = [MzdP, MzdP], MzdP
    lib.mzd_echelonize_m4ri.argtypes, lib.mzd_echelonize_m4ri.restype = [MzdP, c_int, c_int], c_int
    lib.mzd_echelonize_naive.argtypes, lib.mzd_echelonize_naive.restype = [MzdP, c_int], c_int
    lib.mzd_echelonize_pluq.argtypes, lib.mzd_echelonize_pluq.restype = [MzdP, c_int], c_int
    lib.mzd_inv_m4ri.argtypes, lib.mzd_inv_m4ri.restype = [MzdP, MzdP, c_int], MzdP
    for name in ("mzd_trsm_lower_left", "mzd_trsm_upper_left", "mzd_trsm_lower_right", "mzd_trsm_upper_right"):
        getattr(lib, name).argtypes = [MzdP, MzdP, c_int]
        getattr(lib, name).restype = None
    if hasattr(lib, "mzd_mul_mp"):
        lib.mzd_mul_mp.argtypes, lib.mzd_mul_mp.restype = three, MzdP
        lib.mzd_addmul_mp.argtypes, lib.mzd_addmul_mp.restype = three, MzdP
    return lib


def ref(required: bool = False):
    """The unmodified reference compiled into oracle/_ref (None when absent)."""
    global _ref
    if _ref is None and os.path.exists(REF_SO):
        _ref = _declare_ref(ctypes.CDLL(REF_SO, mode=ctypes.RTLD_LOCAL))
    if _ref is None and required:
        raise RuntimeError("oracle/_ref/libm4ri_ref.so missing (run `make -C oracle ref` in the build container)")
    return _ref


def ref_omp():
    if os.path.exists(REF_OMP_SO):
        return _declare_ref(ctypes.CDLL(REF_OMP_SO, mode=ctypes.RTLD_LOCAL))
    return None


# ---- matrices ------------------------------------------------------------------------------

def new(r, c):
    return oracle().orc_init(r, c)


def free(*ms):
    for m in ms:
        if m:
            oracle().orc_free(m)


def window(M, r0, c0, r1, c1):
    return oracle().orc_init_window(M, r0, c0, r1, c1)


def randomize(M):
    """reference fill: three random() draws per word, row-major (m4ri/misc.c:58-71)."""
    oracle().orc_randomize(M)


def random_matrix(r, c):
    M = new(r, c)
    randomize(M)
    return M


def clone(M):
    N = new(M.contents.nrows, M.contents.ncols)
    oracle().orc_copy(N, M)
    return N


def storage(M) -> np.ndarray:
    """all words of a NON-window matrix incl. padding, [nrows, rowstride]"""
    m = M.contents
    if m.nrows == 0 or m.ncols == 0:
        return np.zeros((m.nrows, 0), dtype=np.uint64)
    return np.ctypeslib.as_array(m.data, shape=(m.nrows, m.rowstride))


def digest(M) -> str:
    """sha256 over dims + valid words (excess bits cleared)"""
    m = M.contents
    h = hashlib.sha256()
    h.update(np.array([m.nrows, m.ncols], dtype=np.int64).tobytes())
    h.update(np.ascontiguousarray(m4ri_b200.valid_words(M)).tobytes())
    return h.hexdigest()


# ---- large seeded inputs (BASELINE configs 2-5): the same words here, in bench.py and in
#      tests/golden/make_golden_large.py ------------------------------------------------------------

SEED_A, SEED_B, SEED_C = 101, 102, 103


def seeded_words(seed: int, nrows: int, width: int, row0: int = 0, nrows_total: int | None = None) -> np.ndarray:
    """Rows row0 .. row0+nrows of the [nrows_total, width] uint64 matrix whose words are the raw 64-bit
    outputs of numpy's PCG64(seed) in row-major order (one draw per word, so any row range can be
    produced on its own by advancing the generator: ranks of a sharded run fill only their rows)."""
    bg = np.random.PCG64(seed)
    if row0:
        bg.advance(row0 * width)
    rng = np.random.Generator(bg)
    out = np.empty((nrows, width), dtype=np.uint64)
    step = max(1, (1 << 24) // max(1, width))
    for i in range(0, nrows, step):
        j = min(nrows, i + step)
        out[i:j] = rng.integers(0, 2**64, size=(j - i, width), dtype=np.uint64)
    return out


def fill_seeded(M, seed: int) -> None:
    """Fill a NON-window matrix with seeded_words(seed); excess bits cleared (mzd.h:117-122)."""
    m = M.contents
    st = storage(M)
    st[:, :m.width] = seeded_words(seed, m.nrows, m.width)
    st[:, m.width - 1] &= np.uint64(m.high_bitmask)
    st[:, m.width:] = 0


def block_digest(words2d: np.ndarray) -> str:
    """sha256 of a contiguous copy of a [rows, words] block (shape first)"""
    h = hashlib.sha256()
    h.update(np.array(words2d.shape, dtype=np.int64).tobytes())
    h.update(np.ascontiguousarray(words2d).tobytes())
    return h.hexdigest()


# C of every large case is recorded as 8 row-blocks x 8 column-blocks, so that every rank of a
# 1/2/4/8-GPU partition (row-blocks, or any pr x pc grid with pr, pc dividing 8) can check its own block
# against the reference.
LARGE_BLOCK_ROWS, LARGE_BLOCK_COLS = 8, 8


def large_block_digests(words2d: np.ndarray) -> list:
    r, w = words2d.shape
    br, bw = r // LARGE_BLOCK_ROWS, w // LARGE_BLOCK_COLS
    return [[block_digest(words2d[i * br:(i + 1) * br, j * bw:(j + 1) * bw]) for j in range(LARGE_BLOCK_COLS)]
            for i in range(LARGE_BLOCK_ROWS)]


def equal(A, B) -> bool:
    return bool(oracle().orc_equal(A, B))


# the reference's own shape lists: tests/test_multiplication.c:251-322
MUL_SHAPES = [
    (1, 1, 1, 0, 1024), (1, 128, 128, 0, 0), (3, 131, 257, 0, 0), (64, 64, 64, 0, 64),
    (128, 128, 128, 0, 64), (21, 171, 31, 0, 63), (21, 171, 31, 0, 131), (193, 65, 65, 8, 64),
    (1025, 1025, 1025, 3, 256), (2048, 2048, 4096, 0, 1024), (4096, 3528, 4096, 0, 1024),
    (1024, 1025, 1, 0, 1024), (1000, 1000, 1000, 0, 256), (1000, 10, 20, 0, 64),
    (1710, 1290, 1000, 0, 256), (1290, 1710, 200, 0, 64), (1290, 1710, 2000, 0, 256),
    (1290, 1290, 2000, 0, 64), (1000, 210, 200, 0, 64),
]
ADDMUL_SHAPES = [
    (1, 128, 128, 0, 0), (3, 131, 257, 0, 0), (64, 64, 64, 0, 64), (128, 128, 128, 0, 64),
    (21, 171, 31, 0, 63), (21, 171, 31, 0, 131), (193, 65, 65, 8, 64), (1025, 1025, 1025, 3, 256),
    (4096, 4096, 4096, 0, 2048), (1000, 1000, 1000, 0, 256), (1000, 10, 20, 0, 64),
    (1710, 1290, 1000, 0, 256), (1290, 1710, 200, 0, 64), (1290, 1710, 2000, 0, 256),
    (1290, 1290, 2000, 0, 64), (1000, 210, 200, 0, 64),
]
SQR_SHAPES = [
    (1, 0, 1024), (128, 0, 0), (131, 0, 0), (64, 0, 64), (128, 0, 64), (171, 0, 63), (171, 0, 131),
    (193, 8, 64), (1025, 3, 256), (2048, 0, 1024), (3528, 0, 1024), (1000, 0, 256), (1000, 0, 64),
    (1710, 0, 256), (1290, 0, 64), (2000, 0, 256), (2000, 0, 64), (210, 0, 64),
]
ADDSQR_SHAPES = [
    (1, 0, 0), (131, 0, 0), (64, 0, 64), (128, 0, 64), (171, 0, 63), (171, 0, 131), (193, 8, 64),
    (1025, 3, 256), (4096, 0, 2048), (1000, 0, 256), (1000, 0, 64), (1710, 0, 256), (1290, 0, 64),
    (2000, 0, 256), (2000, 0, 64), (210, 0, 64),
]
