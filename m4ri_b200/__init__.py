"""m4ri_b200 — host-side mirror of M4RI's multiplication interface over libm4ri_b200.so.

The product is the C-ABI shared library built from ``m4ri_b200/csrc`` (declared in
``include/m4ri_b200.h``).  This module is the thin ctypes binding used by the tests and the
benchmark: same function names, argument meaning and error behaviour (stderr + abort) as the
reference's ``mzd_mul`` / ``mzd_addmul`` / ``mzd_mul_m4rm`` family (m4ri/strassen.h:52-126,
m4ri/brilliantrussian.h:274-317).  It contains no compute and no CPU fallback: importing is
cheap, but the first call raises if the CUDA library has not been built.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_int, c_int64, c_uint8, c_uint64, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libm4ri_b200.so")

MZD_FLAG_NONZERO_EXCESS = 0x2  # m4ri/mzd.h:144
MZD_FLAG_WINDOWED = 0x4        # m4ri/mzd.h:150


class MzdT(ctypes.Structure):
    """64-byte ``mzd_t`` header (m4ri/mzd.h:68-99); the one layout every libm4ri-compatible library uses."""

    _fields_ = [
        ("nrows", c_int),
        ("ncols", c_int),
        ("width", c_int64),
        ("rowstride", c_int64),
        ("flags", c_uint8),
        ("padding", c_uint8 * 23),
        ("high_bitmask", c_uint64),
        ("data", POINTER(c_uint64)),
    ]


assert ctypes.sizeof(MzdT) == 64
MzdP = POINTER(MzdT)


class DMat(ctypes.Structure):
    """``m4ri_b200_dmat``: a device-resident bit-packed matrix."""

    _fields_ = [
        ("data", c_void_p),
        ("pitch", c_int64),
        ("nrows", c_int),
        ("ncols", c_int),
        ("owner", c_int),
    ]


DMatP = POINTER(DMat)


class MzpT(ctypes.Structure):
    """``mzp_t`` (m4ri/mzp.h:37-49): a permutation as LAPACK-style transpositions."""

    _fields_ = [("values", POINTER(c_int)), ("length", c_int)]


MzpP = POINTER(MzpT)

HookFn = ctypes.CFUNCTYPE(None, c_void_p, c_int)


class Hooks(ctypes.Structure):
    """``m4ri_b200_hooks``: transfer callbacks of ``m4ri_b200_dmul_quads``."""

    _fields_ = [("need_a", HookFn), ("need_b", HookFn), ("need_c", HookFn), ("done_c", HookFn), ("user", c_void_p)]

_lib = None


def _declare(lib):
    three = [MzdP, MzdP, MzdP, c_int]
    for name in ("mzd_mul", "mzd_addmul", "_mzd_addmul", "_mzd_mul_even", "_mzd_addmul_even",
                 "mzd_mul_m4rm", "mzd_addmul_m4rm", "mzd_mul_mp", "mzd_addmul_mp"):
        fn = getattr(lib, name)
        fn.argtypes, fn.restype = three, MzdP
    lib._mzd_mul_m4rm.argtypes, lib._mzd_mul_m4rm.restype = [MzdP, MzdP, MzdP, c_int, c_int], MzdP
    for side in ("lower_left", "upper_left", "lower_right", "upper_right"):
        for prefix in ("mzd_trsm_", "_mzd_trsm_"):
            fn = getattr(lib, prefix + side)
            fn.argtypes, fn.restype = [MzdP, MzdP, c_int], None
    lib.m4ri_b200_dtrsm.argtypes = [DMatP, DMatP, c_int, c_int, c_int, c_void_p]
    lib.m4ri_b200_version.restype = c_int
    lib.m4ri_b200_device_count.restype = c_int
    lib.m4ri_b200_set_device.argtypes = [c_int]
    lib.m4ri_b200_set_num_devices.argtypes = [c_int]
    lib.m4ri_b200_set_default_cutoff.argtypes = [c_int]
    lib.m4ri_b200_get_default_cutoff.restype = c_int
    lib.m4ri_b200_last_path.restype = c_char_p
    lib.m4ri_b200_set_leaf_variant.argtypes, lib.m4ri_b200_set_leaf_variant.restype = [c_int], c_int
    lib.m4ri_b200_last_leaf_variant.restype = c_int
    lib.m4ri_b200_kernel_launches.restype = c_uint64
    lib.m4ri_b200_profile_end.argtypes = [POINTER(ctypes.c_double), POINTER(ctypes.c_double)]
    lib.m4ri_b200_profile_end.restype = c_uint64
    lib.m4ri_b200_mzd_init.argtypes, lib.m4ri_b200_mzd_init.restype = [c_int, c_int], MzdP
    lib.m4ri_b200_mzd_init_window.argtypes = [MzdP, c_int, c_int, c_int, c_int]
    lib.m4ri_b200_mzd_init_window.restype = MzdP
    lib.m4ri_b200_mzd_free.argtypes = [MzdP]
    lib.m4ri_b200_dmat_alloc.argtypes, lib.m4ri_b200_dmat_alloc.restype = [c_int, c_int], DMatP
    lib.m4ri_b200_dmat_wrap.argtypes = [c_void_p, c_int64, c_int, c_int]
    lib.m4ri_b200_dmat_wrap.restype = DMatP
    lib.m4ri_b200_dmat_free.argtypes = [DMatP]
    lib.m4ri_b200_upload.argtypes = [DMatP, MzdP, c_void_p]
    lib.m4ri_b200_download.argtypes = [MzdP, DMatP, c_void_p]
    lib.m4ri_b200_sync.argtypes = [c_void_p]
    lib.m4ri_b200_dmul_m4rm.argtypes = [DMatP, DMatP, DMatP, c_int, c_void_p]
    lib.m4ri_b200_dmul.argtypes = [DMatP, DMatP, DMatP, c_int, c_int, c_void_p]
    lib.m4ri_b200_dmul_levels.argtypes = [DMatP, DMatP, DMatP, c_int, c_int, c_void_p]
    lib.m4ri_b200_dmul_quads.argtypes = [DMatP * 4, DMatP * 4, DMatP * 4, c_int, c_int, c_void_p, POINTER(Hooks)]
    lib.m4ri_b200_result_free.argtypes = [MzdP]
    lib.m4ri_b200_dmul_tc.argtypes = [DMatP, DMatP, DMatP, c_int, c_void_p]
    lib.m4ri_b200_dmul_tc2.argtypes = [DMatP, DMatP, DMatP, c_void_p]
    lib.m4ri_b200_from_str.argtypes, lib.m4ri_b200_from_str.restype = [c_int, c_int, c_char_p], MzdP
    lib.m4ri_b200_from_jcf.argtypes, lib.m4ri_b200_from_jcf.restype = [c_char_p, c_int], MzdP
    lib.m4ri_b200_to_jcf.argtypes, lib.m4ri_b200_to_jcf.restype = [MzdP, c_char_p], c_int
    lib.m4ri_b200_to_pbm.argtypes, lib.m4ri_b200_to_pbm.restype = [MzdP, c_char_p], c_int
    lib.m4ri_b200_from_pbm.argtypes, lib.m4ri_b200_from_pbm.restype = [c_char_p], MzdP
    lib.m4ri_b200_fprint_row.argtypes = [c_void_p, MzdP, c_int]
    lib.m4ri_b200_dmat_from_jcf.argtypes, lib.m4ri_b200_dmat_from_jcf.restype = [c_char_p, c_int], DMatP
    for name in ("mzd_ple", "_mzd_ple"):
        getattr(lib, name).argtypes, getattr(lib, name).restype = [MzdP, MzpP, MzpP, c_int], c_int
    lib.m4ri_b200_dple.argtypes, lib.m4ri_b200_dple.restype = [DMatP, POINTER(c_int), POINTER(c_int), c_void_p], c_int
    for name in ("_mzd_mul_mp4", "_mzd_addmul_mp4"):
        getattr(lib, name).argtypes, getattr(lib, name).restype = three, MzdP
    lib.m4ri_b200_dtranspose.argtypes = [DMatP, DMatP, c_void_p]
    lib.m4ri_b200_transpose.argtypes, lib.m4ri_b200_transpose.restype = [MzdP, MzdP], MzdP
    lib.m4ri_b200_dadd.argtypes = [DMatP, DMatP, DMatP, c_void_p]
    lib.m4ri_b200_dechelonize.argtypes, lib.m4ri_b200_dechelonize.restype = [DMatP, c_int, c_void_p], c_int
    lib.m4ri_b200_echelonize.argtypes, lib.m4ri_b200_echelonize.restype = [MzdP, c_int], c_int
    lib.m4ri_b200_inv_m4ri.argtypes, lib.m4ri_b200_inv_m4ri.restype = [MzdP, MzdP], MzdP
    lib.mzd_echelonize_m4ri.argtypes, lib.mzd_echelonize_m4ri.restype = [MzdP, c_int, c_int], c_int
    lib.mzd_echelonize.argtypes, lib.mzd_echelonize.restype = [MzdP, c_int], c_int
    lib.mzd_inv_m4ri.argtypes, lib.mzd_inv_m4ri.restype = [MzdP, MzdP, c_int], MzdP
    return lib


def load_library():
    """Load libm4ri_b200.so (built in-tree by ``__graft_entry__.build()``).  Fails loudly."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(m4ri_b200 has no CPU fallback)")
        _lib = _declare(ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_LOCAL))
    return _lib


# ---- host matrices ---------------------------------------------------------------------------

def mzd_init(r: int, c: int):
    """Stand-alone ``mzd_init`` (m4ri/mzd.c:142-157)."""
    return load_library().m4ri_b200_mzd_init(r, c)


def mzd_init_window(M, lowr: int, lowc: int, highr: int, highc: int):
    """``mzd_init_window`` (m4ri/mzd.c:159-177)."""
    return load_library().m4ri_b200_mzd_init_window(M, lowr, lowc, highr, highc)


def mzd_free(M) -> None:
    load_library().m4ri_b200_mzd_free(M)


def words(M) -> np.ndarray:
    """numpy uint64 view [nrows, rowstride] of a matrix' storage (no copy)."""
    m = M.contents
    if m.nrows == 0 or m.ncols == 0:
        return np.zeros((m.nrows, 0), dtype=np.uint64)
    flat = np.ctypeslib.as_array(m.data, shape=((m.nrows - 1) * m.rowstride + m.width,))
    return np.lib.stride_tricks.as_strided(flat, shape=(m.nrows, m.width), strides=(m.rowstride * 8, 8))


def valid_words(M) -> np.ndarray:
    """Copy of the valid bits: [nrows, width] with the excess bits of the last word cleared."""
    w = words(M).copy()
    if w.shape[1]:
        w[:, -1] &= np.uint64(M.contents.high_bitmask)
    return w


# ---- the reference-named entry points ------------------------------------------------------------

def mzd_mul(C, A, B, cutoff: int = 0):
    return load_library().mzd_mul(C, A, B, cutoff)


def mzd_addmul(C, A, B, cutoff: int = 0):
    return load_library().mzd_addmul(C, A, B, cutoff)


def mzd_mul_m4rm(C, A, B, k: int = 0):
    return load_library().mzd_mul_m4rm(C, A, B, k)


def mzd_addmul_m4rm(C, A, B, k: int = 0):
    return load_library().mzd_addmul_m4rm(C, A, B, k)


def _mzd_mul_m4rm(C, A, B, k: int = 0, clear: int = 1):
    return load_library()._mzd_mul_m4rm(C, A, B, k, clear)


def mzd_mul_mp(C, A, B, cutoff: int = 0):
    return load_library().mzd_mul_mp(C, A, B, cutoff)


def mzd_addmul_mp(C, A, B, cutoff: int = 0):
    return load_library().mzd_addmul_mp(C, A, B, cutoff)


def mzd_trsm_lower_left(L, B, cutoff: int = 0) -> None:
    """L X = B, X overwrites B (m4ri/triangular.c:394-455)."""
    load_library().mzd_trsm_lower_left(L, B, cutoff)


def mzd_trsm_upper_left(U, B, cutoff: int = 0) -> None:
    """U X = B, X overwrites B (m4ri/triangular.c:457-516)."""
    load_library().mzd_trsm_upper_left(U, B, cutoff)


def mzd_trsm_lower_right(L, B, cutoff: int = 0) -> None:
    """X L = B, X overwrites B (m4ri/triangular.c:300-392)."""
    load_library().mzd_trsm_lower_right(L, B, cutoff)


def mzd_trsm_upper_right(U, B, cutoff: int = 0) -> None:
    """X U = B, X overwrites B (m4ri/triangular.c:29-148)."""
    load_library().mzd_trsm_upper_right(U, B, cutoff)
