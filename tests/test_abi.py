"""CPU tests (-m "not gpu") of the drop-in boundary: the C-ABI library loads without a GPU,
exports every symbol include/m4ri_b200.h declares, and its host-side helpers follow the
reference's container semantics.  No compute entry point is called here."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import m4ri_b200
from tests import harness as H

HEADER = os.path.join(H.ROOT, "include", "m4ri_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(_?mzd_\w+|m4ri_b200_\w+)\s*\(", src)
    return sorted(set(n for n in names if not n.endswith("_dmat") and n != "m4ri_b200_dmat"))


def test_header_declares_the_reference_entry_points():
    names = declared_functions()
    for required in ("mzd_mul", "mzd_addmul", "_mzd_addmul", "_mzd_mul_even", "_mzd_addmul_even", "mzd_mul_m4rm",
                     "mzd_addmul_m4rm", "_mzd_mul_m4rm", "mzd_mul_mp", "mzd_addmul_mp"):
        assert required in names


def test_library_loads_and_exports_every_declared_symbol():
    lib = m4ri_b200.load_library()
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.m4ri_b200_version() >= 100


def test_header_compiles_as_c_and_mzd_t_is_64_bytes(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "m4ri_b200.h"\n#include <stddef.h>\n'
                   '_Static_assert(sizeof(mzd_t) == 64, "size");\n'
                   '_Static_assert(offsetof(mzd_t, rowstride) == 16 && offsetof(mzd_t, flags) == 24 && '
                   'offsetof(mzd_t, high_bitmask) == 48 && offsetof(mzd_t, data) == 56, "layout");\n'
                   'int main(void) { return 0; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(H.ROOT, "include"),
                           "-c", str(src), "-o", str(tmp_path / "t.o")])
    assert ctypes.sizeof(m4ri_b200.MzdT) == 64


@pytest.mark.parametrize("r,c", [(1, 1), (3, 64), (5, 65), (7, 128), (9, 200), (0, 10), (10, 0)])
def test_standalone_mzd_init_matches_reference_container(r, c):
    """m4ri/mzd.c:142-157: width, even rowstride, high_bitmask, flags, zero fill."""
    M = m4ri_b200.mzd_init(r, c)
    m = M.contents
    assert (m.nrows, m.ncols) == (r, c)
    assert m.width == (c + 63) // 64
    assert m.rowstride == m.width + (m.width & 1)
    assert m.high_bitmask == (2**64 - 1 if c % 64 == 0 else (1 << (c % 64)) - 1)
    assert bool(m.flags & m4ri_b200.MZD_FLAG_NONZERO_EXCESS) == (c % 64 != 0)
    assert not (m.flags & m4ri_b200.MZD_FLAG_WINDOWED)
    O = H.new(r, c).contents
    assert (O.width, O.rowstride, O.high_bitmask, O.flags) == (m.width, m.rowstride, m.high_bitmask, m.flags)
    if r and c:
        assert not np.any(m4ri_b200.words(M))
    m4ri_b200.mzd_free(M)


def test_standalone_window_matches_reference_container():
    """m4ri/mzd.c:159-177"""
    P = m4ri_b200.mzd_init(10, 300)
    W = m4ri_b200.mzd_init_window(P, 2, 64, 7, 64 + 100)
    w = W.contents
    assert (w.nrows, w.ncols, w.width, w.rowstride) == (5, 100, 2, P.contents.rowstride)
    assert w.flags & m4ri_b200.MZD_FLAG_WINDOWED and w.flags & m4ri_b200.MZD_FLAG_NONZERO_EXCESS
    assert ctypes.addressof(w.data.contents) == ctypes.addressof(P.contents.data.contents) + 8 * (2 * w.rowstride + 1)
    m4ri_b200.mzd_free(W)
    m4ri_b200.mzd_free(P)


def test_product_package_does_not_reference_the_oracle():
    """the product must never route through oracle/ (or any CPU implementation)"""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(H.ROOT, "m4ri_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cpp", ".h", ".cuh", "Makefile")):
                text = open(os.path.join(dirpath, fn), errors="ignore").read()
                if re.search(r"oracle|libm4ri_ref|orc_", text):
                    bad.append(os.path.join(dirpath, fn))
    assert not bad, bad


def test_compute_call_without_gpu_dies_loudly():
    """No CPU fallback: on a box without a CUDA device a product call aborts with a message."""
    if m4ri_b200.load_library().m4ri_b200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    code = ("import sys; sys.path.insert(0, %r); import m4ri_b200\n"
            "A = m4ri_b200.mzd_init(8, 8); B = m4ri_b200.mzd_init(8, 8)\n"
            "m4ri_b200.mzd_mul(None, A, B, 0); print('survived')\n" % H.ROOT)
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert p.returncode == -6 and "no usable CUDA device" in p.stderr and "survived" not in p.stdout
