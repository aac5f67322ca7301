"""CPU check of the device RREF (m4ri_b200/csrc/echelon_body.h): the kernel bodies are compiled with g++ into
tests/c/emu_echelon.cpp, which runs the strip loop of echelon.cu with 512 host threads per selection CTA and
compares with a plain Gauss-Jordan elimination.  The device build is covered by tests/test_zz2_echelon_gpu.py."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("emu") / "emu_echelon")
    subprocess.check_call(["g++", "-O2", "-std=c++20", "-pthread", "-Wno-unknown-pragmas",
                           "-I", os.path.join(ROOT, "m4ri_b200", "csrc"),
                           os.path.join(ROOT, "tests", "c", "emu_echelon.cpp"), "-o", exe])
    return exe


def test_builtin_cases(emu):
    out = subprocess.run([emu], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "FAIL" not in out.stdout and out.stdout.count("ok  ") >= 10, out.stdout + out.stderr


@pytest.mark.parametrize("m,n,rank_bound,chunk", [(128, 128, 0, 8192), (31, 121, 0, 8), (320, 192, 70, 40)])
def test_extra_shapes(emu, m, n, rank_bound, chunk):
    out = subprocess.run([emu, str(m), str(n), str(rank_bound), str(chunk)], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "FAIL" not in out.stdout, out.stdout + out.stderr
