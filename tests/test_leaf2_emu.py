"""CPU check of the tall-tile M4RM leaf (m4ri_b200/csrc/m4rm_leaf2_body.h): the kernel body is compiled
with g++ into tests/c/emu_leaf2.cpp, where a CTA is 256 host threads and TMA / mbarriers / shared
memory are emulated, and its result is compared with a definition-level GF(2) product.  No GPU needed;
the device build of the same source is covered by tests/test_zz_leaf2_gpu.py."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", params=[(0, 0), (1, 0), (0, 1), (1, 1)], ids=["a-lds64", "a-lds128", "split-build", "a-lds128-split-build"])
def emu(request, tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("emu") / "emu_leaf2")
    subprocess.check_call(["g++", "-O2", "-std=c++20", "-pthread", "-Wno-unknown-pragmas", f"-DEMU_AWIDE={request.param[0]}", f"-DEMU_SPLIT={request.param[1]}",
                           "-I", os.path.join(ROOT, "m4ri_b200", "csrc"),
                           os.path.join(ROOT, "tests", "c", "emu_leaf2.cpp"), "-o", exe])
    return exe


def test_builtin_cases_bit_exact(emu):
    out = subprocess.run([emu], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "FAIL" not in out.stdout and out.stdout.count("ok  ") >= 10, out.stdout


@pytest.mark.parametrize("hybrid", ["0", "1"])
@pytest.mark.parametrize("count,m,l,n,blocks", [
    (3, 4096, 384, 1024, 5),         # 12 tiles on 5 CTAs: two whole-tile rounds + a stream-K tail of 2 tiles
    (7, 4096, 256, 512, 4),          # 14 tiles on 4 CTAs: three rounds + tail
    (1, 8192, 256, 768, 6),          # 6 tiles on 6 CTAs: one round, empty tail
    (2, 5000, 300, 300, 3),          # ragged rows/columns, 8 tiles on 3 CTAs
])
def test_hybrid_partition(emu, hybrid, count, m, l, n, blocks):
    """whole tiles round-robin first, stream-K over the rest (the launcher's default) and pure stream-K give the same bits"""
    out = subprocess.run([emu, str(count), str(m), str(l), str(n), str(blocks)], capture_output=True, text=True,
                         timeout=600, env=dict(os.environ, EMU_HYBRID=hybrid))
    assert out.returncode == 0 and "FAIL" not in out.stdout, out.stdout + out.stderr


@pytest.mark.parametrize("count,m,l,n,blocks", [
    (3, 4096, 384, 1024, 5), (7, 4096, 256, 512, 4), (1, 8192, 256, 768, 6), (2, 5000, 300, 384, 3), (49, 4096, 128, 256, 9),
])
def test_store_mode_overwrites_uninitialised_tiles(emu, count, m, l, n, blocks):
    """C = A*B: the whole-tile rounds store their tile over garbage, only the products with stream-K tail tiles are zeroed"""
    out = subprocess.run([emu, str(count), str(m), str(l), str(n), str(blocks)], capture_output=True, text=True,
                         timeout=900, env=dict(os.environ, EMU_STORE="1"))
    assert out.returncode == 0 and "FAIL" not in out.stdout, out.stdout + out.stderr


@pytest.mark.parametrize("count,m,l,n,blocks", [
    (1, 1, 1, 1, 1),                 # the smallest product
    (1, 4097, 129, 257, 2),          # one past every tile / slab / word edge
    (2, 300, 2000, 100, 9),          # long K, short and narrow C
    (1, 12288, 256, 256, 5),         # three row tiles
])
def test_extra_shapes(emu, count, m, l, n, blocks):
    out = subprocess.run([emu, str(count), str(m), str(l), str(n), str(blocks)], capture_output=True, text=True,
                         timeout=600)
    assert out.returncode == 0 and "FAIL" not in out.stdout, out.stdout + out.stderr
