"""GPU test: the reference's OWN test programs (tests/test_*.c, compiled unmodified into
oracle/_ref/tests/ against a default, interposable build of libm4ri) run with
LD_PRELOAD=libm4ri_b200.so.  Every mzd_mul / mzd_addmul / _mzd_mul_m4rm the reference makes — from
its tests directly and from inside PLE, PLUQ, TRSM, solve, kernel, inversion and elimination
(m4ri/ple.c:126, triangular.c:55,100,348,439,503, solve.c:89) — is then served by the CUDA library,
and the programs' own assertions are the parity check."""
import os
import re
import subprocess

import pytest

import m4ri_b200
from tests import harness as H

pytestmark = pytest.mark.gpu

BIN_DIR = os.path.join(H.ORACLE_DIR, "_ref", "tests")
# programs whose code path reaches the multiplication symbols (the others still must pass untouched)
USES_MUL = {"test_multiplication", "test_smallops", "test_trsm", "test_ple", "test_pluq", "test_solve",
            "test_kernel", "test_invert"}
ALL = ["test_multiplication", "test_smallops", "test_elimination", "test_trsm", "test_ple", "test_pluq", "test_solve",
       "test_kernel", "test_invert", "test_random", "test_transpose", "test_colswap", "test_misc", "test_alignment",
       "test_djb"]


@pytest.mark.parametrize("prog", ALL)
def test_reference_program_passes_with_gpu_library_preloaded(prog):
    exe = os.path.join(BIN_DIR, prog)
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/tests not built (run `make -C oracle ref` in the build container)")
    env = dict(os.environ, LD_PRELOAD=m4ri_b200.LIB_PATH, M4RI_B200_REPORT="1")
    p = subprocess.run([exe], capture_output=True, text=True, timeout=1800, env=env, cwd=BIN_DIR)
    assert p.returncode == 0, (p.stdout[-2000:], p.stderr[-2000:])
    m = re.search(r"m4ri_b200: served (\d+) products with (\d+) CUDA kernel launches", p.stderr)
    if prog in USES_MUL:
        assert m and int(m.group(1)) > 0 and int(m.group(2)) > 0, "the preloaded library was never called: " + p.stderr[-500:]
