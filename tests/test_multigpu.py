"""GPU tests of mzd_mul_mp / mzd_addmul_mp over several GPUs in one process (csrc/multi.cu).
Skipped on a 1-GPU box; run with `gpurun --gpus 2|4|8 -- python -m pytest tests/test_multigpu.py`.
T5 of SURVEY.md §7: same bits for every G, including m not divisible by 64*G."""
import numpy as np
import pytest

import m4ri_b200
from tests import harness as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    L = m4ri_b200.load_library()
    if L.m4ri_b200_device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    yield L
    L.m4ri_b200_set_num_devices(1)


@pytest.mark.parametrize("shape,cutoff", [((300, 200, 260), 0), ((130, 1000, 70), 0), ((1, 64, 64), 0),
                                          ((4100, 3000, 2500), 1024), ((2048, 2048, 2048), 512), ((70, 4096, 4096), 0),
                                          ((9000, 8200, 8300), 2048)])   # large enough for the per-GPU staging rings
def test_mul_mp_matches_oracle_for_every_device_count(lib, shape, cutoff):
    m, l, n = shape
    H.libc.srandom(21)
    A, B, C0 = H.random_matrix(m, l), H.random_matrix(l, n), H.random_matrix(m, n)
    want_mul = H.oracle().orc_mul(None, A, B, 0)
    want_add = H.oracle().orc_addmul(H.clone(C0), A, B, 0)
    for G in sorted({2, lib.m4ri_b200_device_count()}):
        lib.m4ri_b200_set_num_devices(G)
        C = H.clone(C0)
        lib.mzd_mul_mp(C, A, B, cutoff)
        assert lib.m4ri_b200_last_path().decode().startswith(f"mp{G}:")
        assert np.array_equal(H.storage(C), H.storage(want_mul))
        D = H.clone(C0)
        lib.mzd_addmul_mp(D, A, B, cutoff)
        assert np.array_equal(H.storage(D), H.storage(want_add))
        H.free(C, D)
    H.free(A, B, C0, want_mul, want_add)


def test_switching_the_device_between_large_pageable_products(lib):
    """m4ri_b200_set_device on an initialised library must tear down everything that belongs to the old GPU
    (workspace, streams, the staging ring with its events, the cached SM count) — a large pageable product on GPU 0,
    then on GPU 1, then on GPU 0 again."""
    rng = np.random.default_rng(8)
    m, l, n = 3000, 9000, 5000            # operands above the 4 MiB staging threshold
    A, B = H.new(m, l), H.new(l, n)
    for M in (A, B):
        st = H.storage(M)
        st[:, :] = rng.integers(0, 2**64, size=st.shape, dtype=np.uint64)
        st[:, M.contents.width - 1] &= np.uint64(M.contents.high_bitmask)
        st[:, M.contents.width:] = 0
    want = H.oracle().orc_mul(None, A, B, 0)
    lib.m4ri_b200_set_num_devices(1)
    for device in (0, 1, 0):
        lib.m4ri_b200_set_device(device)
        C = H.new(m, n)
        lib.mzd_mul(C, A, B, 0)
        assert np.array_equal(H.storage(C), H.storage(want)), device
        H.free(C)
    lib.m4ri_b200_set_device(0)
    H.free(A, B, want)
