"""GPU parity of the tall-tile M4RM leaf (leaf variant 2: 4096 x 256-bit C tiles, m4rm_leaf2.cu) — runs
last (file name) so that a failure here cannot hide the results of the established path.

The variant is forced with m4ri_b200_set_leaf_variant(2), so every shape — also ones far smaller than a
tile — goes through it; results are compared bit for bit with the oracle (small shapes), with the
1024-row leaf on the same device-resident operands (leaf-size shapes), and through the Strassen
scheduler (batched seven-product launches)."""
import pytest

import m4ri_b200
from tests import harness as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    L = m4ri_b200.load_library()
    assert L.m4ri_b200_device_count() > 0, "no CUDA device: the product has no CPU fallback"
    prev = L.m4ri_b200_set_leaf_variant(2)
    yield L
    L.m4ri_b200_set_leaf_variant(prev)


@pytest.mark.parametrize("m,l,n", [(1, 1, 1), (100, 64, 64), (513, 511, 300), (4096, 128, 256), (4096, 256, 512),
                                   (4097, 129, 257), (5000, 300, 700), (4100, 130, 260), (8192, 1024, 320),
                                   (300, 2000, 100), (12288, 256, 1000)])
def test_addmul_m4rm_matches_oracle(lib, m, l, n):
    H.libc.srandom(1000 + m + l + n)
    A, B, C = H.random_matrix(m, l), H.random_matrix(l, n), H.random_matrix(m, n)
    want = H.oracle().orc_mul_m4rm(H.clone(C), A, B, 0, 0)
    before = lib.m4ri_b200_kernel_launches()
    lib.mzd_addmul_m4rm(C, A, B, 0)
    assert lib.m4ri_b200_kernel_launches() > before
    assert H.equal(C, want)
    H.free(A, B, C, want)


@pytest.mark.parametrize("n,cutoff", [(1024, 256), (2048, 512), (4096, 1024)])
def test_strassen_batches_match_oracle(lib, n, cutoff):
    H.libc.srandom(77 + n)
    A, B = H.random_matrix(n, n), H.random_matrix(n, n)
    want = H.oracle().orc_mul(None, A, B, 0)
    got = lib.mzd_mul(None, A, B, cutoff)
    assert lib.m4ri_b200_last_path().decode().startswith("strassen:")
    assert H.equal(got, want)
    H.free(A, B, want)
    lib.m4ri_b200_mzd_free(got)


@pytest.mark.parametrize("m,l,n", [(8192, 8192, 8192), (4096, 16384, 8192), (16384, 4096, 4096 + 128)])
def test_leaf_variants_agree_on_device_resident_operands(lib, m, l, n):
    torch = pytest.importorskip("torch")
    g = torch.Generator(device="cuda").manual_seed(m + l + n)

    def rnd(r, c):
        t = torch.randint(-2**62, 2**62, (r, c // 64), dtype=torch.int64, device="cuda", generator=g)
        t ^= torch.randint(-2**62, 2**62, (r, c // 64), dtype=torch.int64, device="cuda", generator=g) << 2
        return t

    tA, tB, tC = rnd(m, l), rnd(l, n), rnd(m, n)
    out = []
    for variant in (1, 2):
        lib.m4ri_b200_set_leaf_variant(variant)
        tX = tC.clone()
        dA = lib.m4ri_b200_dmat_wrap(tA.data_ptr(), l // 64, m, l)
        dB = lib.m4ri_b200_dmat_wrap(tB.data_ptr(), n // 64, l, n)
        dX = lib.m4ri_b200_dmat_wrap(tX.data_ptr(), n // 64, m, n)
        torch.cuda.synchronize()
        lib.m4ri_b200_dmul_m4rm(dX, dA, dB, 0, None)
        lib.m4ri_b200_sync(None)
        for d in (dA, dB, dX):
            lib.m4ri_b200_dmat_free(d)
        out.append(tX)
    lib.m4ri_b200_set_leaf_variant(2)
    assert torch.equal(out[0], out[1])
    assert not torch.equal(out[0], tC)
