"""GPU parity of the tensor-core leaf (tc_leaf.cu: tcgen05.mma kind::mxf4 on bits expanded to e2m1, fp32 accumulators in
TMEM, parity epilogue) — the automatic leaf for C = A*B and C ^= A*B products whose dimensions are in its tile units (m % 128,
l % 1024, n % 256), i.e. the Strassen leaves of every large product.

It is compared bit for bit with the oracle (small shapes), with the M4RM leaves on the same device-resident operands
(leaf-size shapes), on the adversarial inputs of an arithmetic-in-floating-point scheme (all-ones operands: every
accumulator reaches the largest sum a K chunk can produce; identity; zero), through views with a wider pitch, and through
the Strassen scheduler in batches of 7 and 49 products against leaf variant 2 (M4RM only)."""
import numpy as np
import pytest

import m4ri_b200
from tests import harness as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    L = m4ri_b200.load_library()
    assert L.m4ri_b200_device_count() > 0, "no CUDA device: the product has no CPU fallback"
    prev = L.m4ri_b200_set_leaf_variant(0)
    yield L
    L.m4ri_b200_set_leaf_variant(prev)


@pytest.mark.parametrize("m,l,n", [(128, 1024, 256), (256, 2048, 512), (384, 1024, 768), (1280, 3072, 256), (128, 4096, 1024)])
def test_mul_m4rm_on_tile_unit_shapes_matches_oracle(lib, m, l, n):
    """mzd_mul_m4rm with C = A*B semantics on shapes the tensor-core leaf takes; the oracle is the checker."""
    H.libc.srandom(300 + m + l + n)
    A, B = H.random_matrix(m, l), H.random_matrix(l, n)
    want = H.oracle().orc_mul_m4rm(None, A, B, 0, 1)
    got = lib.mzd_mul_m4rm(None, A, B, 0)
    assert lib.m4ri_b200_last_leaf_variant() == 3, "the tensor-core leaf should have served this product"
    assert H.equal(got, want)
    H.free(A, B, want)
    lib.m4ri_b200_mzd_free(got)


def _wrap(lib, t, rows, cols, pitch=None):
    return lib.m4ri_b200_dmat_wrap(t.data_ptr(), pitch if pitch is not None else cols // 64, rows, cols)


@pytest.mark.parametrize("m,l,n", [(4096, 4096, 4096), (8192, 2048, 1024), (2048, 8192, 4096 + 256), (128 * 37, 1024 * 3, 256 * 5)])
def test_tensor_leaf_agrees_with_m4rm_leaves_on_device_operands(lib, m, l, n):
    torch = pytest.importorskip("torch")
    g = torch.Generator(device="cuda").manual_seed(m + 3 * l + 5 * n)

    def rnd(r, c):
        t = torch.randint(-2**62, 2**62, (r, c // 64), dtype=torch.int64, device="cuda", generator=g)
        t ^= torch.randint(-2**62, 2**62, (r, c // 64), dtype=torch.int64, device="cuda", generator=g) << 2
        return t

    tA, tB = rnd(m, l), rnd(l, n)
    out = []
    for variant in (2, 0):
        lib.m4ri_b200_set_leaf_variant(variant)
        tX = torch.full((m, n // 64), -1, dtype=torch.int64, device="cuda")
        dA, dB, dX = _wrap(lib, tA, m, l), _wrap(lib, tB, l, n), _wrap(lib, tX, m, n)
        torch.cuda.synchronize()
        lib.m4ri_b200_dmul_m4rm(dX, dA, dB, 1, None)
        lib.m4ri_b200_sync(None)
        assert lib.m4ri_b200_last_leaf_variant() == (3 if variant == 0 else 2)
        for d in (dA, dB, dX):
            lib.m4ri_b200_dmat_free(d)
        out.append(tX)
    lib.m4ri_b200_set_leaf_variant(0)
    assert torch.equal(out[0], out[1])


def test_extreme_sums_identity_and_zero(lib):
    """All-ones operands drive every fp32 accumulator to the largest value a 1024-element K chunk can produce (1024, then
    parity 0 per chunk); the identity and the zero matrix check the other end."""
    torch = pytest.importorskip("torch")
    m, l, n = 512, 4096, 1024
    ones_a = torch.full((m, l // 64), -1, dtype=torch.int64, device="cuda")
    ones_b = torch.full((l, n // 64), -1, dtype=torch.int64, device="cuda")
    tX = torch.full((m, n // 64), 0x5555, dtype=torch.int64, device="cuda")
    dA, dB, dX = _wrap(lib, ones_a, m, l), _wrap(lib, ones_b, l, n), _wrap(lib, tX, m, n)
    lib.m4ri_b200_dmul_m4rm(dX, dA, dB, 1, None)
    lib.m4ri_b200_sync(None)
    assert lib.m4ri_b200_last_leaf_variant() == 3
    assert not bool(tX.any()), "4096 ones sum to an even number: the product of all-ones operands is zero"
    # l - 1 ones in every row of A (clear bit 0 of each row): odd sums -> all-ones product
    ones_a[:, 0] = -2
    lib.m4ri_b200_dmul_m4rm(dX, dA, dB, 1, None)
    lib.m4ri_b200_sync(None)
    assert bool((tX == -1).all())
    # identity times random = random
    k = 2048
    ident = torch.zeros((k, k // 64), dtype=torch.int64, device="cuda")
    idx = torch.arange(k, device="cuda")
    ident[idx, idx // 64] = torch.where(idx % 64 == 63, torch.tensor(-2**63, device="cuda"), torch.tensor(1, device="cuda") << (idx % 64))
    g = torch.Generator(device="cuda").manual_seed(9)
    R = torch.randint(-2**62, 2**62, (k, k // 64), dtype=torch.int64, device="cuda", generator=g)
    Y = torch.zeros_like(R)
    dI, dR, dY = _wrap(lib, ident, k, k), _wrap(lib, R, k, k), _wrap(lib, Y, k, k)
    lib.m4ri_b200_dmul_m4rm(dY, dI, dR, 1, None)
    lib.m4ri_b200_sync(None)
    assert torch.equal(Y, R)
    lib.m4ri_b200_dmul_m4rm(dY, dR, dI, 1, None)
    lib.m4ri_b200_sync(None)
    assert torch.equal(Y, R)
    Z = torch.zeros_like(R)
    dZ = _wrap(lib, Z, k, k)
    lib.m4ri_b200_dmul_m4rm(dY, dZ, dR, 1, None)
    lib.m4ri_b200_sync(None)
    assert not bool(Y.any())
    for d in (dA, dB, dX, dI, dR, dY, dZ):
        lib.m4ri_b200_dmat_free(d)


def test_views_with_a_wider_pitch(lib):
    """Operands and result as windows of wider device matrices (the Strassen scheduler hands the leaf such views)."""
    torch = pytest.importorskip("torch")
    m, l, n, pad = 1024, 2048, 512, 4          # pitches 2048/64 + 4, 512/64 + 4 words (even: 16-byte aligned rows)
    g = torch.Generator(device="cuda").manual_seed(21)
    bigA = torch.randint(-2**62, 2**62, (m, l // 64 + pad), dtype=torch.int64, device="cuda", generator=g)
    bigB = torch.randint(-2**62, 2**62, (l, n // 64 + pad), dtype=torch.int64, device="cuda", generator=g)
    bigX = torch.full((m, n // 64 + pad), 7, dtype=torch.int64, device="cuda")
    out = []
    for variant in (2, 0):
        lib.m4ri_b200_set_leaf_variant(variant)
        X = bigX.clone()
        dA = _wrap(lib, bigA, m, l, l // 64 + pad)
        dB = _wrap(lib, bigB, l, n, n // 64 + pad)
        dX = _wrap(lib, X, m, n, n // 64 + pad)
        lib.m4ri_b200_dmul_m4rm(dX, dA, dB, 1, None)
        lib.m4ri_b200_sync(None)
        for d in (dA, dB, dX):
            lib.m4ri_b200_dmat_free(d)
        out.append(X)
    lib.m4ri_b200_set_leaf_variant(0)
    assert torch.equal(out[0], out[1])
    assert bool((out[1][:, n // 64:] == 7).all()), "words beyond the view must stay untouched"


@pytest.mark.parametrize("n,cutoff,launches", [(4096, 2048, 7), (8192, 2048, 49)])
def test_strassen_batches_of_tensor_leaves(lib, n, cutoff, launches):
    """7 and 49 tensor-core leaf products per launch under the Strassen scheduler against the M4RM-only schedule (bit for
    bit) and against Freivalds' check on the host."""
    rng = np.random.default_rng(n)
    A, B = H.new(n, n), H.new(n, n)
    for M in (A, B):
        H.storage(M)[:, :] = rng.integers(0, 2**64, size=H.storage(M).shape, dtype=np.uint64)
    res = []
    for variant in (2, 0):
        lib.m4ri_b200_set_leaf_variant(variant)
        C = H.new(n, n)
        lib.mzd_mul(C, A, B, cutoff)
        assert lib.m4ri_b200_last_path().decode().startswith("strassen:")
        assert lib.m4ri_b200_last_leaf_variant() == (3 if variant == 0 else 2)
        res.append(C)
    lib.m4ri_b200_set_leaf_variant(0)
    assert np.array_equal(H.storage(res[0]), H.storage(res[1]))
    Aw, Bw, Cw = (m4ri_b200.valid_words(M) for M in (A, B, res[1]))

    def matvec(M, x):
        f = np.bitwise_xor.reduce(M & x[None, :], axis=1)
        for sft in (32, 16, 8, 4, 2, 1):
            f ^= f >> np.uint64(sft)
        bits = (f & np.uint64(1)).astype(np.uint8)
        return np.packbits(bits.reshape(-1, 64)[:, ::-1], axis=1, bitorder="big").view(">u8").astype(np.uint64).ravel()

    for _ in range(4):
        x = rng.integers(0, 2**64, size=n // 64, dtype=np.uint64)
        assert np.array_equal(matvec(Aw, matvec(Bw, x)), matvec(Cw, x))
    H.free(A, B, *res)


@pytest.mark.parametrize("m,l,n,cutoff,path", [(5000, 9000, 3000, 0, "m4rm"), (9000, 18000, 6000, 4096, "strassen:1"),
                                               (5003, 8999, 3001, 0, "m4rm")])
def test_odd_shapes_are_padded_to_the_tile_units(lib, m, l, n, cutoff, path):
    """Sizes that are no multiples of anything: the host path pads the device operands with zeros to the tensor leaf's tile
    units (capi.cu: padded_dims), the result window is what the M4RM-only schedule gives, bit for bit, and passes
    Freivalds' check."""
    rng = np.random.default_rng(m + n)
    A, B = H.new(m, l), H.new(l, n)
    for M in (A, B):
        st = H.storage(M)
        st[:, :] = rng.integers(0, 2**64, size=st.shape, dtype=np.uint64)
        st[:, M.contents.width - 1] &= np.uint64(M.contents.high_bitmask)
        st[:, M.contents.width:] = 0
    res = []
    for variant in (2, 0):
        lib.m4ri_b200_set_leaf_variant(variant)
        C = H.new(m, n)
        lib.mzd_mul(C, A, B, cutoff)
        if variant == 0:
            assert lib.m4ri_b200_last_path().decode() == path
            assert lib.m4ri_b200_last_leaf_variant() == 3
        res.append(C)
    lib.m4ri_b200_set_leaf_variant(0)
    assert np.array_equal(H.storage(res[0]), H.storage(res[1]))
    H.free(A, B, *res)


@pytest.mark.parametrize("m,l,n", [(256, 1024, 256), (640, 3072, 768)])
def test_accumulating_products_on_the_tensor_leaf(lib, m, l, n):
    """C ^= A*B at leaf level (mzd_addmul_m4rm): the kernel XORs its partial results into C, so the accumulating form is
    the same kernel without the clearing of C."""
    H.libc.srandom(5 + m)
    A, B, C = H.random_matrix(m, l), H.random_matrix(l, n), H.random_matrix(m, n)
    want = H.oracle().orc_mul_m4rm(H.clone(C), A, B, 0, 0)
    lib.mzd_addmul_m4rm(C, A, B, 0)
    assert lib.m4ri_b200_last_leaf_variant() == 3
    assert H.equal(C, want)
    H.free(A, B, C, want)


def test_experimental_entry_points_agree(lib):
    """m4ri_b200_dmul_tc (the simple output-stationary form kept for cross-checking) and m4ri_b200_dmul_tc2."""
    torch = pytest.importorskip("torch")
    m, l, n = 256, 1024, 512
    g = torch.Generator(device="cuda").manual_seed(4)
    tA = torch.randint(-2**62, 2**62, (m, l // 64), dtype=torch.int64, device="cuda", generator=g)
    tB = torch.randint(-2**62, 2**62, (l, n // 64), dtype=torch.int64, device="cuda", generator=g)
    tBt = torch.zeros((n, l // 64), dtype=torch.int64, device="cuda")
    X1 = torch.zeros((m, n // 64), dtype=torch.int64, device="cuda")
    X2 = torch.full((m, n // 64), -1, dtype=torch.int64, device="cuda")
    dA, dB, dBt, d1, d2 = _wrap(lib, tA, m, l), _wrap(lib, tB, l, n), _wrap(lib, tBt, n, l), _wrap(lib, X1, m, n), _wrap(lib, X2, m, n)
    lib.m4ri_b200_dtranspose(dBt, dB, None)
    lib.m4ri_b200_dmul_tc(d1, dA, dBt, 1, None)
    lib.m4ri_b200_dmul_tc2(d2, dA, dB, None)
    lib.m4ri_b200_sync(None)
    assert torch.equal(X1, X2)
    lib.m4ri_b200_dmul_tc(d1, dA, dBt, 0, None)      # accumulate the same product again: zero
    lib.m4ri_b200_sync(None)
    assert not bool(X1.any())
    for d in (dA, dB, dBt, d1, d2):
        lib.m4ri_b200_dmat_free(d)
