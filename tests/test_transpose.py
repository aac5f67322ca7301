"""mzd_transpose as a device op.
CPU part: the oracle's definition-level transpose is pinned against the compiled reference on the shapes of
the reference's tests/test_transpose.c style (tiny, word-boundary +-1, windows).  GPU part: the device
kernel through m4ri_b200_transpose (host mzd_t in/out, reference semantics) against the oracle, plus
involution at a size the oracle does not need to see."""
import numpy as np
import pytest

import m4ri_b200
from tests import harness as H

SHAPES = [(1, 1), (1, 64), (64, 1), (3, 131), (63, 65), (64, 64), (65, 63), (128, 200), (1000, 70), (1025, 1023),
          (2048, 513), (31, 4100)]


def _case(m, n, window, seed, zero_excess=False):
    H.libc.srandom(seed)
    if not window:
        return None, H.random_matrix(m, n)
    P = H.random_matrix(m + 2, (n + 63) // 64 * 64 + 128)
    W = H.window(P, 1, 64, 1 + m, 64 + n)
    if zero_excess and n % 64:      # clear the window's excess bits inside the parent
        H.storage(P)[1:1 + m, 1 + n // 64] &= np.uint64((1 << (n % 64)) - 1)
    return P, W


@pytest.mark.skipif(H.ref() is None, reason="oracle/_ref not built")
@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("window", [False, True])
def test_oracle_transpose_matches_compiled_reference(shape, window):
    """The reference's block transpose reads whole source words (m4ri/mzd.c:1104-1116): for a windowed
    source its result is only right when the window's excess bits are zero (observed: random excess bits
    leak into valid bits of the result).  The oracle — and the device kernel — follow the
    definition DST[j][i] = A[i][j] on valid bits for every source."""
    m, n = shape
    P, A = _case(m, n, window, 31 + m, zero_excess=True)
    want = H.ref().mzd_transpose(None, A)
    got = H.oracle().orc_transpose(None, A)
    assert np.array_equal(H.storage(got)[:, : got.contents.width], H.storage(want)[:, : want.contents.width])
    H.ref().mzd_free(want)
    H.free(got, A, P)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", SHAPES + [(3000, 5000), (4096, 4096)])
@pytest.mark.parametrize("window", [False, True])
def test_device_transpose_matches_oracle(shape, window):
    lib = m4ri_b200.load_library()
    m, n = shape
    P, A = _case(m, n, window, 57 + n)
    want = H.oracle().orc_transpose(None, A)
    # destination: a window of a pattern-filled parent, to prove nothing outside is written
    DP = H.new(n + 2, (m + 63) // 64 * 64 + 128)
    H.storage(DP)[:, :] = np.uint64(0x0123456789ABCDEF)
    before = H.storage(DP).copy()
    D = H.window(DP, 1, 64, 1 + n, 64 + m)
    launches = lib.m4ri_b200_kernel_launches()
    lib.m4ri_b200_transpose(D, A)
    assert lib.m4ri_b200_kernel_launches() > launches
    assert H.equal(D, want)
    expect = before.copy()
    expect_window = m4ri_b200.words(D)          # valid region as written
    after = H.storage(DP)
    # rows outside and words outside the window are untouched
    assert np.array_equal(after[0], before[0]) and np.array_equal(after[1 + n:], before[1 + n:])
    assert np.array_equal(after[:, 0], before[:, 0])
    wlast = 1 + (m + 63) // 64
    assert np.array_equal(after[:, wlast:], before[:, wlast:])
    if m % 64:   # excess bits of the window's last word keep the pattern
        mask = np.uint64(~((1 << (m % 64)) - 1) & (2**64 - 1))
        assert np.all((after[1:1 + n, wlast - 1] & mask) == (before[1:1 + n, wlast - 1] & mask))
    out = lib.m4ri_b200_transpose(None, A)    # DST == NULL allocates
    assert H.equal(out, want)
    lib.m4ri_b200_mzd_free(out)
    H.free(want, D, DP, A, P)


@pytest.mark.gpu
def test_device_transpose_involution_large():
    lib = m4ri_b200.load_library()
    m, n = 20000, 33000
    A = H.new(m, n)
    rng = np.random.default_rng(5)
    st = H.storage(A)
    st[:, :] = rng.integers(0, 2**64, size=st.shape, dtype=np.uint64)
    st[:, A.contents.width - 1] &= np.uint64(A.contents.high_bitmask)
    st[:, A.contents.width:] = 0
    T = lib.m4ri_b200_transpose(None, A)
    # spot-check 2000 random entries against the definition, then transpose back
    aw, tw = m4ri_b200.words(A), m4ri_b200.words(T)
    ii, jj = rng.integers(0, m, 2000), rng.integers(0, n, 2000)
    abits = (aw[ii, jj // 64] >> (jj % 64).astype(np.uint64)) & np.uint64(1)
    tbits = (tw[jj, ii // 64] >> (ii % 64).astype(np.uint64)) & np.uint64(1)
    assert np.array_equal(abits, tbits)
    back = lib.m4ri_b200_transpose(None, T)
    assert np.array_equal(m4ri_b200.valid_words(back), m4ri_b200.valid_words(A))
    lib.m4ri_b200_mzd_free(T); lib.m4ri_b200_mzd_free(back)
    H.free(A)
