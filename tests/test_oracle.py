"""CPU tests (-m "not gpu"): pin the oracle (oracle/m4rm_oracle.c) against
  (1) the committed golden fixtures generated from the unmodified reference, always;
  (2) the compiled reference itself (oracle/_ref/libm4ri_ref.so) when it is present.
"""
import ctypes
import json
import os

import numpy as np
import pytest

import m4ri_b200
from tests import harness as H

with open(os.path.join(H.GOLDEN_DIR, "mul_golden.json")) as f:
    GOLD = json.load(f)
SMALL = np.load(os.path.join(H.GOLDEN_DIR, "golden_small.npz"))

needs_ref = pytest.mark.skipif(H.ref() is None, reason="oracle/_ref not built")


def test_random_word_stream_matches_reference_fixture():
    H.libc.srandom(17)
    M = H.random_matrix(2, 128)
    got = [int(x) for x in H.storage(M).ravel()]
    H.free(M)
    assert got == GOLD["random_first_words_seed17_2x128"]
    assert got[0] == 0xE87F3CB3E4308BB9 and got[1] == 0x5862505BEA655103  # SURVEY.md §4 [probed]


@pytest.mark.parametrize("k", range(1, 11))
def test_gray_code_book_matches_reference_fixture(k):
    O = H.oracle()
    ord_ = (ctypes.c_int * (1 << k))()
    inc_ = (ctypes.c_int * (1 << k))()
    O.orc_build_code(ord_, inc_, k)
    assert list(ord_) == GOLD["graycodes"][str(k)]["ord"]
    # the reference leaves inc[2^k - 1] as written by its last level; compare all slots
    assert list(inc_) == GOLD["graycodes"][str(k)]["inc"]
    assert list(ord_) == [i ^ (i >> 1) for i in range(1 << k)]


def _inputs(case):
    H.libc.srandom(case["seed"])
    if case["kind"] in ("mul", "addmul"):
        A = H.random_matrix(case["m"], case["l"])
        B = H.random_matrix(case["l"], case["n"])
    else:
        A = H.random_matrix(case["m"], case["m"])
        B = A
    C = H.random_matrix(case["m"], case["n"]) if case["kind"] in ("addmul", "addsqr") else None
    return A, B, C


@pytest.mark.parametrize("case", GOLD["cases"], ids=lambda c: f'{c["kind"]}-{c["m"]}x{c["l"]}x{c["n"]}-k{c["k"]}-c{c["cutoff"]}')
def test_oracle_reproduces_reference_golden(case):
    O = H.oracle()
    A, B, C = _inputs(case)
    assert H.digest(A) == case["A"] and H.digest(B) == case["B"]
    if C is not None:
        assert H.digest(C) == case["C_in"]
        C2 = H.clone(C)
        O.orc_addmul(C, A, B, case["cutoff"])
        O.orc_mul_m4rm(C2, A, B, case["k"], 0)
        outs = [C, C2]
    else:
        outs = [O.orc_mul(None, A, B, case["cutoff"]), O.orc_mul_m4rm(None, A, B, case["k"], 1),
                O.orc_mul_naive(None, A, B, 1)]
    for out in outs:
        assert H.digest(out) == case["C"]
        # non-window results must keep zero excess bits (m4ri/mzd.h:117-122)
        st = H.storage(out)
        if st.size:
            assert not np.any(st[:, out.contents.width - 1] & ~np.uint64(out.contents.high_bitmask))
    if "small" in case:
        assert np.array_equal(m4ri_b200.valid_words(outs[0]), SMALL[case["small"] + "_C"])
        assert np.array_equal(m4ri_b200.valid_words(A), SMALL[case["small"] + "_A"])
    H.free(*outs)
    if B is not A:
        H.free(B)
    H.free(A)


@needs_ref
@pytest.mark.parametrize("k", [1, 2, 3, 5, 8])
@pytest.mark.parametrize("ncols", [64, 65, 200, 1000])
def test_make_table_matches_compiled_reference(k, ncols):
    O, R = H.oracle(), H.ref()
    H.libc.srandom(5 + k + ncols)
    B = H.random_matrix(40, ncols)
    r = 7
    To, Tr = H.new(1 << k, ncols), R.mzd_init(1 << k, ncols)
    Lo, Lr = (ctypes.c_int * (1 << k))(), (ctypes.c_int * (1 << k))()
    O.orc_make_table(B, r, k, To, Lo)
    R.mzd_make_table(B, r, 0, k, Tr, Lr)
    assert list(Lo) == list(Lr)
    assert np.array_equal(H.storage(To), H.storage(Tr))
    # semantic check: T[L[x]] = XOR of rows r+b for bits b of x
    bw = m4ri_b200.valid_words(B)
    tw = m4ri_b200.valid_words(To)
    for x in range(1 << k):
        want = np.zeros(bw.shape[1], dtype=np.uint64)
        for b in range(k):
            if x >> b & 1:
                want ^= bw[r + b]
        assert np.array_equal(tw[Lo[x]], want)
    H.free(To, B)
    R.mzd_free(Tr)


@needs_ref
@pytest.mark.parametrize("shape", [(70, 130, 200, 0), (513, 511, 300, 4), (300, 1030, 1100, 8), (17, 64, 54, 0),
                                   (100, 100, 53, 0), (15, 200, 200, 0)])
def test_m4rm_matches_compiled_reference_same_process(shape):
    m, l, n, k = shape
    O, R = H.oracle(), H.ref()
    H.libc.srandom(99)
    A, B, C0 = H.random_matrix(m, l), H.random_matrix(l, n), H.random_matrix(m, n)
    for clear in (1, 0):
        Co, Cr = H.clone(C0), H.clone(C0)
        O.orc_mul_m4rm(Co, A, B, k, clear)
        R._mzd_mul_m4rm(Cr, A, B, k, clear)
        assert np.array_equal(H.storage(Co), H.storage(Cr))
        H.free(Co, Cr)
    H.free(A, B, C0)


@needs_ref
@pytest.mark.parametrize("dims", [(2, 11, 12, 13), (64, 64, 64, 64), (100, 70, 90, 64), (513, 511, 300, 128),
                                  (600, 1030, 1100, 256)])
def test_window_pattern_preserved_like_reference(dims):
    """tests/testing.c:3-37 + tests/test_smallops.c:70-91: products written into windows of a
    pattern-filled matrix must not touch a bit outside the window."""
    cutoff, m, l, n = dims[3], dims[0], dims[1], dims[2]
    O, R = H.oracle(), H.ref()
    pat = np.uint64(0xAAAAAAAAAAAAAAAA)
    H.libc.srandom(3)

    def windowed(rows, cols):
        P = H.new(rows + 3, (cols + 63) // 64 * 64 + 64)  # parent without excess bits
        H.storage(P)[:, : P.contents.width] = pat
        W = H.window(P, 0, 0, rows, cols)
        H.randomize(W)
        return P, W

    PA, A = windowed(m, l)
    PB, B = windowed(l, n)
    PC, C = windowed(m, n)
    PCr = H.clone(PC)
    Cr = H.window(PCr, 0, 0, m, n)
    for fn_o, fn_r in ((O.orc_mul, R.mzd_mul), (O.orc_addmul, R.mzd_addmul)):
        fn_o(C, A, B, cutoff)
        fn_r(Cr, A, B, cutoff)
        assert np.array_equal(H.storage(PC), H.storage(PCr))
    # outside the window the pattern is intact
    st = H.storage(PC)
    assert np.all(st[m:, : PC.contents.width] == pat)
    H.free(A, B, C, Cr, PA, PB, PC, PCr)


def test_large_golden_fixture_is_complete_and_inputs_reproduce():
    """tests/golden/large_golden.json (reference digests at the BASELINE sizes): every case present with 8 x 2
    block digests, and the seeded input generator reproduces the recorded input digest (config 2 checked here;
    the larger ones are checked by the GPU tests that use them)."""
    import json
    with open(os.path.join(H.GOLDEN_DIR, "large_golden.json")) as f:
        g = json.load(f)
    assert set(g["cases"]) >= {"cfg2_16384", "mid_32768", "cfg3_65536", "cfg5_32768x131072x32768"}
    for c in g["cases"].values():
        assert len(c["C_blocks"]) == H.LARGE_BLOCK_ROWS and all(len(r) == H.LARGE_BLOCK_COLS for r in c["C_blocks"])
    c = g["cases"]["cfg2_16384"]
    A = H.new(c["m"], c["l"])
    H.fill_seeded(A, H.SEED_A)
    assert H.digest(A) == c["A"]
    # a row range produced on its own (what a rank of a sharded run does) equals the same rows of the whole
    part = H.seeded_words(H.SEED_A, 100, c["l"] // 64, row0=5000)
    assert np.array_equal(part, H.storage(A)[5000:5100, :c["l"] // 64])
    H.free(A)
