"""GPU parity tests (-m gpu): libm4ri_b200.so, called through its C-ABI with HOST mzd_t
operands, against the oracle / the compiled reference / the golden fixtures.

Structure follows the reference's tests/test_multiplication.c (mul_/addmul_/sqr_/addsqr_
test_equality over its shape list) and tests/test_smallops.c + tests/testing.c (window
pattern preservation).  Bit-exact everywhere: this is integer work.
"""
import ctypes
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import m4ri_b200
from tests import harness as H

pytestmark = pytest.mark.gpu

with open(os.path.join(H.GOLDEN_DIR, "mul_golden.json")) as f:
    GOLD = json.load(f)


@pytest.fixture(scope="module")
def lib():
    L = m4ri_b200.load_library()
    assert L.m4ri_b200_device_count() > 0, "no CUDA device: the product has no CPU fallback"
    return L


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


def _no_excess(M):
    st = H.storage(M)
    if st.size:
        assert not np.any(st[:, M.contents.width - 1] & ~np.uint64(M.contents.high_bitmask))
        assert not np.any(st[:, M.contents.width:])


@pytest.mark.parametrize("case", GOLD["cases"], ids=lambda c: f'{c["kind"]}-{c["m"]}x{c["l"]}x{c["n"]}-k{c["k"]}-c{c["cutoff"]}')
def test_equality_on_reference_shape_list(lib, case):
    """mul_test_equality / addmul_test_equality / sqr_ / addsqr_ (test_multiplication.c:17-244):
    Strassen entry, M4RM entry and the oracle must agree; the result must also equal the digest
    the unmodified reference produced for the same inputs."""
    A, B, C = _inputs(case)
    assert H.digest(A) == case["A"]
    launches0 = lib.m4ri_b200_kernel_launches()
    if C is None:
        C1 = H.new(case["m"], case["n"])
        H.randomize(C1)  # must be overwritten
        lib.mzd_mul(C1, A, B, case["cutoff"])
        C2 = H.new(case["m"], case["n"])
        lib.mzd_mul_m4rm(C2, A, B, case["k"])
        C3 = H.new(case["m"], case["n"])
        H.randomize(C3)
        lib.mzd_mul_mp(C3, A, B, case["cutoff"])
        want = H.oracle().orc_mul(None, A, B, case["cutoff"])
        outs = [C1, C2, C3]
    else:
        C1, C2, C3 = H.clone(C), H.clone(C), H.clone(C)
        lib.mzd_addmul(C1, A, B, case["cutoff"])
        lib.mzd_addmul_m4rm(C2, A, B, case["k"])
        lib._mzd_mul_m4rm(C3, A, B, case["k"], 0)
        want = H.oracle().orc_addmul(H.clone(C), A, B, case["cutoff"])
        outs = [C1, C2, C3]
    assert lib.m4ri_b200_kernel_launches() > launches0, "no CUDA kernel was launched"
    assert H.digest(want) == case["C"]
    for out in outs:
        assert H.equal(out, want)
        assert H.digest(out) == case["C"]
        _no_excess(out)
    H.free(want, *outs)
    if C is not None:
        H.free(C)
    if B is not A:
        H.free(B)
    H.free(A)


@pytest.mark.parametrize("dims", [(2, 11, 12, 13), (64, 64, 64, 64), (100, 70, 90, 64), (513, 511, 300, 128),
                                  (2048, 1030, 1100, 256), (65, 3, 129, 0), (1, 1, 1, 0)])
def test_window_operands_and_pattern_preservation(lib, dims):
    """tests/testing.c:3-37, tests/test_smallops.c:70-91: A, B and C are windows (foreign
    rowstride, non-zero excess) of pattern-filled parents; no bit outside C's window may change."""
    m, l, n, cutoff = dims
    pat = np.uint64(0x5555AAAA3333CCCC)
    H.libc.srandom(1234)

    def windowed(rows, cols):
        P = H.new(rows + 3, (cols + 63) // 64 * 64 + 128)
        H.storage(P)[:, :] = pat
        W = H.window(P, 1, 64, 1 + rows, 64 + cols)   # odd word offset: data only 8-byte aligned
        H.randomize(W)
        return P, W

    PA, A = windowed(m, l)
    PB, B = windowed(l, n)
    PC, C = windowed(m, n)
    PC_before = H.storage(PC).copy()
    for name, ofn in (("mzd_mul", H.oracle().orc_mul), ("mzd_addmul", H.oracle().orc_addmul),
                      ("mzd_mul_m4rm", None), ("mzd_addmul_m4rm", None)):
        PW = H.clone(PC)
        H.storage(PW)[:, :] = H.storage(PC)
        Cw = H.window(PW, 1, 64, 1 + m, 64 + n)
        if ofn is not None:
            ofn(Cw, A, B, cutoff)
        elif name == "mzd_mul_m4rm":
            H.oracle().orc_mul_m4rm(Cw, A, B, 0, 1)
        else:
            H.oracle().orc_mul_m4rm(Cw, A, B, 0, 0)
        PG = H.clone(PC)
        H.storage(PG)[:, :] = H.storage(PC)
        Cg = H.window(PG, 1, 64, 1 + m, 64 + n)
        getattr(lib, name)(Cg, A, B, cutoff)
        assert np.array_equal(H.storage(PG), H.storage(PW)), name
        # rows/words outside the window still carry the pattern
        assert np.all(H.storage(PG)[0, :] == pat) and np.all(H.storage(PG)[1 + m:, :] == pat)
        assert np.all(H.storage(PG)[:, 0] == pat)
        H.free(Cw, Cg, PW, PG)
    assert np.array_equal(H.storage(PC), PC_before)
    H.free(A, B, C, PA, PB, PC)


def test_null_result_is_allocated_and_squaring_alias(lib):
    H.libc.srandom(5)
    A = H.random_matrix(300, 300)
    C = lib.mzd_mul(None, A, A, 0)            # C == NULL, A == B (strassen.c:356-363)
    want = H.oracle().orc_mul(None, A, A, 0)
    assert C.contents.nrows == 300 and C.contents.ncols == 300 and H.equal(C, want)
    D = lib.mzd_addmul(None, A, A, 0)         # fresh zero C (strassen.c:686-687)
    assert H.equal(D, want)
    E = lib.mzd_mul_m4rm(None, A, A, 0)
    assert H.equal(E, want)
    lib.m4ri_b200_mzd_free(C); lib.m4ri_b200_mzd_free(D); lib.m4ri_b200_mzd_free(E)
    H.free(A, want)


@pytest.mark.parametrize("k", [0, 1, 2, 5, 8, 9, 16])
def test_any_k_gives_identical_bits(lib, k):
    H.libc.srandom(77)
    A, B = H.random_matrix(193, 65), H.random_matrix(65, 65)
    C = H.new(193, 65)
    lib.mzd_mul_m4rm(C, A, B, k)
    want = H.oracle().orc_mul_naive(None, A, B, 1)
    assert H.equal(C, want)
    H.free(A, B, C, want)


@pytest.mark.parametrize("shape", [(0, 10, 10), (10, 0, 10), (10, 10, 0), (5, 0, 0)])
def test_empty_dimensions(lib, shape):
    m, l, n = shape
    A, B = H.new(m, l), H.new(l, n)
    C = H.new(m, n)
    if m and n:
        H.randomize(C)
    keep = H.clone(C)
    lib.mzd_addmul(C, A, B, 0)                # early-out, C unchanged (strassen.c:692-695)
    assert H.equal(C, keep)
    lib.mzd_mul(C, A, B, 0)                   # empty inner dimension: the product is zero
    zero = H.new(m, n)
    assert H.equal(C, zero)
    H.free(A, B, C, keep, zero)


_DIE_SNIPPET = r"""
import sys
sys.path.insert(0, {root!r})
import m4ri_b200
from tests import harness as H
L = m4ri_b200.load_library()
A, B, C = H.new(4, 5), H.new({brows}, 7), H.new({crows}, 7)
getattr(L, {fn!r})(C, A, B, {cutoff})
print("survived")
"""


@pytest.mark.parametrize("fn,brows,crows,cutoff", [("mzd_mul", 6, 4, 0), ("mzd_mul", 5, 3, 0), ("mzd_mul", 5, 4, -1),
                                                    ("mzd_addmul", 6, 4, 0), ("mzd_addmul", 5, 4, -1),
                                                    ("mzd_mul_m4rm", 6, 4, 0), ("mzd_addmul_m4rm", 5, 3, 0)])
def test_invalid_arguments_die_like_m4ri_die(fn, brows, crows, cutoff):
    """dimension mismatch / wrong C / cutoff < 0 -> stderr + abort() (misc.c:36-42)."""
    code = _DIE_SNIPPET.format(root=H.ROOT, fn=fn, brows=brows, crows=crows, cutoff=cutoff)
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert p.returncode == -6, (p.returncode, p.stdout, p.stderr)   # SIGABRT
    assert "survived" not in p.stdout and "mzd_" in p.stderr


def test_device_resident_api_matches_host_api(lib):
    H.libc.srandom(11)
    m, l, n = 700, 900, 1100
    A, B, C0 = H.random_matrix(m, l), H.random_matrix(l, n), H.random_matrix(m, n)
    dA, dB, dC = (lib.m4ri_b200_dmat_alloc(m, l), lib.m4ri_b200_dmat_alloc(l, n), lib.m4ri_b200_dmat_alloc(m, n))
    lib.m4ri_b200_upload(dA, A, None); lib.m4ri_b200_upload(dB, B, None); lib.m4ri_b200_upload(dC, C0, None)
    out = H.new(m, n)
    # C ^= A*B (leaf), then C ^= A*B again through Strassen with a tiny cutoff -> back to C0
    lib.m4ri_b200_dmul_m4rm(dC, dA, dB, 0, None)
    lib.m4ri_b200_download(out, dC, None)
    want = H.oracle().orc_addmul(H.clone(C0), A, B, 0)
    assert H.equal(out, want)
    lib.m4ri_b200_dmul(dC, dA, dB, 128, 0, None)
    assert lib.m4ri_b200_last_path().decode().startswith("strassen:")
    lib.m4ri_b200_download(out, dC, None)
    assert H.equal(out, C0)
    # C = A*B (clear) and the device _mzd_add
    lib.m4ri_b200_dmul(dC, dA, dB, 256, 1, None)
    lib.m4ri_b200_download(out, dC, None)
    prod = H.oracle().orc_mul(None, A, B, 0)
    assert H.equal(out, prod)
    dD = lib.m4ri_b200_dmat_alloc(m, n)
    lib.m4ri_b200_upload(dD, C0, None)
    lib.m4ri_b200_dadd(dD, dD, dC, None)
    lib.m4ri_b200_download(out, dD, None)
    assert H.equal(out, want)
    for d in (dA, dB, dC, dD):
        lib.m4ri_b200_dmat_free(d)
    H.free(A, B, C0, out, want, prod)


def _fill_fast(M, seed):
    """large inputs: numpy generator instead of 3 random() calls per word"""
    rng = np.random.default_rng(seed)
    st = H.storage(M)
    st[:, :] = rng.integers(0, 2**64, size=st.shape, dtype=np.uint64)
    w = M.contents.width
    st[:, w - 1] &= np.uint64(M.contents.high_bitmask)
    st[:, w:] = 0


@pytest.mark.parametrize("n,cutoff", [(4096, 0), (4096, 1024), (8192, 2048)])
def test_config_scale_bit_exact_against_oracle(lib, n, cutoff):
    """BASELINE config 0 scale (4096^3) and one size up, full compare against the oracle."""
    A, B = H.new(n, n), H.new(n, n)
    _fill_fast(A, 1); _fill_fast(B, 2)
    C = H.new(n, n)
    lib.mzd_mul(C, A, B, cutoff)
    want = H.oracle().orc_mul(None, A, B, 0)
    assert np.array_equal(H.storage(C), H.storage(want))
    lib.mzd_mul_m4rm(C, A, B, 0)
    assert np.array_equal(H.storage(C), H.storage(want))
    H.free(A, B, C, want)


def _gf2_matvec_rows(M, x_bits):
    """(M @ x) over GF(2) for bit-packed M [rows, words] and packed vector x [words] -> bool[rows]"""
    anded = M & x_bits[None, :]
    folded = np.bitwise_xor.reduce(anded, axis=1)
    for s in (32, 16, 8, 4, 2, 1):
        folded ^= folded >> np.uint64(s)
    return (folded & np.uint64(1)).astype(bool)


def _pack(bits):
    pad = (-len(bits)) % 64
    b = np.concatenate([bits, np.zeros(pad, dtype=bool)])
    return np.packbits(b.reshape(-1, 64)[:, ::-1], axis=1, bitorder="big").view(">u8").astype(np.uint64).ravel()


def _unpack(words, n):
    b = np.unpackbits(words.astype("<u8").view(np.uint8), bitorder="little")
    return b[:n].astype(bool)


@pytest.mark.parametrize("shape,cutoff,fn,nvec", [((16384, 16384, 16384), 0, "mzd_mul_m4rm", 40),   # BASELINE config 2
                                                  ((16384, 16384, 16384), 4096, "mzd_mul", 40),
                                                  ((4100, 20000, 9000), 2048, "mzd_mul", 40),
                                                  ((65536, 65536, 65536), 0, "mzd_mul", 16),          # config 3
                                                  ((32768, 131072, 32768), 0, "mzd_mul", 16)])        # config 5 shape
def test_full_size_freivalds_and_linearity(lib, shape, cutoff, fn, nvec):
    """Sizes the oracle cannot finish quickly: size-independent properties.
    (1) Freivalds over GF(2): for random x, (A*B)x == A(Bx), nvec vectors -> error prob 2^-nvec.
    (2) linearity: (A ^ A')*B == A*B ^ A'*B through mzd_addmul (config 5 is an addmul)."""
    m, l, n = shape
    A, B = H.new(m, l), H.new(l, n)
    _fill_fast(A, 10); _fill_fast(B, 20)
    C = H.new(m, n)
    getattr(lib, fn)(C, A, B, cutoff)
    Aw, Bw, Cw = m4ri_b200.valid_words(A), m4ri_b200.valid_words(B), m4ri_b200.valid_words(C)
    rng = np.random.default_rng(3)
    for _ in range(nvec):
        x = rng.integers(0, 2, size=n).astype(bool)
        bx = _gf2_matvec_rows(Bw, _pack(x))
        abx = _gf2_matvec_rows(Aw, _pack(bx))
        cx = _gf2_matvec_rows(Cw, _pack(x))
        assert np.array_equal(abx, cx)
    # linearity via addmul: C ^= A2*B must equal (A^A2)*B
    A2 = H.new(m, l)
    _fill_fast(A2, 30)
    lib.mzd_addmul(C, A2, B, cutoff)
    H.storage(A2)[:, :] ^= H.storage(A)
    D = H.new(m, n)
    getattr(lib, fn)(D, A2, B, cutoff)
    assert np.array_equal(H.storage(C), H.storage(D))
    H.free(A, B, C, A2, D)


def test_c_program_linked_against_both_libraries(tmp_path):
    """a C program in the style of tests/test_multiplication.c, linked with
    libm4ri_b200.so first and the unmodified reference second."""
    if not os.path.exists(H.REF_SO):
        pytest.skip("oracle/_ref/libm4ri_ref.so not present")
    exe = tmp_path / "dropin"
    libdir = os.path.dirname(m4ri_b200.LIB_PATH)
    refdir = os.path.dirname(H.REF_SO)
    subprocess.check_call(["gcc", "-O1", "-std=gnu99", "-D_DEFAULT_SOURCE", "-I", os.path.join(H.ROOT, "include"),
                           os.path.join(H.ROOT, "tests", "c", "dropin.c"), "-o", str(exe),
                           "-L", libdir, "-lm4ri_b200", "-L", refdir, "-l:libm4ri_ref.so", "-lm",
                           f"-Wl,-rpath,{libdir}", f"-Wl,-rpath,{refdir}"])
    p = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "All tests passed." in p.stdout


def test_fuzz_random_shapes_windows_and_entry_points(lib):
    """Seeded fuzz over shapes the fixed lists do not hit: every entry point, random dims from 1 to
    ~2500 (biased to word/tile boundaries +-1), A/B/C randomly windows at random word offsets."""
    rng = np.random.default_rng(20261017)
    edges = [1, 2, 63, 64, 65, 127, 128, 129, 255, 256, 257, 511, 512, 513, 1023, 1024, 1025, 2047, 2048, 2049]
    fns = ["mzd_mul", "mzd_addmul", "mzd_mul_m4rm", "mzd_addmul_m4rm", "_mzd_mul_m4rm", "_mzd_mul_even",
           "_mzd_addmul_even", "_mzd_addmul", "mzd_mul_mp", "mzd_addmul_mp"]
    O = H.oracle()

    def dim():
        return int(rng.choice(edges)) if rng.random() < 0.5 else int(rng.integers(1, 2500))

    def operand(rows, cols):
        if rng.random() < 0.5:
            M = H.new(rows, cols)
            H.storage(M)[:, :] = rng.integers(0, 2**64, size=H.storage(M).shape, dtype=np.uint64)
            H.storage(M)[:, M.contents.width - 1] &= np.uint64(M.contents.high_bitmask)
            H.storage(M)[:, M.contents.width:] = 0
            return None, M
        off_w, off_r = int(rng.integers(0, 3)), int(rng.integers(0, 3))
        P = H.new(rows + off_r + 2, (cols + 63) // 64 * 64 + 64 * (off_w + 1))
        H.storage(P)[:, :] = rng.integers(0, 2**64, size=H.storage(P).shape, dtype=np.uint64)
        return P, H.window(P, off_r, 64 * off_w, off_r + rows, 64 * off_w + cols)

    for it in range(60):
        m, l, n = dim(), dim(), dim()
        fn = fns[it % len(fns)]
        cutoff = int(rng.choice([0, 64, 128, 256, 512, 1024]))
        PA, A = operand(m, l)
        PB, B = operand(l, n)
        PC, C = operand(m, n)
        ref_parent = H.clone(PC) if PC is not None else None
        if ref_parent is not None:
            H.storage(ref_parent)[:, :] = H.storage(PC)
            Cw = H.window(ref_parent, 0, 0, 1, 1)  # placeholder, replaced below
            H.free(Cw)
            d = C.contents
            off_words = (ctypes.addressof(d.data.contents) - ctypes.addressof(PC.contents.data.contents)) // 8
            r0, w0 = divmod(off_words, PC.contents.rowstride)
            Cw = H.window(ref_parent, r0, 64 * w0, r0 + m, 64 * w0 + n)
        else:
            Cw = H.clone(C)
        accumulate = "add" in fn
        if accumulate:
            O.orc_addmul(Cw, A, B, 0)
        else:
            O.orc_mul(Cw, A, B, 0)
        if fn == "_mzd_mul_m4rm":
            lib._mzd_mul_m4rm(C, A, B, 0, 1)
        else:
            getattr(lib, fn)(C, A, B, cutoff)
        if PC is not None:
            assert np.array_equal(H.storage(PC), H.storage(ref_parent)), (it, fn, m, l, n, cutoff)
        else:
            assert np.array_equal(H.storage(C), H.storage(Cw)), (it, fn, m, l, n, cutoff)
        for parent, win in ((PA, A), (PB, B), (PC, C)):
            H.free(win)
            if parent is not None:
                H.free(parent)
        H.free(Cw)
        if ref_parent is not None:
            H.free(ref_parent)


@pytest.mark.parametrize("fn", ["mzd_mul", "mzd_addmul"])
def test_large_pageable_windows_go_through_the_staging_ring(lib, fn):
    """Operands big enough (> 4 MiB) to take the pinned staging path (csrc/staging.cu), as WINDOWS of
    pageable parents: foreign row stride, odd word offsets, partial last words — every bit outside C's
    window must survive, the result must equal the oracle's."""
    m, l, n = 6100, 7000, 6500
    rng = np.random.default_rng(77)

    def windowed(rows, cols, off_r, off_w):
        P = H.new(rows + off_r + 1, (cols + 63) // 64 * 64 + 64 * (off_w + 2))
        H.storage(P)[:, :] = rng.integers(0, 2**64, size=H.storage(P).shape, dtype=np.uint64)
        return P, H.window(P, off_r, 64 * off_w, off_r + rows, 64 * off_w + cols)

    PA, A = windowed(m, l, 1, 1)
    PB, B = windowed(l, n, 2, 3)
    PC, C = windowed(m, n, 1, 2)
    want_parent = H.clone(PC)
    H.storage(want_parent)[:, :] = H.storage(PC)
    Cw = H.window(want_parent, 1, 128, 1 + m, 128 + n)
    (H.oracle().orc_mul if fn == "mzd_mul" else H.oracle().orc_addmul)(Cw, A, B, 0)
    getattr(lib, fn)(C, A, B, 2048)
    assert np.array_equal(H.storage(PC), H.storage(want_parent))
    H.free(A, B, C, Cw, PA, PB, PC, want_parent)


def test_entry_points_from_four_host_threads(lib):
    """an OpenMP libm4ri calls _mzd_mul_even / _mzd_addmul_even from four concurrent sections
    (m4ri/mp.c:87-108) and those calls bind to this library under LD_PRELOAD: the context (workspace stack, staging
    ring, streams) is serialised by a process-wide lock, so concurrent callers get correct products."""
    import threading
    H.libc.srandom(123)
    jobs = []
    for t in range(4):
        m, l, n = 900 + 130 * t, 1100 + 64 * t, 700 + 257 * t
        A, B, C = H.random_matrix(m, l), H.random_matrix(l, n), H.random_matrix(m, n)
        want = H.oracle().orc_addmul(H.clone(C), A, B, 0)
        jobs.append((A, B, C, want))
    errors = []

    def work(job, fn):
        A, B, C, want = job
        for _ in range(6):
            D = H.clone(C)
            getattr(lib, fn)(D, A, B, 256)
            if not np.array_equal(H.storage(D), H.storage(want)):
                errors.append(fn)
            H.free(D)

    threads = [threading.Thread(target=work, args=(job, fn)) for job, fn in zip(jobs, ["mzd_addmul", "_mzd_addmul_even", "_mzd_addmul_mp4", "mzd_addmul_m4rm"])]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors
    for job in jobs:
        H.free(*job)


def test_staging_ring_survives_release(lib):
    """m4ri_b200_release() frees the pinned staging ring and stops its copy threads; the next large pageable operand
    must start a fresh ring (a re-created worker once picked up the stale job of the released ring)."""
    m, l, n = 3000, 9000, 5000          # A: 3.4 MB, B: 5.6 MB > the 4 MiB staging threshold
    A, B = H.new(m, l), H.new(l, n)
    _fill_fast(A, 5); _fill_fast(B, 6)
    want = H.oracle().orc_mul(None, A, B, 0)
    for _ in range(3):
        C = H.new(m, n)
        lib.mzd_mul(C, A, B, 0)
        assert np.array_equal(H.storage(C), H.storage(want))
        H.free(C)
        lib.m4ri_b200_release()
    H.free(A, B, want)
