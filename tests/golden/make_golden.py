"""Generate tests/golden/*.json|npz from the UNMODIFIED reference (oracle/_ref/libm4ri_ref.so,
compiled from /root/reference by `make -C oracle ref`).  Run in the build container:

    python tests/golden/make_golden.py

Inputs are the reference's own: srandom(seed) then mzd_randomize(A), mzd_randomize(B)
[, mzd_randomize(C) for addmul] in that order (tests/test_multiplication.c:17-33, 133-150).
For every case of the reference's shape list the fixture records sha256 digests of the inputs
and of the reference's result; a few tiny cases are stored in full (golden_small.npz).
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests import harness as H  # noqa: E402
import m4ri_b200  # noqa: E402

SEED = 17


def ref_matrix(R, r, c, fill=True):
    M = R.mzd_init(r, c)
    if fill:
        R.mzd_randomize(M)
    return M


def main():
    R = H.ref(required=True)
    cases = []
    small = {}

    def record(kind, m, l, n, k, cutoff, idx):
        H.libc.srandom(SEED + idx)
        if kind in ("mul", "addmul"):
            A = ref_matrix(R, m, l)
            B = ref_matrix(R, l, n)
        else:  # sqr / addsqr: B is A
            A = ref_matrix(R, m, m)
            B = A
            l = n = m
        entry = dict(kind=kind, m=m, l=l, n=n, k=k, cutoff=cutoff, seed=SEED + idx,
                     A=H.digest(A), B=H.digest(B))
        if kind in ("addmul", "addsqr"):
            C = ref_matrix(R, m, n)
            entry["C_in"] = H.digest(C)
            c_in = m4ri_b200.valid_words(C)
            R.mzd_addmul(C, A, B, cutoff)
        else:
            C = R.mzd_mul(None, A, B, cutoff)
            C2 = R.mzd_mul_m4rm(None, A, B, k)
            assert R.mzd_equal(C, C2), "reference disagrees with itself?!"
            R.mzd_free(C2)
            c_in = None
        entry["C"] = H.digest(C)
        cases.append(entry)
        if max(m, l, n) <= 257:
            key = f"{kind}_{idx}"
            small[key + "_A"] = m4ri_b200.valid_words(A)
            small[key + "_B"] = m4ri_b200.valid_words(B)
            small[key + "_C"] = m4ri_b200.valid_words(C)
            if c_in is not None:
                small[key + "_Cin"] = c_in
            entry["small"] = key
        R.mzd_free(C)
        if B is not A:
            R.mzd_free(B)
        R.mzd_free(A)

    idx = 0
    for (m, l, n, k, cutoff) in H.MUL_SHAPES:
        record("mul", m, l, n, k, cutoff, idx); idx += 1
    for (m, l, n, k, cutoff) in H.ADDMUL_SHAPES:
        record("addmul", m, l, n, k, cutoff, idx); idx += 1
    for (n, k, cutoff) in H.SQR_SHAPES:
        record("sqr", n, n, n, k, cutoff, idx); idx += 1
    for (n, k, cutoff) in H.ADDSQR_SHAPES:
        record("addsqr", n, n, n, k, cutoff, idx); idx += 1

    # first words of the reference generator after srandom(17)
    H.libc.srandom(17)
    M = ref_matrix(R, 2, 128)
    first = [int(x) for x in H.storage(M).ravel()]
    R.mzd_free(M)

    # Gray code book as the reference builds it (graycode.c:42-50)
    codes = {}
    for k in range(1, 11):
        ord_ = (H.c_int * (1 << k))()
        inc_ = (H.c_int * (1 << k))()
        R.m4ri_build_code(ord_, inc_, k)
        codes[str(k)] = dict(ord=list(ord_), inc=list(inc_))

    with open(os.path.join(H.GOLDEN_DIR, "mul_golden.json"), "w") as f:
        json.dump(dict(source="malb/m4ri@5d0d0ce via oracle/_ref/libm4ri_ref.so",
                       random_first_words_seed17_2x128=first, graycodes=codes, cases=cases), f, indent=1)
    np.savez_compressed(os.path.join(H.GOLDEN_DIR, "golden_small.npz"), **small)
    print(f"wrote {len(cases)} cases, {len(small)} small arrays")


if __name__ == "__main__":
    main()
