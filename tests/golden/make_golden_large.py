"""Generate tests/golden/large_golden.json: digests of the UNMODIFIED reference's outputs on the
BASELINE.json configurations the oracle port cannot finish in seconds.  Run in the
build container (needs /root/reference compiled into oracle/_ref by `make -C oracle ref`):

    python tests/golden/make_golden_large.py [--only NAME]

Inputs: harness.fill_seeded (raw PCG64 words, seeds 101/102/103 for A/B/C_in) — a generator
both sides can run, as SURVEY §8d allows for the 65536-size cases.  Outputs: the reference's
mzd_mul_mp / mzd_addmul_mp (OpenMP build, the reference's fastest CPU path; identical bits to
mzd_mul — checked here at 16384^3 against the serial mzd_mul and mzd_mul_m4rm).  For every case
the fixture holds the sha256 of A, B, (C_in) and C, and of C's 8 x 2 blocks (harness.
large_block_digests) so that sharded runs can check per-rank blocks.  The reference's wall time
and thread count in THIS container are recorded too (a baseline of the build box, not the GPU box).
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ["OMP_NUM_THREADS"] = str(len(os.sched_getaffinity(0)))
from tests import harness as H  # noqa: E402
import m4ri_b200  # noqa: E402

CASES = [  # name, kind, m, l, n, BASELINE config
    ("cfg2_16384", "mul", 16384, 16384, 16384, 2),
    ("mid_32768", "mul", 32768, 32768, 32768, None),
    ("cfg5_32768x131072x32768", "addmul", 32768, 131072, 32768, 5),
    ("cfg3_65536", "mul", 65536, 65536, 65536, 3),
]
OUT = os.path.join(H.GOLDEN_DIR, "large_golden.json")


def fill(R, M, seed):
    m = M.contents
    st = H.storage(M)
    st[:, :m.width] = H.seeded_words(seed, m.nrows, m.width)
    st[:, m.width:] = 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    R = H.ref_omp()
    assert R is not None, "oracle/_ref/libm4ri_ref_omp.so missing (make -C oracle ref)"
    S = H.ref(required=True)
    out = {}
    if os.path.exists(OUT):
        with open(OUT) as f:
            out = json.load(f)
    out["source"] = "malb/m4ri@5d0d0ce via oracle/_ref/libm4ri_ref_omp.so (mzd_mul_mp / mzd_addmul_mp, cutoff 0)"
    out["inputs"] = "harness.fill_seeded: PCG64 raw words, seeds A=101 B=102 C_in=103"
    out.setdefault("cases", {})
    for name, kind, m, l, n, cfg in CASES:
        if args.only and args.only != name:
            continue
        A, B = R.mzd_init(m, l), R.mzd_init(l, n)
        fill(R, A, H.SEED_A)
        fill(R, B, H.SEED_B)
        entry = dict(kind=kind, m=m, l=l, n=n, baseline_config=cfg, A=H.digest(A), B=H.digest(B))
        if kind == "addmul":
            C = R.mzd_init(m, n)
            fill(R, C, H.SEED_C)
            entry["C_in"] = H.digest(C)
            t0 = time.perf_counter()
            R.mzd_addmul_mp(C, A, B, 0)
            dt = time.perf_counter() - t0
        else:
            t0 = time.perf_counter()
            C = R.mzd_mul_mp(None, A, B, 0)
            dt = time.perf_counter() - t0
        if max(m, l, n) <= 16384:   # the OpenMP block split and the serial paths give the same bits
            # (loaded as separate libraries: matrices do not cross allocators)
            A2, B2 = S.mzd_init(m, l), S.mzd_init(l, n)
            fill(S, A2, H.SEED_A)
            fill(S, B2, H.SEED_B)
            C2 = S.mzd_mul(None, A2, B2, 0)
            C3 = S.mzd_mul_m4rm(None, A2, B2, 0)
            assert H.digest(C2) == H.digest(C) == H.digest(C3)
            for M in (A2, B2, C2, C3):
                S.mzd_free(M)
        entry["C"] = H.digest(C)
        entry["C_blocks"] = H.large_block_digests(m4ri_b200.valid_words(C))
        entry["reference_seconds_build_container"] = round(dt, 3)
        entry["reference_threads_build_container"] = int(os.environ["OMP_NUM_THREADS"])
        out["cases"][name] = entry
        print(name, f"{dt:.2f} s", f"{2.0 * m * l * n / dt:.3e} bit-ops/s", entry["C"][:16], flush=True)
        for M in (A, B, C):
            R.mzd_free(M)
        with open(OUT, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
