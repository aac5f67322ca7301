#!/bin/bash
# Round 2, GPU session 6 (1 GPU): device PLE — parity tests, the reference's own test programs with _mzd_ple served by
# the library, timing; quick re-check of the large golden tests with the 8 x 8 block digests.
set -u
OUT=gpurun_out/r2s6; mkdir -p $OUT
stage() { echo "=== $1 ($(date +%T))" | tee -a $OUT/session.log; }
stage "PLE parity"
timeout 900 python -m pytest tests/test_zz4_ple_gpu.py -m gpu -x -q > $OUT/pytest_ple.log 2>&1; tail -15 $OUT/pytest_ple.log | tee -a $OUT/session.log
stage "reference programs under preload (PLE/PLUQ/solve/kernel/invert/elimination)"
timeout 1500 python -m pytest tests/test_reference_suite_dropin.py -m gpu -q > $OUT/pytest_dropin.log 2>&1; tail -6 $OUT/pytest_dropin.log | tee -a $OUT/session.log
stage "PLE timing"
timeout 600 python tools/ple_time.py 4096 8192 16384 32768 65536 ref:4096 ref:8192 > $OUT/ple_time.log 2>&1; cat $OUT/ple_time.log | tee -a $OUT/session.log
stage "large golden (8 x 8 block digests)"
timeout 600 python -m pytest tests/test_large_golden_gpu.py -m gpu -x -q 2>&1 | tail -3 | tee -a $OUT/session.log
stage "compute-sanitizer memcheck on a small PLE"
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_zz4_ple_gpu.py -m gpu -x -q -k "513 or 65-129 or identity" > $OUT/sanitize_ple.log 2>&1; tail -4 $OUT/sanitize_ple.log | tee -a $OUT/session.log
stage "done"
