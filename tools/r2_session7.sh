#!/bin/bash
# Round 2, GPU session 7 (1 GPU): PLE with the cluster strip kernel — parity (both strip variants), phase profile, timing;
# new tests (exported elimination symbols, wide pitch, four host threads).
set -u
OUT=gpurun_out/r2s7; mkdir -p $OUT
stage() { echo "=== $1 ($(date +%T))" | tee -a $OUT/session.log; }
stage "PLE parity, cluster strips"
timeout 900 python -m pytest tests/test_zz4_ple_gpu.py -m gpu -x -q > $OUT/pytest_ple.log 2>&1; tail -8 $OUT/pytest_ple.log | tee -a $OUT/session.log
stage "PLE parity, one-CTA strips"
M4RI_B200_PLE_STRIP=0 timeout 900 python -m pytest tests/test_zz4_ple_gpu.py -m gpu -x -q 2>&1 | tail -2 | tee -a $OUT/session.log
stage "PLE phase profile + timing"
M4RI_B200_PLE_PROFILE=1 timeout 600 python tools/ple_time.py 16384 65536 > $OUT/ple_profile.log 2>&1; cat $OUT/ple_profile.log | tee -a $OUT/session.log
M4RI_B200_PLE_STRIP=0 M4RI_B200_PLE_PROFILE=1 timeout 600 python tools/ple_time.py 16384 > $OUT/ple_profile_v0.log 2>&1; cat $OUT/ple_profile_v0.log | tee -a $OUT/session.log
timeout 600 python tools/ple_time.py 4096 8192 16384 32768 65536 > $OUT/ple_time.log 2>&1; cat $OUT/ple_time.log | tee -a $OUT/session.log
stage "new tests"
timeout 900 python -m pytest tests/test_zz2_echelon_gpu.py tests/test_parity_gpu.py -m gpu -x -q -k "libm4ri_named or wider_pitch or four_host_threads" 2>&1 | tail -4 | tee -a $OUT/session.log
stage "sanitizer on the cluster strip kernel"
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_zz4_ple_gpu.py -m gpu -x -q -k "513 or 2100 or identity" 2>&1 | tail -3 | tee -a $OUT/session.log
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_zz4_ple_gpu.py -m gpu -x -q -k "2100-2050-random" 2>&1 | tail -3 | tee -a $OUT/session.log
stage "done"
