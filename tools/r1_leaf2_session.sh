#!/bin/bash
# One GPU session for the tall-tile leaf (variant 2): parity first, then timings against the 1024-row leaf,
# the bench line in automatic mode, ncu captures, and — with what is left — the whole GPU suite in automatic mode.
# Every stage has its own timeout and logs to gpurun_out/; a failing stage does not stop the later ones
# except that the ncu/bench stages are skipped when the first parity check fails.
set -u
OUT=gpurun_out
mkdir -p $OUT
stage() { echo "=== $1 ($(date +%T))" | tee -a $OUT/session.log; }

stage "smoke of leaf variant 2 (one tile)"
M4RI_B200_LEAF=2 timeout 150 python -m pytest tests/test_zz_leaf2_gpu.py -x -q -k "4096-128-256 or 100-64-64" > $OUT/leaf2_first.log 2>&1
first=$?
tail -3 $OUT/leaf2_first.log | tee -a $OUT/session.log
if [ $first -ne 0 ]; then
  stage "variant 2 failed: compute-sanitizer on one small product"
  M4RI_B200_LEAF=2 timeout 200 compute-sanitizer --tool memcheck python tools/leaf_run.py 4096 128 256 1 > $OUT/leaf2_memcheck.log 2>&1
  tail -30 $OUT/leaf2_memcheck.log | tee -a $OUT/session.log
  stage "established path: full GPU suite"
  timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_default.log 2>&1
  tail -5 $OUT/pytest_gpu_default.log | tee -a $OUT/session.log
  exit 0
fi

stage "parity of leaf variant 2"
timeout 300 python -m pytest tests/test_zz_leaf2_gpu.py -x -q > $OUT/leaf2_parity.log 2>&1
tail -3 $OUT/leaf2_parity.log | tee -a $OUT/session.log

stage "leaf timings: variant 1, variant 2 (A by LDS.64), variant 2 (A by LDS.128)"
SHAPES="8192,8192,8192,-1 16384,16384,16384,-1 4096,8192,8192,-1 16384,16384,16384,1 65536,65536,65536,8192"
M4RI_B200_LEAF=1 timeout 150 python tools/leaf_time.py $SHAPES > $OUT/time_leaf1.log 2>&1
M4RI_B200_LEAF=0 timeout 150 python tools/leaf_time.py $SHAPES > $OUT/time_leaf2.log 2>&1
M4RI_B200_LEAF=0 M4RI_B200_LEAF2_AWIDE=1 timeout 150 python tools/leaf_time.py $SHAPES > $OUT/time_leaf2_awide.log 2>&1
for f in time_leaf1 time_leaf2 time_leaf2_awide; do echo "--- $f"; cat $OUT/$f.log; done | tee -a $OUT/session.log

stage "bench.py, automatic leaf selection"
M4RI_B200_LEAF=0 timeout 300 python bench.py > $OUT/bench_leaf_auto.json 2> $OUT/bench_leaf_auto.err
tail -c 1500 $OUT/bench_leaf_auto.json | tee -a $OUT/session.log

stage "ncu --set full: one 8192^3 launch and one batched 7 x 8192^3 launch of variant 2"
M4RI_B200_LEAF=2 timeout 240 ncu --set full --clock-control none --import-source on -k regex:m4rm_leaf2 -c 1 -f \
  -o $OUT/leaf2_8192 python tools/leaf_run.py 8192 8192 8192 1 > $OUT/ncu_leaf2_8192.log 2>&1
M4RI_B200_LEAF=2 timeout 240 ncu --set full --clock-control none --import-source on -k regex:m4rm_leaf2 -c 1 -f \
  -o $OUT/leaf2_7x8192 python tools/leaf_run.py 16384 16384 16384 1 8192 > $OUT/ncu_leaf2_7x8192.log 2>&1
ls -la $OUT/*.ncu-rep 2>&1 | tee -a $OUT/session.log

stage "whole GPU suite, automatic leaf selection"
M4RI_B200_LEAF=0 timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_auto.log 2>&1
tail -5 $OUT/pytest_gpu_auto.log | tee -a $OUT/session.log

stage "ncu launch list of the bench, automatic leaf selection"
M4RI_B200_LEAF=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
  --log-file $OUT/launches_r1c.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $OUT/bench_under_ncu.log 2>&1
wc -l $OUT/launches_r1c.csv | tee -a $OUT/session.log
stage "done"
