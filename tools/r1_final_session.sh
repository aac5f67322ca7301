#!/bin/bash
# Final 1-GPU session of the round: whole GPU suite, bench lines (own + reference arm), ncu evidence, sanitizer.
set -u
OUT=gpurun_out; mkdir -p $OUT
stage() { echo "=== $1 ($(date +%T))" | tee -a $OUT/final.log; }
stage "whole GPU suite (default = automatic leaf selection)"
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_final.log 2>&1
tail -4 $OUT/pytest_gpu_final.log | tee -a $OUT/final.log
stage "bench.py (own arm)"
timeout 300 python bench.py > $OUT/bench_final_n1.json 2> $OUT/bench_final_n1.err
tail -c 600 $OUT/bench_final_n1.json | tee -a $OUT/final.log
stage "ncu --set full: batched 7 x 8192^3 launch (the bench's leaf launch) and 16384^3"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:m4rm_leaf2 -c 1 -f \
  -o $OUT/final_leaf2_7x8192 python tools/leaf_run.py 16384 16384 16384 1 8192 > $OUT/ncu_final_7x8192.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:m4rm_leaf2 -c 1 -f \
  -o $OUT/final_leaf2_16384 python tools/leaf_run.py 16384 16384 16384 1 > $OUT/ncu_final_16384.log 2>&1
ls -la $OUT/final_*.ncu-rep 2>&1 | tee -a $OUT/final.log
stage "ncu launch list of the bench"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
  --log-file $OUT/launches_final.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $OUT/bench_under_ncu_final.log 2>&1
wc -l $OUT/launches_final.csv | tee -a $OUT/final.log
stage "compute-sanitizer, leaf variant 2 forced for every shape"
M4RI_B200_LEAF=2 timeout 200 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_run.py > $OUT/sanitize_memcheck_leaf2.log 2>&1
tail -3 $OUT/sanitize_memcheck_leaf2.log | tee -a $OUT/final.log
M4RI_B200_LEAF=2 timeout 200 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_run.py > $OUT/sanitize_racecheck_leaf2.log 2>&1
tail -3 $OUT/sanitize_racecheck_leaf2.log | tee -a $OUT/final.log
stage "bench.py --impl reference"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_final_ref.json 2> $OUT/bench_final_ref.err
tail -c 400 $OUT/bench_final_ref.json | tee -a $OUT/final.log
stage "done"
