#!/bin/bash
# Round 2, GPU session 9 (1 GPU): profiling evidence for the final kernels — launch list of the bench command, ncu --set
# full of the dominant launch (49 x 4096^3 on the tall-tile leaf) and of the PLE strip kernel; bench lines; PLE timing.
set -u
OUT=gpurun_out/r2s9; mkdir -p $OUT
stage() { echo "=== $1 ($(date +%T))" | tee -a $OUT/session.log; }
stage "bench lines (final code)"
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/bench_cfg3_n1.json 2> $OUT/bench_cfg3_n1.err; tail -c 400 $OUT/bench_cfg3_n1.json | tee -a $OUT/session.log
timeout 300 python bench.py --workload cfg2 --steps 10 --warmup 3 > $OUT/bench_cfg2_n1.json 2> /dev/null
timeout 600 python bench.py --workload cfg5 --steps 5 --warmup 3 > $OUT/bench_cfg5_n1.json 2> /dev/null
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> /dev/null; tail -c 300 $OUT/bench_ref.json | tee -a $OUT/session.log
stage "launch list of the bench command"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/launches_bench.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-check > $OUT/bench_under_ncu.log 2>&1
wc -l $OUT/launches_bench.csv | tee -a $OUT/session.log
stage "ncu --set full: 49 x 4096^3 (the bench's leaf launch), one launch"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:m4rm_leaf2 -s 1 -c 1 -f -o $OUT/leaf2_49x4096 \
  python tools/leaf_run.py 16384 16384 16384 2 4096 > $OUT/ncu_leaf2.log 2>&1; tail -2 $OUT/ncu_leaf2.log | tee -a $OUT/session.log
stage "ncu --set full: PLE cluster strip kernel (3 launches)"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:ple_strip_cluster -s 2 -c 3 -f -o $OUT/ple_strip \
  python tools/ple_time.py 8192 > $OUT/ncu_ple.log 2>&1; tail -2 $OUT/ncu_ple.log | tee -a $OUT/session.log
stage "PLE timing + profile"
M4RI_B200_PLE_PROFILE=1 timeout 600 python tools/ple_time.py 65536 2>&1 | tee -a $OUT/session.log
timeout 600 python tools/ple_time.py 4096 8192 16384 32768 65536 ref:8192 2>&1 | tee $OUT/ple_time.log | tee -a $OUT/session.log
stage "PLE + IO + echelon tests"
timeout 900 python -m pytest tests/test_zz4_ple_gpu.py tests/test_zz2_echelon_gpu.py -m gpu -x -q 2>&1 | tail -2 | tee -a $OUT/session.log
stage "done"
