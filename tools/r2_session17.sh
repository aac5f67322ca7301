#!/bin/bash
# Round 2, GPU session 17 (1 GPU): upper Strassen level pairs fused as well (A/B), final default leaf variant; parity; bench
# lines of all workloads; ncu --set full of the final leaf launch; whole GPU suite.
set -u
OUT=gpurun_out/r2s17; mkdir -p $OUT
stage() { echo "=== $1 ($(date +%T))" | tee -a $OUT/session.log; }
SHAPES="65536,65536,65536,4 65536,65536,65536,3 32768,32768,32768,3 16384,16384,16384,2 16384,16384,16384,-1 32768,131072,32768,3 32768,65536,16384,3"
stage "default"
timeout 300 python tools/leaf_time.py $SHAPES 2>&1 | cut -c1-200 | tee -a $OUT/session.log
stage "M4RI_B200_NO_UPPER_PAIRS=1"
M4RI_B200_NO_UPPER_PAIRS=1 timeout 300 python tools/leaf_time.py 65536,65536,65536,4 2>&1 | cut -c1-200 | tee -a $OUT/session.log
stage "whole GPU suite"
timeout 1800 python -m pytest tests -m gpu -x -q -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1; tail -3 $OUT/pytest_gpu.log | tee -a $OUT/session.log
stage "bench"
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/bench_cfg3_n1.json 2> $OUT/bench_cfg3_n1.err
timeout 300 python bench.py --workload cfg2 --steps 10 --warmup 3 > $OUT/bench_cfg2_n1.json 2> /dev/null
timeout 600 python bench.py --workload cfg5 --steps 5 --warmup 3 > $OUT/bench_cfg5_n1.json 2> /dev/null
python - <<'PY' | tee -a $OUT/session.log
import json
for f in ['bench_cfg3_n1','bench_cfg2_n1','bench_cfg5_n1']:
    d=json.loads(open(f'gpurun_out/r2s17/{f}.json').read().strip().splitlines()[-1])
    print(f, 'res %.2f ms %.3e | e2e %s %.1f ms | pinned %.1f ms | %s leaf %.3e share %.3f roof %.3f/%.3f launches %d verified %s' % (d['ms_per_step'], d['value'], d['e2e']['host_memory'], d['e2e']['ms_per_step'], d['e2e_pinned']['ms_per_step'], d['config']['path'], d['roofline']['leaf_bitops_per_s'], d['roofline']['leaf_share_of_step'], d['roofline']['frac'], d['roofline']['frac_lookup_only'], d['gpu_launches'], d['verified']))
PY
stage "ncu --set full: 49 x 4096^3 launch, final default kernel; launch list of the bench command"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:m4rm_leaf2 -s 1 -c 1 -f -o $OUT/leaf2_49x4096_final \
  python tools/leaf_run.py 16384 16384 16384 2 4096 > $OUT/ncu_leaf2.log 2>&1; tail -1 $OUT/ncu_leaf2.log | tee -a $OUT/session.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/launches_bench.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-check > $OUT/bench_under_ncu.log 2>&1
wc -l $OUT/launches_bench.csv | tee -a $OUT/session.log
stage "done"
