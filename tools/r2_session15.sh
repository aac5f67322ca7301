#!/bin/bash
# Round 2, GPU session 15 (1 GPU): store mode of the tall-tile leaf (C = A*B without the zero fill of the product
# temporaries), LDS.128 A loads now that the kernel compiles without spills; parity; bench.
set -u
OUT=gpurun_out/r2s15; mkdir -p $OUT
stage() { echo "=== $1 ($(date +%T))" | tee -a $OUT/session.log; }
stage "parity"
timeout 900 python -m pytest tests/test_large_golden_gpu.py tests/test_parity_gpu.py tests/test_zz_leaf2_gpu.py tests/test_trsm_gpu.py tests/test_zz4_ple_gpu.py -m gpu -x -q 2>&1 | tail -3 | tee -a $OUT/session.log
M4RI_B200_LEAF2_AWIDE=1 timeout 900 python -m pytest tests/test_large_golden_gpu.py tests/test_zz_leaf2_gpu.py -m gpu -x -q 2>&1 | tail -2 | tee -a $OUT/session.log
SHAPES="65536,65536,65536,4 32768,32768,32768,3 16384,16384,16384,2 16384,16384,16384,-1 32768,131072,32768,3 32768,65536,16384,3 8192,8192,8192,1"
stage "default (store mode on)"
timeout 300 python tools/leaf_time.py $SHAPES 2>&1 | cut -c1-200 | tee -a $OUT/session.log
stage "M4RI_B200_LEAF2_STORE=0"
M4RI_B200_LEAF2_STORE=0 timeout 300 python tools/leaf_time.py $SHAPES 2>&1 | cut -c1-200 | tee -a $OUT/session.log
stage "M4RI_B200_LEAF2_AWIDE=1"
M4RI_B200_LEAF2_AWIDE=1 timeout 300 python tools/leaf_time.py $SHAPES 2>&1 | cut -c1-200 | tee -a $OUT/session.log
stage "M4RI_B200_LEAF2_SPLIT=0"
M4RI_B200_LEAF2_SPLIT=0 timeout 300 python tools/leaf_time.py $SHAPES 2>&1 | cut -c1-200 | tee -a $OUT/session.log
stage "bench"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_cfg3_n1.json 2> $OUT/bench_cfg3_n1.err
python - <<'PY' | tee -a $OUT/session.log
import json
d=json.loads(open('gpurun_out/r2s15/bench_cfg3_n1.json').read().strip().splitlines()[-1])
print('res %.2f ms %.3e | e2e %s %.1f ms | pinned %.1f ms | %s leaf %.3e share %.3f roof %.3f/%.3f launches %d verified %s' % (d['ms_per_step'], d['value'], d['e2e']['host_memory'], d['e2e']['ms_per_step'], d['e2e_pinned']['ms_per_step'], d['config']['path'], d['roofline']['leaf_bitops_per_s'], d['roofline']['leaf_share_of_step'], d['roofline']['frac'], d['roofline']['frac_lookup_only'], d['gpu_launches'], d['verified']))
PY
stage "done"
