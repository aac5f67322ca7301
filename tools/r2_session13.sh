#!/bin/bash
# Round 2, GPU session 13 (1 GPU): two-level fused Winograd additions A/B, parity, bench lines.
set -u
OUT=gpurun_out/r2s13; mkdir -p $OUT
stage() { echo "=== $1 ($(date +%T))" | tee -a $OUT/session.log; }
stage "parity with the fused two-level additions"
timeout 900 python -m pytest tests/test_large_golden_gpu.py tests/test_parity_gpu.py tests/test_zz_leaf2_gpu.py -m gpu -x -q 2>&1 | tail -3 | tee -a $OUT/session.log
for F in fused twopass; do
  stage "additions: $F"
  if [ $F = twopass ]; then export M4RI_B200_NO_FUSED2=1; else unset M4RI_B200_NO_FUSED2; fi
  timeout 300 python tools/leaf_time.py 65536,65536,65536,4 32768,32768,32768,3 16384,16384,16384,2 32768,131072,32768,3 32768,65536,16384,3 2>&1 | cut -c1-200 | tee -a $OUT/session.log
done
unset M4RI_B200_NO_FUSED2
stage "bench lines"
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/bench_cfg3_n1.json 2> $OUT/bench_cfg3_n1.err
timeout 600 python bench.py --workload cfg5 --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_cfg5_n1.json 2> /dev/null
python - <<'PY' | tee -a $OUT/session.log
import json
for f in ['bench_cfg3_n1','bench_cfg5_n1']:
    d=json.loads(open(f'gpurun_out/r2s13/{f}.json').read().strip().splitlines()[-1])
    print(f, 'res %.2f ms %.3e | e2e %s %.1f ms | pinned %.1f ms | %s leaf %.3e share %.3f roof %.3f/%.3f traffic %s launches %d verified %s' % (d['ms_per_step'], d['value'], d['e2e']['host_memory'], d['e2e']['ms_per_step'], d['e2e_pinned']['ms_per_step'], d['config']['path'], d['roofline']['leaf_bitops_per_s'], d['roofline']['leaf_share_of_step'], d['roofline']['frac'], d['roofline']['frac_lookup_only'], d['roofline']['traffic'], d['gpu_launches'], d['verified']))
PY
stage "rest of the GPU suite"
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_large_golden_gpu.py --deselect tests/test_parity_gpu.py --deselect tests/test_zz_leaf2_gpu.py 2>&1 | tail -3 | tee -a $OUT/session.log
stage "done"
