#!/bin/bash
# Round 2, GPU session 10 (2 GPUs): in-process mzd_mul_mp with the quadrant-hooks pipeline (tests, timing, bench e2e_inproc);
# L2 residency hints of the tall-tile leaf A/B (time + DRAM bytes); PLE timing repeat.
set -u
OUT=gpurun_out/r2s10; mkdir -p $OUT
stage() { echo "=== $1 ($(date +%T))" | tee -a $OUT/session.log; }
stage "in-process mzd_mul_mp (quadrant hooks), 2 GPUs: tests + timing"
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -x -q > $OUT/pytest_multigpu_g2.log 2>&1; tail -5 $OUT/pytest_multigpu_g2.log | tee -a $OUT/session.log
timeout 300 python tools/mp_time.py 65536 2 2>&1 | tee -a $OUT/session.log
stage "bench N=2 with the in-process leg"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561"
timeout 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 > $OUT/n2_cfg3.json 2> $OUT/n2_cfg3.err
python - <<'PY' | tee -a $OUT/session.log
import json
d=json.loads(open('gpurun_out/r2s10/n2_cfg3.json').read().strip().splitlines()[-1])
print('resident %.2f ms e2e %s %.1f pinned %.1f inproc %s verified %s' % (d['ms_per_step'], d['e2e']['host_memory'], d['e2e']['ms_per_step'], d['e2e_pinned']['ms_per_step'], d.get('e2e_inproc'), d['verified']))
PY
grep -E "Error|error" $OUT/n2_cfg3.err | tail -3
stage "leaf L2 hints A/B on GPU 0"
for H in 1 0; do
  echo "L2HINT=$H" | tee -a $OUT/session.log
  CUDA_VISIBLE_DEVICES=0 M4RI_B200_LEAF2_L2HINT=$H timeout 300 python tools/leaf_time.py 65536,65536,65536,4 16384,16384,16384,2 16384,16384,16384,-1 32768,131072,32768,3 2>&1 | tee -a $OUT/session.log
  CUDA_VISIBLE_DEVICES=0 M4RI_B200_LEAF2_L2HINT=$H timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:m4rm_leaf2 -s 1 -c 1 --csv \
     python tools/leaf_run.py 16384 16384 16384 2 4096 2>/dev/null | grep -E "m4rm_leaf2" | cut -d, -f5,12- | tee -a $OUT/session.log
done
stage "PLE timing repeat"
CUDA_VISIBLE_DEVICES=0 timeout 600 python tools/ple_time.py 16384 32768 32768 65536 2>&1 | tee -a $OUT/session.log
stage "done"
