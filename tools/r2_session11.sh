#!/bin/bash
# Round 2, GPU session 11 (2 GPUs): which L2 hint form is legal on sm_100a (bisect), in-process mzd_mul_mp timing, bench N=2
# with the in-process leg, PLE timing repeat.
set -u
OUT=gpurun_out/r2s11; mkdir -p $OUT
stage() { echo "=== $1 ($(date +%T))" | tee -a $OUT/session.log; }
stage "leaf L2 hint forms on GPU 0"
for H in 0 4 1 2 6; do
  echo "L2HINT=$H" | tee -a $OUT/session.log
  CUDA_VISIBLE_DEVICES=0 M4RI_B200_LEAF2_L2HINT=$H timeout 120 python tools/leaf_time.py 16384,16384,16384,2 65536,65536,65536,4 2>&1 | tail -2 | cut -c1-200 | tee -a $OUT/session.log
  CUDA_VISIBLE_DEVICES=0 M4RI_B200_LEAF2_L2HINT=$H timeout 120 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:m4rm_leaf2 -s 1 -c 1 --csv \
     python tools/leaf_run.py 16384 16384 16384 2 4096 2>/dev/null | grep -E "m4rm_leaf2" | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | tee -a $OUT/session.log
done
stage "in-process mzd_mul_mp (quadrant hooks), 2 GPUs"
timeout 300 python tools/mp_time.py 65536 2 2>&1 | tee -a $OUT/session.log
stage "bench N=2 with the in-process leg"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562"
timeout 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 > $OUT/n2_cfg3.json 2> $OUT/n2_cfg3.err
python - <<'PY' | tee -a $OUT/session.log
import json
d=json.loads(open('gpurun_out/r2s11/n2_cfg3.json').read().strip().splitlines()[-1])
print('resident %.2f ms e2e %s %.1f pinned %.1f inproc %s verified %s' % (d['ms_per_step'], d['e2e']['host_memory'], d['e2e']['ms_per_step'], d['e2e_pinned']['ms_per_step'], d.get('e2e_inproc'), d['verified']))
PY
grep -E "Error|error" $OUT/n2_cfg3.err | tail -3
stage "PLE timing repeat"
CUDA_VISIBLE_DEVICES=0 timeout 600 python tools/ple_time.py 16384 32768 32768 65536 2>&1 | tee -a $OUT/session.log
stage "done"
