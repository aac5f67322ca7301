#!/bin/bash
# Round 2, GPU session 3 (2 GPUs): two-level Winograd node (49 products per leaf launch) — parity + depth sweep on GPU 0;
# in-process mzd_mul_mp (grid + pipeline, peer copies) tests and timing on 2 GPUs; bench N=2 with forced chunk depth.
set -u
OUT=gpurun_out/r2s3; mkdir -p $OUT
stage() { echo "=== $1 ($(date +%T))" | tee -a $OUT/session.log; }
stage "parity: large golden + parity suite"
CUDA_VISIBLE_DEVICES=0 timeout 900 python -m pytest tests/test_large_golden_gpu.py tests/test_parity_gpu.py -m gpu -x -q > $OUT/pytest_parity.log 2>&1
tail -5 $OUT/pytest_parity.log | tee -a $OUT/session.log
stage "depth sweep with the two-level node"
CUDA_VISIBLE_DEVICES=0 timeout 300 python tools/leaf_time.py 65536,65536,65536,3 65536,65536,65536,4 32768,32768,32768,2 32768,32768,32768,3 \
   16384,16384,16384,1 16384,16384,16384,2 16384,65536,32768,2 16384,16384,32768,1 16384,16384,32768,2 32768,32768,65536,2 32768,32768,65536,3 \
   32768,131072,32768,2 32768,131072,32768,3 > $OUT/depth_sweep.log 2>&1
cat $OUT/depth_sweep.log | tee -a $OUT/session.log
stage "same with M4RI_B200_NO_NODE2=1 (one-level nodes only)"
CUDA_VISIBLE_DEVICES=0 M4RI_B200_NO_NODE2=1 timeout 300 python tools/leaf_time.py 65536,65536,65536,3 65536,65536,65536,4 > $OUT/depth_sweep_nonode2.log 2>&1
cat $OUT/depth_sweep_nonode2.log | tee -a $OUT/session.log
stage "in-process mzd_mul_mp on 2 GPUs: tests"
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -x -q > $OUT/pytest_multigpu_g2.log 2>&1
tail -8 $OUT/pytest_multigpu_g2.log | tee -a $OUT/session.log
stage "in-process mzd_mul_mp timing (pageable host matrices)"
timeout 300 python tools/mp_time.py 65536 2 > $OUT/mp_time_g2.log 2>&1; cat $OUT/mp_time_g2.log | tee -a $OUT/session.log
M4RI_B200_MP_KSUB=2 timeout 300 python tools/mp_time.py 65536 2 > $OUT/mp_time_g2_k2.log 2>&1; cat $OUT/mp_time_g2_k2.log | tee -a $OUT/session.log
stage "bench N=2 (pipeline, chunk levels 3)"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29556"
timeout 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 --chunk-levels 3 > $OUT/n2_cfg3_cl3.json 2> $OUT/n2_cfg3_cl3.err
python - <<'PY' | tee -a $OUT/session.log
import json
d=json.loads(open('gpurun_out/r2s3/n2_cfg3_cl3.json').read().strip().splitlines()[-1])
print('resident %.2f ms e2e pageable %.1f pinned %.1f' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e_pinned']['ms_per_step']), d['verified'])
PY
stage "done"
