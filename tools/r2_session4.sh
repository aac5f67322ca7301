#!/bin/bash
# Round 2, GPU session 4 (2 GPUs): new default depth (cutoff 4096: 4096-row leaves, 49 per launch) on one GPU incl. parity;
# hooks-mode end-to-end path at N = 2 (oracle check small, reference digests at full size); staging-thread sweep.
set -u
OUT=gpurun_out/r2s4; mkdir -p $OUT
stage() { echo "=== $1 ($(date +%T))" | tee -a $OUT/session.log; }
sumline() { python - "$1" <<'PY' | tee -a $OUT/session.log
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    e=d.get('e2e',{}); p=d.get('e2e_pinned',{})
    print('  value %.3e (%.2f ms) e2e %s %.1f ms pinned %.1f ms path %s verified %s leaf %.3e share %.3f launches %d' % (d['value'], d['ms_per_step'], e.get('host_memory'), e.get('ms_per_step',0), p.get('ms_per_step',0), d['config']['path'], d.get('verified'), d['roofline']['leaf_bitops_per_s'], d['roofline']['leaf_share_of_step'], d['gpu_launches']))
except Exception as ex:
    print('  no line:', ex)
PY
}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29557"
stage "N=1 bench, default depth"
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/n1_cfg3.json 2> $OUT/n1_cfg3.err; sumline $OUT/n1_cfg3.json; tail -2 $OUT/n1_cfg3.err
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --workload cfg5 --steps 5 --warmup 3 --no-cpu-baseline > $OUT/n1_cfg5.json 2> $OUT/n1_cfg5.err; sumline $OUT/n1_cfg5.json
stage "N=1 pageable e2e vs staging threads"
for T in 4 8 12; do
  M4RI_B200_STAGE_THREADS=$T CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --steps 5 --warmup 2 --pageable --no-cpu-baseline --no-check > $OUT/n1_stage_t$T.json 2>/dev/null
  echo "threads $T" | tee -a $OUT/session.log; sumline $OUT/n1_stage_t$T.json
done
stage "N=2 hooks small (oracle verify)"
timeout 300 $TR bench.py --gpus 2 --size 8192 --steps 2 --warmup 1 --verify --no-cpu-baseline > $OUT/n2_small_hooks.json 2> $OUT/n2_small_hooks.err
sumline $OUT/n2_small_hooks.json; grep -E "verify|Error|error" $OUT/n2_small_hooks.err | tail -6 | tee -a $OUT/session.log
stage "N=2 hooks cfg3 / cfg5"
timeout 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 > $OUT/n2_cfg3_hooks.json 2> $OUT/n2_cfg3_hooks.err; sumline $OUT/n2_cfg3_hooks.json; grep -E "Error|error" $OUT/n2_cfg3_hooks.err | tail -4
timeout 600 $TR bench.py --gpus 2 --workload cfg5 --steps 5 --warmup 3 > $OUT/n2_cfg5_hooks.json 2> $OUT/n2_cfg5_hooks.err; sumline $OUT/n2_cfg5_hooks.json; grep -E "Error|error" $OUT/n2_cfg5_hooks.err | tail -4
stage "N=2 kchunk / serial for comparison"
timeout 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 --e2e-mode kchunk --ksub 2 --no-check > $OUT/n2_cfg3_kchunk.json 2> /dev/null; sumline $OUT/n2_cfg3_kchunk.json
timeout 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 --e2e-mode serial --no-check > $OUT/n2_cfg3_serial.json 2> /dev/null; sumline $OUT/n2_cfg3_serial.json
stage "in-process mzd_mul_mp, 2 GPUs"
timeout 300 python -m pytest tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -2 | tee -a $OUT/session.log
timeout 300 python tools/mp_time.py 65536 2 2>&1 | tee -a $OUT/session.log
M4RI_B200_MP_KSUB=4 timeout 300 python tools/mp_time.py 65536 2 2>&1 | tee -a $OUT/session.log
stage "GPU suite on one GPU (new default depth)"
CUDA_VISIBLE_DEVICES=0 timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; tail -3 $OUT/pytest_gpu.log | tee -a $OUT/session.log
stage "done"
