#!/bin/bash
# Round 2, GPU session 5 (8 GPUs, kept short — charged 8x): bench at N = 8 and N = 4 (2 x N/2 grid, hooks-mode end to end,
# reference-digest verification on every rank), config 5 at N = 8, in-process mzd_mul_mp tests + timing over 8 GPUs.
set -u
OUT=gpurun_out/r2s5; mkdir -p $OUT
stage() { echo "=== $1 ($(date +%T))" | tee -a $OUT/session.log; }
sumline() { python - "$1" <<'PY' | tee -a $OUT/session.log
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    e=d.get('e2e',{}); p=d.get('e2e_pinned',{})
    print('  value %.3e (%.2f ms) e2e %s %.1f ms pinned %.1f ms path %s %s verified %s leaf %.3e share %.3f' % (d['value'], d['ms_per_step'], e.get('host_memory'), e.get('ms_per_step',0), p.get('ms_per_step',0), d['config']['path'], d['config']['local_product'], d.get('verified'), d['roofline']['leaf_bitops_per_s'], d['roofline']['leaf_share_of_step']))
except Exception as ex:
    print('  no line:', ex)
PY
}
tr() { echo "python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2"; }
nproc | tee -a $OUT/session.log
stage "N=8 cfg3"
timeout 400 $(tr 8 29601) bench.py --gpus 8 --steps 5 --warmup 3 > $OUT/n8_cfg3.json 2> $OUT/n8_cfg3.err; sumline $OUT/n8_cfg3.json; grep -E "Error|error|MISMATCH" $OUT/n8_cfg3.err | tail -4
stage "N=8 cfg5"
timeout 400 $(tr 8 29602) bench.py --gpus 8 --workload cfg5 --steps 5 --warmup 3 > $OUT/n8_cfg5.json 2> $OUT/n8_cfg5.err; sumline $OUT/n8_cfg5.json; grep -E "Error|error|MISMATCH" $OUT/n8_cfg5.err | tail -4
stage "N=4 cfg3"
timeout 400 $(tr 4 29603) bench.py --gpus 4 --steps 5 --warmup 3 > $OUT/n4_cfg3.json 2> $OUT/n4_cfg3.err; sumline $OUT/n4_cfg3.json; grep -E "Error|error|MISMATCH" $OUT/n4_cfg3.err | tail -4
stage "N=8 cfg3, 4 x 2 grid of round 1 for comparison (resident only)"
timeout 400 $(tr 8 29604) bench.py --gpus 8 --steps 5 --warmup 3 --no-e2e --no-check --grid rows > $OUT/n8_cfg3_rows.json 2> /dev/null; sumline $OUT/n8_cfg3_rows.json
stage "in-process mzd_mul_mp over 8 GPUs: tests + timing"
timeout 400 python -m pytest tests/test_multigpu.py -m gpu -x -q > $OUT/pytest_multigpu_g8.log 2>&1; tail -3 $OUT/pytest_multigpu_g8.log | tee -a $OUT/session.log
timeout 300 python tools/mp_time.py 65536 8 2>&1 | tee -a $OUT/session.log
stage "reference arm under torchrun (OMP threads check)"
timeout 300 $(tr 8 29605) bench.py --impl reference --gpus 8 --steps 1 --warmup 0 --ref-sample 16384 2>/dev/null | cut -c1-200 | tee -a $OUT/session.log
stage "done"
