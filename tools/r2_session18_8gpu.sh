#!/bin/bash
# Round 2, GPU session 18 (8 GPUs, final kernels): bench at N = 8, 4, 2 (config 4) and N = 8 config 5, each with the
# in-process mzd_mul_mp leg on rank 0 and reference-digest verification on every rank; in-process tests on 8 GPUs.
set -u
OUT=gpurun_out/r2s18; mkdir -p $OUT
stage() { echo "=== $1 ($(date +%T))" | tee -a $OUT/session.log; }
sumline() { python - "$1" <<'PY' | tee -a $OUT/session.log
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    e=d.get('e2e',{}); p=d.get('e2e_pinned',{}); i=d.get('e2e_inproc') or {}
    print('  value %.3e (%.2f ms) e2e %s %.1f ms pinned %.1f ms inproc %.1f ms (%s, digest %s) path %s %s verified %s leaf %.3e share %.3f' % (d['value'], d['ms_per_step'], e.get('host_memory'), e.get('ms_per_step',0), p.get('ms_per_step',0), i.get('ms_per_step',0), i.get('path'), i.get('reference_digest'), d['config']['path'], d['config']['local_product'], d.get('verified'), d['roofline']['leaf_bitops_per_s'], d['roofline']['leaf_share_of_step']))
except Exception as ex:
    print('  no line:', ex)
PY
}
tr() { echo "python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2"; }
nproc | tee -a $OUT/session.log
stage "N=8 cfg3"
timeout 400 $(tr 8 29621) bench.py --gpus 8 --steps 5 --warmup 3 > $OUT/n8_cfg3.json 2> $OUT/n8_cfg3.err; sumline $OUT/n8_cfg3.json; grep -E "Error|error|MISMATCH" $OUT/n8_cfg3.err | tail -4
stage "N=8 cfg5"
timeout 400 $(tr 8 29622) bench.py --gpus 8 --workload cfg5 --steps 5 --warmup 3 > $OUT/n8_cfg5.json 2> $OUT/n8_cfg5.err; sumline $OUT/n8_cfg5.json; grep -E "Error|error|MISMATCH" $OUT/n8_cfg5.err | tail -4
stage "N=4 cfg3"
timeout 400 $(tr 4 29623) bench.py --gpus 4 --steps 5 --warmup 3 > $OUT/n4_cfg3.json 2> $OUT/n4_cfg3.err; sumline $OUT/n4_cfg3.json; grep -E "Error|error|MISMATCH" $OUT/n4_cfg3.err | tail -4
stage "N=2 cfg3"
timeout 400 $(tr 2 29624) bench.py --gpus 2 --steps 5 --warmup 3 > $OUT/n2_cfg3.json 2> $OUT/n2_cfg3.err; sumline $OUT/n2_cfg3.json; grep -E "Error|error|MISMATCH" $OUT/n2_cfg3.err | tail -4
stage "in-process mzd_mul_mp over 8 GPUs: tests"
timeout 400 python -m pytest tests/test_multigpu.py -m gpu -x -q > $OUT/pytest_multigpu_g8.log 2>&1; tail -3 $OUT/pytest_multigpu_g8.log | tee -a $OUT/session.log
stage "done"
