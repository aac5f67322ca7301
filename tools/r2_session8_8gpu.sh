#!/bin/bash
# Round 2, GPU session 8 (8 GPUs, short): bench at N = 8 — config 4 (65536^3) and config 5 — with the 8 x 8 block digests.
set -u
OUT=gpurun_out/r2s8; mkdir -p $OUT
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
stage "N=8 cfg3"
timeout 400 $(tr 8 29611) bench.py --gpus 8 --steps 5 --warmup 3 > $OUT/n8_cfg3.json 2> $OUT/n8_cfg3.err; sumline $OUT/n8_cfg3.json; grep -E "Error|error|MISMATCH" $OUT/n8_cfg3.err | tail -4
stage "N=8 cfg5"
timeout 400 $(tr 8 29612) bench.py --gpus 8 --workload cfg5 --steps 5 --warmup 3 > $OUT/n8_cfg5.json 2> $OUT/n8_cfg5.err; sumline $OUT/n8_cfg5.json; grep -E "Error|error|MISMATCH" $OUT/n8_cfg5.err | tail -4
stage "done"
