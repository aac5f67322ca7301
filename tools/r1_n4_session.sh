#!/bin/bash
# 4-GPU session: the 2 x 2 grid path of bench.py at full size with the device-side check.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 4 --steps 5 --warmup 3 --check > $OUT/bench_n4_grid.json 2> $OUT/bench_n4_grid.err
echo "rc=$?" | tee -a $OUT/bench_n4_grid.err
grep -a "check\|Error\|error" $OUT/bench_n4_grid.err | tail -20
tail -c 2500 $OUT/bench_n4_grid.json
