#!/bin/bash
# Round 2, GPU session 2 (2 GPUs): K-chunk pipelined end-to-end path at N = 2 (oracle check at small n, reference digests at
# full size), serial form beside it; fp4 library probe on GPU 0.
set -u
OUT=gpurun_out/r2s2; mkdir -p $OUT
stage() { echo "=== $1 ($(date +%T))" | tee -a $OUT/session.log; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555"
stage "N=2 small, oracle verify (pipeline, ksub 1 and 2; cfg5-like accumulate via --workload cfg5 is full size only)"
timeout 300 $TR bench.py --gpus 2 --size 8192 --steps 2 --warmup 1 --verify --no-cpu-baseline > $OUT/n2_small_k1.json 2> $OUT/n2_small_k1.err
tail -c 300 $OUT/n2_small_k1.json | tee -a $OUT/session.log; grep -E "verify|check|Error|error" $OUT/n2_small_k1.err | tail -6 | tee -a $OUT/session.log
timeout 300 $TR bench.py --gpus 2 --size 8192 --steps 2 --warmup 1 --verify --ksub 2 --no-cpu-baseline > $OUT/n2_small_k2.json 2> $OUT/n2_small_k2.err
tail -c 300 $OUT/n2_small_k2.json | tee -a $OUT/session.log; grep -E "verify|check|Error|error" $OUT/n2_small_k2.err | tail -6 | tee -a $OUT/session.log
stage "N=2 full size cfg3 (pipeline)"
timeout 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 > $OUT/n2_cfg3.json 2> $OUT/n2_cfg3.err
tail -c 1200 $OUT/n2_cfg3.json | tee -a $OUT/session.log; grep -E "check|Error|error" $OUT/n2_cfg3.err | tail -6 | tee -a $OUT/session.log
stage "N=2 full size cfg3 (pipeline, ksub 2)"
timeout 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 --ksub 2 > $OUT/n2_cfg3_k2.json 2> $OUT/n2_cfg3_k2.err
tail -c 1200 $OUT/n2_cfg3_k2.json | tee -a $OUT/session.log; grep -E "check|Error|error" $OUT/n2_cfg3_k2.err | tail -6 | tee -a $OUT/session.log
stage "N=2 full size cfg3 (serial e2e, round-1 form)"
timeout 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 --serial-e2e > $OUT/n2_cfg3_serial.json 2> $OUT/n2_cfg3_serial.err
tail -c 1200 $OUT/n2_cfg3_serial.json | tee -a $OUT/session.log
stage "N=2 cfg5 (pipeline)"
timeout 600 $TR bench.py --gpus 2 --workload cfg5 --steps 5 --warmup 3 > $OUT/n2_cfg5.json 2> $OUT/n2_cfg5.err
tail -c 1200 $OUT/n2_cfg5.json | tee -a $OUT/session.log; grep -E "check|Error|error" $OUT/n2_cfg5.err | tail -6 | tee -a $OUT/session.log
stage "fp4 probe"
timeout 300 python tools/tc/lib_probe.py 8192 16384 > $OUT/tc_lib_probe2.jsonl 2> $OUT/tc_lib_probe2.err; grep -i "fp4\|mxfp4" $OUT/tc_lib_probe2.jsonl | cut -c1-420 | tee -a $OUT/session.log
stage "done"
