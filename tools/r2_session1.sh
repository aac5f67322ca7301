#!/bin/bash
# Round 2, GPU session 1 (1 GPU): whole GPU suite incl. the reference-digest parity at the BASELINE sizes,
# bench lines of all three workloads, the full-size reference CPU timing, tensor-core probes, leaf shape sweep.
set -u
OUT=gpurun_out/r2s1; mkdir -p $OUT
stage() { echo "=== $1 ($(date +%T))" | tee -a $OUT/session.log; }
stage "nproc / mem"; nproc | tee -a $OUT/session.log; free -g | head -2 | tee -a $OUT/session.log
stage "large golden parity first"
timeout 900 python -m pytest tests/test_large_golden_gpu.py -m gpu -x -q > $OUT/pytest_large_golden.log 2>&1
tail -5 $OUT/pytest_large_golden.log | tee -a $OUT/session.log
stage "bench cfg3"
timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/bench_cfg3_n1.json 2> $OUT/bench_cfg3_n1.err
tail -c 1500 $OUT/bench_cfg3_n1.json | tee -a $OUT/session.log; tail -3 $OUT/bench_cfg3_n1.err | tee -a $OUT/session.log
stage "bench cfg2"
timeout 300 python bench.py --workload cfg2 --steps 10 --warmup 3 > $OUT/bench_cfg2_n1.json 2> $OUT/bench_cfg2_n1.err
tail -c 600 $OUT/bench_cfg2_n1.json | tee -a $OUT/session.log; tail -3 $OUT/bench_cfg2_n1.err | tee -a $OUT/session.log
stage "bench cfg5"
timeout 600 python bench.py --workload cfg5 --steps 5 --warmup 3 > $OUT/bench_cfg5_n1.json 2> $OUT/bench_cfg5_n1.err
tail -c 600 $OUT/bench_cfg5_n1.json | tee -a $OUT/session.log; tail -3 $OUT/bench_cfg5_n1.err | tee -a $OUT/session.log
stage "tensor-core probes"
timeout 120 tools/tc/b1_probe > $OUT/tc_b1_probe.jsonl 2>&1; cat $OUT/tc_b1_probe.jsonl | tee -a $OUT/session.log
timeout 600 python tools/tc/lib_probe.py 8192 16384 > $OUT/tc_lib_probe.jsonl 2> $OUT/tc_lib_probe.err; cat $OUT/tc_lib_probe.jsonl | tee -a $OUT/session.log; tail -3 $OUT/tc_lib_probe.err
stage "leaf shape sweep (levels forced)"
timeout 300 python tools/leaf_time.py 16384,16384,16384,-1 16384,16384,16384,1 16384,16384,16384,2 8192,8192,8192,1 \
   16384,65536,32768,2 16384,65536,32768,3 16384,16384,32768,1 16384,16384,32768,2 16384,8192,32768,1 16384,8192,32768,2 \
   32768,65536,16384,2 32768,65536,16384,3 65536,65536,65536,3 65536,65536,65536,4 > $OUT/leaf_sweep.log 2>&1
cat $OUT/leaf_sweep.log | tee -a $OUT/session.log
stage "reference arm (driver form) and one full-size reference run"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref_32768.json 2> $OUT/bench_ref.err
tail -c 700 $OUT/bench_ref_32768.json | tee -a $OUT/session.log
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 --ref-sample 65536 > $OUT/bench_ref_65536.json 2>> $OUT/bench_ref.err
tail -c 700 $OUT/bench_ref_65536.json | tee -a $OUT/session.log
stage "rest of the GPU suite"
timeout 1200 python -m pytest tests -m gpu -x -q --deselect tests/test_large_golden_gpu.py > $OUT/pytest_gpu.log 2>&1
tail -5 $OUT/pytest_gpu.log | tee -a $OUT/session.log
stage "done"
