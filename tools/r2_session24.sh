#!/bin/bash
# round 2, session 24: padding to tensor-leaf units for non-power-of-two sizes; whole GPU suite; final 1-GPU bench lines
mkdir -p gpurun_out
timeout 200 python tools/leaf_time.py 50048,50048,50048,8192 40064,60032,30080,8192 2>&1 | tail -2 | tee gpurun_out/tc_odd_sizes.log
M4RI_B200_LEAF=2 timeout 200 python tools/leaf_time.py 50048,50048,50048,4096 2>&1 | tail -1 | tee -a gpurun_out/tc_odd_sizes.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_suite_tc2.log 2>&1; tail -3 gpurun_out/gpu_suite_tc2.log
timeout 900 python bench.py > gpurun_out/bench_final_cfg3.json 2> gpurun_out/bench_final.err; tail -c 400 gpurun_out/bench_final_cfg3.json
timeout 600 python bench.py --workload cfg2 > gpurun_out/bench_final_cfg2.json 2>> gpurun_out/bench_final.err
timeout 600 python bench.py --workload cfg5 > gpurun_out/bench_final_cfg5.json 2>> gpurun_out/bench_final.err
