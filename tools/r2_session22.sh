#!/bin/bash
# round 2, session 22: tensor-core leaf as the default — bench (all three workloads) and the whole GPU suite
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_tc_cfg3.json 2> gpurun_out/bench_tc_cfg3.err; tail -c 3000 gpurun_out/bench_tc_cfg3.json
timeout 600 python bench.py --workload cfg2 > gpurun_out/bench_tc_cfg2.json 2>> gpurun_out/bench_tc_cfg3.err
timeout 600 python bench.py --workload cfg5 > gpurun_out/bench_tc_cfg5.json 2>> gpurun_out/bench_tc_cfg3.err
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_suite_tc.log 2>&1; tail -5 gpurun_out/gpu_suite_tc.log
