#!/bin/bash
# round 2, session 23: tensor-core leaf tests, compute-sanitizer on a small tensor-leaf product, ncu full set of one 49 x 8192^3 launch
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_zz5_tensor_leaf_gpu.py -x -q 2>&1 | tail -5 | tee gpurun_out/tc_tests.log
timeout 300 compute-sanitizer --tool memcheck python tools/tc_leaf_check.py 256,1024,256 512,2048,512 > gpurun_out/tc_sanitizer_memcheck.log 2>&1; tail -4 gpurun_out/tc_sanitizer_memcheck.log
timeout 300 compute-sanitizer --tool racecheck python tools/tc_leaf_check.py 256,1024,256 > gpurun_out/tc_sanitizer_racecheck.log 2>&1; tail -3 gpurun_out/tc_sanitizer_racecheck.log
timeout 400 ncu --set full --import-source on --clock-control none -k regex:tc_leaf2 -c 1 -s 1 -o gpurun_out/r02_tc_leaf_49x8192_ncu -f python tools/leaf_time.py 32768,32768,32768,8192 > gpurun_out/tc_ncu49.log 2>&1; tail -2 gpurun_out/tc_ncu49.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_tc_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/bench_under_ncu.log 2>&1; tail -1 gpurun_out/r02_tc_launches.csv | cut -c1-200
