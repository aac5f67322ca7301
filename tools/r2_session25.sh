#!/bin/bash
# round 2, session 25: ncu full set (with source counters) of the final tensor-core leaf, 49 x 8192^3 in one launch
mkdir -p gpurun_out
timeout 400 ncu --set full --import-source on --clock-control none -k regex:tc_leaf2 -c 1 -s 1 -o gpurun_out/r02_tc_leaf_final_ncu -f python tools/leaf_time.py 32768,32768,32768,8192 > gpurun_out/tc_ncu_final.log 2>&1; tail -2 gpurun_out/tc_ncu_final.log
