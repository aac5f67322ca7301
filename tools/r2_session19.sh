#!/bin/bash
# round 2, session 19: the pipelined tensor-core leaf — time without the pre-pass, ncu full set of the main kernel
mkdir -p gpurun_out
timeout 120 python tools/tc_leaf_check.py 16384,16384,16384 2>&1 | grep tc2 > gpurun_out/tc2_time.log
M4RI_B200_TC_REUSE=1 timeout 120 python tools/tc_leaf_check.py 16384,16384,16384 4096,16384,4096 2>&1 | grep tc2 >> gpurun_out/tc2_time.log
cat gpurun_out/tc2_time.log
M4RI_B200_TC_REUSE=1 timeout 300 ncu --set full --clock-control none -k regex:tc_leaf2 -c 1 -s 2 -o gpurun_out/r02_tc2_ncu -f python tools/tc_leaf_check.py 16384,16384,16384 > gpurun_out/tc2_ncu.log 2>&1
tail -3 gpurun_out/tc2_ncu.log
