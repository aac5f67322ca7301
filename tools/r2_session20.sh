#!/bin/bash
# round 2, session 20: ncu of the pipelined tensor-core leaf (main kernel only, images reused)
mkdir -p gpurun_out
M4RI_B200_TC_REUSE=1 timeout 300 ncu --set full --import-source on --clock-control none -k regex:tc_leaf2 -c 1 -s 2 -o gpurun_out/r02_tc2b_ncu -f python tools/tc_leaf_check.py 16384,16384,16384 > gpurun_out/tc2b_ncu.log 2>&1
tail -2 gpurun_out/tc2b_ncu.log
