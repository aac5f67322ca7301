#!/bin/bash
# Short GPU session: parity of leaf variant 2, timings against variant 1, one ncu --set full capture.
set -u
OUT=gpurun_out
mkdir -p $OUT
TAG=${1:-s2}
stage() { echo "=== $1 ($(date +%T))" | tee -a $OUT/session_$TAG.log; }
stage "parity of leaf variant 2"
timeout 300 python -m pytest tests/test_zz_leaf2_gpu.py -x -q > $OUT/leaf2_parity_$TAG.log 2>&1
tail -3 $OUT/leaf2_parity_$TAG.log | tee -a $OUT/session_$TAG.log
stage "leaf timings"
SHAPES="8192,8192,8192,-1 16384,16384,16384,-1 4096,8192,8192,-1 16384,16384,16384,1 65536,65536,65536,8192"
M4RI_B200_LEAF=1 timeout 150 python tools/leaf_time.py $SHAPES > $OUT/time_leaf1_$TAG.log 2>&1
M4RI_B200_LEAF=0 timeout 150 python tools/leaf_time.py $SHAPES > $OUT/time_leaf2_$TAG.log 2>&1
M4RI_B200_LEAF=0 M4RI_B200_LEAF2_AWIDE=1 timeout 150 python tools/leaf_time.py $SHAPES > $OUT/time_leaf2_awide_$TAG.log 2>&1
for f in time_leaf1 time_leaf2 time_leaf2_awide; do echo "--- $f"; cat $OUT/${f}_$TAG.log; done | tee -a $OUT/session_$TAG.log
stage "ncu --set full: one 16384^3 launch of variant 2"
M4RI_B200_LEAF=2 timeout 240 ncu --set full --clock-control none --import-source on -k regex:m4rm_leaf2 -c 1 -f \
  -o $OUT/leaf2_16384_$TAG python tools/leaf_run.py 16384 16384 16384 1 > $OUT/ncu_leaf2_16384_$TAG.log 2>&1
ls -la $OUT/*_$TAG.ncu-rep 2>&1 | tee -a $OUT/session_$TAG.log
stage "done"
