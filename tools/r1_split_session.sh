#!/bin/bash
# A/B timing of the tall-tile leaf: tables built by all warps vs by alternating halves of the CTA.
OUT=gpurun_out; mkdir -p $OUT
SHAPES="16384,16384,16384,-1 16384,16384,16384,1 65536,65536,65536,8192"
for v in 0 1; do
  echo "--- M4RI_B200_LEAF2_SPLIT=$v"
  M4RI_B200_LEAF2_SPLIT=$v timeout 100 python tools/leaf_time.py $SHAPES 2>&1
done | tee $OUT/split_ab.log
M4RI_B200_LEAF=2 M4RI_B200_LEAF2_SPLIT=1 timeout 100 python -m pytest tests/test_zz_leaf2_gpu.py -x -q 2>&1 | tail -2 | tee -a $OUT/split_ab.log
