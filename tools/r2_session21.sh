#!/bin/bash
# round 2, session 21: the tensor-core leaf (M4RI_B200_LEAF=3) under the Strassen scheduler: reference digests + timing
mkdir -p gpurun_out
M4RI_B200_LEAF=3 timeout 600 python -m pytest tests/test_large_golden_gpu.py -x -q 2>&1 | tail -3 | tee gpurun_out/tc_golden.log
for v in 0 3; do
  M4RI_B200_LEAF=$v timeout 200 python tools/leaf_time.py 16384,16384,16384,-1 32768,32768,32768,0 65536,65536,65536,0 2>&1 | tail -3 | tee -a gpurun_out/tc_strassen.log
done
