#!/bin/bash
# Round 2, GPU session 12 (1 GPU): hybrid (whole tiles round-robin + stream-K tail) vs pure stream-K partition of the
# tall-tile leaf — time and DRAM bytes; parity of the new partition.
set -u
OUT=gpurun_out/r2s12; mkdir -p $OUT
stage() { echo "=== $1 ($(date +%T))" | tee -a $OUT/session.log; }
for H in 1 0; do
  stage "M4RI_B200_LEAF2_HYBRID=$H"
  M4RI_B200_LEAF2_HYBRID=$H timeout 300 python tools/leaf_time.py 65536,65536,65536,4 65536,65536,65536,3 16384,16384,16384,2 16384,16384,16384,-1 32768,131072,32768,3 32768,65536,16384,3 8192,8192,8192,1 2>&1 | cut -c1-200 | tee -a $OUT/session.log
  M4RI_B200_LEAF2_HYBRID=$H timeout 120 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:m4rm_leaf2 -s 1 -c 1 --csv \
     python tools/leaf_run.py 16384 16384 16384 2 4096 2>/dev/null | grep -E "m4rm_leaf2" | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | tee -a $OUT/session.log
  M4RI_B200_LEAF2_HYBRID=$H timeout 120 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:m4rm_leaf2 -s 1 -c 1 --csv \
     python tools/leaf_run.py 16384 16384 16384 2 -1 2>/dev/null | grep -E "m4rm_leaf2" | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | tee -a $OUT/session.log
done
stage "parity with the hybrid partition"
timeout 900 python -m pytest tests/test_zz_leaf2_gpu.py tests/test_large_golden_gpu.py tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -3 | tee -a $OUT/session.log
stage "done"
