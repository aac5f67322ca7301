#!/bin/bash
# Round 2, GPU session 14 (1 GPU): the whole GPU suite with the full log kept.
set -u
OUT=gpurun_out/r2s14; mkdir -p $OUT
timeout 1800 python -m pytest tests -m gpu -x -q -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1
echo "rc=$?" >> $OUT/pytest_gpu.log
grep -n "Fatal\|Abort\|abort\|m4ri_b200:\|Segmentation\|FAILED\|passed\|failed\|Current thread" $OUT/pytest_gpu.log | head -20
grep -n "File \"/tmp/code\|File \"/root/repo" $OUT/pytest_gpu.log | head -12
tail -5 $OUT/pytest_gpu.log
