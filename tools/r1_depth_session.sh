#!/bin/bash
# Strassen depth sweep with the tall-tile leaf (automatic leaf selection is the default now).
OUT=gpurun_out; mkdir -p $OUT
timeout 200 python tools/leaf_time.py 8192,8192,8192,1 8192,8192,8192,-1 16384,16384,16384,2 16384,16384,16384,1 65536,65536,65536,3 65536,65536,65536,4 32768,131072,32768,3 32768,131072,32768,2 > $OUT/depth_sweep.log 2>&1
cat $OUT/depth_sweep.log
