#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_last.log 2>&1; tail -3 $OUT/pytest_gpu_last.log
timeout 100 python bench.py > $OUT/bench_last_n1.json 2> $OUT/bench_last_n1.err; python -c "
import json; d=json.load(open('$OUT/bench_last_n1.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic'], d['roofline']['leaf_avg_ms'])"
