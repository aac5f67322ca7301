#!/bin/bash
# round 2, session 27 (final state): whole GPU suite, smoke, the three bench workloads, ncu launch list of the bench
# command, and the widened rows (PLE, RREF, TRSM) now that their large updates run on the tensor-core leaf
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_suite_final.log 2>&1; tail -3 gpurun_out/gpu_suite_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_final2_cfg3.json 2> gpurun_out/bench_final2.err
timeout 600 python bench.py --workload cfg2 > gpurun_out/bench_final2_cfg2.json 2>> gpurun_out/bench_final2.err
timeout 600 python bench.py --workload cfg5 > gpurun_out/bench_final2_cfg5.json 2>> gpurun_out/bench_final2.err
python - <<'PY'
import json
for w in ("cfg3","cfg2","cfg5"):
    d=json.loads(open(f"gpurun_out/bench_final2_{w}.json").read().strip().splitlines()[-1])
    print(w, "%.2f ms %.3e | e2e %.1f pinned %.1f | %s | leaf %.3e frac %.3f" % (d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["e2e_pinned"]["ms_per_step"], d["verified"]["reference_digest"], d["roofline"]["leaf_bitops_per_s"], d["roofline"]["frac"]))
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_final_launches.csv python bench.py --steps 2 --warmup 1 --no-check > gpurun_out/bench_under_ncu2.log 2>&1
timeout 300 python tools/ple_time.py 16384 32768 65536 2>&1 | tail -3 | tee gpurun_out/widened_final.log
timeout 300 python tools/echelon_time.py 16384 32768 2>&1 | tail -2 | tee -a gpurun_out/widened_final.log
