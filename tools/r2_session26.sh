#!/bin/bash
# round 2, session 26: pipelined leaf nodes (LeafPipe, M4RI_B200_NO_PIPE) — an experiment that was REMOVED again:
# this script documents how it was measured and no longer toggles anything
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_large_golden_gpu.py tests/test_zz5_tensor_leaf_gpu.py -x -q 2>&1 | tail -2
for np in 1 0; do
  if [ $np = 1 ]; then export M4RI_B200_NO_PIPE=1; else unset M4RI_B200_NO_PIPE; fi
  timeout 200 python tools/leaf_time.py 65536,65536,65536,8192 32768,32768,32768,4096 2>&1 | tail -2
done
timeout 600 python bench.py --steps 3 > gpurun_out/bench_pipe_cfg3.json 2> gpurun_out/bench_pipe.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_pipe_cfg3.json").read().strip().splitlines()[-1])
print("bench cfg3: %.2f ms %.3e | e2e %.1f ms pinned %.1f | verified %s | leaf share %.3f" % (d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["e2e_pinned"]["ms_per_step"], d["verified"], d["roofline"]["leaf_share_of_step"]))
PY
tail -2 gpurun_out/bench_pipe.err
