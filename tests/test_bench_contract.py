"""CPU test of bench.py's contract for the reference arm (no GPU needed): one JSON line with the keys the
driver reads, produced by the reference's own CPU implementation (oracle/_ref) on a bounded sample."""
import json
import os
import subprocess
import sys

from tests import harness as H

REQUIRED = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
            "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"}


def test_reference_arm_prints_one_valid_json_line():
    env = dict(os.environ, OMP_NUM_THREADS="1")   # as under torch.distributed.run
    p = subprocess.run([sys.executable, os.path.join(H.ROOT, "bench.py"), "--impl", "reference", "--gpus", "1",
                        "--steps", "3", "--warmup", "1", "--ref-sample", "8192"], capture_output=True, text=True, timeout=900, env=env)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert REQUIRED <= set(d), REQUIRED - set(d)
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["unit"] == "bit-ops/s" and d["value"] > 1e11
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert "workload" in d["config"]
    # torchrun exports OMP_NUM_THREADS=1; the arm must still use every host thread it may run on
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0)) or d["cpu_baseline"]["kind"] != "reference"


def test_non_zero_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(H.ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=120, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""
