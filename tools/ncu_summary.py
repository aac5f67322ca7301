"""Summarise an .ncu-rep (read here, no GPU needed): the metrics DESIGN.md / bench.py quote, as JSON.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep [label] > profiles/rNN_x_summary.json
"""
import csv
import io
import json
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.max",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "smsp__cycles_active.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def main():
    rep = sys.argv[1]
    label = sys.argv[2] if len(sys.argv) > 2 else rep
    out = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True)
    rows = list(csv.reader(io.StringIO(out)))
    header, units = rows[0], rows[1]
    res = {}
    for k, row in enumerate(rows[2:]):
        d = dict(zip(header, row))
        entry = {"kernel": d.get("Kernel Name", "?")}
        for name in WANT:
            if name in d:
                entry[name] = {"value": d[name], "unit": units[header.index(name)]}
        res[f"{label} #{k}"] = entry
    json.dump(res, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
