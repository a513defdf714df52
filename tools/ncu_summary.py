#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into a small JSON / text record.

    python tools/ncu_summary.py gpurun_out/prof_rollout.ncu-rep|prof_rollout_raw.csv [--json out.json]
"""
import csv
import io
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed_op_shared_ld.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    rep = sys.argv[1]
    if rep.endswith(".csv"):                 # an `ncu --page raw --csv` export made on the GPU box
        out = open(rep).read()
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {}
        for i, h in enumerate(hdr):
            if h == "Kernel Name" or h in KEEP:
                d[h + (" [%s]" % units[i] if units[i] else "")] = r[i]
            elif h.startswith(STALL) and h.endswith("_per_issue_active.ratio"):
                v = float(r[i] or 0)
                if v >= 0.05:
                    d.setdefault("stalls_per_issue", {})[h[len(STALL):-len("_per_issue_active.ratio")]] = round(v, 3)
        res.append(d)
    text = json.dumps(res, indent=1)
    if "--json" in sys.argv:
        open(sys.argv[sys.argv.index("--json") + 1], "w").write(text + "\n")
    print(text)


if __name__ == "__main__":
    main()
