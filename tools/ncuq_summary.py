#!/usr/bin/env python
"""Summarise a tools/ncu_quick.sh CSV: one line per launch."""
import collections
import csv
import sys


def main(path, limit=12):
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if r and r[0] == 'ID':
            hdr, start = r, i + 1
            break
    idx = {h: i for i, h in enumerate(hdr)}
    data = collections.OrderedDict()
    for r in rows[start:]:
        if len(r) < len(hdr):
            continue
        key = (int(r[idx['ID']]), r[idx['Kernel Name']][:34], r[idx['Grid Size']])
        data.setdefault(key, {})[r[idx['Metric Name']]] = float(r[idx['Metric Value']].replace(',', ''))
    seen = collections.Counter()
    for k, v in data.items():
        seen[k[1]] += 1
        if seen[k[1]] > limit:
            continue
        g = lambda m: v.get(m, float('nan'))
        print("%3d %-34s %-13s t=%8.1fus inst=%.4e ipc=%.2f alu=%2.0f%% fma=%2.0f%% issue=%2.0f%% lanes=%.1f smem_wf=%.2e dram=%.1f+%.1fMB" % (
            k[0], k[1], k[2], g('gpu__time_duration.sum') / 1e3, g('smsp__inst_executed.sum'),
            g('sm__inst_executed.avg.per_cycle_elapsed'),
            g('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active'),
            g('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active'),
            g('smsp__issue_active.avg.pct_of_peak_sustained_active'),
            g('smsp__thread_inst_executed_per_inst_executed.ratio'),
            g('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum'), g('dram__bytes_read.sum') / 1e6, g('dram__bytes_write.sum') / 1e6))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 12)
