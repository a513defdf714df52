#!/usr/bin/env python
"""Quick CUDA-graph timing of the step / afterstates / env_step kernels (for A/B experiments).

    python tools/time_kernels.py [afterstates] [step] [env]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rein48_b200 as r48
import bench


def main():
    which = sys.argv[1:] or ["afterstates", "step", "env"]
    torch.cuda.set_device(0)
    r48._native.check(r48._native.lib().r48_init(0))
    peak, _ = bench.measured_peaks()
    if "afterstates" in which:
        r = bench.bench_afterstates_kernel(torch, r48, peak)
        print("afterstates 8M: %.2f us  %.1f%% of HBM peak" % (r["us_per_launch"], 100 * r["roofline"]["frac"]))
    if "step" in which:
        r = bench.bench_step_kernel(torch, r48, peak)
        print("step 1M: %.2f us (%.1f%%)   8M: %.2f us (%.1f%%)" % (
            r["us_per_launch"], 100 * r["roofline"]["frac"], r["at_8M_boards"]["us_per_launch"],
            100 * r["at_8M_boards"]["frac"]))
    if "env" in which:
        r = bench.bench_env_step_kernel(torch, r48, peak)
        print("env_step 1M: %.2f us  %.1f%%" % (r["us_per_launch"], 100 * r["roofline"]["frac"]))


if __name__ == "__main__":
    main()
