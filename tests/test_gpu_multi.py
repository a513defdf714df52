# -*- coding: utf-8 -*-
"""N > 1 on real GPUs: one process per GPU, contiguous global episode ranges, ONE NCCL all-reduce
of the statistics vector -- the reduced vector must equal the single-GPU run bit for bit.
Skipped on boxes with fewer than 2 GPUs (the CPU suite covers the same logic over gloo)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import torch, torch.distributed as dist
import rein48_b200 as r48
rank, world = int(sys.argv[1]), int(sys.argv[2])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", init_method="tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                        device_id=torch.device("cuda", rank))
n, seed = 1_000_003, 4242
res = r48.sharded_rollouts(n, seed=seed)                    # this rank's slice + all-reduce(SUM)
whole = r48.random_rollouts(n, seed=seed, device="cuda:%d" % rank)   # every episode on this GPU
assert bool((res.stats == whole.stats).all()), "sharded + all-reduced statistics differ from the single-GPU run"
lo, hi = r48.shard_range(n, rank, world)
assert bool((res.final_boards == whole.final_boards[lo:hi]).all())
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok", r48.EpisodeStats(res.stats).episodes)
"""


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_rollouts_over_nccl_are_bit_identical(tmp_path):
    world = min(torch.cuda.device_count(), 8)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r), str(world)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(world)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
