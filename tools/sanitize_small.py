#!/usr/bin/env python
"""Tiny pass over every kernel for compute-sanitizer (memcheck / racecheck): small batches,
boards up to 32768 tiles (serial path), odd sizes (tail paths)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import rein48_b200 as r48
from rein48_b200.batched import spawn


def boards(n, seed, max_exp):
    rng = np.random.default_rng(seed)
    e = rng.integers(1, max_exp + 1, (n, 16)).astype(np.uint64)
    e[rng.random((n, 16)) < 0.25] = 0
    return torch.from_numpy((e << (np.arange(16, dtype=np.uint64) * np.uint64(4))).sum(1).astype(np.uint64).view(np.int64)).cuda()


def main():
    torch.cuda.set_device(0)
    n = 3001
    for mode in ("reference", "merge_sum"):
        env = r48.BatchedGame(n, seed=1, board_base=(1 << 64) - 100, reward_mode=mode, id_stride=n)
        env.boards.copy_(boards(n, 3, 15))
        for t in range(3):
            env.step(torch.randint(0, 4, (n,), device="cuda"))
        env.afterstates()
        env.env_step(torch.randint(0, 4, (n,), device="cuda"))
        env.step_injected(torch.randint(0, 4, (n,), device="cuda"), torch.randint(0, 16, (n,), device="cuda"),
                          torch.randint(1, 3, (n,), device="cuda"))
    b = boards(n, 5, 15)
    spawn(b, 7, 0, 3)
    r48.decode(b); r48.decode(b, log2=True); r48.encode(r48.decode(b, dtype=torch.int32)); r48.scores(b); r48.blank_counts(b)
    for pol in ("random", "greedy_blanks"):
        r48.random_rollouts(1500, seed=2, policy=pol)
        r48.rollout_trajectories(700, seed=2, policy=pol)
    r48.random_rollouts_host(900, seed=4)
    r48.random_rollouts_host(900, seed=4, records=True)
    ring = r48.ReplayRing(1000, seed=3)
    env = r48.BatchedGame(700, seed=9)
    for t in range(4):
        env.env_step(torch.randint(0, 4, (700,), device="cuda"), ring=ring)       # wraps the ring
    ring.store(b[:333], torch.randint(0, 4, (333,), device="cuda"), None, b[333:666])
    ring.sample(128); ring.sample(128, replace=True, obs=True); ring.sample(5000)
    ref = r48.ReplayRing(100, mode="reference")
    ref.store(b[:333], torch.randint(0, 4, (333,), device="cuda"), None, b[333:666])
    ref.sample()
    torch.cuda.synchronize()
    print("sanitize pass done")


if __name__ == "__main__":
    main()
