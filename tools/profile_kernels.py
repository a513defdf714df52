#!/usr/bin/env python
"""Small driver for ncu captures: launches each hot kernel a few times at a fixed size.

    python tools/profile_kernels.py [rollout|step|afterstates|env|greedy|traj|all] [--rollout-log2 22]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rein48_b200 as r48


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("which", nargs="?", default="all")
    ap.add_argument("--rollout-log2", type=int, default=22)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    torch.cuda.set_device(0)
    L = r48._native.lib()
    stream = torch.cuda.current_stream().cuda_stream
    if a.which == "quick":
        # 2 rollout launches at 2^22 and 2^24, 4 step launches at 2^20 (rotating windows), 2 at 2^23
        for lg in (22, 24, 26):
            buf = r48.RolloutBuffers(1 << lg)
            for i in range(2):
                r48.random_rollouts(1 << lg, seed=2048 + i, buffers=buf, with_stats=False)
            torch.cuda.synchronize()
            st = int(buf.lengths.to(torch.int64).sum().item())
            print("rollout 2^%d env_steps %d" % (lg, st))
            buf = None
    if a.which in ("rollout", "all"):
        n = 1 << a.rollout_log2
        buf = r48.RolloutBuffers(n)
        for i in range(a.reps):
            r48.random_rollouts(n, seed=2048 + i, buffers=buf)
        torch.cuda.synchronize()
        print("rollout", r48.EpisodeStats(buf.stats).summary())
    if a.which in ("step", "afterstates", "all", "quick"):
        n = 1 << 23
        env = r48.BatchedGame(n, seed=2048)
        for _ in range(64):
            env.step(torch.randint(0, 4, (n,), device="cuda", dtype=torch.uint8))
        boards = env.boards.clone()
        acts = torch.randint(0, 4, (n,), device="cuda", dtype=torch.uint8)
        out = torch.empty_like(boards)
        rw = torch.empty(n, dtype=torch.int32, device="cuda")
        dn = torch.empty(n, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
    if a.which in ("step", "all", "quick"):
        m = 1 << 20
        for i in range(a.reps + 5):          # PROFILE_STEP_1M launches (rotating 1M windows)
            o = (i % 8) * m
            r48._native.check(L.r48_step(boards[o:].data_ptr(), acts[o:].data_ptr(), out[o:].data_ptr(),
                                         rw[o:].data_ptr(), dn[o:].data_ptr(), m, 2048, o, 64, 0, None, stream))
        for i in range(a.reps):              # one 8M-board launch
            r48._native.check(L.r48_step(boards.data_ptr(), acts.data_ptr(), out.data_ptr(), rw.data_ptr(),
                                         dn.data_ptr(), n, 2048, 0, 64, 0, None, stream))
        torch.cuda.synchronize()
        print("step done")
    if a.which in ("afterstates", "all"):
        for i in range(a.reps):
            r48.afterstates(boards)
        torch.cuda.synchronize()
        print("afterstates done")
    if a.which in ("env", "all"):
        m = 1 << 20
        env = r48.BatchedGame(m, seed=7)
        act = torch.randint(0, 4, (m,), device="cuda", dtype=torch.uint8)
        for i in range(a.reps + 40):
            env.env_step(act)
        torch.cuda.synchronize()
        print("env_step done")
    if a.which in ("greedy", "all"):
        n = 1 << a.rollout_log2
        buf = r48.RolloutBuffers(n)
        for i in range(a.reps):
            r48.random_rollouts(n, seed=2048 + i, buffers=buf, policy="greedy_blanks")
        torch.cuda.synchronize()
        print("greedy", r48.EpisodeStats(buf.stats).summary())
    if a.which in ("ring", "all"):
        m, cap = 1 << 20, 1 << 24
        ring = r48.ReplayRing(cap, seed=7)
        src = [torch.randint(0, 1 << 62, (m,), device="cuda", dtype=torch.int64) for _ in range(2)]
        act = torch.randint(0, 4, (m,), device="cuda", dtype=torch.uint8)
        rw = torch.zeros(m, dtype=torch.int32, device="cuda")
        for i in range(17 + a.reps):
            ring.store(src[i & 1], act, rw, src[1 - (i & 1)], act)
        for i in range(a.reps):
            ring.sample(m)
        env = r48.BatchedGame(m, seed=7)
        for i in range(a.reps + 40):
            env.env_step(act, ring=ring)
        torch.cuda.synchronize()
        print("ring done")
    if a.which in ("traj", "all"):
        for i in range(a.reps):
            tr = r48.rollout_trajectories(1 << 20, seed=2048 + i)
        torch.cuda.synchronize()
        print("trajectories", tr.transitions)


if __name__ == "__main__":
    main()
