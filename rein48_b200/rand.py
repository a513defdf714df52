# -*- coding: utf-8 -*-
"""Random-policy driver: the reference's control/rand.py + main.play loop, batched.

`Rand.random_action` keeps the reference's signature (control/rand.py:9-11).  The data path
is `random_rollouts`: n episodes from reset to game over inside ONE fused CUDA kernel
(r48_rollout), replacing the Python loop at main.py:36-42; multi-GPU runs give each rank a
contiguous slice of global episode ids and all-reduce only the statistics vector.
"""
import random
from collections import namedtuple

import torch

from . import _native
from .batched import _require_cuda, _stream, scores
from .stats import STATS_WORDS, EpisodeStats, allreduce_stats, shard_range

ACTION_NAMES = ("UP", "DOWN", "LEFT", "RIGHT")


class Rand:
    """control/rand.py: a uniformly random action NAME, ignoring its arguments."""

    @staticmethod
    def random_action(*args):
        return ACTION_NAMES[random.randint(0, 3)]


RolloutResult = namedtuple("RolloutResult", "final_boards lengths stats")


class RolloutBuffers:
    """Device buffers for repeated rollouts of up to `n` episodes (allocated once)."""

    def __init__(self, n, device="cuda"):
        self.device = _require_cuda(device)
        self.n = int(n)
        with torch.cuda.device(self.device):
            self.final_boards = torch.empty(self.n, dtype=torch.int64, device=self.device)
            self.lengths = torch.empty(self.n, dtype=torch.int32, device=self.device)
            self.stats = torch.zeros(STATS_WORDS, dtype=torch.int64, device=self.device)
            self.workspace = torch.zeros(_native.ROLLOUT_WORKSPACE_BYTES // 8, dtype=torch.int64,
                                         device=self.device)


POLICIES = {"random": 0, "greedy_blanks": 1}


def random_rollouts(n, seed=0, device="cuda", board_base=0, buffers=None, with_stats=True, policy="random"):
    """Play global episodes board_base .. board_base+n-1 to game over with the uniform random
    policy (or policy="greedy_blanks": 1-ply greedy on the number of blank cells, random
    tie-break order).  Returns RolloutResult(final_boards int64[n], lengths int32[n],
    stats int64[4120]); everything stays on the device and the call does not synchronise."""
    if buffers is None:
        buffers = RolloutBuffers(n, device)
    if n > buffers.n:
        raise ValueError("buffers hold %d episodes, asked for %d" % (buffers.n, n))
    dev = buffers.device
    L = _native.lib()
    with torch.cuda.device(dev):
        if with_stats:
            buffers.stats.zero_()
        _native.check(L.r48_rollout_policy(
            int(n), int(seed) & 0xFFFFFFFFFFFFFFFF, int(board_base), POLICIES[policy],
            buffers.final_boards.data_ptr(),
            buffers.lengths.data_ptr(), buffers.stats.data_ptr() if with_stats else None,
            buffers.workspace.data_ptr(), _stream(dev)))
    return RolloutResult(buffers.final_boards[:n], buffers.lengths[:n], buffers.stats if with_stats else None)


class Trajectories(namedtuple("Trajectories", "offsets boards actions final_boards lengths stats")):
    """Flat (state, action) log of a batch of episodes; see rollout_trajectories."""

    def valid_mask(self):
        """bool[T]: True on slots that hold a step (False on the alignment padding)."""
        t = self.boards.numel()
        slot = torch.arange(t, device=self.boards.device)
        ep = torch.searchsorted(self.offsets[1:].contiguous(), slot, right=True)
        return slot - self.offsets[ep] < self.lengths.to(torch.int64)[ep]

    @property
    def transitions(self):
        return int(self.lengths.sum().item())


def rollout_trajectories(n, seed=0, device="cuda", board_base=0, policy="random"):
    """Play n episodes and keep every transition: data generation for the learners (the
    reference's README: "run tens of thousands of games on the GPU to make data").
    Returns Trajectories(offsets int64[n+1], boards int64[T], actions uint8[T], final_boards,
    lengths, stats).  Episode i occupies slots offsets[i] .. offsets[i] + lengths[i] - 1 (every
    episode starts on a multiple of 4 slots; the up-to-3 slots of padding behind it are
    unspecified); boards[] holds the state BEFORE each step, the state after the last one is
    final_boards[i].  `valid_mask()` on the result marks the real slots."""
    res = random_rollouts(n, seed=seed, device=device, board_base=board_base, policy=policy)
    dev = res.final_boards.device
    L = _native.lib()
    with torch.cuda.device(dev):
        offsets = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        torch.cumsum((res.lengths.to(torch.int64) + 3) & ~3, 0, out=offsets[1:])
        total = int(offsets[-1].item())
        boards = torch.empty(total, dtype=torch.int64, device=dev)
        actions = torch.empty(total, dtype=torch.uint8, device=dev)
        ws = torch.zeros(_native.ROLLOUT_WORKSPACE_BYTES // 8, dtype=torch.int64, device=dev)
        _native.check(L.r48_rollout_trajectories(
            int(n), int(seed) & 0xFFFFFFFFFFFFFFFF, int(board_base), POLICIES[policy], res.lengths.data_ptr(),
            offsets.data_ptr(), boards.data_ptr(), actions.data_ptr(), ws.data_ptr(), _stream(dev)))
    return Trajectories(offsets, boards, actions, res.final_boards, res.lengths, res.stats)


def sharded_rollouts(n_total, seed=0, device=None, buffers=None, group=None):
    """One rank's share of n_total episodes + the SUM all-reduce of the statistics vector.
    Call from every rank of an initialised torch.distributed job (or standalone)."""
    import torch.distributed as dist
    rank = dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    lo, hi = shard_range(n_total, rank, world)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    res = random_rollouts(hi - lo, seed=seed, device=device, board_base=lo, buffers=buffers)
    allreduce_stats(res.stats, group=group)
    return res


HostRecords = namedtuple("HostRecords", "records stats")


def record_scores(records):
    """score of every episode of a packed record array (main.py:48's np.sum(state_matrix))."""
    return (records.to(torch.int64) & 0xFFFFFFFF) >> 13 << 1


def record_lengths(records):
    """min(length, 8191) of every episode of a packed record array."""
    return records.to(torch.int64) & 8191


def random_rollouts_host(n, seed=0, device=0, board_base=0, out=None, policy="random", records=False):
    """End-to-end form over HOST buffers (r48_rollout_host_ex): results land in pinned host
    tensors.  records=False: RolloutResult(final_boards int64[n], lengths int32[n], stats) --
    12 bytes per episode cross PCIe; records=True: HostRecords(records int32[n], stats) -- one
    packed word per episode (score / 2 in bits 31..13, min(length, 8191) below; see
    record_scores / record_lengths), 4 bytes per episode."""
    if not torch.cuda.is_available():
        raise RuntimeError("rein48_b200 needs a CUDA device (B200, sm_100a); none is visible")
    if out is None:
        stats = torch.empty(STATS_WORDS, dtype=torch.int64).pin_memory()
        if records:
            out = HostRecords(torch.empty(n, dtype=torch.int32).pin_memory(), stats)
        else:
            out = RolloutResult(torch.empty(n, dtype=torch.int64).pin_memory(),
                                torch.empty(n, dtype=torch.int32).pin_memory(), stats)
    index = torch.device(device).index if not isinstance(device, int) else device
    if records:
        ptrs = (None, None, out.records.data_ptr())
    else:
        ptrs = (out.final_boards.data_ptr(), out.lengths.data_ptr(), None)
    _native.check(_native.lib().r48_rollout_host_ex(
        int(n), int(seed) & 0xFFFFFFFFFFFFFFFF, int(board_base), POLICIES[policy], ptrs[0], ptrs[1], ptrs[2],
        out.stats.data_ptr() if out.stats is not None else None, int(index or 0)))
    return out


def play(game, control="rand", show_state=False, show_result=False):
    """main.play (main.py:11-48) for the adapter `Game`: loop the chosen policy until game
    over and return the sum of tiles.  Only the random policy is data-parallel; "hand" is
    the reference's interactive keyboard loop and is not provided."""
    import numpy as np
    if control != "rand":
        raise NotImplementedError("only control='rand' is provided (control/hand.py is interactive)")
    over = False
    while not over:
        if show_state:
            game.print_terminal(game.state_matrix)
        _, _, over = game.step(Rand.random_action(game.state_matrix))
    if show_result:
        game.print_terminal(game.state_matrix)
    return np.sum(game.state_matrix)


__all__ = ["Rand", "RolloutBuffers", "RolloutResult", "random_rollouts", "sharded_rollouts",
           "rollout_trajectories", "Trajectories",
           "random_rollouts_host", "HostRecords", "record_scores", "record_lengths", "play", "EpisodeStats", "scores"]
