# -*- coding: utf-8 -*-
"""rein48_b200 -- B200-native batched 2048 environment and random-rollout path.

A drop-in for the env / rollout hot path of nevertiree/Rein48 (game/GameClient.py,
control/rand.py, main.play): hand-written sm_100a CUDA kernels behind a C ABI
(include/r48.h, libr48.so), called from Python with zero-copy torch tensors.
"""
from . import _native
from ._native import R48Error, build
from .batched import BatchedGame, HostPlayerViews, afterstates, blank_counts, decode, encode, scores, spawn_injected
from .game import Game, action_code
from .rand import (HostRecords, Rand, RolloutBuffers, RolloutResult, Trajectories, play, random_rollouts,
                   random_rollouts_host, record_lengths, record_scores, rollout_trajectories, sharded_rollouts)
from .replay import ReplayRing
from .stats import STATS_WORDS, EpisodeStats, allreduce_stats, shard_range

__all__ = [
    "BatchedGame", "HostPlayerViews", "Game", "Rand", "play", "random_rollouts", "random_rollouts_host", "HostRecords", "record_scores", "record_lengths",
    "sharded_rollouts", "rollout_trajectories", "Trajectories", "ReplayRing", "RolloutBuffers", "RolloutResult", "EpisodeStats", "allreduce_stats",
    "shard_range", "afterstates", "decode", "encode", "scores", "blank_counts", "spawn_injected",
    "action_code", "build", "R48Error", "STATS_WORDS",
]
__version__ = "0.2.0"
