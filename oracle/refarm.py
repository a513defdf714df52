# -*- coding: utf-8 -*-
"""The reference's own CPU path, timed: `random.seed(s); Game(); main.play(game, "rand")` over all
host cores.  TEST / BASELINE INFRASTRUCTURE (same rule as the rest of oracle/).

Runs the unmodified reference from oracle/_ref/ (byte-compiled by oracle/build_ref.py) when that
directory exists -- kind "reference" -- and oracle/pyport.py otherwise -- kind "port".
"""
import os
import random
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available():
    return os.path.exists(os.path.join(REF, "game", "GameClient.pyc")) and \
        os.path.exists(os.path.join(REF, "main.pyc"))


def _load():
    """(Game, play) of the unmodified reference.  The reference uses top-level packages `game`
    and `control`, so its build directory goes first on sys.path of the (worker) process."""
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import importlib
    main = importlib.import_module("main")             # main.py:3-8 imports game.GameClient, control.*
    return main.Game, main.play


def play_seeded(seed):
    """One config-1 episode through the reference's own loop.  -> (score, steps, max tile)"""
    Game, play = _load()
    random.seed(seed)
    game = Game()
    counted = {"n": 0}
    step = game.step

    def counting_step(action):                          # main.play does not report the step count
        counted["n"] += 1
        return step(action)
    game.step = counting_step
    score = play(game, "rand", show_state=False, show_result=False)
    return int(score), counted["n"], max(max(row) for row in game.state_matrix)


def _worker(seed_range):
    lo, hi = seed_range
    return sum(play_seeded(s)[1] for s in range(lo, hi))


def timed_rollouts(episodes, processes):
    """`episodes` seeded games over a multiprocessing pool -> (total_steps, seconds, kind)."""
    if not available():
        from oracle import pyport
        steps, dt = pyport.timed_rollouts(episodes, processes)
        return steps, dt, "port"
    import multiprocessing as mp
    chunk = max(1, episodes // (processes * 8))
    ranges = [(lo, min(lo + chunk, episodes)) for lo in range(0, episodes, chunk)]
    ctx = mp.get_context("fork")
    with ctx.Pool(processes) as pool:
        pool.map(_worker, [(0, 1)] * processes)         # import + spin the workers up outside the clock
        t0 = time.perf_counter()
        total = sum(pool.map(_worker, ranges))
        dt = time.perf_counter() - t0
    return total, dt, "reference"
