# -*- coding: utf-8 -*-
"""The reference's own CPU path, timed: `random.seed(s); Game(); main.play(game, "rand")` over all
host cores.  TEST / BASELINE INFRASTRUCTURE (same rule as the rest of oracle/).

Runs the unmodified reference from oracle/_ref/ (compiled by oracle/build_ref.py) when that
directory exists and matches this interpreter -- kind "reference" -- and oracle/pyport.py otherwise
-- kind "port".
"""
import marshal
import os
import random
import sys
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
# module name -> compiled file, in import order (main.py:3-8 imports the other three)
MODULES = (("control.rand", "control/rand.r48c"), ("control.hand", "control/hand.r48c"),
           ("game.GameClient", "game/GameClient.r48c"), ("r48_reference_main", "main.r48c"))
_HEADER = b"R48C%d.%d\n" % sys.version_info[:2]


def available():
    try:
        for _, rel in MODULES:
            with open(os.path.join(REF, rel), "rb") as fh:
                if fh.read(len(_HEADER)) != _HEADER:
                    return False
        return True
    except OSError:
        return False


_loaded = None


def _load():
    """(Game, play) of the unmodified reference: every compiled module is executed into a module
    object registered under the name the reference imports it by (`game.GameClient`, `control.rand`,
    `control.hand`; main.py itself under a private name)."""
    global _loaded
    if _loaded:
        return _loaded
    for pkg in ("game", "control"):
        if pkg not in sys.modules:
            m = types.ModuleType(pkg)
            m.__path__ = []
            sys.modules[pkg] = m
    main = None
    for name, rel in MODULES:
        with open(os.path.join(REF, rel), "rb") as fh:
            fh.read(len(_HEADER))
            code = marshal.load(fh)
        mod = types.ModuleType(name)
        mod.__file__ = os.path.join(REF, rel)
        sys.modules[name] = mod
        if "." in name:
            setattr(sys.modules[name.split(".")[0]], name.split(".")[1], mod)
        exec(code, mod.__dict__)                       # __name__ != "__main__": main() does not run
        main = mod
    _loaded = (main.Game, main.play)
    return _loaded


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
