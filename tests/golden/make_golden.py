# -*- coding: utf-8 -*-
"""Generate the golden fixtures in this directory by running the UNMODIFIED reference.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py [--episodes 200000] [--procs 8]

It imports game/GameClient.py, control/rand.py and main.play from /root/reference as they
are and records their behaviour; nothing from the reference is copied into this repo, only
the OUTPUTS below (all arrays, deterministic given the seeds written in each file):

  rows_ref.npz      exhaustive 65536 rows x 4 actions through Game.update_matrix on 1x4 / 4x1
                    matrices, the way game/GameClientTest.py:49-331 drives it
  boards_ref.npz    random + mid-game 4x4 boards x 4 actions through Game.update_matrix,
                    Game.has_game_over, Game.has_table_filled
  episodes_ref.npz  seeded whole episodes (random.seed(s); Game(); main.play loop) with every
                    random draw recorded, for draw-injected transition parity, plus the
                    per-seed fingerprints (steps, score, max tile)
  dist_ref.npz      random-policy rollout distributions (length / score / max tile) and spawn
                    statistics over --episodes reference games
  testvectors_ref.npz  the literal vectors of game/GameClientTest.py, replayed through the
                    reference (so the expected values are the reference's, not retyped)
"""
import argparse
import multiprocessing as mp
import os
import random
import sys

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def ref_modules():
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import game.GameClient as gc
    from control.rand import Rand
    return gc, Rand


def exp_of(v):
    """tile value -> exponent (0 stays 0); exact, raises on non powers of two."""
    v = int(v)
    if v == 0:
        return 0
    e = v.bit_length() - 1
    assert v == 1 << e and e >= 1, v
    return e


def pack(matrix):
    """4x4 tile values -> uint64, cell (i,j) = nibble 4i+j; exponent 16 (65536) is kept in a
    side flag by the callers, here it saturates to 15."""
    b = 0
    for i in range(4):
        for j in range(4):
            b |= min(exp_of(matrix[i][j]), 15) << (4 * (4 * i + j))
    return b


def unpack(b):
    return [[(1 << ((b >> (4 * (4 * i + j))) & 15)) & ~1 for j in range(4)] for i in range(4)]


# ---------------------------------------------------------------- rows

def gen_rows():
    gc, _ = ref_modules()
    Game = gc.Game
    out = np.zeros((4, 65536, 4), np.uint8)       # exponents, 16 possible after 15+15
    changed = np.zeros((4, 65536), np.uint8)
    for r in range(65536):
        cells = [(1 << ((r >> (4 * t)) & 15)) & ~1 for t in range(4)]
        for a in range(4):
            if a < 2:
                m = [[c] for c in cells]          # 4x1 column, as GameClientTest.py:51
            else:
                m = [list(cells)]                 # 1x4 row, as GameClientTest.py:193
            res, reward, ch = Game.update_matrix(m, a)
            assert reward == 0
            flat = [row[0] for row in res] if a < 2 else res[0]
            out[a, r] = [exp_of(v) for v in flat]
            changed[a, r] = ch
    np.savez_compressed(os.path.join(HERE, "rows_ref.npz"), out_exp=out, changed=changed)
    print("rows_ref.npz", out.shape, int(changed.sum()))


# ---------------------------------------------------------------- boards

def midgame_boards(n, seed):
    """Boards sampled from reference random play (every step of seeded games)."""
    gc, Rand = ref_modules()
    boards = []
    s = seed
    while len(boards) < n:
        random.seed(s)
        s += 1
        g = gc.Game()
        over = False
        while not over and len(boards) < n:
            boards.append(pack(g.state_matrix))
            _, _, over = g.step(Rand.random_action())
        if len(boards) < n:
            boards.append(pack(g.state_matrix))   # the terminal board too
    return boards


def gen_boards(n_random=12000, n_mid=12000):
    gc, _ = ref_modules()
    Game = gc.Game
    rng = np.random.default_rng(20481)
    boards = []
    # random boards: exponent range and zero density vary per board so that full boards,
    # sparse boards and high tiles (up to 14, no 15+15 merges) are all present
    for _ in range(n_random):
        hi = int(rng.integers(2, 15))
        pz = float(rng.choice([0.0, 0.1, 0.3, 0.6]))
        e = rng.integers(1, hi + 1, 16)
        e[rng.random(16) < pz] = 0
        boards.append(sum(int(e[p]) << (4 * p) for p in range(16)))
    boards += midgame_boards(n_mid, seed=777)
    # hand-made edge cases
    boards += [0, 0x1111111111111111, 0x2121121221211212, 0xFFFFFFFFFFFFFFFF & 0xEEEEEEEEEEEEEEEE,
               0x0000000000000001, 0x1000000000000000, 0x1234123412341234, 0x4321432143214321]
    boards = np.array(boards, np.uint64)
    n = boards.size
    after = np.zeros((n, 4), np.uint64)
    changed = np.zeros((n, 4), np.uint8)
    over = np.zeros(n, np.uint8)
    filled = np.zeros(n, np.uint8)
    for i, b in enumerate(boards.tolist()):
        m = unpack(b)
        over[i] = Game.has_game_over(m)
        filled[i] = Game.has_table_filled(m)
        for a in range(4):
            res, reward, ch = Game.update_matrix(unpack(b), a)
            assert reward == 0
            after[i, a] = pack(res)
            changed[i, a] = ch
    np.savez_compressed(os.path.join(HERE, "boards_ref.npz"), boards=boards, after=after,
                        changed=changed, over=over, filled=filled)
    print("boards_ref.npz", n, int(over.sum()), int(filled.sum()))


# ---------------------------------------------------------------- episodes with recorded draws

class DrawLog:
    """Stands in for the `random` module inside game.GameClient: forwards to the real
    global generator (same MT19937 stream) and keeps what it returned."""

    def __init__(self):
        self.ints = []
        self.unis = []

    def randint(self, a, b):
        k = random.randint(a, b)
        self.ints.append((b + 1, k))
        return k

    def uniform(self, a, b):
        u = random.uniform(a, b)
        self.unis.append(u)
        return u


def record_episode(seed):
    gc, Rand = ref_modules()
    log = DrawLog()
    real = gc.random
    gc.random = log
    try:
        random.seed(seed)
        g = gc.Game()
        n0, k0 = log.ints[-1]
        assert n0 == 16
        reset = (k0, 1 if log.unis[-1] > 0.1 else 2, pack(g.state_matrix))
        rows = []
        over = False
        while not over:
            before = pack(g.state_matrix)
            n_int = len(log.ints)
            action = Rand.random_action(g.state_matrix)      # draws from the real module
            code = {"UP": 0, "DOWN": 1, "LEFT": 2, "RIGHT": 3}[action]
            _, reward, over = g.step(action)
            assert reward == 0
            if len(log.ints) > n_int:                        # a spawn happened
                nb, k = log.ints[-1]
                vexp = 1 if log.unis[-1] > 0.1 else 2
                ch = 1
            else:
                nb, k, vexp, ch = 0, 0, 0, 0
            rows.append((before, code, ch, nb, k, vexp, pack(g.state_matrix), int(over)))
    finally:
        gc.random = real
    score = int(np.sum(g.state_matrix))
    mx = int(np.max(g.state_matrix))
    return reset, rows, (len(rows), score, mx)


def gen_episodes(n_seeds=150):
    cols = {k: [] for k in ("seed", "before", "action", "changed", "n_blank", "k", "vexp",
                            "after", "done")}
    resets, prints = [], []
    for s in range(n_seeds):
        reset, rows, fp = record_episode(s)
        resets.append(reset)
        prints.append(fp)
        for r in rows:
            cols["seed"].append(s)
            for name, v in zip(("before", "action", "changed", "n_blank", "k", "vexp", "after",
                                "done"), r):
                cols[name].append(v)
    # cross-check the recording against an un-instrumented run of main.play
    gc, _ = ref_modules()
    sys.argv = ["main.py"]
    import main as ref_main
    for s in (0, 1, 2, 3, 4):
        random.seed(s)
        g = gc.Game()
        score = ref_main.play(g, "rand", show_result=False)
        assert int(score) == prints[s][1], (s, score, prints[s])
    np.savez_compressed(
        os.path.join(HERE, "episodes_ref.npz"),
        seed=np.array(cols["seed"], np.int32),
        before=np.array(cols["before"], np.uint64),
        action=np.array(cols["action"], np.uint8),
        changed=np.array(cols["changed"], np.uint8),
        n_blank=np.array(cols["n_blank"], np.uint8),
        k=np.array(cols["k"], np.uint8),
        vexp=np.array(cols["vexp"], np.uint8),
        after=np.array(cols["after"], np.uint64),
        done=np.array(cols["done"], np.uint8),
        reset_k=np.array([r[0] for r in resets], np.uint8),
        reset_vexp=np.array([r[1] for r in resets], np.uint8),
        reset_board=np.array([r[2] for r in resets], np.uint64),
        fingerprint=np.array(prints, np.int64),      # (steps, score, max tile) per seed
    )
    print("episodes_ref.npz", n_seeds, len(cols["seed"]), prints[:5])


# ---------------------------------------------------------------- distributions

def _dist_worker(args):
    lo, hi = args
    gc, Rand = ref_modules()
    log = DrawLog()
    gc.random = log
    len_h = np.zeros(2048, np.int64)
    score_h = np.zeros(2048, np.int64)
    max_h = np.zeros(17, np.int64)
    pos_h = np.zeros((17, 16), np.int64)
    val_h = np.zeros(3, np.int64)
    eff = 0
    for s in range(lo, hi):
        random.seed(s)
        log.ints.clear()
        log.unis.clear()
        g = gc.Game()
        steps = 0
        over = False
        while not over:
            _, _, over = g.step(Rand.random_action(g.state_matrix))
            steps += 1
        score = int(np.sum(g.state_matrix))
        len_h[min(steps, 2047)] += 1
        score_h[min(score // 2, 2047)] += 1
        max_h[exp_of(int(np.max(g.state_matrix)))] += 1
        for nb, k in log.ints:
            pos_h[nb, k] += 1
        for u in log.unis:
            val_h[1 if u > 0.1 else 2] += 1
        eff += len(log.ints) - 1
    return len_h, score_h, max_h, pos_h, val_h, eff


def gen_dist(episodes, procs):
    chunk = 500
    jobs = [(lo, min(lo + chunk, episodes)) for lo in range(0, episodes, chunk)]
    with mp.get_context("fork").Pool(procs) as pool:
        parts = pool.map(_dist_worker, jobs)
    len_h = sum(p[0] for p in parts)
    score_h = sum(p[1] for p in parts)
    max_h = sum(p[2] for p in parts)
    pos_h = sum(p[3] for p in parts)
    val_h = sum(p[4] for p in parts)
    eff = sum(p[5] for p in parts)
    np.savez_compressed(os.path.join(HERE, "dist_ref.npz"), episodes=np.int64(episodes),
                        len_hist=len_h, score_hist=score_h, maxexp_hist=max_h,
                        spawn_pos_hist=pos_h, spawn_val_hist=val_h, effective_moves=np.int64(eff))
    n = len_h.sum()
    mean_len = (len_h * np.arange(2048)).sum() / n
    mean_score = (score_h * np.arange(2048) * 2).sum() / n
    print("dist_ref.npz episodes", n, "mean len %.2f mean score %.2f" % (mean_len, mean_score),
          "P(4)=%.5f" % (val_h[2] / val_h.sum()), "maxexp", max_h.tolist())


# ---------------------------------------------------------------- the reference's own test vectors

def gen_testvectors():
    """Replay game/GameClientTest.py's inputs through the reference.  The ten input lines
    are the ones listed at GameClientTest.py:49-331 (same ten per direction)."""
    gc, _ = ref_modules()
    Game = gc.Game
    lines = [[0, 0, 1, 0], [1, 0, 1, 0], [2, 0, 1, 0], [2, 2, 1, 0], [2, 2, 2, 2],
             [8, 8, 4, 0], [8, 4, 4, 4], [2, 0, 0, 2], [0, 4, 2, 2], [8, 8, 8, 0]]
    out = np.zeros((4, len(lines), 4), np.int64)
    for a in range(4):
        for t, ln in enumerate(lines):
            m = [[c] for c in ln] if a < 2 else [list(ln)]
            res, _, _ = Game.update_matrix(m, a)
            out[a, t] = [row[0] for row in res] if a < 2 else res[0]
    over_boards = np.array([
        [[0, 0, 0, 0], [0, 2, 0, 0], [0, 0, 0, 0], [0, 0, 0, 0]],
        [[2, 4, 2, 4], [4, 2, 4, 2], [2, 4, 2, 4], [4, 2, 4, 2]],
        [[2, 4, 2, 4], [2, 4, 2, 4], [2, 4, 2, 4], [2, 4, 2, 4]]], np.int64)
    over = np.array([Game.has_game_over(b.tolist()) for b in over_boards], np.uint8)
    filled_boards = np.ones((3, 4, 4), np.int64)
    filled_boards[1, 1, 2] = 4
    filled_boards[2, 1, 1] = 0
    filled = np.array([Game.has_table_filled(b.tolist()) for b in filled_boards], np.uint8)
    np.savez_compressed(os.path.join(HERE, "testvectors_ref.npz"), lines=np.array(lines, np.int64),
                        moved=out, over_boards=over_boards, over=over,
                        filled_boards=filled_boards, filled=filled)
    print("testvectors_ref.npz", out[2].tolist())


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--episodes", type=int, default=200000)
    ap.add_argument("--procs", type=int, default=os.cpu_count())
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    todo = args.only.split(",") if args.only else ["testvectors", "rows", "boards", "episodes", "dist"]
    if "testvectors" in todo:
        gen_testvectors()
    if "rows" in todo:
        gen_rows()
    if "boards" in todo:
        gen_boards()
    if "episodes" in todo:
        gen_episodes()
    if "dist" in todo:
        gen_dist(args.episodes, args.procs)
