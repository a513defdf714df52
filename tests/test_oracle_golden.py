# -*- coding: utf-8 -*-
"""Pins the CPU oracle (oracle/r48_oracle.c) and the Python port (oracle/pyport.py) to the
reference: every fixture under tests/golden/ was produced by running the unmodified
reference (tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest

REF = "/root/reference"


def exps_to_values(e):
    e = np.asarray(e, np.int64)
    return np.where(e > 0, np.int64(1) << e, 0)


def row_cells(r):
    return [(1 << ((r >> (4 * t)) & 15)) & ~1 for t in range(4)]


def line_matrix(cells, action):
    """Embed one line in an otherwise empty 4x4 board: column 0 for UP/DOWN, row 0 for LEFT/RIGHT."""
    m = np.zeros((4, 4), np.int64)
    if action < 2:
        m[:, 0] = cells
    else:
        m[0, :] = cells
    return m


def line_of(m, action):
    return m[:, 0] if action < 2 else m[0, :]


# ------------------------------------------------------------------ the reference's own test vectors

def test_reference_test_vectors(orc, golden):
    """game/GameClientTest.py:49-331 -- 10 lines x (U, D, L, R); tile '1' is legal here because
    the oracle's rules work on tile VALUES, like the reference."""
    g = golden("testvectors_ref.npz")
    for a in range(4):
        for t, cells in enumerate(g["lines"]):
            out, _, _ = orc.update_matrix(line_matrix(cells, a), a)
            assert line_of(out, a).tolist() == g["moved"][a, t].tolist(), (a, cells)
    # literal expectations of GameClientTest.py:191-260 (LEFT block) as a guard on the fixture
    assert g["moved"][2].tolist() == [[1, 0, 0, 0], [2, 0, 0, 0], [2, 1, 0, 0], [4, 1, 0, 0],
                                      [4, 4, 0, 0], [16, 4, 0, 0], [8, 8, 4, 0], [4, 0, 0, 0],
                                      [4, 4, 0, 0], [16, 8, 0, 0]]


def test_reference_game_over_and_filled_vectors(orc, golden):
    """GameClientTest.py:10-21 and :23-31."""
    g = golden("testvectors_ref.npz")
    assert g["over"].tolist() == [0, 1, 0]
    assert g["filled"].tolist() == [1, 1, 0]
    for b, want in zip(g["over_boards"], g["over"]):
        assert orc.has_game_over(b) == bool(want)
    for b, want in zip(g["filled_boards"], g["filled"]):
        assert orc.has_table_filled(b) == bool(want)


def test_spawn_on_full_board_is_noop(orc):
    """GameClientTest.py:43-44."""
    full = np.array([[2, 4, 2, 4], [4, 2, 4, 2], [2, 4, 2, 4], [4, 2, 4, 2]])
    out, n = orc.random_fill_grid(full, 0, 2)
    assert n == 0 and (out == full).all()


# ------------------------------------------------------------------ exhaustive rows

def test_rows_exhaustive(orc, golden):
    """All 65536 rows x 4 actions against the reference's outputs (incl. 15+15 -> 65536)."""
    g = golden("rows_ref.npz")
    want, want_changed = g["out_exp"], g["changed"]
    bad = 0
    for r in range(0, 65536):
        cells = row_cells(r)
        for a in range(4):
            out, _, changed = orc.update_matrix(line_matrix(cells, a), a)
            got = line_of(out, a)
            if got.tolist() != exps_to_values(want[a, r]).tolist() or changed != bool(want_changed[a, r]):
                bad += 1
    assert bad == 0


def test_rows_packed_saturation(orc, golden):
    """Packed move == reference with exponent 16 clamped to 15 (documented domain edge)."""
    g = golden("rows_ref.npz")
    want = np.minimum(g["out_exp"], 15)
    for r in range(0, 65536, 7):
        for a in (2, 3):
            out, changed, _ = orc.move(r, a)
            exp = sum(int(want[a, r, t]) << (4 * t) for t in range(4))
            assert out == exp and changed == bool(g["changed"][a, r])


# ------------------------------------------------------------------ full boards

def test_boards_against_reference(orc, golden):
    g = golden("boards_ref.npz")
    boards = g["boards"]
    out, _, valid, done = orc.afterstates_batch(boards)
    assert (out.T == g["after"]).all()
    got_changed = (valid[:, None] >> np.arange(4)) & 1
    assert (got_changed == g["changed"]).all()
    assert (done == g["over"]).all()
    filled = np.array([orc.has_table_filled(orc.decode(b)) for b in boards[:4000]], np.uint8)
    assert (filled == g["filled"][:4000]).all()
    # game over <=> no action changes the board (the empty board is the one exception:
    # nothing moves, yet it is not full)
    nonempty = boards != 0
    assert ((valid == 0) == (g["over"] == 1))[nonempty].all()


# ------------------------------------------------------------------ recorded episodes, draws injected

def test_episode_transitions_with_injected_draws(orc, golden):
    g = golden("episodes_ref.npz")
    out, reward, done = orc.step_injected_batch(g["before"], g["action"], g["k"], g["vexp"])
    assert (out == g["after"]).all()
    assert (reward == 0).all()                      # the reference's reward is 0
    assert (done == g["done"]).all()
    # reset = one tile
    for k, v, b in zip(g["reset_k"], g["reset_vexp"], g["reset_board"]):
        m, n = orc.random_fill_grid(np.zeros((4, 4), np.int64), int(k), 1 << int(v))
        assert n == 16 and orc.encode(m) == int(b)


def test_pyport_matches_reference_episodes(golden):
    """The Python port consumes MT19937 in the reference's order: same seed, same episode."""
    from oracle import pyport
    fp = golden("episodes_ref.npz")["fingerprint"]
    for s in range(0, 40):
        score, steps, mx = pyport.play_seeded(s)
        assert (steps, score, mx) == tuple(fp[s]), s
    # SURVEY section 6 / 8c pin 3
    assert tuple(fp[0]) == (180, 352, 128) and tuple(fp[1]) == (109, 194, 64)


def test_pyport_slide_matches_oracle(orc):
    from oracle import pyport
    rng = np.random.default_rng(5)
    for _ in range(3000):
        e = rng.integers(0, 8, 16)
        vals = exps_to_values(e).reshape(4, 4)
        a = int(rng.integers(0, 4))
        grid, reward, changed = pyport.slide(vals.tolist(), pyport.ACTION_NAMES[a])
        out, _, ch = orc.update_matrix(vals, a)
        assert reward == 0 and grid == out.tolist() and changed == ch
        assert pyport.game_over(vals.tolist()) == orc.has_game_over(vals)


def test_pyport_action_spellings():
    from oracle import pyport
    for code, names in enumerate((("UP", "Up", "U", "up", "u", 0), ("DOWN", "Down", "D", "down", "d", 1),
                                  ("LEFT", "Left", "L", "left", "l", 2),
                                  ("RIGHT", "Right", "R", "right", "r", 3))):
        for nm in names:
            assert pyport.action_index(nm) == code
    with pytest.raises(ValueError):
        pyport.action_index("X")


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present (GPU box)")
def test_live_reference_spot_check(orc):
    """Where the reference is importable (build container) compare live, not just fixtures."""
    import sys
    sys.path.insert(0, REF)
    try:
        from game.GameClient import Game
    finally:
        sys.path.remove(REF)
    rng = np.random.default_rng(99)
    for _ in range(2000):
        e = rng.integers(0, 12, 16)
        vals = exps_to_values(e).reshape(4, 4)
        a = int(rng.integers(0, 4))
        res, reward, ch = Game.update_matrix(vals.tolist(), a)
        out, _, och = orc.update_matrix(vals, a)
        assert res == out.tolist() and ch == och and reward == 0
        assert Game.has_game_over(vals.tolist()) == orc.has_game_over(vals)


# ------------------------------------------------------------------ Philox + draw spec

def test_philox_known_answers(orc):
    """Random123 philox4x32-10 KATs (SURVEY 8c pin 5)."""
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
         (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]
    for ctr, key, want in kat:
        assert tuple(int(x) for x in orc.philox(ctr, key)) == want


def test_draw_spec(orc):
    seed, bid = 0x1234567890ABCDEF, (5 << 32) | 77
    for tick in range(6):
        w = orc.philox((bid & 0xffffffff, bid >> 32, tick >> 1, 0), (seed & 0xffffffff, seed >> 32))
        a, v = orc.draw(seed, bid, tick)
        assert (a, v) == (int(w[2 * (tick & 1)]), int(w[2 * (tick & 1) + 1]))


def test_step_equals_rollout(orc):
    """orc_rollout == reset followed by orc_step with the Philox actions, tick by tick."""
    seed = 2048
    fb, ln = orc.rollout(64, seed, board_base=1000)
    boards = orc.reset_batch(64, seed, 1000)
    alive = np.ones(64, bool)
    length = np.zeros(64, np.uint32)
    step = 0
    while alive.any():
        acts = np.array([orc.draw(seed, 1000 + i, step + 1)[0] >> 30 for i in range(64)], np.uint8)
        nb, rw, dn = orc.step_batch(boards, acts, seed, 1000, step)
        boards = np.where(alive, nb, boards)
        length[alive] += 1
        alive &= dn == 0
        step += 1
    assert (boards == fb).all() and (length == ln).all()


def test_stats_vector(orc):
    fb, ln = orc.rollout(500, 7)
    st = orc.episode_stats(fb, ln)
    sc = orc.scores(fb)
    assert st[orc.ST_EPISODES] == 500
    assert st[orc.ST_SUM_LEN] == ln.sum() and st[orc.ST_SUM_SCORE] == sc.sum()
    assert st[orc.ST_SUM_SCORE2] == (sc.astype(np.uint64) ** 2).sum()
    assert st[orc.ST_HIST_MAXEXP:orc.ST_HIST_MAXEXP + 16].sum() == 500
    assert st[orc.ST_HIST_LEN:orc.ST_HIST_LEN + 2048].sum() == 500
    assert st[orc.ST_HIST_SCORE:orc.ST_HIST_SCORE + 2048].sum() == 500
    # two halves accumulate to the whole (what the multi-GPU all-reduce relies on)
    a = orc.episode_stats(fb[:200], ln[:200])
    b = orc.episode_stats(fb[200:], ln[200:])
    assert ((a + b) == st).all()
