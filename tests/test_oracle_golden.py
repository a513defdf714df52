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
    """Random123 kat_vectors for philox4x32 with 10 rounds (SURVEY 8c pin 5) and with 7 rounds (the
    count the draw spec uses)."""
    kat = [
        (10, (0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        (10, (0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        (10, (0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
         (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
        (7, (0, 0, 0, 0), (0, 0), (0x5f6fb709, 0x0d893f64, 0x4f121f81, 0x4f730a48)),
        (7, (0xffffffff,) * 4, (0xffffffff,) * 2, (0x5207ddc2, 0x45165e59, 0x4d8ee751, 0x8c52f662)),
        (7, (0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
         (0x4dfccaba, 0x190a87f0, 0xc47362ba, 0xb6b5242a)),
    ]
    for rounds, ctr, key, want in kat:
        assert tuple(int(x) for x in orc.philox(ctr, key, rounds)) == want
    assert orc.DRAW_ROUNDS == 7


def test_draw_spec(orc):
    """tick t -> word t & 3 of the Philox4x32-7 call (id lo, id hi, t >> 2, 0)."""
    seed, bid = 0x1234567890ABCDEF, (5 << 32) | 77
    for tick in range(11):
        w = orc.philox((bid & 0xffffffff, bid >> 32, tick >> 2, 0), (seed & 0xffffffff, seed >> 32))
        assert orc.draw(seed, bid, tick) == int(w[tick & 3])


def test_spawn_follows_the_axis_of_the_move(orc):
    """The k-th blank is counted row-major after LEFT/RIGHT and column-major after UP/DOWN; the
    value is 4 iff word * 0x9E3779B1 mod 2^32 < ceil(0.1 * 2^32)."""
    rng = np.random.default_rng(5)
    seed, base = 99, 1 << 40
    for trial in range(400):
        vals = np.where(rng.random(16) < 0.45, 0, 2 ** rng.integers(1, 8, 16)).reshape(4, 4)
        board = orc.encode(vals)
        action = int(rng.integers(0, 4))
        step = int(rng.integers(0, 50))
        moved, _, changed = orc.update_matrix(vals, action)
        got = orc.decode(int(orc.step_batch(np.array([board], np.uint64), np.array([action], np.uint8), seed,
                                            base + trial, step)[0][0]))
        if not changed:
            assert (got == moved).all()
            continue
        a = orc.draw(seed, base + trial, step + 1)
        n = int((moved == 0).sum())
        k = (((a << 2) & 0xffffffff) * n) >> 32
        value = 4 if (a * orc.VALUE_HASH) & 0xffffffff < orc.SPAWN4_THRESHOLD else 2
        fill = orc.random_fill_grid_colmajor if action < 2 else orc.random_fill_grid
        want, _ = fill(moved, k, value)
        assert (got == want).all()
    # column-major really is the transpose of row-major
    m = np.array([[0, 2, 0, 0], [4, 0, 0, 2], [0, 0, 8, 0], [2, 0, 0, 0]])
    for k in range(int((m == 0).sum())):
        a, _ = orc.random_fill_grid_colmajor(m, k, 2)
        b, _ = orc.random_fill_grid(m.T, k, 2)
        assert (a == b.T).all()


def test_draw_fields_are_jointly_uniform():
    """action, cell and value all come from ONE word: action = top 2 bits, cell from the next 30,
    value from the word times an odd constant.  Chi-square of the joint (action, cell, value) table
    against independence with the exact marginals, alpha = 1e-3 per n (Bonferroni over 8 values of n)."""
    from scipy.stats import chi2
    from oracle import oracle as o
    rng = np.random.default_rng(2048)
    t = o.SPAWN4_THRESHOLD
    for n in (1, 2, 3, 5, 7, 11, 13, 16):
        a = rng.integers(0, 1 << 32, size=1 << 24, dtype=np.uint64)
        k = (((a << np.uint64(2)) & np.uint64(0xffffffff)) * np.uint64(n)) >> np.uint64(32)
        four = ((a * np.uint64(o.VALUE_HASH)) & np.uint64(0xffffffff)) < t
        idx = ((a >> np.uint64(30)) * np.uint64(n) + k) * np.uint64(2) + four.astype(np.uint64)
        cnt = np.bincount(idx.astype(np.int64), minlength=8 * n).astype(float)
        exp = np.empty(8 * n)
        exp[0::2] = a.size * (1 - t / 2 ** 32) / (4 * n)
        exp[1::2] = a.size * (t / 2 ** 32) / (4 * n)
        x2 = ((cnt - exp) ** 2 / exp).sum()
        assert chi2.sf(x2, 8 * n - 1) > 1e-3 / 8, (n, x2)


def test_step_equals_rollout(orc):
    """orc_rollout == reset followed by orc_step with the Philox actions, tick by tick."""
    seed = 2048
    fb, ln = orc.rollout(64, seed, board_base=1000)
    boards = orc.reset_batch(64, seed, 1000)
    alive = np.ones(64, bool)
    length = np.zeros(64, np.uint32)
    step = 0
    while alive.any():
        acts = np.array([orc.draw(seed, 1000 + i, step + 1) >> 30 for i in range(64)], np.uint8)
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


def test_episode_records(orc):
    fb, ln = orc.rollout(300, 5)
    ln[7] = 9000                                        # beyond the 13-bit field: clamps
    rec = orc.episode_records(fb, ln)
    assert ((rec >> 13) << 1 == orc.scores(fb)).all()
    assert ((rec & 8191) == np.minimum(ln, 8191)).all()


# ------------------------------------------------------------------ transition ring (replay.py:8-47)

def test_ring_follows_the_reference_replay(orc):
    """The oracle ring against the reference's Replay semantics restated with a Python list:
    store drops when full (replay.py:18-21), sample draws distinct items or everything (:30-33)."""
    cap = 100
    ring = orc.Ring(cap)
    rng = np.random.default_rng(3)
    ref = []
    for _ in range(7):
        n = int(rng.integers(1, 40))
        s = rng.integers(0, 1 << 62, n).astype(np.uint64)
        a = rng.integers(0, 4, n).astype(np.uint8)
        r = rng.integers(0, 99, n).astype(np.int32)
        nx = rng.integers(0, 1 << 62, n).astype(np.uint64)
        d = rng.integers(0, 2, n).astype(np.uint8)
        ring.append(s, a, r, nx, d, drop_when_full=True)
        for i in range(n):
            if len(ref) < cap:                          # Replay.store
                ref.append((s[i], a[i], r[i], nx[i], d[i]))
    assert ring.size() == len(ref) == cap
    assert [tuple(x) for x in zip(ring.state, ring.action, ring.reward, ring.next_state, ring.done)] == ref
    idx = ring.sample_indices(10, seed=1, draw_id=0)
    assert len(set(idx.tolist())) == 10 and idx.min() >= 0 and idx.max() < cap
    assert (ring.sample_indices(10, seed=1, draw_id=0) == idx).all()          # keyed, reproducible
    assert (ring.sample_indices(10, seed=1, draw_id=1) != idx).any()
    allidx = ring.sample_indices(130, seed=1, draw_id=2)                      # asks for more than it holds
    assert sorted(allidx[:cap].tolist()) == list(range(cap)) and (allidx[cap:] == -1).all()


def test_ring_overwrites_oldest(orc):
    cap = 37
    ring = orc.Ring(cap)
    total = 0
    for n in (10, 30, 90, 5):                           # 90 > capacity: only its last 37 survive
        s = np.arange(total, total + n, dtype=np.uint64)
        ring.append(s, (s & 3).astype(np.uint8), s.astype(np.int32), s + 1000, (s & 1).astype(np.uint8))
        total += n
    assert int(ring.cursor[0]) == total and ring.size() == cap
    want = np.zeros(cap, np.uint64)
    for v in range(total - cap, total):
        want[v % cap] = v
    assert (ring.state == want).all() and (ring.next_state == want + 1000).all()
    assert (ring.reward == want.astype(np.int32)).all()


def test_ring_sample_indices_are_uniform(orc):
    """Chi-square of slot frequencies over many keyed draws, both samplers; alpha = 1e-3."""
    from scipy.stats import chi2
    cap, batch, draws = 211, 32, 4000
    ring = orc.Ring(cap)
    z = np.zeros(cap + 5, np.uint64)
    ring.append(z, z.astype(np.uint8), z.astype(np.int32), z, z.astype(np.uint8))
    for repl in (False, True):
        cnt = np.zeros(cap)
        for d in range(draws):
            idx = ring.sample_indices(batch, seed=77, draw_id=d, with_replacement=repl)
            if not repl:
                assert len(set(idx.tolist())) == batch
            cnt += np.bincount(idx, minlength=cap)
        exp = draws * batch / cap
        x2 = ((cnt - exp) ** 2 / exp).sum()
        assert chi2.sf(x2, cap - 1) > 1e-3, (repl, x2)
    # position within the sample is uniform too (element 0 of a without-replacement draw)
    first = np.array([ring.sample_indices(1, seed=78, draw_id=d)[0] for d in range(20000)])
    cnt = np.bincount(first, minlength=cap)
    assert chi2.sf(((cnt - 20000 / cap) ** 2 / (20000 / cap)).sum(), cap - 1) > 1e-3


def test_compiled_reference_replays_its_own_fingerprints(golden):
    """oracle/_ref (the unmodified reference, compiled by oracle/build_ref.py; what bench.py's CPU arm
    runs) plays the seeded episodes recorded in tests/golden from the reference checkout."""
    from oracle import build_ref, refarm
    if not refarm.available() and build_ref.build(quiet=True) is None:
        pytest.skip("no reference checkout and no prebuilt oracle/_ref")
    fp = golden("episodes_ref.npz")["fingerprint"]
    for s in range(5):
        score, steps, mx = refarm.play_seeded(s)
        assert (steps, score, mx) == tuple(int(x) for x in fp[s])
    assert refarm.timed_rollouts(16, 2)[2] == "reference"
