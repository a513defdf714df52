# -*- coding: utf-8 -*-
"""The reference's own unit tests (game/GameClientTest.py: TestGameClient) re-run against the
GPU-backed drop-in `rein48_b200.Game`, same test names and same structure.  The input lines and
expected outputs are not retyped: they come from tests/golden/testvectors_ref.npz, which was
produced by replaying the reference's vectors through the unmodified reference.  The reference
uses tile "1" in several vectors, which is not a 2048 tile; those lines are doubled (merging only
tests equality, so doubling every tile doubles every result)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def Game():
    import rein48_b200
    return rein48_b200.Game


@pytest.fixture(scope="module")
def vectors(golden):
    return golden("testvectors_ref.npz")


def column(cells):
    m = [[0] * 4 for _ in range(4)]
    for t, c in enumerate(cells):
        m[t][0] = int(c)
    return m


def row(cells):
    m = [[0] * 4 for _ in range(4)]
    m[0] = [int(c) for c in cells]
    return m


class TestGameClient:

    def test_has_matrix_filled(self, Game, vectors):           # GameClientTest.py:10-21
        for board, want in zip(vectors["filled_boards"], vectors["filled"]):
            assert Game.has_table_filled((2 * board).tolist()) == bool(want)

    def test_is_game_over(self, Game, vectors):                # GameClientTest.py:23-31
        for board, want in zip(vectors["over_boards"], vectors["over"]):
            assert Game.has_game_over(board.tolist()) == bool(want)

    def test_random_fill_grid(self, Game):                     # GameClientTest.py:33-44
        for m in ([[0] * 4 for _ in range(4)],
                  [[0, 0, 0, 0], [0, 2, 0, 0], [0, 0, 0, 0], [0, 0, 0, 0]],
                  [[2, 4, 2, 4], [4, 2, 4, 2], [2, 4, 2, 4], [4, 2, 4, 0]]):
            before = np.sum(m)
            assert np.sum(Game.random_fill_grid(m)) - before in (2, 4)     # the reference's `== 2 or 4`
        full = [[2, 4, 2, 4], [4, 2, 4, 2], [2, 4, 2, 4], [4, 2, 4, 2]]
        assert np.sum(Game.random_fill_grid([r[:] for r in full])) - np.sum(full) == 0

    def test_random_action(self):                              # GameClientTest.py:46-47
        from rein48_b200 import Rand
        assert Rand.random_action() in ("UP", "DOWN", "LEFT", "RIGHT")

    def _check(self, Game, vectors, action_index, spelling, embed, extract):
        for cells, want in zip(vectors["lines"], vectors["moved"][action_index]):
            actual_matrix, reward, _ = Game.update_matrix(embed(2 * cells), spelling)   # 3-tuple, as at :124
            assert extract(actual_matrix) == [int(2 * w) for w in want]
            assert reward == 0

    def test_update_matrix_up(self, Game, vectors):            # GameClientTest.py:49-118
        self._check(Game, vectors, 0, "U", column, lambda m: [m[t][0] for t in range(4)])

    def test_update_matrix_down(self, Game, vectors):          # GameClientTest.py:120-189
        self._check(Game, vectors, 1, "D", column, lambda m: [m[t][0] for t in range(4)])

    def test_update_matrix_left(self, Game, vectors):          # GameClientTest.py:191-260
        self._check(Game, vectors, 2, "L", row, lambda m: m[0])

    def test_update_matrix_right(self, Game, vectors):         # GameClientTest.py:262-331
        self._check(Game, vectors, 3, "R", row, lambda m: m[0])

    def test_print_terminal(self, Game, capsys):               # GameClientTest.py:333-335
        Game.print_terminal([[2, 0, 0, 0], [0, 4, 0, 0], [0, 0, 8, 0], [0, 0, 0, 16]])
        out = capsys.readouterr().out
        assert out.count("|") == 20 and " 16 " in out
