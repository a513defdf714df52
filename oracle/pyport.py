# -*- coding: utf-8 -*-
"""Pure-Python restatement of the reference environment and random-policy loop.

TEST INFRASTRUCTURE, NOT PRODUCT CODE (same rule as r48_oracle.c): only tests/,
__graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.

Why it exists next to the C oracle: the reference is a Python program and cannot travel
to the GPU box (/root/reference is not there, and its sources must not be copied), so the
"reference's own CPU path" that bench.py times beside the GPU numbers is this port.  It
keeps the reference's data structures and cost profile on purpose -- a list of four lists
of tile VALUES, a deep copy per move for the has_changed test (GameClient.py:137,180),
CPython's global ``random`` for the three draws (rand.py:11, GameClient.py:121,125) -- and
consumes the MT19937 stream in exactly the reference's order, so that
``random.seed(s); PortGame(); play(...)`` reproduces the reference's episode for seed s
(pinned by tests/golden/episodes_ref.npz, generated from the unmodified reference).

Reference map:
    PortGame.reset / step        game/GameClient.py:33-38 / 40-51
    slide                        game/GameClient.py:129-254
    spawn                        game/GameClient.py:102-127
    filled / game_over           game/GameClient.py:96-100 / 65-94
    random_action                control/rand.py:9-11
    play                         main.py:36-42,48
"""
import copy
import random

SIZE = 4
_NAMES = (
    ("UP", "Up", "U", "up", "u", 0),
    ("DOWN", "Down", "D", "down", "d", 1),
    ("LEFT", "Left", "L", "left", "l", 2),
    ("RIGHT", "Right", "R", "right", "r", 3),
)
ACTION_NAMES = ("UP", "DOWN", "LEFT", "RIGHT")


def action_index(action):
    """Map the spellings GameClient.py:140,182,206,230 accept to 0..3; else ValueError (:254)."""
    for code, names in enumerate(_NAMES):
        if action in names:      # == semantics, so 0.0 / np.int64(0) / True behave as there
            return code
    raise ValueError("Input action signal is wrong")


def _cells(code, line):
    """Coordinates of one line, first element = the wall the tiles slide toward."""
    if code == 0:
        return [(r, line) for r in range(SIZE)]
    if code == 1:
        return [(r, line) for r in range(SIZE - 1, -1, -1)]
    if code == 2:
        return [(line, c) for c in range(SIZE)]
    return [(line, c) for c in range(SIZE - 1, -1, -1)]


def slide(grid, action):
    """update_matrix: returns (grid, 0, has_changed); mutates grid like the reference."""
    code = action_index(action)
    snapshot = copy.deepcopy(grid)
    for line in range(SIZE):
        where = _cells(code, line)
        dst, src = 0, 1
        while src < SIZE:
            while src < SIZE and grid[where[src][0]][where[src][1]] == 0:
                src += 1
            if src == SIZE:
                break
            dr, dc = where[dst]
            sr, sc = where[src]
            if grid[dr][dc] == 0:
                grid[dr][dc] += grid[sr][sc]
                grid[sr][sc] = 0
            elif grid[dr][dc] == grid[sr][sc]:
                grid[dr][dc] += grid[sr][sc]
                grid[sr][sc] = 0
                dst += 1
            else:
                if dst + 1 != src:
                    nr, nc = where[dst + 1]
                    grid[nr][nc] += grid[sr][sc]
                    grid[sr][sc] = 0
                dst += 1
            src += 1
    return grid, 0, snapshot != grid


def spawn(grid):
    """random_fill_grid: k-th blank in row-major order, then 2 w.p. 0.9 / 4 w.p. 0.1."""
    blanks = [(r, c) for r in range(SIZE) for c in range(SIZE) if grid[r][c] == 0]
    if not blanks:
        return grid
    r, c = blanks[random.randint(0, len(blanks) - 1)]
    grid[r][c] = 2 if random.uniform(0, 1) > 0.1 else 4
    return grid


def filled(grid):
    return all(v != 0 for row in grid for v in row)


def game_over(grid):
    if not filled(grid):
        return False
    for r in range(SIZE):
        for c in range(SIZE):
            if r + 1 < SIZE and grid[r][c] == grid[r + 1][c]:
                return False
            if c + 1 < SIZE and grid[r][c] == grid[r][c + 1]:
                return False
    return True


class PortGame:
    """4x4 only; same attribute names as the reference's Game."""

    def __init__(self):
        self.reward_space_size = 1
        self.action_space_size = 4
        self.state_space_size = SIZE
        self.state_matrix = None
        self.reset()

    def reset(self):
        self.state_matrix = spawn([[0] * SIZE for _ in range(SIZE)])
        return self.state_matrix

    def step(self, action):
        self.state_matrix, reward, moved = slide(self.state_matrix, action)
        if moved:
            self.state_matrix = spawn(self.state_matrix)
        return self.state_matrix, reward, game_over(self.state_matrix)


def random_action(*_):
    return ACTION_NAMES[random.randint(0, 3)]


def play(game):
    """The main.py:36-42 loop with printing off.  Returns (score, steps)."""
    steps = 0
    over = False
    while not over:
        _, _, over = game.step(random_action(game.state_matrix))
        steps += 1
    return sum(sum(row) for row in game.state_matrix), steps


def play_seeded(seed):
    """One config-1 episode: random.seed(s); Game(); play.  Returns (score, steps, max_tile)."""
    random.seed(seed)
    game = PortGame()
    score, steps = play(game)
    return score, steps, max(max(row) for row in game.state_matrix)


def _worker(seed_range):
    lo, hi = seed_range
    steps = 0
    for s in range(lo, hi):
        steps += play_seeded(s)[1]
    return steps


def timed_rollouts(episodes, processes):
    """CPU baseline: `episodes` seeded games over a multiprocessing pool.
    Returns (total_steps, seconds)."""
    import multiprocessing as mp
    import time
    chunk = max(1, episodes // (processes * 8))
    ranges = [(lo, min(lo + chunk, episodes)) for lo in range(0, episodes, chunk)]
    ctx = mp.get_context("fork")
    with ctx.Pool(processes) as pool:
        pool.map(_worker, [(0, 1)] * processes)  # spin the workers up outside the clock
        t0 = time.perf_counter()
        total = sum(pool.map(_worker, ranges))
        dt = time.perf_counter() - t0
    return total, dt
