# -*- coding: utf-8 -*-
"""`Game`: a drop-in for the reference's game/GameClient.py `Game` over a batch of one.

Same constructor, `reset`, `step`, `state_matrix` (a list of four lists of tile VALUES that is
mutated in place and returned by every call, as there), `*_space_size` attributes, the same
static helpers, the same accepted action spellings and the same ValueError -- so main.play,
a3c.Worker.work (a3c.py:187-243) and ddpg() (ddpg.py:12-70) run against it unchanged.  Every
transition (slide, merge, spawn placement, game over) is computed by the CUDA kernels through
the C ABI; the host only converts formats.

Random draws come from one of two sources:
  rng="python" (default)  the three draws are taken from Python's global `random` in exactly
                          the reference's order (rand.py:11, GameClient.py:121, :125) and
                          INJECTED into the kernels, so `random.seed(s); Game(); play(...)`
                          replays the reference's episode for seed s move for move;
  rng="philox"            draws are made on the GPU from the Philox stream keyed by
                          (seed, board_id, tick) -- the stream the batched kernels use.
"""
import random

import torch

from . import batched as B

_ACTIONS = (
    ("UP", "Up", "U", "up", "u", 0),
    ("DOWN", "Down", "D", "down", "d", 1),
    ("LEFT", "Left", "L", "left", "l", 2),
    ("RIGHT", "Right", "R", "right", "r", 3),
)


def action_code(action):
    """GameClient.py:140,182,206,230: the accepted spellings; anything else -> ValueError (:254)."""
    for code, names in enumerate(_ACTIONS):
        if action in names:
            return code
    raise ValueError("Input action signal is wrong:\n You must input valid inputs, such as  [U] [D] [L] [R]... ")


def _device(device):
    return B._require_cuda(device if device is not None else "cuda")


def _to_board(matrix, device):
    """list-of-lists of tile values -> int64[1] packed board on the GPU (r48_encode_i32)."""
    vals = torch.as_tensor(matrix, dtype=torch.int64)
    if vals.numel() != 16:
        raise NotImplementedError("only 4x4 boards are supported (16 exponent nibbles per uint64)")
    return B.encode(vals.to(torch.int32).reshape(1, 4, 4).to(device))


def _to_matrix(board):
    """int64[1] packed board -> list of four lists of python ints (r48_decode_i32)."""
    return B.decode(board, dtype=torch.int32)[0].cpu().tolist()


class Game:

    state_matrix, state_space_size = None, 0
    default_device = None          # statics use this (or the current CUDA device)

    def __init__(self, table_matrix_size=4, seed=0, board_id=0, device=None, rng="python"):
        self.reward_space_size = 1
        self.action_space_size = 4
        if table_matrix_size > 4:
            raise NotImplementedError("only 4x4 boards are supported (16 exponent nibbles per uint64)")
        self.state_space_size = 4               # the reference clamps sizes < 4 up to 4
        # the names algorithm/ddpg/agent.py:12-14 reads
        self.state_size, self.action_size, self.reward_size = 4, 4, 1
        if rng not in ("python", "philox"):
            raise ValueError("rng must be 'python' or 'philox'")
        self.rng = rng
        self.device = _device(device)
        self._env = B.BatchedGame(1, seed=seed, device=self.device, board_base=board_id)
        # one launch + one synchronize per move: action, draws and result go through pinned buffers
        self._views = B.HostPlayerViews(self._env.boards)
        self.state_matrix = [[0] * 4 for _ in range(4)]
        self.reset()

    # ------------------------------------------------------------ public, as in the reference

    def _read_view(self):
        v = self._views.views[0]
        self._valid = int(v["valid"])
        self._blanks = v["blanks"].tolist()
        return v["cells"].reshape(4, 4).tolist(), int(v["reward"]), bool(v["done"])

    def reset(self, display=False):
        env = self._env
        if self.rng == "philox":
            env.reset()
            self._views.step(B.ACTION_NONE, 0, 0)                        # read the board out
        else:
            env.boards.zero_()
            env.done.zero_()
            env.steps = 0
            k = random.randint(0, 15)                                   # 16 blanks
            vexp = 1 if random.uniform(0, 1) > 0.1 else 2
            self._views.step(B.ACTION_NONE, k, vexp)
        # the reference REPLACES the list in reset (GameClient.py:34) and mutates it in step
        self.state_matrix = self._read_view()[0]
        if display:
            Game.print_terminal(self.state_matrix)
        return self.state_matrix

    def step(self, action):
        code = action_code(action)
        env = self._env
        if self.rng == "philox":
            env.step([code])
            self._views.step(B.ACTION_NONE, 0, 0)                        # read the board out
            new, _, done = self._read_view()
            reward = int(env.reward.item())
        else:
            k, vexp = 0, 0
            if (self._valid >> code) & 1:                                # has_changed
                k = random.randint(0, self._blanks[code] - 1)            # GameClient.py:121
                vexp = 1 if random.uniform(0, 1) > 0.1 else 2            # GameClient.py:125
            self._views.step(code, k, vexp)
            env.steps += 1
            new, reward, done = self._read_view()
        for i in range(4):                       # same list objects, updated in place
            self.state_matrix[i][:] = new[i]
        return self.state_matrix, reward, done

    # ------------------------------------------------------------ statics, as in the reference

    @staticmethod
    def create_matrix(table_size=4):
        return [[0] * table_size for _ in range(table_size)]

    @staticmethod
    def has_game_over(game_matrix):
        dev = _device(Game.default_device)
        return bool(B.afterstates(_to_board(game_matrix, dev))[3].item())

    @staticmethod
    def has_table_filled(game_matrix):
        dev = _device(Game.default_device)
        return int(B.blank_counts(_to_board(game_matrix, dev)).item()) == 0

    @staticmethod
    def random_fill_grid(game_matrix):
        """Mutates and returns game_matrix; draws from Python's `random` like the reference."""
        dev = _device(Game.default_device)
        board = _to_board(game_matrix, dev)
        n_blank = int(B.blank_counts(board).item())
        if n_blank == 0:
            return game_matrix
        k = random.randint(0, n_blank - 1)
        vexp = 1 if random.uniform(0, 1) > 0.1 else 2
        new = _to_matrix(B.spawn_injected(board, [k], [vexp]))
        for i in range(4):
            game_matrix[i][:] = new[i]
        return game_matrix

    @staticmethod
    def update_matrix(matrix, action):
        """-> (matrix, reward, has_changed); mutates matrix like the reference."""
        code = action_code(action)
        dev = _device(Game.default_device)
        after, reward, valid, _ = B.afterstates(_to_board(matrix, dev))
        new = _to_matrix(after[code])
        for i in range(4):
            matrix[i][:] = new[i]
        return matrix, int(reward[code, 0].item()), bool((int(valid.item()) >> code) & 1)

    @staticmethod
    def print_terminal(matrix):
        width = len(matrix[0])
        rule = "-" * (1 + 7 * width)
        print(rule)
        for row in matrix:
            cells = [(str(int(v)).center(6) if v != 0 else " " * 6) for v in row]
            print("|" + "|".join(cells) + "|")
            print(rule)
