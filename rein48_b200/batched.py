# -*- coding: utf-8 -*-
"""BatchedGame: n 2048 boards resident on one B200, stepped by libr48's CUDA kernels.

Mirrors the surface of the reference's `Game` (game/GameClient.py:15-51) for a whole batch:
`reset()`, `step(actions)`, `state_matrix()` readout, `*_space_size` attributes.  Tensors go
in and out zero-copy (data_ptr on the caller's CUDA stream); nothing here computes a
transition on the host, and there is no fallback if the CUDA library is unavailable.
"""
import torch

from . import _native

REWARD_MODES = {"reference": 0, "merge_sum": 1, 0: 0, 1: 1}


def _stream(device):
    return torch.cuda.current_stream(device).cuda_stream


def _require_cuda(device):
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("rein48_b200 runs on CUDA devices only (got %s); there is no CPU path" % device)
    if not torch.cuda.is_available():
        raise RuntimeError("rein48_b200 needs a CUDA device (B200, sm_100a); none is visible")
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    return device


def from_any(x):
    """A torch tensor for `x` WITHOUT copying when x is a tensor, an object that speaks DLPack
    (`__dlpack__`: cupy / jax / numba arrays, other frameworks' tensors) or a DLPack capsule
    (torch.utils.dlpack.to_dlpack); anything else is returned as it came."""
    if torch.is_tensor(x):
        return x
    if hasattr(x, "__dlpack__") or type(x).__name__ == "PyCapsule":
        return torch.utils.dlpack.from_dlpack(x)
    return x


def as_actions(actions, device):
    """Any integer tensor/sequence -> uint8 device tensor (values are checked on the GPU)."""
    actions = from_any(actions)
    if not torch.is_tensor(actions):
        actions = torch.as_tensor(actions)
    if actions.dtype != torch.uint8:
        # clamp so that e.g. 256 or -1 cannot wrap into a legal action byte
        actions = actions.to(torch.int64).clamp(-1, 255).to(torch.uint8)
    return actions.to(device, non_blocking=True).contiguous()


class BatchedGame:
    """n independent 4x4 games.  Boards are int64 tensors holding the uint64 bit pattern
    (16 exponent nibbles, cell (i,j) = nibble 4*i+j).

    seed / board_base key the Philox stream: board i of this batch is global board
    board_base + i, so shards of one logical batch on different GPUs draw disjoint streams.
    id_stride = the size of the WHOLE logical batch (all shards): env_step() gives episode e of
    env i the id board_base + i + e * id_stride, so it must be the same on every shard and at
    least the world batch, or shards replay each other's streams one episode apart.  It
    defaults to n for an unsharded batch (board_base == 0).

    Every reset() after the first starts a new EPOCH: the Philox key becomes
    seed + epoch * 0x9E3779B97F4A7C15 (mod 2^64), so a loop that calls reset() per episode (the
    a3c / ddpg workers do) sees fresh games each time, reproducibly from (seed, epoch).
    """

    state_space_size = 4          # GameClient.py:21-27
    action_space_size = 4
    reward_space_size = 1
    # names algorithm/ddpg/agent.py:12-14 looks for
    state_size = 4
    action_size = 4
    reward_size = 1

    EPOCH_KEY_STEP = 0x9E3779B97F4A7C15

    def __init__(self, n, seed=0, device="cuda", board_base=0, reward_mode="reference", id_stride=None):
        self.device = _require_cuda(device)
        self.n = int(n)
        self.seed0 = int(seed) & 0xFFFFFFFFFFFFFFFF
        self.seed = self.seed0
        self.epoch = -1
        self.board_base = int(board_base)
        if id_stride is not None and int(id_stride) < self.n:
            raise ValueError("id_stride must be at least the batch size")
        self.id_stride = int(id_stride) if id_stride is not None else (self.n if self.board_base == 0 else None)
        self.reward_mode = REWARD_MODES[reward_mode]
        self._lib = _native.lib()
        with torch.cuda.device(self.device):
            self.boards = torch.zeros(self.n, dtype=torch.int64, device=self.device)
            self.reward = torch.zeros(self.n, dtype=torch.int32, device=self.device)
            self.done = torch.zeros(self.n, dtype=torch.uint8, device=self.device)
            self.status = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.steps = 0
        # the four state tensors live as long as the object: their addresses are fetched once
        self._ptrs = (self.boards.data_ptr(), self.reward.data_ptr(), self.done.data_ptr(), self.status.data_ptr())
        self.reset()

    # -- Game.reset (GameClient.py:33-38): one tile per board
    def reset(self, epoch=None):
        """New games on every board.  `epoch` (default: the next one) selects the draw stream."""
        self.epoch = self.epoch + 1 if epoch is None else int(epoch)
        self.seed = (self.seed0 + self.epoch * self.EPOCH_KEY_STEP) & 0xFFFFFFFFFFFFFFFF
        with torch.cuda.device(self.device):
            _native.check(self._lib.r48_reset(self.boards.data_ptr(), self.n, self.seed,
                                              self.board_base, _stream(self.device)))
            self.done.zero_()
            if hasattr(self, "env_steps"):            # the per-env counters of env_step() restart too
                self.env_steps.zero_()
                self.env_episodes.zero_()
        self.steps = 0
        return self.boards

    # -- Game.step (GameClient.py:40-51)
    def step(self, actions):
        """actions: integer tensor [n] in 0..3 (UP, DOWN, LEFT, RIGHT).
        Returns (boards, reward, done) -- the same tensors every call (updated in place), as
        the reference returns the same list object every call."""
        # fast path: a uint8 tensor already on this device goes straight to the C call (the
        # Python side of a call is ~5 us; the kernel for 2^20 boards is ~11 us)
        if not (torch.is_tensor(actions) and actions.dtype == torch.uint8 and actions.device == self.device
                and actions.is_contiguous()):
            actions = as_actions(actions, self.device)      # DLPack producers come in here, zero-copy
        if actions.numel() != self.n:
            raise ValueError("expected %d actions, got %d" % (self.n, actions.numel()))
        if torch.cuda.current_device() != self.device.index:
            with torch.cuda.device(self.device):
                return self.step(actions)
        p = self._ptrs
        rc = self._lib.r48_step(p[0], actions.data_ptr(), p[0], p[1], p[2], self.n, self.seed, self.board_base,
                                self.steps, self.reward_mode, p[3], torch.cuda.current_stream().cuda_stream)
        if rc:
            _native.check(rc)
        self.steps += 1
        return self.boards, self.reward, self.done

    def step_injected(self, actions, spawn_k, spawn_exp):
        """step() with the spawn draws supplied (parity mode, see r48_step_injected)."""
        a = as_actions(actions, self.device)
        k = as_actions(spawn_k, self.device)
        v = as_actions(spawn_exp, self.device)
        with torch.cuda.device(self.device):
            _native.check(self._lib.r48_step_injected(
                self.boards.data_ptr(), a.data_ptr(), k.data_ptr(), v.data_ptr(), self.boards.data_ptr(),
                self.reward.data_ptr(), self.done.data_ptr(), self.n, self.reward_mode,
                self.status.data_ptr(), _stream(self.device)))
        self.steps += 1
        return self.boards, self.reward, self.done

    # -- vectorised-env form: per-env counters, auto-reset, fused observation
    def env_step(self, actions, auto_reset=True, obs=True, log2=False, id_stride=None, ring=None):
        """Game.step for every env with its OWN tick/episode counters; finished games are reset
        in place when auto_reset (done[i] = 1 then comes with the new episode's first board,
        the finished board is kept in `self.final_boards`).  Returns (obs or boards, reward,
        done).  `ring` (a ReplayRing) receives every env's transition in the same kernel.
        Do not mix with step() on the same object: that one keys all boards by one shared step
        counter."""
        a = as_actions(actions, self.device)
        if a.numel() != self.n:
            raise ValueError("expected %d actions, got %d" % (self.n, a.numel()))
        stride = id_stride if id_stride is not None else self.id_stride
        if stride is None:
            raise ValueError("this batch is a shard (board_base != 0): pass id_stride = the size of the whole "
                             "logical batch to the constructor, so that shards do not reuse each other's ids")
        if ring is not None:
            if ring.device != self.device:
                raise ValueError("the ring lives on %s, the environments on %s" % (ring.device, self.device))
            if ring.mode != "ring":
                raise ValueError("the fused append overwrites the oldest slot; a mode='reference' ring takes "
                                 "transitions through ReplayRing.store")
        with torch.cuda.device(self.device):
            if not hasattr(self, "env_steps"):
                self.env_steps = torch.zeros(self.n, dtype=torch.int32, device=self.device)
                self.env_episodes = torch.zeros(self.n, dtype=torch.int32, device=self.device)
                self.final_boards = torch.zeros(self.n, dtype=torch.int64, device=self.device)
                self.obs = torch.empty((self.n, 4, 4), dtype=torch.float32, device=self.device)
            _native.check(self._lib.r48_env_step_ring(
                self.boards.data_ptr(), a.data_ptr(), self.env_steps.data_ptr(), self.env_episodes.data_ptr(),
                self.reward.data_ptr(), self.done.data_ptr(), self.obs.data_ptr() if obs else None,
                int(bool(log2)), self.final_boards.data_ptr(), self.n, self.seed, self.board_base,
                int(stride), self.reward_mode, int(bool(auto_reset)),
                self.status.data_ptr(), ring._ref() if ring is not None else None, _stream(self.device)))
            if ring is not None:
                ring._appended(self.n)
        return (self.obs if obs else self.boards), self.reward, self.done

    def check_actions(self):
        """Synchronising check that no step() so far saw an action outside 0..3; raises the
        ValueError the reference raises at GameClient.py:254."""
        if int(self.status.item()) & 1:
            self.status.zero_()
            raise ValueError("Input action signal is wrong: actions must be 0 (UP), 1 (DOWN), 2 (LEFT), 3 (RIGHT)")

    # -- 1-ply expansion (Game.update_matrix x 4 + has_game_over)
    def afterstates(self):
        return afterstates(self.boards, self.reward_mode)

    # -- readout
    def state_matrix(self, dtype=torch.float32, log2=False):
        """[n, 4, 4] tile values (what np.array(state) gives the learners), or exponents."""
        return decode(self.boards, dtype=dtype, log2=log2)

    def scores(self):
        return scores(self.boards)


# ---------------------------------------------------------------------- functional forms

def _i64(boards):
    boards = from_any(boards)
    if not torch.is_tensor(boards):
        raise TypeError("boards must be a CUDA tensor (or a DLPack producer) of int64/uint64 bit patterns")
    if boards.dtype not in (torch.int64, torch.uint64):
        raise TypeError("boards must be int64/uint64 bit patterns")
    if not boards.is_cuda:
        raise RuntimeError("boards must live on a CUDA device; there is no CPU path")
    return boards.contiguous()


def afterstates(boards, reward_mode=0):
    """-> (after [4,n] int64, reward [4,n] int32, valid [n] uint8 bitmask, done [n] uint8);
    after[a] is the contiguous batch of boards after action a (0 UP, 1 DOWN, 2 LEFT, 3 RIGHT)."""
    boards = _i64(boards)
    n = boards.numel()
    dev = boards.device
    with torch.cuda.device(dev):
        out = torch.empty((4, n), dtype=torch.int64, device=dev)
        reward = torch.empty((4, n), dtype=torch.int32, device=dev)
        valid = torch.empty(n, dtype=torch.uint8, device=dev)
        done = torch.empty(n, dtype=torch.uint8, device=dev)
        _native.check(_native.lib().r48_afterstates(
            boards.data_ptr(), out.data_ptr(), reward.data_ptr(), valid.data_ptr(), done.data_ptr(), n,
            REWARD_MODES[reward_mode], _stream(dev)))
    return out, reward, valid, done


def decode(boards, dtype=torch.float32, log2=False):
    boards = _i64(boards)
    n = boards.numel()
    dev = boards.device
    with torch.cuda.device(dev):
        if dtype == torch.float32:
            out = torch.empty((n, 4, 4), dtype=torch.float32, device=dev)
            _native.check(_native.lib().r48_decode_f32(boards.data_ptr(), out.data_ptr(), n, int(bool(log2)),
                                                      _stream(dev)))
        elif dtype == torch.int32:
            if log2:
                raise ValueError("log2 planes are float32 only")
            out = torch.empty((n, 4, 4), dtype=torch.int32, device=dev)
            _native.check(_native.lib().r48_decode_i32(boards.data_ptr(), out.data_ptr(), n, _stream(dev)))
        else:
            raise TypeError("dtype must be torch.float32 or torch.int32")
    return out


def encode(values):
    """[n,4,4] int32 tile values on the GPU -> packed boards; ValueError on illegal tiles."""
    values = from_any(values)
    if not values.is_cuda:
        raise RuntimeError("values must live on a CUDA device; there is no CPU path")
    values = values.to(torch.int32).contiguous()
    n = values.numel() // 16
    dev = values.device
    with torch.cuda.device(dev):
        out = torch.empty(n, dtype=torch.int64, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        _native.check(_native.lib().r48_encode_i32(values.data_ptr(), out.data_ptr(), n, status.data_ptr(),
                                                  _stream(dev)))
        if int(status.item()) & 2:
            raise ValueError("tile values must be 0 or a power of two in 2..32768")
    return out


def scores(boards):
    """-> (score [n] int32 = sum of tiles, max_exp [n] uint8)"""
    boards = _i64(boards)
    n = boards.numel()
    dev = boards.device
    with torch.cuda.device(dev):
        sc = torch.empty(n, dtype=torch.int32, device=dev)
        mx = torch.empty(n, dtype=torch.uint8, device=dev)
        _native.check(_native.lib().r48_scores(boards.data_ptr(), sc.data_ptr(), mx.data_ptr(), n, _stream(dev)))
    return sc, mx


def blank_counts(boards):
    boards = _i64(boards)
    n = boards.numel()
    dev = boards.device
    with torch.cuda.device(dev):
        out = torch.empty(n, dtype=torch.uint8, device=dev)
        _native.check(_native.lib().r48_blank_counts(boards.data_ptr(), out.data_ptr(), n, _stream(dev)))
    return out


def spawn_injected(boards, spawn_k, spawn_exp):
    """In place: put exponent spawn_exp[i] into the spawn_k[i]-th blank of boards[i]."""
    boards = _i64(boards)
    dev = boards.device
    k = as_actions(spawn_k, dev)
    v = as_actions(spawn_exp, dev)
    with torch.cuda.device(dev):
        _native.check(_native.lib().r48_spawn_injected(boards.data_ptr(), k.data_ptr(), v.data_ptr(),
                                                      boards.numel(), _stream(dev)))
    return boards


ACTION_NONE = _native.ACTION_NONE          # r48_step_injected_view: no move, spawn only
VIEW_DTYPE = None


def _view_dtype():
    """numpy layout of struct r48_game_view (include/r48.h)."""
    global VIEW_DTYPE
    if VIEW_DTYPE is None:
        import numpy as np
        VIEW_DTYPE = np.dtype([("cells", "<i4", (16,)), ("reward", "<i4"), ("done", "u1"), ("valid", "u1"),
                               ("blanks", "u1", (4,)), ("reserved", "u1", (2,))])
        assert VIEW_DTYPE.itemsize == 76
    return VIEW_DTYPE


class HostPlayerViews:
    """r48_step_injected_view for a player whose policy and random draws live on the host (the
    reference's main.play over a `Game`): the action, the two injected draws and the result travel
    through PINNED host buffers the kernel reads and writes directly, so a move is one launch and
    one stream synchronize -- no copies, no allocation.

        views = HostPlayerViews(boards)          # boards: int64[n] on the GPU, stepped in place
        v = views.step(actions, spawn_k, spawn_exp)     # numpy record array [n]: cells, reward, done,
                                                        # valid (bit a: action a changes the new board),
                                                        # blanks[a] (blank cells after action a)
    """

    def __init__(self, boards, reward_mode=0):
        import numpy as np
        self.boards = _i64(boards)
        self.device = self.boards.device
        self.n = self.boards.numel()
        self.reward_mode = int(reward_mode)
        self._in = torch.zeros((3, self.n), dtype=torch.uint8).pin_memory()
        self._out = torch.zeros(self.n * _view_dtype().itemsize, dtype=torch.uint8).pin_memory()
        self._status = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.inputs = self._in.numpy()                     # [0] action, [1] spawn_k, [2] spawn_exp
        self.views = self._out.numpy().view(_view_dtype())
        base = self._in.data_ptr()
        self._args = (self.boards.data_ptr(), base, base + self.n, base + 2 * self.n, self.n, self.reward_mode,
                      self._out.data_ptr(), self._status.data_ptr())
        self._call = _native.lib().r48_step_injected_view
        self._index = self.device.index if self.device.index is not None else torch.cuda.current_device()

    def step(self, actions, spawn_k, spawn_exp):
        """Apply Game.step with injected draws to every board (action 255 = spawn only) and return
        the views; the returned array is the pinned buffer itself, valid until the next call."""
        self.inputs[0] = actions
        self.inputs[1] = spawn_k
        self.inputs[2] = spawn_exp
        if torch.cuda.current_device() != self._index:
            with torch.cuda.device(self.device):
                return self._launch()
        return self._launch()

    def _launch(self):
        stream = torch.cuda.current_stream(self.device)
        _native.check(self._call(*self._args, stream.cuda_stream))
        stream.synchronize()
        return self.views

    def illegal_action_seen(self):
        """True if any call since the last check carried an action byte other than 0..3 / 255."""
        bad = bool(self._status.item())
        if bad:
            self._status.zero_()
        return bad


def spawn(boards, seed, board_base=0, tick=0):
    """In place: Game.random_fill_grid with the GPU's Philox draws of (seed, board, tick)."""
    boards = _i64(boards)
    dev = boards.device
    with torch.cuda.device(dev):
        _native.check(_native.lib().r48_spawn(boards.data_ptr(), boards.numel(), int(seed) & 0xFFFFFFFFFFFFFFFF,
                                             int(board_base), int(tick), _stream(dev)))
    return boards
