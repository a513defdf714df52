# -*- coding: utf-8 -*-
"""numpy-facing binding of the CPU oracle (oracle/r48_oracle.c).

TEST INFRASTRUCTURE, NOT PRODUCT CODE: importable only from tests/, __graft_entry__.smoke()
and bench.py's CPU-baseline legs.  rein48_b200/ never imports this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libr48_oracle.so")

u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")

STATS_WORDS = 4120
ST_EPISODES, ST_SUM_LEN, ST_SUM_SCORE, ST_SUM_SCORE2, ST_SUM_LEN2 = 0, 1, 2, 3, 4
ST_HIST_MAXEXP, ST_HIST_LEN, ST_HIST_SCORE = 8, 24, 2072


def build(force=False):
    """Compile the oracle with the committed Makefile (gcc only)."""
    src = os.path.join(_HERE, "r48_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(build())
    L.orc_encode.argtypes = [i64p, C.POINTER(C.c_uint64)]
    L.orc_encode.restype = C.c_int
    L.orc_decode.argtypes = [C.c_uint64, i64p]
    L.orc_decode.restype = None
    L.orc_update_matrix.argtypes = [i64p, C.c_int, C.POINTER(C.c_int64)]
    L.orc_update_matrix.restype = C.c_int
    L.orc_has_table_filled.argtypes = [i64p]
    L.orc_has_table_filled.restype = C.c_int
    L.orc_has_game_over.argtypes = [i64p]
    L.orc_has_game_over.restype = C.c_int
    L.orc_random_fill_grid.argtypes = [i64p, C.c_int, C.c_int64]
    L.orc_random_fill_grid.restype = C.c_int
    L.orc_move.argtypes = [C.c_uint64, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int64)]
    L.orc_move.restype = C.c_uint64
    L.orc_game_over.argtypes = [C.c_uint64]
    L.orc_game_over.restype = C.c_int
    L.orc_random_fill_grid_colmajor.argtypes = [i64p, C.c_int, C.c_int64]
    L.orc_random_fill_grid_colmajor.restype = C.c_int
    L.orc_philox4x32.argtypes = [u32p, u32p, C.c_int, u32p]
    L.orc_philox4x32.restype = None
    L.orc_draw.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32]
    L.orc_draw.restype = C.c_uint32
    L.orc_spawn.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32]
    L.orc_spawn.restype = C.c_uint64
    L.orc_episode_records.argtypes = [u64p, u32p, C.c_int64, u32p]
    L.orc_episode_records.restype = None
    L.orc_ring_append.argtypes = [u64p, u8p, i32p, u64p, u8p, u64p, C.c_uint64, u64p, u8p, C.c_void_p, u64p,
                                  C.c_void_p, C.c_int64, C.c_int]
    L.orc_ring_append.restype = None
    L.orc_ring_sample_indices.argtypes = [C.c_uint64, C.c_uint64, C.c_int64, C.c_uint64, C.c_uint64, C.c_int, i64p]
    L.orc_ring_sample_indices.restype = None
    L.orc_reset.argtypes = [C.c_uint64, C.c_uint64]
    L.orc_reset.restype = C.c_uint64
    L.orc_reset_batch.argtypes = [u64p, C.c_int64, C.c_uint64, C.c_uint64]
    L.orc_reset_batch.restype = None
    L.orc_step_batch.argtypes = [u64p, u8p, u64p, i32p, u8p, C.c_int64, C.c_uint64, C.c_uint64,
                                 C.c_uint32, C.c_int]
    L.orc_step_batch.restype = C.c_int
    L.orc_step_injected_batch.argtypes = [u64p, u8p, u8p, u8p, u64p, i32p, u8p, C.c_int64, C.c_int]
    L.orc_step_injected_batch.restype = C.c_int
    L.orc_env_step_batch.argtypes = [u64p, u8p, u32p, u32p, i32p, u8p, u64p, C.c_int64, C.c_uint64,
                                     C.c_uint64, C.c_uint64, C.c_int, C.c_int]
    L.orc_env_step_batch.restype = C.c_int
    L.orc_afterstates_batch.argtypes = [u64p, u64p, i32p, u8p, u8p, C.c_int64, C.c_int]
    L.orc_afterstates_batch.restype = None
    L.orc_decode_batch_f32.argtypes = [u64p, f32p, C.c_int64, C.c_int]
    L.orc_decode_batch_f32.restype = None
    L.orc_decode_batch_i32.argtypes = [u64p, i32p, C.c_int64]
    L.orc_decode_batch_i32.restype = None
    L.orc_play_episode.argtypes = [C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64)]
    L.orc_play_episode.restype = C.c_uint32
    L.orc_rollout.argtypes = [C.c_int64, C.c_uint64, C.c_uint64, u64p, u32p]
    L.orc_rollout.restype = None
    L.orc_rollout_greedy.argtypes = [C.c_int64, C.c_uint64, C.c_uint64, u64p, u32p]
    L.orc_rollout_greedy.restype = None
    L.orc_rollout_trajectories.argtypes = [C.c_int64, C.c_uint64, C.c_uint64, C.c_int, i64p, u64p, u8p,
                                           C.c_int64, u64p]
    L.orc_rollout_trajectories.restype = C.c_int64
    L.orc_rollout_mt.argtypes = [C.c_int64, C.c_uint64, C.c_uint64, u64p, u32p, C.c_int]
    L.orc_rollout_mt.restype = C.c_int
    L.orc_stats_words.argtypes = []
    L.orc_stats_words.restype = C.c_int
    L.orc_score.argtypes = [C.c_uint64]
    L.orc_score.restype = C.c_uint32
    L.orc_max_exp.argtypes = [C.c_uint64]
    L.orc_max_exp.restype = C.c_uint
    L.orc_episode_stats.argtypes = [u64p, u32p, C.c_int64, u64p]
    L.orc_episode_stats.restype = None
    assert L.orc_stats_words() == STATS_WORDS
    _lib = L
    return L


# ---------------------------------------------------------------- matrix-level helpers

def _flat(matrix):
    m = np.ascontiguousarray(np.asarray(matrix, dtype=np.int64).reshape(-1))
    assert m.size == 16, "4x4 boards only"
    return m


def encode(matrix):
    """4x4 tile values -> packed uint64 (raises on non power-of-two tiles)."""
    out = C.c_uint64(0)
    if lib().orc_encode(_flat(matrix), C.byref(out)) != 0:
        raise ValueError("tile value is not 0 or a power of two >= 2")
    return int(out.value)


def decode(board):
    m = np.zeros(16, np.int64)
    lib().orc_decode(int(board), m)
    return m.reshape(4, 4)


def update_matrix(matrix, action):
    """-> (new 4x4 int64 matrix, merge_sum, changed).  Accepts any tile values (incl. 1)."""
    m = _flat(matrix).copy()
    gained = C.c_int64(0)
    changed = lib().orc_update_matrix(m, int(action), C.byref(gained))
    if changed < 0:
        raise ValueError("bad action")
    return m.reshape(4, 4), int(gained.value), bool(changed)


def has_game_over(matrix):
    return bool(lib().orc_has_game_over(_flat(matrix)))


def has_table_filled(matrix):
    return bool(lib().orc_has_table_filled(_flat(matrix)))


def random_fill_grid(matrix, k, value):
    m = _flat(matrix).copy()
    n = lib().orc_random_fill_grid(m, int(k), int(value))
    if n < 0:
        raise ValueError("k out of range")
    return m.reshape(4, 4), n


def move(board, action):
    changed = C.c_int(0)
    gained = C.c_int64(0)
    out = lib().orc_move(int(board), int(action), C.byref(changed), C.byref(gained))
    return int(out), bool(changed.value), int(gained.value)


def game_over(board):
    return bool(lib().orc_game_over(int(board)))


def random_fill_grid_colmajor(matrix, k, value):
    m = _flat(matrix).copy()
    n = lib().orc_random_fill_grid_colmajor(m, int(k), int(value))
    if n < 0:
        raise ValueError("k out of range")
    return m.reshape(4, 4), n


DRAW_ROUNDS = 7
SPAWN4_THRESHOLD = 0x1999999A
VALUE_HASH = 0x9E3779B1


def philox(ctr, key, rounds=DRAW_ROUNDS):
    out = np.zeros(4, np.uint32)
    lib().orc_philox4x32(np.asarray(ctr, np.uint32), np.asarray(key, np.uint32), int(rounds), out)
    return out


def draw(seed, board_id, tick):
    """The tick's Philox word: action = a >> 30, cell from a << 2, value from a * VALUE_HASH."""
    return int(lib().orc_draw(int(seed), int(board_id), int(tick)))


def spawn_batch(boards, seed, board_base, tick):
    L = lib()
    return np.array([L.orc_spawn(int(b), int(seed), int(board_base) + i, int(tick))
                     for i, b in enumerate(boards)], np.uint64)


# ---------------------------------------------------------------- batch forms

def reset_batch(n, seed, board_base=0):
    boards = np.zeros(n, np.uint64)
    lib().orc_reset_batch(boards, n, int(seed), int(board_base))
    return boards


def step_batch(boards, actions, seed, board_base, step, reward_mode=0):
    boards = np.ascontiguousarray(boards, np.uint64)
    actions = np.ascontiguousarray(actions, np.uint8)
    n = boards.size
    out = np.zeros(n, np.uint64)
    reward = np.zeros(n, np.int32)
    done = np.zeros(n, np.uint8)
    rc = lib().orc_step_batch(boards, actions, out, reward, done, n, int(seed), int(board_base),
                              int(step), int(reward_mode))
    if rc != 0:
        raise ValueError("bad action")
    return out, reward, done


def step_injected_batch(boards, actions, spawn_k, spawn_exp, reward_mode=0):
    boards = np.ascontiguousarray(boards, np.uint64)
    n = boards.size
    out = np.zeros(n, np.uint64)
    reward = np.zeros(n, np.int32)
    done = np.zeros(n, np.uint8)
    rc = lib().orc_step_injected_batch(boards, np.ascontiguousarray(actions, np.uint8),
                                       np.ascontiguousarray(spawn_k, np.uint8),
                                       np.ascontiguousarray(spawn_exp, np.uint8),
                                       out, reward, done, n, int(reward_mode))
    if rc != 0:
        raise ValueError("bad action")
    return out, reward, done


def env_step_batch(boards, actions, steps, episodes, seed, board_base, id_stride, reward_mode=0,
                   auto_reset=True):
    """In place on copies: -> (boards, steps, episodes, reward, done, final_boards)"""
    boards = np.array(boards, np.uint64)
    steps = np.array(steps, np.uint32)
    episodes = np.array(episodes, np.uint32)
    n = boards.size
    reward = np.zeros(n, np.int32)
    done = np.zeros(n, np.uint8)
    final = np.zeros(n, np.uint64)
    rc = lib().orc_env_step_batch(boards, np.ascontiguousarray(actions, np.uint8), steps, episodes, reward,
                                  done, final, n, int(seed), int(board_base), int(id_stride),
                                  int(reward_mode), int(bool(auto_reset)))
    if rc != 0:
        raise ValueError("bad action")
    return boards, steps, episodes, reward, done, final


def afterstates_batch(boards, reward_mode=0):
    boards = np.ascontiguousarray(boards, np.uint64)
    n = boards.size
    out = np.zeros((4, n), np.uint64)            # planar: one row per action
    reward = np.zeros((4, n), np.int32)
    valid = np.zeros(n, np.uint8)
    done = np.zeros(n, np.uint8)
    lib().orc_afterstates_batch(boards, out.reshape(-1), reward.reshape(-1), valid, done, n,
                                int(reward_mode))
    return out, reward, valid, done


def decode_batch(boards, dtype="float32", log2_planes=False):
    boards = np.ascontiguousarray(boards, np.uint64)
    n = boards.size
    if dtype == "float32":
        out = np.zeros((n, 4, 4), np.float32)
        lib().orc_decode_batch_f32(boards, out.reshape(-1), n, int(bool(log2_planes)))
    else:
        out = np.zeros((n, 4, 4), np.int32)
        lib().orc_decode_batch_i32(boards, out.reshape(-1), n)
    return out


def rollout(n, seed, board_base=0, threads=1):
    """-> (final_boards u64[n], lengths u32[n])"""
    fb = np.zeros(n, np.uint64)
    ln = np.zeros(n, np.uint32)
    if threads <= 1:
        lib().orc_rollout(n, int(seed), int(board_base), fb, ln)
    else:
        lib().orc_rollout_mt(n, int(seed), int(board_base), fb, ln, int(threads))
    return fb, ln


def rollout_greedy(n, seed, board_base=0):
    fb = np.zeros(n, np.uint64)
    ln = np.zeros(n, np.uint32)
    lib().orc_rollout_greedy(n, int(seed), int(board_base), fb, ln)
    return fb, ln


def rollout_trajectories(n, seed, board_base=0, policy=0, capacity=None):
    """-> (offsets int64[n+1], boards u64[T], actions u8[T], final_boards u64[n])"""
    capacity = capacity or 1200 * n + 4096
    offsets = np.zeros(n + 1, np.int64)
    boards = np.zeros(capacity, np.uint64)
    actions = np.zeros(capacity, np.uint8)
    final = np.zeros(n, np.uint64)
    total = lib().orc_rollout_trajectories(n, int(seed), int(board_base), int(policy), offsets, boards,
                                           actions, capacity, final)
    if total < 0:
        raise MemoryError("trajectory capacity too small")
    return offsets, boards[:total], actions[:total], final


def episode_stats(final_boards, lengths, stats=None):
    if stats is None:
        stats = np.zeros(STATS_WORDS, np.uint64)
    fb = np.ascontiguousarray(final_boards, np.uint64)
    ln = np.ascontiguousarray(lengths, np.uint32)
    lib().orc_episode_stats(fb, ln, fb.size, stats)
    return stats


def scores(final_boards):
    L = lib()
    return np.array([L.orc_score(int(b)) for b in final_boards], np.uint32)


def max_exps(final_boards):
    L = lib()
    return np.array([L.orc_max_exp(int(b)) for b in final_boards], np.uint8)


def episode_records(final_boards, lengths):
    fb = np.ascontiguousarray(final_boards, np.uint64)
    ln = np.ascontiguousarray(lengths, np.uint32)
    out = np.zeros(fb.size, np.uint32)
    lib().orc_episode_records(fb, ln, fb.size, out)
    return out


# ---------------------------------------------------------------- transition ring (Replay)

class Ring:
    """CPU restatement of the device transition ring (algorithm/ddpg/replay.py:8-47)."""

    def __init__(self, capacity):
        self.capacity = int(capacity)
        self.state = np.zeros(capacity, np.uint64)
        self.action = np.zeros(capacity, np.uint8)
        self.reward = np.zeros(capacity, np.int32)
        self.next_state = np.zeros(capacity, np.uint64)
        self.done = np.zeros(capacity, np.uint8)
        self.cursor = np.zeros(1, np.uint64)

    def append(self, state, action, reward, next_state, done, drop_when_full=False):
        state = np.ascontiguousarray(state, np.uint64)
        reward = np.ascontiguousarray(reward, np.int32)
        done = np.ascontiguousarray(done, np.uint8)
        lib().orc_ring_append(self.state, self.action, self.reward, self.next_state, self.done, self.cursor,
                              self.capacity, state, np.ascontiguousarray(action, np.uint8),
                              reward.ctypes.data, np.ascontiguousarray(next_state, np.uint64),
                              done.ctypes.data, state.size, int(bool(drop_when_full)))

    def size(self):
        return int(min(int(self.cursor[0]), self.capacity))

    def sample_indices(self, batch, seed, draw_id, with_replacement=False):
        out = np.zeros(batch, np.int64)
        lib().orc_ring_sample_indices(int(self.cursor[0]), self.capacity, batch, int(seed), int(draw_id),
                                      int(bool(with_replacement)), out)
        return out
