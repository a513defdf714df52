# -*- coding: utf-8 -*-
"""ReplayRing: the reference's replay buffer (algorithm/ddpg/replay.py:8-47) for batches of
transitions, resident on the GPU.

The reference keeps a Python list of [state, action, reward, next_state] (ddpg.py:31), refuses
new items once `max_size` are stored (replay.py:18-21), samples `batch_size` of them without
replacement (random.sample, :33) and empties itself after every sample (:26).  Here the items
are `capacity` 32-byte records (state, next_state, reward, action, done: one DRAM sector each, so
a sampled slot costs one sector) in one device array, filled by `r48_ring_append` -- or directly
by the env-step kernel, `BatchedGame.env_step(ring=...)` -- and gathered by `r48_ring_sample`
into separate batch tensors; nothing crosses to the host.

  mode="ring"       overwrite the oldest slot when full, sampling does not clear  (the usual
                    DQN replay; what the fused env-step append does)
  mode="reference"  replay.py's rules: store drops when full, sample() clears

`next_state` is a value of its own: ddpg.py:29-31 stores the same mutated list object as state
and next_state, an aliasing bug that is deliberately not reproduced.
"""
import ctypes as C

import torch

from . import _native
from .batched import _require_cuda, _stream, as_actions, from_any, _i64

MINI_BATCH_SIZE = 10          # replay.py:5


class ReplayRing:

    def __init__(self, capacity=100, device="cuda", mode="ring", seed=0):
        if mode not in ("ring", "reference"):
            raise ValueError("mode must be 'ring' or 'reference'")
        if int(capacity) < 1:
            raise ValueError("capacity must be positive")
        self.device = _require_cuda(device)
        self.max_size = self.capacity = int(capacity)          # replay.py:11
        self.mode = mode
        self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self.draws = 0                 # sample() calls so far: keys the next sample
        self.appended = 0              # host mirror of the device cursor (no sync needed to know the size)
        c = self.capacity
        with torch.cuda.device(self.device):
            # [capacity][4] int64 = r48_transition[capacity]: state, next_state, reward|action<<32|done<<40, 0
            self.slots = torch.zeros((c, 4), dtype=torch.int64, device=self.device)
            self.cursor = torch.zeros(2, dtype=torch.int64, device=self.device)
        if self.slots.data_ptr() % 32:
            raise RuntimeError("ring storage is not 32-byte aligned")
        self._struct = _native.Ring(self.slots.data_ptr(), self.cursor.data_ptr(), c)
        self._lib = _native.lib()

    def _ref(self):
        return C.byref(self._struct)

    # the stored fields as tensors of `capacity` elements (views / unpacked copies of the records;
    # for inspection and tests -- learners take batches from sample())
    @property
    def state(self):
        return self.slots[:, 0].contiguous()

    @property
    def next_state(self):
        return self.slots[:, 1].contiguous()

    @property
    def reward(self):
        return (self.slots[:, 2] & 0xFFFFFFFF).to(torch.int32)

    @property
    def action(self):
        return ((self.slots[:, 2] >> 32) & 0xFF).to(torch.uint8)

    @property
    def done(self):
        return ((self.slots[:, 2] >> 40) & 0xFF).to(torch.uint8)

    def _appended(self, n):
        if self.mode == "reference":
            self.appended = min(self.capacity, self.appended + n)
        else:
            self.appended += n

    # -- replay.py:14-16
    @property
    def cur_size(self):
        return min(self.appended, self.capacity)

    def __len__(self):
        return self.cur_size

    def filled(self):
        return self.max_size <= self.cur_size

    # -- replay.py:18-21, for a batch
    def store(self, state, action, reward, next_state, done=None):
        """Append n transitions: state / next_state int64[n] packed boards, action uint8[n],
        reward int32[n] (None = the reference's constant 0), done uint8[n] (optional).  Accepts
        torch CUDA tensors or DLPack producers, zero-copy."""
        state, next_state = _i64(state), _i64(next_state)
        n = state.numel()
        if next_state.numel() != n:
            raise ValueError("state and next_state differ in length")
        action = as_actions(action, self.device)
        if reward is not None:
            reward = from_any(reward)
            reward = reward.to(self.device, torch.int32).contiguous()
        if done is not None:
            done = as_actions(done, self.device)
        for t in (action, reward, done):
            if t is not None and t.numel() != n:
                raise ValueError("all transition arrays must have %d elements" % n)
        with torch.cuda.device(self.device):
            _native.check(self._lib.r48_ring_append(
                self._ref(), state.data_ptr(), action.data_ptr(), reward.data_ptr() if reward is not None else None,
                next_state.data_ptr(), done.data_ptr() if done is not None else None, n,
                int(self.mode == "reference"), _stream(self.device)))
        self._appended(n)

    add = store

    # -- replay.py:23-27
    def sample(self, batch_size=MINI_BATCH_SIZE, replace=False, obs=False, log2=False):
        """-> dict with the reference's keys ('state', 'action', 'reward', 'next_state') plus 'done'
        and 'index'.  Without replacement a buffer holding fewer than batch_size items returns all
        of them (replay.py:30-31).  obs=True returns the boards as float32 [B,4,4] tile values
        (np.array of the state lists, as the learners take them) instead of packed words.  In
        mode='reference' the buffer is cleared afterwards (replay.py:26)."""
        b = int(batch_size)
        count = b if (replace and self.cur_size > 0) else min(b, self.cur_size)
        dev = self.device
        with torch.cuda.device(dev):
            out = {
                "state": torch.empty(b, dtype=torch.int64, device=dev),
                "action": torch.empty(b, dtype=torch.uint8, device=dev),
                "reward": torch.empty(b, dtype=torch.int32, device=dev),
                "next_state": torch.empty(b, dtype=torch.int64, device=dev),
                "done": torch.empty(b, dtype=torch.uint8, device=dev),
                "index": torch.empty(b, dtype=torch.int64, device=dev),
            }
            so = torch.empty((b, 4, 4), dtype=torch.float32, device=dev) if obs else None
            no = torch.empty((b, 4, 4), dtype=torch.float32, device=dev) if obs else None
            if b:
                _native.check(self._lib.r48_ring_sample(
                    self._ref(), b, self.seed, self.draws, int(bool(replace)), out["index"].data_ptr(),
                    out["state"].data_ptr(), out["action"].data_ptr(), out["reward"].data_ptr(),
                    out["next_state"].data_ptr(), out["done"].data_ptr(),
                    so.data_ptr() if obs else None, no.data_ptr() if obs else None, int(bool(log2)), _stream(dev)))
        self.draws += 1
        if obs:
            out["state_packed"], out["next_state_packed"] = out["state"], out["next_state"]
            out["state"], out["next_state"] = so, no
        out = {k: v[:count] for k, v in out.items()}
        if self.mode == "reference":
            self.clear()
        return out

    # -- replay.py:45-47
    def clear(self):
        with torch.cuda.device(self.device):
            _native.check(self._lib.r48_ring_clear(self._ref(), _stream(self.device)))
        self.appended = 0
