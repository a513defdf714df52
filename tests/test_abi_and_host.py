# -*- coding: utf-8 -*-
"""CPU-side checks: the C-ABI library loads and exports exactly what include/r48.h declares,
argument validation works without a device, host-side logic (sharding, statistics view,
action spellings) and the world-size-2 statistics all-reduce over gloo."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def native():
    from rein48_b200 import _native
    _native.build()
    return _native


def header_symbols():
    text = open(os.path.join(ROOT, "include", "r48.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(r48_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(native):
    declared = header_symbols()
    assert declared == sorted(native.SYMBOLS)
    L = native.lib()
    for name in declared:
        assert hasattr(L, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", native.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r"\bT (r48_[a-z0-9_]+)", out)))
    assert exported == declared
    assert L.r48_version() == 201 == native.VERSION
    assert native.built_id() == native.source_id() and L.r48_build_id().decode().endswith(native.source_id())


def test_header_constants_match_python_and_oracle(native):
    text = open(os.path.join(ROOT, "include", "r48.h")).read()
    consts = {k: int(v) for k, v in re.findall(r"#define (R48_[A-Z0-9_]+)\s+(-?\d+)\b", text)}
    from rein48_b200 import stats
    from oracle import oracle as orc
    assert consts["R48_STATS_WORDS"] == stats.STATS_WORDS == orc.STATS_WORDS == native.STATS_WORDS
    assert consts["R48_STATS_HIST_MAXEXP"] == stats.HIST_MAXEXP == orc.ST_HIST_MAXEXP
    assert consts["R48_STATS_HIST_LEN"] == stats.HIST_LEN == orc.ST_HIST_LEN
    assert consts["R48_STATS_HIST_SCORE"] == stats.HIST_SCORE == orc.ST_HIST_SCORE
    assert consts["R48_ROLLOUT_WORKSPACE_BYTES"] == native.ROLLOUT_WORKSPACE_BYTES
    assert (consts["R48_ERR_NULL"], consts["R48_ERR_ALIGN"], consts["R48_ERR_ARG"], consts["R48_ERR_CUDA"],
            consts["R48_ERR_ACTION"]) == (native.ERR_NULL, native.ERR_ALIGN, native.ERR_ARG, native.ERR_CUDA,
                                          native.ERR_ACTION)


def test_argument_validation_needs_no_device(native):
    """Errors are return codes, never exceptions across the ABI; these paths return before any
    CUDA call so they run on the CPU box."""
    L = native.lib()
    buf = (ctypes.c_uint64 * 4)()
    p = ctypes.addressof(buf)
    assert L.r48_reset(p, -1, 0, 0, None) == native.ERR_ARG
    assert L.r48_reset(None, 4, 0, 0, None) == native.ERR_NULL
    assert L.r48_reset(p + 4, 2, 0, 0, None) == native.ERR_ALIGN
    assert L.r48_reset(None, 0, 0, 0, None) == native.OK               # empty batch is a no-op
    assert L.r48_step(p, p, p, None, None, 4, 0, 0, 0, 2, None, None) == native.ERR_ARG
    assert L.r48_step(None, p, p, None, None, 4, 0, 0, 0, 0, None, None) == native.ERR_NULL
    assert L.r48_afterstates(p, p + 4, None, None, None, 1, 0, None) == native.ERR_ALIGN
    assert L.r48_rollout(4, 0, 0, p, None, None, p, None) == native.ERR_NULL
    assert L.r48_rollout(0, 0, 0, None, None, None, None, None) == native.OK
    assert b"NULL" in L.r48_last_error()
    # transition ring: struct and argument checks
    big = (ctypes.c_uint64 * 16)()
    q = (ctypes.addressof(big) + 31) & ~31                                      # a 32-byte aligned record
    ring = native.Ring(q, p, 0)
    assert L.r48_ring_append(ctypes.byref(ring), p, p, None, p, None, 1, 0, None) == native.ERR_ARG     # capacity 0
    ring.capacity = 4
    ring.slots = None
    assert L.r48_ring_append(ctypes.byref(ring), p, p, None, p, None, 1, 0, None) == native.ERR_NULL
    ring.slots = q + 8
    assert L.r48_ring_append(ctypes.byref(ring), p, p, None, p, None, 1, 0, None) == native.ERR_ALIGN
    ring.slots = q
    assert L.r48_ring_append(ctypes.byref(ring), p, p, None, p, None, 0, 0, None) == native.OK          # empty append
    assert L.r48_ring_append(ctypes.byref(ring), None, p, None, p, None, 2, 0, None) == native.ERR_NULL
    assert L.r48_ring_sample(ctypes.byref(ring), 2, 0, 0, 0, None, None, p, p, p, p, None, None, 0, None) == native.ERR_NULL
    assert L.r48_ring_sample(ctypes.byref(ring), 2, 0, 0, 0, None, p, p, p, p, p, None, None, 7, None) == native.ERR_ARG
    assert L.r48_ring_clear(None, None) == native.ERR_NULL
    assert L.r48_rollout_host_ex(-1, 0, 0, 0, None, None, None, None, 0) == native.ERR_ARG
    assert L.r48_rollout_host_ex(4, 0, 0, 9, None, None, None, None, 0) == native.ERR_ARG             # unknown policy
    assert L.r48_episode_records(p, p, None, 4, None) == native.ERR_NULL
    with pytest.raises(native.R48Error):
        native.check(L.r48_reset(None, 4, 0, 0, None))


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_product_path_fails_loudly_without_gpu():
    import rein48_b200 as r48
    with pytest.raises(RuntimeError):
        r48.BatchedGame(8)
    with pytest.raises(RuntimeError):
        r48.random_rollouts(8)
    with pytest.raises(RuntimeError):
        r48.Game()
    with pytest.raises(RuntimeError):
        r48.afterstates(torch.zeros(4, dtype=torch.int64))


def test_product_code_never_imports_the_oracle():
    """No file of the product package imports, loads or names the oracle library."""
    pkg = os.path.join(ROOT, "rein48_b200")
    pat = re.compile(r"^\s*(from|import)\s+oracle|libr48_oracle|r48_oracle\.c|orc_[a-z_]+\s*\(", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not pat.search(text), f


def test_action_spellings_and_shards():
    from rein48_b200.game import action_code
    from rein48_b200.stats import shard_range
    for code, names in enumerate((("UP", "Up", "U", "up", "u", 0), ("DOWN", "Down", "D", "down", "d", 1),
                                  ("LEFT", "Left", "L", "left", "l", 2), ("RIGHT", "Right", "R", "right", "r", 3))):
        for nm in names:
            assert action_code(nm) == code
        assert action_code(np.int64(code)) == code
    for bad in ("X", 4, -1, None, "UPP"):
        with pytest.raises(ValueError):
            action_code(bad)
    n = 1_000_003
    for world in (1, 2, 4, 8):
        spans = [shard_range(n, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_rand_policy_matches_reference_stream():
    """Rand.random_action draws random.randint(0, 3) like control/rand.py:11."""
    import random
    from rein48_b200.rand import Rand
    random.seed(123)
    got = [Rand.random_action([[0]]) for _ in range(50)]
    random.seed(123)
    want = [("UP", "DOWN", "LEFT", "RIGHT")[random.randint(0, 3)] for _ in range(50)]
    assert got == want


def test_episode_stats_view():
    from oracle import oracle as orc
    from rein48_b200.stats import EpisodeStats
    fb, ln = orc.rollout(3000, 11)
    st = EpisodeStats(torch.from_numpy(orc.episode_stats(fb, ln).view(np.int64)))
    sc = orc.scores(fb)
    assert st.episodes == 3000 and st.steps == int(ln.sum())
    assert abs(st.mean_score - sc.mean()) < 1e-9 and abs(st.std_score - sc.std()) < 1e-6
    assert st.max_length == int(ln.max()) and st.max_tile == 1 << int(orc.max_exps(fb).max())
    assert st.summary()["episodes"] == 3000


_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
from oracle import oracle as orc
from rein48_b200.stats import shard_range, allreduce_stats, STATS_WORDS
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
n, seed = 4001, 77
lo, hi = shard_range(n, dist.get_rank(), 2)
fb, ln = orc.rollout(hi - lo, seed, lo)                 # this rank's slice of GLOBAL episode ids
st = torch.from_numpy(orc.episode_stats(fb, ln).view(np.int64).copy())
allreduce_stats(st)
fbw, lnw = orc.rollout(n, seed, 0)
whole = orc.episode_stats(fbw, lnw).view(np.int64)
assert (st.numpy() == whole).all(), "sharded + all-reduced stats differ from the single-process run"
assert st.numel() == STATS_WORDS
dist.destroy_process_group()
print("rank", sys.argv[1], "ok")
"""


def test_two_rank_gloo_allreduce_is_exact(tmp_path):
    """The N>1 path on CPU: two ranks own disjoint global-id ranges, one SUM all-reduce of the
    statistics vector, result identical to one process playing every episode."""
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o


def test_integration_md_stub_is_valid_python_and_binds_declared_symbols():
    """The ctypes stub in INTEGRATION.md compiles, and every r48_* name it uses is declared in r48.h
    (the GPU suite executes it against the library)."""
    import re
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = re.search(r"```python\n(# game/r48_binding\.py.*?)```", text, re.S).group(1)
    compile(block, "INTEGRATION.md", "exec")
    header = open(os.path.join(ROOT, "include", "r48.h")).read()
    used = set(re.findall(r"_lib\.(r48_\w+)", block))
    assert used and all(re.search(r"\b%s\s*\(" % name, header) for name in used), used
