# -*- coding: utf-8 -*-
"""The on-device transition ring (algorithm/ddpg/replay.py:8-47) against the oracle
ring: wrap-around, contents after 3 x capacity appends, the reference's store/sample/clear rules,
the append fused into the env-step kernel, and chi-square of the sampled slots."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

SEED = 0xFEED_2048


@pytest.fixture(scope="module")
def r48():
    import rein48_b200
    rein48_b200._native.lib()
    return rein48_b200


def u64(t):
    return t.detach().cpu().numpy().view(np.uint64)


def ring_arrays(ring):
    return (u64(ring.state), ring.action.cpu().numpy(), ring.reward.cpu().numpy(), u64(ring.next_state),
            ring.done.cpu().numpy())


def assert_same(ring, oring):
    got = ring_arrays(ring)
    want = (oring.state, oring.action, oring.reward, oring.next_state, oring.done)
    for g, w in zip(got, want):
        assert (g == w).all()
    assert int(ring.cursor[0].item()) == int(oring.cursor[0]) and int(ring.cursor[1].item()) == 0
    assert ring.cur_size == oring.size()


def batch(rng, n):
    s = rng.integers(0, 1 << 62, n).astype(np.uint64)
    nx = rng.integers(0, 1 << 62, n).astype(np.uint64)
    return (s, rng.integers(0, 4, n).astype(np.uint8), rng.integers(0, 1000, n).astype(np.int32), nx,
            rng.integers(0, 2, n).astype(np.uint8))


def to_dev(arrs):
    return [torch.from_numpy(a.view(np.int64) if a.dtype == np.uint64 else a).cuda() for a in arrs]


@pytest.mark.parametrize("cap", [1000, 4097])
def test_append_wraps_like_the_oracle(r48, orc, cap):
    """3 x capacity transitions in ragged batches (one of them larger than the ring)"""
    rng = np.random.default_rng(cap)
    ring, oring = r48.ReplayRing(cap, mode="ring"), orc.Ring(cap)
    total = 0
    sizes = [1, cap // 3, cap - 1, cap + 17, 5, cap // 2, 2 * cap + 3]
    for n in sizes:
        b = batch(rng, n)
        ring.store(*to_dev(b))
        oring.append(*b)
        total += n
        assert_same(ring, oring)
    assert total > 3 * cap and ring.filled() and len(ring) == cap


def test_reference_mode_is_replay_py(r48, orc):
    """mode='reference': store drops once full (replay.py:18-21), sample returns distinct items or
    everything it holds (:30-33) and clears the buffer (:26)."""
    cap = 100
    rng = np.random.default_rng(9)
    ring, oring = r48.ReplayRing(cap, mode="reference", seed=SEED), orc.Ring(cap)
    assert not ring.filled()
    for n in (30, 30, 30, 30):
        b = batch(rng, n)
        ring.store(*to_dev(b))
        oring.append(*b, drop_when_full=True)
    assert_same(ring, oring)
    assert ring.filled() and ring.cur_size == cap == int(oring.cursor[0])
    want = oring.sample_indices(10, SEED, 0)
    got = ring.sample()                                   # MINI_BATCH_SIZE = 10 (replay.py:5)
    assert set(got) >= {"state", "action", "reward", "next_state"}       # the reference's keys (replay.py:38-43)
    assert (got["index"].cpu().numpy() == want).all() and len(set(want.tolist())) == 10
    assert (u64(got["state"]) == oring.state[want]).all() and (u64(got["next_state"]) == oring.next_state[want]).all()
    assert (got["action"].cpu().numpy() == oring.action[want]).all()
    assert (got["reward"].cpu().numpy() == oring.reward[want]).all()
    assert (got["done"].cpu().numpy() == oring.done[want]).all()
    assert ring.cur_size == 0 and int(ring.cursor[0].item()) == 0 and not ring.filled()    # cleared
    b = batch(rng, 4)
    ring.store(*to_dev(b))
    got = ring.sample(10)                                 # fewer than asked for: everything, once each
    assert got["state"].numel() == 4 and sorted(u64(got["state"]).tolist()) == sorted(b[0].tolist())


def test_sample_gathers_and_decodes(r48, orc):
    cap, n = 5000, 3777
    rng = np.random.default_rng(1)
    ring, oring = r48.ReplayRing(cap, seed=SEED), orc.Ring(cap)
    e = rng.integers(0, 12, (n, 16)).astype(np.uint64)
    s = (e << (np.arange(16, dtype=np.uint64) * np.uint64(4))).sum(1).astype(np.uint64)
    b = (s, rng.integers(0, 4, n).astype(np.uint8), rng.integers(0, 9, n).astype(np.int32), np.roll(s, 1),
         rng.integers(0, 2, n).astype(np.uint8))
    ring.store(*to_dev(b))
    oring.append(*b)
    for draw, (bs, repl, log2) in enumerate(((256, False, False), (4096, False, True), (4096, True, False))):
        want = oring.sample_indices(bs, SEED, draw, with_replacement=repl)
        keep = want[want >= 0]
        got = ring.sample(bs, replace=repl, obs=True, log2=log2)
        assert got["index"].numel() == keep.size == (bs if repl else min(bs, n))
        assert (got["index"].cpu().numpy() == keep).all()
        if not repl:
            assert len(set(keep.tolist())) == keep.size
        assert (u64(got["state_packed"]) == oring.state[keep]).all()
        assert (got["state"].cpu().numpy() == orc.decode_batch(oring.state[keep], "float32", log2)).all()
        assert (got["next_state"].cpu().numpy() == orc.decode_batch(oring.next_state[keep], "float32", log2)).all()
        assert (got["reward"].cpu().numpy() == oring.reward[keep]).all()
    assert ring.cur_size == n                             # mode='ring': sampling does not clear


def test_sampled_slots_are_uniform(r48):
    """chi-square of slot frequencies over 2000 keyed draws of 64, alpha = 1e-3, both samplers"""
    from scipy.stats import chi2
    cap = 997
    ring = r48.ReplayRing(cap, seed=3)
    z = torch.zeros(cap, dtype=torch.int64, device="cuda")
    ring.store(z, z.to(torch.uint8), z.to(torch.int32), z, z.to(torch.uint8))
    for repl in (False, True):
        cnt = torch.zeros(cap, dtype=torch.int64, device="cuda")
        for _ in range(2000):
            cnt += torch.bincount(ring.sample(64, replace=repl)["index"], minlength=cap)
        c = cnt.cpu().numpy().astype(float)
        exp = 2000 * 64 / cap
        assert chi2.sf(((c - exp) ** 2 / exp).sum(), cap - 1) > 1e-3, repl


@pytest.mark.parametrize("mode", [0, 1])
def test_env_step_appends_its_transitions(r48, orc, mode):
    """the append fused into the env-step kernel: (s, a, r, s', done) of every env, s' being the
    board the step produced (the finished board when done), distinct from s"""
    n, cap = 1500, 4000
    env = r48.BatchedGame(n, seed=SEED, board_base=77, reward_mode=mode, id_stride=n)
    ring, oring = r48.ReplayRing(cap), orc.Ring(cap)
    boards = orc.reset_batch(n, SEED, 77)
    steps = np.zeros(n, np.uint32)
    eps = np.zeros(n, np.uint32)
    rng = np.random.default_rng(6)
    dones = 0
    for t in range(120):
        a = rng.integers(0, 4, n).astype(np.uint8)
        env.env_step(torch.from_numpy(a).cuda(), ring=ring)
        before = boards
        boards, steps, eps, o_r, o_d, o_f = orc.env_step_batch(boards, a, steps, eps, SEED, 77, n, mode)
        nxt = np.where(o_d.astype(bool), o_f, boards)       # the finished board, not the auto-reset one
        oring.append(before, a, o_r, nxt, o_d)
        dones += int(o_d.sum())
        if t % 17 == 0 or t == 119:
            assert_same(ring, oring)
    assert dones > 0 and int(oring.cursor[0]) == 120 * n > 3 * cap
    with pytest.raises(ValueError):
        env.env_step(torch.zeros(n, dtype=torch.uint8, device="cuda"), ring=r48.ReplayRing(8, mode="reference"))


def test_ring_append_in_a_cuda_graph(r48, orc):
    """the cursor lives on the device: a captured append replayed k times is k appends"""
    cap, n, k = 1000, 300, 9
    ring, oring = r48.ReplayRing(cap), orc.Ring(cap)
    b = batch(np.random.default_rng(2), n)
    d = to_dev(b)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        ring.store(*d)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        ring.store(*d)
    for _ in range(k):
        g.replay()
    torch.cuda.synchronize()
    ring.appended = (k + 1) * n                           # the host mirror cannot see graph replays
    for _ in range(k + 1):
        oring.append(*b)
    assert_same(ring, oring)


def test_ring_full_size_properties(r48):
    """2^20 envs feeding a 2^22-slot ring for 6 steps (1.5 x capacity), then a 2^18 sample: properties
    that need no oracle.  A transition conserves the tile sum up to the spawned tile (+0, +2 or +4);
    done marks exactly the boards no move changes; sampled slots are distinct and inside the ring;
    next_state is a value of its own (a changed board differs from its predecessor)."""
    n, cap, b = 1 << 20, 1 << 22, 1 << 18
    env = r48.BatchedGame(n, seed=SEED)
    ring = r48.ReplayRing(cap, seed=SEED)
    g = torch.Generator(device="cuda").manual_seed(5)
    for _ in range(60):                                                   # mid-game boards first
        env.env_step(torch.randint(0, 4, (n,), device="cuda", dtype=torch.uint8, generator=g))
    for _ in range(6):
        env.env_step(torch.randint(0, 4, (n,), device="cuda", dtype=torch.uint8, generator=g), ring=ring)
    assert ring.cur_size == cap and int(ring.cursor[0].item()) == 6 * n
    s = ring.sample(b)
    idx = s["index"]
    assert idx.numel() == b and int(idx.min()) >= 0 and int(idx.max()) < cap
    assert torch.unique(idx).numel() == b                                  # without replacement
    sc0, _ = r48.scores(s["state"])
    sc1, _ = r48.scores(s["next_state"])
    gain = sc1 - sc0
    changed = s["state"] != s["next_state"]
    assert bool(((gain == 2) | (gain == 4))[changed].all()) and bool((gain == 0)[~changed].all())
    _, _, valid, over = r48.afterstates(s["next_state"])
    assert bool((over == s["done"]).all())
    assert bool((s["action"] <= 3).all()) and bool((s["reward"] == 0).all())
    # the move recorded is the move that was made: applying it to `state` gives next_state minus one tile
    after, _, _, _ = r48.afterstates(s["state"])
    moved = after.gather(0, s["action"].to(torch.int64)[None, :])[0]
    diff = moved ^ s["next_state"]
    nz = torch.zeros_like(diff)
    for t in range(16):                                                    # number of nibbles that differ
        nz += ((diff >> (4 * t)) & 15) != 0
    assert bool((nz[changed] == 1).all()) and bool((nz[~changed] == 0).all())
