# -*- coding: utf-8 -*-
"""north_star checks 2 and 3, and the size-independent properties at BASELINE.json's full sizes.

Tolerances (stated here, as SURVEY 8c asks):
  spawn statistics   chi-square against the exact reference law (P(4) = 0.1; uniform over the
                     blanks), alpha = 1e-3 per test with a Bonferroni factor for the 16
                     position tests, >= 1e7 GPU spawns in total;
  rollouts           two-sample KS, GPU (2^20 episodes) vs the reference (200 000 episodes in
                     tests/golden/dist_ref.npz) on episode length and score, reject if
                     p < 1e-3; chi-square homogeneity on the max-tile pmf, same alpha.
"""
import numpy as np
import pytest
import torch
from scipy import stats as sps

pytestmark = pytest.mark.gpu

ALPHA = 1e-3


@pytest.fixture(scope="module")
def r48():
    import rein48_b200
    rein48_b200._native.lib()
    return rein48_b200


def ks_from_hists(h1, h2):
    """two-sample KS on binned integer data: D and the asymptotic p-value"""
    n1, n2 = h1.sum(), h2.sum()
    d = np.abs(np.cumsum(h1) / n1 - np.cumsum(h2) / n2).max()
    en = np.sqrt(n1 * n2 / (n1 + n2))
    return d, sps.kstwobign.sf((en + 0.12 + 0.11 / en) * d)


def test_spawn_chi_square(r48):
    from rein48_b200.batched import spawn
    per = 700_000
    rng = np.random.default_rng(1)
    total = 0
    fours = 0
    for nb in range(1, 17):
        cells = rng.permutation(16)
        board = 0
        for p in cells[nb:]:
            board |= int(rng.integers(1, 12)) << (4 * int(p))
        blanks = np.sort(cells[:nb])
        if board >= 1 << 63:
            board -= 1 << 64                                   # int64 bit pattern
        b = torch.full((per,), board, dtype=torch.int64, device="cuda")
        spawn(b, seed=424242 + nb, board_base=nb * 10_000_000, tick=nb)
        diff = (b ^ board)
        assert bool((diff != 0).all())
        # the spawned nibble: isolate the lowest set bit -> position and value
        low = diff & (-diff)
        pos = torch.log2(low.to(torch.float64)).round().to(torch.int64)      # exact for powers of two < 2^53?
        # positions up to bit 61: float64 log2 of an exact power of two is exact
        cell = pos // 4
        vexp = torch.where(pos % 4 == 0, 1, 2)
        assert bool((((diff >> (cell * 4)) & 15) == vexp).all())
        counts = torch.bincount(cell, minlength=16).cpu().numpy()
        assert counts[[c for c in range(16) if c not in set(blanks.tolist())]].sum() == 0
        if nb > 1:
            chi, p = sps.chisquare(counts[blanks])
            assert p > ALPHA / 16, (nb, chi, p)
        fours += int((vexp == 2).sum())
        total += per
    assert total >= 10_000_000
    chi, p = sps.chisquare([fours, total - fours], [0.1 * total, 0.9 * total])
    assert p > ALPHA, (fours / total, chi, p)


def test_rollout_distributions_match_reference(r48, golden):
    ref = golden("dist_ref.npz")
    n = 1 << 20
    res = r48.random_rollouts(n, seed=20480)
    st = r48.EpisodeStats(res.stats)
    assert st.episodes == n
    d_len, p_len = ks_from_hists(st.length_hist.astype(np.float64), ref["len_hist"].astype(np.float64))
    d_sc, p_sc = ks_from_hists(st.score_hist.astype(np.float64), ref["score_hist"].astype(np.float64))
    assert p_len > ALPHA, ("length", d_len, p_len)
    assert p_sc > ALPHA, ("score", d_sc, p_sc)
    # max-tile pmf: chi-square homogeneity over the bins the reference populated well
    g = st.maxexp_hist.astype(np.float64)
    r = ref["maxexp_hist"][:16].astype(np.float64)
    keep = (r + g) > 50
    table = np.stack([g[keep], r[keep]])
    chi, p, _, _ = sps.chi2_contingency(table)
    assert p > ALPHA, ("max tile", chi, p)
    # SURVEY section 6 figures, loosely (means within 1%)
    assert abs(st.mean_length - 142.3) < 1.5 and abs(st.mean_score - 265.1) < 2.7


def test_rollout_config3_properties(r48):
    """16M-episode fused rollout (BASELINE config 3): every final board is dead, every length
    is positive, histograms account for every episode, score is even and equals the spawned
    mass bound, and the result does not depend on how the launch was scheduled (run twice)."""
    n = 1 << 24
    buf = r48.RolloutBuffers(n)
    res = r48.random_rollouts(n, seed=2048, buffers=buf)
    st1 = res.stats.clone()
    fb1 = res.final_boards[:4096].clone()
    _, _, valid, done = r48.afterstates(res.final_boards)
    assert bool((valid == 0).all()) and bool((done == 1).all())
    assert int(res.lengths.min()) >= 1
    st = r48.EpisodeStats(st1)
    assert st.episodes == n and st.length_hist.sum() == n and st.score_hist.sum() == n
    assert st.maxexp_hist.sum() == n
    sc, _ = r48.scores(res.final_boards)
    assert int(sc.to(torch.int64).sum()) == int(st1[2])
    # a 4x4 board holds 16 tiles: length >= number of spawns - 1 >= score/4 - 1
    assert bool((res.lengths.to(torch.int64) * 4 + 4 >= sc.to(torch.int64)).all())
    del valid, done, sc
    res2 = r48.random_rollouts(n, seed=2048, buffers=buf)
    assert bool((res2.stats == st1).all()) and bool((res2.final_boards[:4096] == fb1).all())


def test_step_config2_properties(r48):
    """1M boards, one step (BASELINE config 2): tile mass grows by exactly the spawned tile
    where the board changed and by nothing elsewhere; unchanged boards are bit-identical."""
    n = 1 << 20
    warm = r48.random_rollouts(n, seed=7)            # dead boards ...
    env = r48.BatchedGame(n, seed=11)
    for _ in range(40):                              # ... and live mid-game ones
        env.step(torch.randint(0, 4, (n,), device="cuda", dtype=torch.uint8))
    for boards in (env.boards.clone(), warm.final_boards.clone()):
        env.boards.copy_(boards)
        before, _ = r48.scores(boards)
        acts = torch.randint(0, 4, (n,), device="cuda", dtype=torch.uint8)
        after_all, _, valid, _ = r48.afterstates(boards)
        new, reward, done = env.step(acts)
        after, _ = r48.scores(new)
        changed = ((valid.to(torch.int64) >> acts.to(torch.int64)) & 1).bool()
        gain = after - before
        assert bool(((gain == 2) | (gain == 4))[changed].all())
        assert bool((new == boards)[~changed].all())
        assert bool((reward == 0).all())
        moved = after_all.gather(0, acts.to(torch.int64)[None, :])[0]
        assert bool(((new ^ moved) != 0)[changed].all()) and bool((r48.blank_counts(moved)[changed] >= 1).all())
        _, _, v2, d2 = r48.afterstates(new)
        assert bool((done == d2).all())


def test_afterstates_config4_properties(r48):
    """8M boards, 4 moves each (BASELINE config 4): every move conserves the tile mass, the
    valid mask is exactly 'afterstate differs', done <=> no valid move on a non-empty board."""
    n = 1 << 23
    env = r48.BatchedGame(n, seed=3)
    for _ in range(60):
        env.step(torch.randint(0, 4, (n,), device="cuda", dtype=torch.uint8))
    boards = env.boards
    after, reward, valid, done = r48.afterstates(boards, reward_mode=1)
    base, _ = r48.scores(boards)
    for a in range(4):
        col = after[a]
        sc, _ = r48.scores(col)
        assert bool((sc == base).all())
        differs = col != boards
        assert bool((differs == ((valid >> a) & 1).bool()).all())
        assert bool((reward[a][~differs] == 0).all())
    assert bool(((valid == 0) == (done == 1)).all())
    assert 0 < int(done.sum()) < n


def test_rollout_one_million_episodes_bit_exact(r48, orc):
    """A larger differential run: 2^20 episodes (1.5e8 env-steps) on the GPU against the oracle
    on all host threads -- every final board and every length."""
    import os
    n = 1 << 20
    res = r48.random_rollouts(n, seed=0xDEADBEEF, board_base=1 << 40)
    fb, ln = orc.rollout(n, 0xDEADBEEF, 1 << 40, threads=max(1, (os.cpu_count() or 4)))
    assert (res.final_boards.cpu().numpy().view(np.uint64) == fb).all()
    assert (res.lengths.cpu().numpy().view(np.uint32) == ln).all()
    assert (res.stats.cpu().numpy().view(np.uint64) == orc.episode_stats(fb, ln)).all()
