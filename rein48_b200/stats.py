# -*- coding: utf-8 -*-
"""Episode statistics vector: layout (include/r48.h R48_STATS_*), readers, and the one
collective of the multi-GPU rollout -- a SUM all-reduce of this vector.

Every entry is a count or a sum, so shards reduce with a single SUM and the reduced vector
is bit-identical for 1, 2, 4 or 8 GPUs (max-type quantities are read off the highest
non-empty histogram bin)."""
import torch

STATS_WORDS = 4120
EPISODES, SUM_LEN, SUM_SCORE, SUM_SCORE2, SUM_LEN2 = 0, 1, 2, 3, 4
HIST_MAXEXP, HIST_LEN, HIST_SCORE = 8, 24, 2072
LEN_BINS = SCORE_BINS = 2048


def shard_range(n_total, rank, world_size):
    """Contiguous global episode ids owned by `rank`: [lo, hi)."""
    if not 0 <= rank < world_size:
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    lo = n_total * rank // world_size
    hi = n_total * (rank + 1) // world_size
    return lo, hi


def allreduce_stats(stats, group=None):
    """SUM all-reduce of an int64 statistics tensor in place (NCCL for CUDA tensors, gloo for
    CPU tensors); a no-op when torch.distributed is not initialised."""
    import torch.distributed as dist
    if stats.dtype != torch.int64 or stats.numel() != STATS_WORDS:
        raise ValueError("stats must be int64[%d]" % STATS_WORDS)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats


class EpisodeStats:
    """Read-only view over a statistics vector (torch tensor or numpy array)."""

    def __init__(self, vector):
        if torch.is_tensor(vector):
            vector = vector.detach().cpu().numpy()
        self.v = vector.astype("int64", copy=False)
        if self.v.size != STATS_WORDS:
            raise ValueError("expected %d words" % STATS_WORDS)

    @property
    def episodes(self):
        return int(self.v[EPISODES])

    @property
    def steps(self):
        return int(self.v[SUM_LEN])

    @property
    def mean_length(self):
        return self.v[SUM_LEN] / max(1, self.episodes)

    @property
    def mean_score(self):
        return self.v[SUM_SCORE] / max(1, self.episodes)

    @property
    def std_score(self):
        n = max(1, self.episodes)
        m = self.v[SUM_SCORE] / n
        return max(0.0, float(self.v[SUM_SCORE2]) / n - m * m) ** 0.5

    @property
    def std_length(self):
        n = max(1, self.episodes)
        m = self.v[SUM_LEN] / n
        return max(0.0, float(self.v[SUM_LEN2]) / n - m * m) ** 0.5

    @property
    def maxexp_hist(self):
        return self.v[HIST_MAXEXP:HIST_MAXEXP + 16]

    @property
    def length_hist(self):
        return self.v[HIST_LEN:HIST_LEN + LEN_BINS]

    @property
    def score_hist(self):
        """bin b counts episodes with score in {2b, 2b+1}; the last bin is a clamp"""
        return self.v[HIST_SCORE:HIST_SCORE + SCORE_BINS]

    @property
    def max_tile(self):
        nz = self.maxexp_hist.nonzero()[0]
        return (1 << int(nz[-1])) if nz.size and nz[-1] > 0 else 0

    @property
    def max_length(self):
        nz = self.length_hist.nonzero()[0]
        return int(nz[-1]) if nz.size else 0

    def summary(self):
        return {
            "episodes": self.episodes, "steps": self.steps,
            "mean_length": round(float(self.mean_length), 3), "std_length": round(self.std_length, 3),
            "mean_score": round(float(self.mean_score), 3), "std_score": round(self.std_score, 3),
            "max_tile": self.max_tile, "max_length": self.max_length,
            "max_tile_pmf": {str(1 << e): round(float(c) / max(1, self.episodes), 5)
                             for e, c in enumerate(self.maxexp_hist) if c},
        }
