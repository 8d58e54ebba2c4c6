"""Multi-GPU plumbing: env-index sharding and the optional episode-statistics reduction.

The physics path shards trivially: every env owns its simulator in the
reference (`/root/reference/gym_softrobot/envs/soft_pendulum/soft_pendulum.py:115`), so
rank r simply owns a contiguous range of global env indices and there is NO collective on
the hot path.  The only exchange is an optional all-reduce of a few per-rank episode
counters (NCCL on GPUs, gloo in the CPU tests).
"""
from dataclasses import dataclass


@dataclass(frozen=True)
class EnvShard:
    rank: int
    world_size: int
    n_env_total: int
    start: int   # first global env index owned by this rank
    count: int   # envs owned by this rank

    @property
    def stop(self):
        return self.start + self.count


def shard_envs(n_env_total: int, rank: int, world_size: int) -> EnvShard:
    """Contiguous, balanced ranges: the first `n % world` ranks get one extra env."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    if n_env_total < 0:
        raise ValueError("n_env_total < 0")
    base, extra = divmod(n_env_total, world_size)
    count = base + (1 if rank < extra else 0)
    start = rank * base + min(rank, extra)
    return EnvShard(rank, world_size, n_env_total, start, count)


def env_seed(base_seed: int, global_env_index: int) -> int:
    """Seed of one env: a reference env reset with `seed = base_seed + global index`."""
    return int(base_seed) + int(global_env_index)


class EpisodeStats:
    """Per-rank episode counters with an all-reduce(sum) across ranks (off the hot path)."""

    FIELDS = ("episodes", "return_sum", "length_sum", "nan_count")

    def __init__(self, device="cpu"):
        import torch
        self._t = torch.zeros(len(self.FIELDS), dtype=torch.float64, device=device)

    def add(self, episodes=0, return_sum=0.0, length_sum=0, nan_count=0):
        import torch
        self._t += torch.tensor([episodes, return_sum, length_sum, nan_count], dtype=torch.float64,
                                device=self._t.device)

    def local(self):
        return dict(zip(self.FIELDS, self._t.tolist()))

    def reduce(self):
        """Global sums (all ranks get the result). No-op when torch.distributed is not initialised."""
        import torch.distributed as dist
        t = self._t.clone()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        out = dict(zip(self.FIELDS, t.tolist()))
        out["mean_return"] = out["return_sum"] / out["episodes"] if out["episodes"] else float("nan")
        out["mean_length"] = out["length_sum"] / out["episodes"] if out["episodes"] else float("nan")
        return out
