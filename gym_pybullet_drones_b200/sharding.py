"""Multi-GPU layout of the aviaries: one process per GPU, each owning a contiguous slice of the aviary axis.

The step path needs NO communication -- every aviary is independent (the only couplings in the reference,
`_downwash` and MultiHover's reward/termination reductions, stay inside one aviary, BaseAviary.py:785-811,
MultiHoverAviary.py:75-130), and an aviary never straddles GPUs.  The only collective offered is an optional
all-gather of observations / rewards / flags for a learner that wants single tensors (torch.distributed:
NCCL over NVLink on GPUs, gloo on CPU for the tests).
"""
from dataclasses import dataclass

import torch
import torch.distributed as dist


@dataclass(frozen=True)
class EnvShard:
    """Aviaries [start, stop) of `total` belong to `rank` of `world`."""
    rank: int
    world: int
    total: int
    start: int
    stop: int

    @property
    def count(self):
        return self.stop - self.start


def shard_envs(total_envs: int, rank: int = None, world: int = None) -> EnvShard:
    """Contiguous, balanced partition of the aviary axis (the first `total % world` ranks get one extra)."""
    if world is None:
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world %d" % (rank, world))
    if total_envs < world:
        raise ValueError("fewer aviaries (%d) than ranks (%d)" % (total_envs, world))
    base, extra = divmod(total_envs, world)
    start = rank * base + min(rank, extra)
    return EnvShard(rank, world, total_envs, start, start + base + (1 if rank < extra else 0))


def all_gather_envs(local: torch.Tensor, shard: EnvShard, group=None, out: torch.Tensor = None) -> torch.Tensor:
    """All-gathers a per-aviary tensor [E_local, ...] into [E_total, ...] (same order as a single-GPU run).
    Off the step path: call it only where the learner needs one tensor.  (formation.py uses it with `out=` for the
    positions of a formation sharded by drones: same partition, the unit is then a drone.)"""
    if shard.world == 1:
        if out is None:
            return local
        out.copy_(local)
        return out
    counts = [shard_envs(shard.total, r, shard.world).count for r in range(shard.world)]
    if len(set(counts)) == 1:
        if out is None:
            out = torch.empty((shard.total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    # uneven shards: collectives need equal sizes, so pad every shard to the largest and trim after the gather
    out_arg = out
    m = max(counts)
    padded = torch.zeros((m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[:local.shape[0]] = local
    out = torch.empty((shard.world * m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    res = torch.cat([out[r * m:r * m + c] for r, c in enumerate(counts)], dim=0)
    if out_arg is not None:
        out_arg.copy_(res)
        return out_arg
    return res
