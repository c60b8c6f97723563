"""Environment sharding across GPUs: contiguous blocks per rank, no collective inside a step,
one all-gather of the packed observation block per env.step(). Backend-agnostic (NCCL on GPUs, gloo in tests)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_total: int, rank: int, world: int) -> tuple[int, int]:
    """env ids [begin, end) owned by `rank`: env_id in [r*N/G, (r+1)*N/G)."""
    return (n_total * rank) // world, (n_total * (rank + 1)) // world


def gather_observations(obs_local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """Returns the [n_total, obs_dim] observation block on every rank. Shards may differ by one row."""
    world = dist.get_world_size(group)
    if world == 1:
        return obs_local
    rank = dist.get_rank(group)
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    assert obs_local.shape[0] == sizes[rank][1] - sizes[rank][0]
    if all(e - b == sizes[0][1] - sizes[0][0] for b, e in sizes):
        out = torch.empty((n_total, *obs_local.shape[1:]), dtype=obs_local.dtype, device=obs_local.device)
        dist.all_gather_into_tensor(out.view(-1), obs_local.contiguous().view(-1), group=group)
        return out
    # ragged shards: pad every rank to the largest shard, gather once, strip the padding
    mx = max(e - b for b, e in sizes)
    padded = torch.zeros((mx, *obs_local.shape[1:]), dtype=obs_local.dtype, device=obs_local.device)
    padded[: obs_local.shape[0]] = obs_local
    buf = torch.empty((world * mx, *obs_local.shape[1:]), dtype=obs_local.dtype, device=obs_local.device)
    dist.all_gather_into_tensor(buf.view(-1), padded.view(-1), group=group)
    return torch.cat([buf[r * mx: r * mx + (e - b)] for r, (b, e) in enumerate(sizes)], dim=0)
