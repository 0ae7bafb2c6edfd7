"""Batch sharding across the GPUs of one box: trajectories are independent, so each rank
runs the whole reverse loop on its contiguous slice of the batch (weights, SDF grid, schedule and
start/goal replicated) and the only collective is ONE final all-gather of the sampled plans
[B/G, H, D] over NCCL/NVLink. The reference has no multi-GPU code at all.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n_total: int, world_size: int, rank: int):
    """Contiguous [lo, hi) slice of rank; the remainder goes to the first ranks."""
    base, rem = divmod(n_total, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allgather_plans(x_local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """[b_local, H, D] per rank -> [n_total, H, D] on every rank (rank order = batch order)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        assert x_local.shape[0] == n_total
        return x_local
    world = dist.get_world_size(group)
    sizes = [shard_bounds(n_total, world, r) for r in range(world)]
    counts = [hi - lo for lo, hi in sizes]
    x_local = x_local.contiguous()
    assert x_local.shape[0] == counts[dist.get_rank(group)], "local shard has the wrong size"
    if len(set(counts)) == 1:
        out = torch.empty((n_total, *x_local.shape[1:]), dtype=x_local.dtype, device=x_local.device)
        dist.all_gather_into_tensor(out, x_local, group=group)
        return out
    mx = max(counts)
    pad = torch.zeros((mx, *x_local.shape[1:]), dtype=x_local.dtype, device=x_local.device)
    pad[: x_local.shape[0]] = x_local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)


def sample_sharded(sample_local, n_total: int, noise=None, gather=True, group=None):
    """Runs `sample_local(n_local, noise_local)` on this rank's slice and gathers the plans.

    `noise` (optional, parity mode): the GLOBAL injected noise [S, n_total, H, D]; each rank takes its batch
    slice, so the gathered result equals the single-process run shard by shard — with ONE caveat: the guide
    reproduces `LimitsNormalizer.unnormalize`'s batch-global clip (normalization.py:156-167: if ANY element of
    the batch leaves [-1, 1] the whole batch is clamped) through one device flag per launch, and each rank's
    flag sees only its own shard. A shard with no out-of-range value therefore does not clamp where the
    single-process batch would have (and vice versa the result of the reference itself depends on which
    trajectories share a batch). Unguided sampling has no such coupling and is shard-invariant bit for bit;
    tests/test_gpu_parity.py (shard invariance, guided) counts the batch-dependent clamps of the run it compares.
    """
    if dist.is_available() and dist.is_initialized():
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    else:
        world, rank = 1, 0
    lo, hi = shard_bounds(n_total, world, rank)
    noise_local = noise[:, lo:hi].contiguous() if noise is not None else None
    x_local = sample_local(hi - lo, noise_local)
    if not gather:
        return x_local
    return allgather_plans(x_local, n_total, group)
