"""Multi-GPU: batches of proteins run as independent replicas.

Every trunk op is per-sample (all contractions are per `b`, InstanceNorm is per sample), so a
batch shards over ranks with NO data-path collective: each rank runs the trunk on its slice; an
all-gather reassembles the outputs only if the caller wants them on every rank. One process per
GPU, `torch.distributed` (NCCL on GPUs, gloo in the CPU tests) for the plumbing.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(batch: int, rank: int, world: int):
    """Contiguous, balanced slice [lo, hi) of a batch for `rank` (first ranks get the remainder)."""
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


@torch.no_grad()
def run_replicated(trunk, msa: torch.Tensor, pair: torch.Tensor, group=None, gather: bool = True):
    """Run `trunk(msa, pair)` data-parallel over the batch axis.

    Every rank holds the full (msa, pair) batch (or at least its own slice range); returns the
    full-batch outputs on every rank when `gather`, else this rank's slice.
    """
    if not dist.is_initialized():
        return trunk(msa, pair)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    B = msa.shape[0]
    lo, hi = shard_bounds(B, rank, world)
    if hi > lo:
        m, p = trunk(msa[lo:hi].contiguous(), pair[lo:hi].contiguous())
    else:
        m, p = msa[:0].float(), pair[:0].float()
    if not gather:
        return m, p
    outs = []
    for t, full in ((m, msa), (p, pair)):
        parts = []
        for r in range(world):
            a, b = shard_bounds(B, r, world)
            parts.append(torch.empty((b - a,) + tuple(full.shape[1:]), dtype=torch.float32, device=full.device))
        dist.all_gather(parts, t.contiguous(), group=group) if all(x.shape == parts[0].shape for x in parts) \
            else _all_gather_ragged(parts, t.contiguous(), rank, world, group)
        outs.append(torch.cat(parts, 0))
    return outs[0], outs[1]


def _all_gather_ragged(parts, mine, rank, world, group):
    """Uneven slices: broadcast each rank's slice in turn."""
    for r in range(world):
        if r == rank:
            parts[r].copy_(mine)
        dist.broadcast(parts[r], src=dist.get_global_rank(group, r) if group is not None else r, group=group)
