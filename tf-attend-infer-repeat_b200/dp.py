"""Data-parallel plumbing (one process per GPU, torch.distributed): the AIR batch shards by rows,
weights are replicated, and the only exchange is ONE all-reduce(SUM) of the flat gradient buffer
(SURVEY.md 8e).  Each rank weights its per-item losses by 1/(B_local * world), so the summed
gradient equals the reference's global-batch mean gradient; clip + Adam then run on identical
reduced gradients on every rank."""
from __future__ import annotations

import torch
import torch.distributed as dist


def world_size(group=None) -> int:
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def rank(group=None) -> int:
    return dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0


def shard_rows(t: torch.Tensor, rank_: int, world: int, dim: int = 0) -> torch.Tensor:
    """Contiguous, equal row shard of the global batch (the batch must divide evenly so that
    the mean of shard means equals the global mean)."""
    n = t.shape[dim]
    if n % world:
        raise ValueError(f"global batch {n} is not divisible by world size {world}")
    per = n // world
    return t.narrow(dim, rank_ * per, per).contiguous()


def shard_noise(noise: dict, rank_: int, world: int) -> dict:
    """The five [T, B, ...] noise tensors shard along the batch axis (dim 1), consistently with the images."""
    return {k: shard_rows(v, rank_, world, dim=1) for k, v in noise.items()}


def allreduce_flat(flat_grad: torch.Tensor, group=None) -> torch.Tensor:
    """SUM all-reduce of the flat gradient buffer, in place (NCCL on GPUs, gloo in CPU tests)."""
    if world_size(group) > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
    return flat_grad


def allreduce_async(t: torch.Tensor, group=None):
    """SUM all-reduce of a (contiguous slice of the) flat gradient buffer, asynchronous: the collective is ordered
    after the work already queued on the current stream and runs on the communicator's own stream; ``.wait()`` makes
    the current stream (not the host) wait for it."""
    return dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group, async_op=True)


def allreduce_scalars(values: torch.Tensor, group=None) -> torch.Tensor:
    """Mean of per-rank scalars (logged loss / accuracy)."""
    w = world_size(group)
    if w > 1:
        dist.all_reduce(values, op=dist.ReduceOp.SUM, group=group)
        values /= w
    return values
