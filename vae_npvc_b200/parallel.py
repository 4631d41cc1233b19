"""Data-parallel plumbing of the hot path: frames shard over ranks, every rank holds a
full replica of the 939,162 parameters, and ONE all-reduce(SUM) of the flat gradient buffer per
step (3.76 MB, NCCL over NVLink on GPUs; gloo on CPU for the logic tests) precedes the replicated,
deterministic TF-form Adam step (grad_scale = 1/world).  No other collective exists on the path:
per-frame layer-norm and per-frame losses never couple frames (F5/F6).
"""
import torch
import torch.distributed as dist


def world_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n_items, rank, world):
    """Contiguous, balanced [lo, hi) shard of n_items for `rank` (utterance / frame sharding)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_flat_grad_(grad, world=None):
    """In-place SUM all-reduce of the flat gradient bucket; returns the scale (1/world) the Adam
    kernel applies (mean of per-rank means == global mean for equal per-rank frame counts)."""
    _, w = world_info()
    world = w if world is None else world
    if world > 1:
        dist.all_reduce(grad, op=dist.ReduceOp.SUM)
    return 1.0 / world


def broadcast_params_(theta, src=0):
    _, w = world_info()
    if w > 1:
        dist.broadcast(theta, src=src)


def reduce_loss_scalars(losses):
    """Mean over ranks of the [G, D_KL, logP] scalars (only when logging)."""
    _, w = world_info()
    if w > 1:
        losses = losses.clone()
        dist.all_reduce(losses, op=dist.ReduceOp.SUM)
        losses /= w
    return losses
