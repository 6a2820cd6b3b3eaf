"""Multi-GPU plumbing: one process per GPU, trajectories sharded, weights replicated.

The forward / eval recursion needs NO collective (every reduction is inside one trajectory,
SURVEY.md section 8e); the only exchange on the path is the gradient all-reduce of BPTT training
(NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
from typing import Dict, Tuple

import torch
import torch.distributed as dist


def shard_bounds(num_trajectories: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous, balanced partition of [0, N): the first N % W ranks get one extra trajectory."""
    assert 0 <= rank < world_size
    base, extra = divmod(num_trajectories, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(tree, rank: int, world_size: int, batch_dim: int = 1):
    """Slice every tensor of a (possibly dict-valued) batch along its trajectory axis
    ((T, N, ...) -> batch_dim=1; (N, ...) -> batch_dim=0)."""
    if isinstance(tree, dict):
        return {k: shard_batch(v, rank, world_size, batch_dim) for k, v in tree.items()}
    lo, hi = shard_bounds(tree.shape[batch_dim], rank, world_size)
    return tree.narrow(batch_dim, lo, hi - lo)


def gather_estimates(local: torch.Tensor, num_trajectories: int, batch_dim: int = 1) -> torch.Tensor:
    """All-gather per-rank (T, N_local, sd) estimates back into (T, N, sd) (ragged shards allowed)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [shard_bounds(num_trajectories, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    pad_shape = list(local.shape)
    pad_shape[batch_dim] = width
    padded = local.new_zeros(pad_shape)
    padded.narrow(batch_dim, 0, local.shape[batch_dim]).copy_(local)
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded.contiguous())
    return torch.cat([p.narrow(batch_dim, 0, hi - lo) for p, (lo, hi) in zip(parts, sizes)], dim=batch_dim)


def allreduce_gradients(module: torch.nn.Module, average: bool = True) -> int:
    """One flat fp32 all-reduce of every parameter gradient (the BPTT step's only collective).
    Returns the number of elements reduced."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return 0
    grads = [p.grad for p in module.parameters() if p.grad is not None]
    if not grads:
        return 0
    flat = torch.cat([g.reshape(-1).float() for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    if average:
        flat /= dist.get_world_size()
    offset = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[offset:offset + n].view_as(g))
        offset += n
    return flat.numel()


def max_over_ranks(value: float, device) -> float:
    """Timing reduction used by bench.py: the slowest rank defines the step time."""
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
