"""Multi-GPU plumbing: environments shard across ranks with no data-path collective; the
only exchange is one all-reduce of the policy gradients per SHAC episode plus a few scalars.

One flat bucket: the largest policy (ManageHousehold, [512]x3) is ~0.54 M floats = 2.2 MB, latency-bound
over NVSwitch, so a single NCCL launch beats per-parameter calls."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(num_envs: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block of environments owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(num_envs, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def broadcast_parameters(algo, src: int = 0):
    for p in list(algo.policy.parameters()) + list(algo.target_value_function.parameters()) \
            + list(algo.value_function.parameters()):
        dist.broadcast(p.data, src)


def allreduce_policy_grads(algo):
    """Mean over ranks of d(loss_r)/d(theta): each rank's loss is normalised by ITS env count (shac.py:139-142),
    so the global loss over world*N environments is the mean of the per-rank losses."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    engine = getattr(algo, "_fused_engine", None)
    bufs = engine.static_grad_buffers() if engine is not None else None
    if bufs is not None:
        # graph path: every policy gradient is a view of two static buffers -> two in-place NCCL calls, no packing
        for b in bufs:
            dist.all_reduce(b, op=dist.ReduceOp.AVG)
        return
    grads = [p.grad for p in algo.policy.parameters() if p.grad is not None]
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat.div_(dist.get_world_size())
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()
