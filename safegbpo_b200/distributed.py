"""Multi-GPU plumbing: environments shard across ranks with no data-path collective; the
exchanges are one all-reduce of the policy gradients per SHAC episode (shac.py:146-149), one all-reduce of the critic
gradients per critic batch (shac.py:176-181: every rank fits on ITS shard of the TD targets, the averaged gradient is
the gradient of the mean loss over the ranks' batches), and a few episode scalars (shac.py:119-120, 140-142).

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
    if engine is not None and engine.last and engine.last.get("grads_synced"):
        return          # the episode graph already averaged the static gradient buffers (fused_rollout._episode_body)
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


allreduce_policy_grads.graph_capturable = True      # (one or two in-place NCCL calls on static buffers)


def _allreduce_mean(flat: torch.Tensor):
    if dist.get_backend() == "nccl":
        dist.all_reduce(flat, op=dist.ReduceOp.AVG)
    else:                                   # gloo has no AVG
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(dist.get_world_size())


def allreduce_critic_grads(algo):
    """Mean over ranks of the critic gradients of the current batch, before clip + Adam (shac.py:176-181).  All
    `value_function.*.grad` are made views of ONE flat buffer on first use, so this is a single in-place collective."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    params = [p for p in algo.value_function.parameters() if p.grad is not None]
    flat = getattr(algo, "_vf_flat_grad", None)
    ok = flat is not None and all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(params, algo._vf_flat_views))
    if not ok:
        flat = torch.cat([p.grad.reshape(-1) for p in params])
        views, off = [], 0
        for p in params:
            v = flat[off:off + p.numel()].view_as(p)
            off += p.numel()
            p.grad = v
            views.append(v)
        algo._vf_flat_grad, algo._vf_flat_views = flat, views
    _allreduce_mean(flat)


def allreduce_episode_scalars(policy_loss: float, average_reward: float, interventions: int, device=None) -> tuple[float, float, int]:
    """Global episode scalars: every rank's loss / average reward is normalised by ITS environment count (equal shards), so
    the global values are the means over ranks; interventions add up (safeguard.py:76-80)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return policy_loss, average_reward, interventions
    t = torch.tensor([policy_loss, average_reward, float(interventions)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    w = dist.get_world_size()
    return float(t[0]) / w, float(t[1]) / w, int(round(float(t[2])))


def attach(algo, src: int = 0):
    """Make `algo` one replica of a data-parallel job: same initial networks on every rank, policy AND critic gradients
    averaged over ranks, so policy, critic and target critic stay identical after every `_learn_episode`."""
    broadcast_parameters(algo, src)
    algo.grad_sync = allreduce_policy_grads
    algo.vf_grad_sync = allreduce_critic_grads


def replicas_in_sync(algo) -> bool:
    """True when policy, critic and target critic parameters are bit-identical on every rank."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return True
    ps = list(algo.policy.parameters()) + list(algo.value_function.parameters()) + list(algo.target_value_function.parameters())
    flat = torch.cat([p.detach().reshape(-1).double() for p in ps])
    lo, hi = flat.clone(), flat.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    return bool(torch.equal(lo, hi))
