"""CPU: the N>1 host logic (environment sharding, parameter broadcast, gradient all-reduce) with
world_size 2 on the gloo backend.  The data path itself has no collective."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

from safegbpo_b200 import distributed as sgdist


def test_shard_bounds_partition():
    for n in (1, 7, 64, 4096, 65537):
        for world in (1, 2, 3, 8):
            spans = [sgdist.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


class _Algo:
    def __init__(self, seed):
        torch.manual_seed(seed)
        self.policy = nn.Sequential(nn.Linear(5, 8), nn.ELU(), nn.LayerNorm(8), nn.Linear(8, 1))
        self.value_function = nn.Linear(5, 1)
        self.target_value_function = nn.Linear(5, 1)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    algo = _Algo(seed=rank)                      # different initial weights per rank
    sgdist.broadcast_parameters(algo, src=0)
    ref = _Algo(seed=0)
    same = all(torch.equal(a, b) for a, b in zip(algo.policy.parameters(), ref.policy.parameters()))
    # per-rank loss on the rank's shard of a common batch
    torch.manual_seed(123)
    x = torch.randn(10, 5)
    lo, hi = sgdist.shard_bounds(10, rank, world)
    algo.policy(x[lo:hi]).pow(2).mean().backward()
    sgdist.allreduce_policy_grads(algo)
    full = _Algo(seed=0)
    (sum(full.policy(x[slice(*sgdist.shard_bounds(10, r, world))]).pow(2).mean() for r in range(world)) / world).backward()
    close = all(torch.allclose(a.grad, b.grad, atol=1e-7) for a, b in zip(algo.policy.parameters(), full.policy.parameters()))
    out[rank] = (same, close)
    dist.destroy_process_group()


def test_broadcast_and_gradient_allreduce_world2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert dict(out) == {0: (True, True), 1: (True, True)}


class _ToyEnv:
    """Differentiable linear plant on the CPU with the Simulator surface SHAC's eager loop uses (the real environments
    have no CPU path); truncation every `num_steps` like simulator.py:102-106."""

    def __init__(self, num_envs: int, num_steps: int, seed: int):
        self.num_envs, self.num_steps, self.obs_dim, self.action_dim = num_envs, num_steps, 3, 1
        self.gen = torch.Generator().manual_seed(seed)
        self.state = torch.zeros(num_envs, 3)
        self.steps = 0

    def reset(self, seed=None):
        self.state = torch.rand(self.num_envs, 3, generator=self.gen) * 2 - 1
        self.steps = 0
        return self.state, {}

    def cut_computation_graph(self):
        self.state = self.state.detach()

    def step(self, action):
        if self.steps >= self.num_steps:
            self.reset()
        a = torch.tensor([[0.9, 0.1, 0.0], [0.0, 0.9, 0.1], [0.0, 0.0, 0.8]])
        self.state = self.state @ a.t() + torch.cat([torch.zeros(self.num_envs, 2), 0.1 * action], dim=1)
        self.steps += 1
        reward = -(self.state ** 2).sum(dim=1) - 0.01 * action[:, 0] ** 2
        trunc = torch.full((self.num_envs,), self.steps >= self.num_steps)
        return self.state, reward, torch.zeros_like(trunc), trunc, {}


def _learn_worker(rank, world, port, out):
    from safegbpo_b200.learning_algorithms.shac import SHAC
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)                      # different initial networks and different environment shards
    env = _ToyEnv(num_envs=6, num_steps=5, seed=rank)
    algo = SHAC(env=env, policy_kwargs={"net_arch": [16, 16]}, policy_optim_kwargs={"lr": 1e-2},
                vf_kwargs={"net_arch": [16, 16]}, vf_optim_kwargs={"lr": 1e-2}, regularisation_coefficient=0.0,
                len_trajectories=8, vf_num_fits=2, vf_fit_num_batches=2, fused="off")
    before = sgdist.replicas_in_sync(algo)             # False: per-rank seeds
    sgdist.attach(algo)
    attached = sgdist.replicas_in_sync(algo)
    scal = None
    for ep in range(3):
        r, pl, vl = algo._learn_episode(ep)
        scal = sgdist.allreduce_episode_scalars(pl, r, 3 + rank)
    after = sgdist.replicas_in_sync(algo)
    # without the critic hook the critics drift apart (what round 1 shipped): the probe must notice
    algo.vf_grad_sync = None
    algo._learn_episode(3)
    drift = sgdist.replicas_in_sync(algo)
    out[rank] = (before, attached, after, drift, scal)
    dist.destroy_process_group()


def test_learn_episode_keeps_policy_and_critics_in_sync_world2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_learn_worker, args=(2, port, out), nprocs=2, join=True)
    res = dict(out)
    for rank in (0, 1):
        before, attached, after, drift, scal = res[rank]
        assert (before, attached, after, drift) == (False, True, True, False)
    assert res[0][4] == res[1][4] and res[0][4][2] == 7          # same global scalars on both ranks; interventions add up
