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
