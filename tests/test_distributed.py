"""World-size-2 gloo tests (CPU) of the host-side multi-GPU logic: contiguous sharding of
problems across ranks and the single collective of a restart sweep (gather of per-problem
(best cost, global trajectory index)).  The data path has no other communication."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_problems, n_restarts, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from irm_motion_planning_b200.batch import gather_best, shard_range
    rng = np.random.default_rng(0)                      # every rank draws the same global table
    cost = rng.uniform(1, 3, (n_problems, n_restarts)).astype(np.float32)
    ful = rng.uniform(size=(n_problems, n_restarts)) < 0.4
    lo, hi = shard_range(n_problems, rank, world)       # whole problems per rank
    key = np.where(ful[lo:hi], cost[lo:hi], np.inf)
    key = np.where(np.isinf(key).all(1, keepdims=True), cost[lo:hi], key)
    r = key.argmin(1)
    local_cost = torch.from_numpy(cost[lo:hi][np.arange(hi - lo), r].copy())
    local_idx = torch.from_numpy(((np.arange(lo, hi) * n_restarts) + r).astype(np.int32))
    c, i = gather_best(local_cost, local_idx)
    if rank == 0:
        out.put((c.numpy().copy(), i.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_problems,n_restarts", [(64, 8), (37, 5)])
def test_gather_best_two_ranks_equals_single_process(n_problems, n_restarts):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_problems, n_restarts, out)) for r in range(2)]
    for p in procs:
        p.start()
    c, i = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(0)
    cost = rng.uniform(1, 3, (n_problems, n_restarts)).astype(np.float32)
    ful = rng.uniform(size=(n_problems, n_restarts)) < 0.4
    key = np.where(ful, cost, np.inf)
    key = np.where(np.isinf(key).all(1, keepdims=True), cost, key)
    r = key.argmin(1)
    assert np.array_equal(i, np.arange(n_problems) * n_restarts + r)
    assert np.array_equal(c, cost[np.arange(n_problems), r])


def test_gather_best_is_identity_without_process_group():
    from irm_motion_planning_b200.batch import gather_best
    c, i = torch.rand(5), torch.arange(5, dtype=torch.int32)
    c2, i2 = gather_best(c, i)
    assert c2 is c and i2 is i
