"""World-size-2 gloo tests (CPU) of the host-side multi-GPU logic: sharding of a restart sweep
across ranks (whole problems per rank, or a block of the restart axis per rank) and the single
collective of the sweep (all-gather of the per-problem winners).  The data path has no other
communication."""
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
    c, i = gather_best(local_cost, local_idx, n_total=n_problems)
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


def _sweep_state(n_problems, n_restarts, seed=0):
    """A synthetic finished sweep [P][R]: obstacle costs (with ties), fulfilled flags, one problem without any
    fulfilled restart, one NaN cost."""
    rng = np.random.default_rng(seed)
    cost = rng.uniform(1, 3, (n_problems, n_restarts)).astype(np.float32)
    ful = rng.uniform(size=(n_problems, n_restarts)) < 0.4
    ful[3] = False
    cost[5, 1] = cost[5, n_restarts - 2] = 0.5
    ful[5, 1] = ful[5, n_restarts - 2] = True                  # a tie across two ranks' blocks: the lower index wins
    cost[7, 2] = np.nan
    return cost, ful


def _expected_winners(cost, ful):
    c = np.where(np.isnan(cost), np.inf, cost)
    key = np.where(ful, c, np.inf)
    key = np.where(np.isinf(key).all(1, keepdims=True), c, key)
    return key.argmin(1)


def _keys_worker(rank, world, port, n_problems, n_restarts, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from irm_motion_planning_b200.batch import decode_keys, encode_keys, gather_best_keys, restart_shard
    cost, ful = _sweep_state(n_problems, n_restarts)
    lo, hi = restart_shard(n_restarts, rank, world)            # this rank's block of the restart axis, every problem
    r = _expected_winners(cost[:, lo:hi], ful[:, lo:hi])       # what fgd_argmin_per_problem computes locally
    p = np.arange(n_problems)
    keys = encode_keys(torch.from_numpy(cost[p, lo + r].copy()), torch.from_numpy(ful[p, lo + r].copy()),
                       torch.from_numpy((p * n_restarts + lo + r).astype(np.int64)))
    win = gather_best_keys(keys)
    if rank == 0:
        c, i, f = decode_keys(win)
        out.put((c.numpy().copy(), i.numpy().copy(), f.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_problems,n_restarts", [(64, 8), (37, 5)])
def test_restart_sharded_sweep_two_ranks_equals_single_process(n_problems, n_restarts):
    """Restart-axis sharding: each rank reduces its block, ONE all-gather of int64 order keys, elementwise min."""
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_keys_worker, args=(r, 2, port, n_problems, n_restarts, out)) for r in range(2)]
    for p in procs:
        p.start()
    c, i, f = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cost, ful = _sweep_state(n_problems, n_restarts)
    r = _expected_winners(cost, ful)
    assert np.array_equal(i, np.arange(n_problems) * n_restarts + r)
    assert np.array_equal(c, cost[np.arange(n_problems), r])
    assert np.array_equal(f, ful[np.arange(n_problems), r])


def test_order_keys_round_trip_and_order():
    from irm_motion_planning_b200.batch import decode_keys, encode_keys
    cost = torch.tensor([0.0, 1.5, 1.5, 2.0, 1.0, float("nan")])
    ful = torch.tensor([True, True, True, True, False, True])
    idx = torch.tensor([5, 7, 3, 1, 0, 2])
    k = encode_keys(cost, ful, idx)
    assert (k >= 0).all()
    order = torch.argsort(k).tolist()
    assert order == [0, 2, 1, 3, 5, 4]          # fulfilled first (NaN last among them), then cost, then index
    c, i, f = decode_keys(k[:5])
    assert torch.equal(c, cost[:5]) and torch.equal(i, idx[:5].to(torch.int32)) and torch.equal(f, ful[:5])


def test_restart_shard_covers_the_axis():
    from irm_motion_planning_b200.batch import restart_shard
    for world in (1, 2, 3, 4, 8):
        spans = [restart_shard(256, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == 256 and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
