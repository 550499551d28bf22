"""Two-rank NCCL test of the restart sweep (config 5) on real GPUs: every rank optimises its block of the restart axis
of every problem (strict math), reduces locally with fgd_argmin_per_problem, the ranks exchange the per-problem order
keys with ONE all-gather, and the winners equal the CPU oracle's argmin over the whole sweep.  Needs two GPUs (skipped
on a single-GPU box; the same logic runs under gloo in tests/test_distributed.py and as sequential shards in
tests/test_gpu_parity.py::test_config5_restart_sweep_pipeline_equals_oracle)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
P, R = 10, 32


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from irm_motion_planning_b200.batch import BatchedFGD, decode_keys, gather_best_keys, restart_shard
    from irm_motion_planning_b200.trajectory import Trajectory
    from irm_motion_planning_b200.workloads import initial_alpha, make_workload
    wl = make_workload("c5", B=P * 256, seed=6)
    wl.n_restarts = R
    tr = Trajectory(wl.args, strict_math=True)
    tr.set_obstacles(wl.obstacles)
    lo, hi = restart_shard(R, rank, world)
    a0, s, g = initial_alpha(wl, tr, 6, restarts=(lo, hi))
    eng = BatchedFGD(tr, "bls")
    res = eng.optimize_device(torch.as_tensor(a0, device="cuda"), torch.as_tensor(s, device="cuda").contiguous(),
                              torch.as_tensor(g, device="cuda").contiguous())
    keys = eng.best_keys(res.fstate, res.istate, P, hi - lo, index_offset=lo, problem_stride=R)
    win = gather_best_keys(keys)
    cost, idx, ful = decode_keys(win)
    if rank == 0:
        out.put((cost.cpu().numpy(), idx.cpu().numpy(), ful.cpu().numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_restart_sweep_two_gpus_nccl_equals_oracle():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    from irm_motion_planning_b200.trajectory import Trajectory
    from irm_motion_planning_b200.workloads import initial_alpha, make_workload
    from oracle import mirror as M
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    cost, idx, ful = out.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    wl = make_workload("c5", B=P * 256, seed=6)
    wl.n_restarts = R
    tr = Trajectory(wl.args, create_handle=False)
    a0, s, g = initial_alpha(wl, tr, 6)
    hp = type("HP", (), dict(vars(wl.args)))()
    hp.n_timesteps = 50
    ca, cfs, cis = M.Mirror(hp, tr.km, tr.dkm, tr.jac, wl.obstacles, "bls").optimize(a0, s, g)
    toc, f = cfs[:, M.F_TOC].reshape(P, R), cis[:, M.I_FULFILLED].reshape(P, R).astype(bool)
    c = np.where(f, toc, np.inf)
    c = np.where(np.isinf(c).all(1, keepdims=True), toc, c)
    r = c.argmin(1)
    assert np.array_equal(idx, np.arange(P) * R + r)
    assert np.array_equal(cost, toc[np.arange(P), r]) and np.array_equal(ful, f[np.arange(P), r])
