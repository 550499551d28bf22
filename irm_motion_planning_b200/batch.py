"""Batched FGD engine: device state, launches, sharding and the restart gather.

One process drives one GPU.  A batch of B independent trajectories lives in
device tensors (alpha [B,T,3], start/goal [B,3], loop state [B,8]+[B,8]); one
persistent kernel launch optimises all of them (``Handle.optimize``).  Across
GPUs the batch is sharded by trajectory index with no data-path communication;
the only collective is a gather of per-problem (best cost, global index) pairs
(``gather_best``) over torch.distributed (NCCL on GPUs, gloo in CPU tests).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np

from . import backend


@dataclass
class BatchResult:
    alpha: "object"          # torch tensor [B,T,3] (device) or numpy (host path)
    fstate: "object"         # [B,8] float32
    istate: "object"         # [B,8] int32

    def _np(self, x):
        return x if isinstance(x, np.ndarray) else x.cpu().numpy()

    @property
    def fulfilled(self):
        return self._np(self.istate)[:, backend.I_FULFILLED].astype(bool)

    @property
    def obstacle_cost(self):
        return self._np(self.fstate)[:, backend.F_TOC]

    @property
    def loss(self):
        return self._np(self.fstate)[:, backend.F_LOSS]

    @property
    def inner_iterations(self):
        return self._np(self.istate)[:, backend.I_INNER_TOTAL]

    @property
    def candidate_evals(self):
        return self._np(self.istate)[:, backend.I_CAND_EVALS]

    @property
    def outer_iterations(self):
        # outer-loop bodies executed (I_OUTER counts lambda escalations)
        is_ = self._np(self.istate)
        return np.maximum(1, is_[:, backend.I_OUTER] + is_[:, backend.I_FULFILLED])

    @property
    def done(self):
        return self._np(self.istate)[:, backend.I_STATUS] == backend.ST_DONE


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of n items for `rank` of `world` (sizes differ by at most 1)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class BatchedFGD:
    """Batch front-end over one ``Trajectory`` (which owns the device handle)."""

    def __init__(self, trajectory, mode: str):
        if mode not in ("bls", "gd"):
            raise ValueError(mode)
        self.trajectory = trajectory
        self.mode = mode
        self.T = trajectory.N_timesteps

    @property
    def handle(self) -> backend.Handle:
        return self.trajectory.handle

    # -- state -------------------------------------------------------------
    def new_state(self, B: int):
        import torch
        return (torch.zeros(B, backend.FSTATE, dtype=torch.float32, device="cuda"),
                torch.zeros(B, backend.ISTATE, dtype=torch.int32, device="cuda"))

    # -- device-resident path ---------------------------------------------
    def optimize_device(self, alpha, start, goal, fstate=None, istate=None, max_launch_iters: int = -1) -> BatchResult:
        """alpha [B,T,3], start/goal [B,3]: contiguous float32 CUDA tensors; alpha is updated in place."""
        B = int(alpha.shape[0])
        if fstate is None:
            fstate, istate = self.new_state(B)
        self.handle.optimize(self.mode, B, alpha, start, goal, fstate, istate, max_launch_iters)
        return BatchResult(alpha, fstate, istate)

    # -- host-buffer path (what the reference-facing call looks like) -----
    def optimize_host(self, alpha: np.ndarray, start: np.ndarray, goal: np.ndarray) -> BatchResult:
        """Host float32 arrays in, host arrays out; H2D and D2H copies happen inside the C-ABI call."""
        alpha = np.ascontiguousarray(alpha, np.float32).reshape(-1, self.T, 3)
        B = alpha.shape[0]
        start = np.ascontiguousarray(np.broadcast_to(np.asarray(start, np.float32).reshape(-1, 3), (B, 3)))
        goal = np.ascontiguousarray(np.broadcast_to(np.asarray(goal, np.float32).reshape(-1, 3), (B, 3)))
        out = np.empty_like(alpha)
        fs = np.empty((B, backend.FSTATE), np.float32)
        is_ = np.empty((B, backend.ISTATE), np.int32)
        self.handle.optimize_host_io(self.mode, B, alpha, out, start, goal, fs, is_)
        return BatchResult(out, fs, is_)

    def optimize_pinned(self, alpha_pin, start_pin, goal_pin, out_alpha_pin, out_f_pin, out_i_pin) -> None:
        """Same as optimize_host but on caller-owned pinned torch tensors (no allocation in the timed path)."""
        B = int(alpha_pin.shape[0])
        self.handle.optimize_host_io(self.mode, B, alpha_pin, out_alpha_pin, start_pin, goal_pin, out_f_pin, out_i_pin)

    # -- restart sweep: local argmin + one gather ---------------------------
    def best_per_problem(self, result: BatchResult, n_problems: int, n_restarts: int, index_offset: int = 0):
        import torch
        cost = torch.empty(n_problems, dtype=torch.float32, device="cuda")
        idx = torch.empty(n_problems, dtype=torch.int32, device="cuda")
        self.handle.argmin_per_problem(n_problems, n_restarts, result.fstate, result.istate, index_offset, cost, idx)
        return cost, idx


def gather_best(cost, idx, group=None):
    """The single collective of a sharded sweep: all-gather the per-problem
    (best cost, global trajectory index) pairs of every rank.  Works on CUDA
    tensors over NCCL and on CPU tensors over gloo.  Returns concatenated
    (cost [P_total], idx [P_total]) ordered by rank (= by problem index for
    contiguous shards)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return cost, idx
    world = dist.get_world_size(group)
    # shards may differ by one problem: pad to the max and trim after the gather
    n = torch.tensor([cost.numel()], dtype=torch.int64, device=cost.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    m = max(sizes)
    packed = torch.zeros(m, 2, dtype=torch.float32, device=cost.device)
    packed[: cost.numel(), 0] = cost
    packed[: cost.numel(), 1] = idx.view(torch.float32) if idx.dtype == torch.int32 else idx.to(torch.int32).view(torch.float32)
    outs = [torch.empty_like(packed) for _ in range(world)]
    dist.all_gather(outs, packed, group=group)
    costs = torch.cat([o[:s, 0] for o, s in zip(outs, sizes)])
    idxs = torch.cat([o[:s, 1].contiguous().view(torch.int32) for o, s in zip(outs, sizes)])
    return costs, idxs
