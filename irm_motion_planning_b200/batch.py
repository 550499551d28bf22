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
        self._side = None          # side stream of the live-obstacle publisher

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

    # -- dynamic environment: one launch, obstacle sets published while it runs ---------------
    def optimize_live(self, alpha, start, goal, fstate, istate, obstacle_sets, poll_every: int = 8, period_us: float = 200.0,
                      switch_log=None, max_sets: Optional[int] = None) -> int:
        """Config 4 without relaunches: ONE persistent launch on the current stream; meanwhile this thread publishes
        ``obstacle_sets[k % len]`` (host arrays or pinned tensors) every ``period_us`` on a side stream with
        ``fgd_set_obstacles_async`` until the launch has finished (or ``max_sets`` were published).  Every team polls the
        generation counter when it picks a trajectory up and every ``poll_every`` inner iterations.  Returns the number
        of sets published.  The call returns when the kernel has finished (the host is the publisher)."""
        import time
        import torch
        B = int(alpha.shape[0])
        h = self.handle
        if self._side is None:
            self._side = torch.cuda.Stream()
        done = torch.cuda.Event()
        h.optimize_live(self.mode, B, alpha, start, goal, fstate, istate, poll_every, switch_log)
        done.record()
        k, side = 0, self._side.cuda_stream
        t_next = time.perf_counter()
        while not done.query() and (max_sets is None or k < max_sets):
            now = time.perf_counter()
            if now < t_next:
                continue
            t_next = now + period_us * 1e-6
            k += 1
            h.set_obstacles(obstacle_sets[k % len(obstacle_sets)], stream=side)
        done.synchronize()
        return k

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
    def best_per_problem(self, result: BatchResult, n_problems: int, n_restarts: int, index_offset: int = 0,
                         problem_stride: int = 0):
        """(cost [P] f32, global index [P] i32) of the best local restart of every problem."""
        import torch
        cost = torch.empty(n_problems, dtype=torch.float32, device="cuda")
        idx = torch.empty(n_problems, dtype=torch.int32, device="cuda")
        self.handle.argmin_per_problem(n_problems, n_restarts, result.fstate, result.istate, index_offset, cost, idx,
                                       problem_stride=problem_stride)
        return cost, idx

    def best_keys(self, fstate, istate, n_problems: int, n_restarts: int, index_offset: int = 0, problem_stride: int = 0,
                  out=None):
        """Order keys [P] int64 of the best local restart of every problem (see ``decode_keys``)."""
        import torch
        if out is None:
            out = torch.empty(n_problems, dtype=torch.int64, device="cuda")
        self.handle.argmin_per_problem(n_problems, n_restarts, fstate, istate, index_offset, best_key=out,
                                       problem_stride=problem_stride)
        return out


# ---------------------------------------------------------------------------
# restart sweep across ranks
# ---------------------------------------------------------------------------

def restart_shard(n_restarts: int, rank: int, world: int) -> Tuple[int, int]:
    """Restarts [lo, hi) of EVERY problem that `rank` optimises.  Sharding the restart axis (not the problem
    axis) gives every rank the same mix of easy and hard problems: the per-rank work differs by the sampling
    noise of single trajectories (~0.2 % at 131 072 per rank) instead of that of whole problems (~1.5 %)."""
    return shard_range(n_restarts, rank, world)


def encode_keys(cost, fulfilled, index):
    """Host/torch restatement of the argmin kernel's order key: (unfulfilled << 62) | (cost bits << 31) | index."""
    import torch
    cb = cost.to(torch.float32).contiguous().view(torch.int32).to(torch.int64)
    cb = torch.where(cost >= 0, cb, torch.full_like(cb, 0x7FFFFFFF))
    unful = (~fulfilled.to(torch.bool)).to(torch.int64)
    return (unful << 62) | (cb << 31) | index.to(torch.int64)


def decode_keys(keys):
    """keys [P] int64 -> (cost f32 [P], global index i32 [P], fulfilled bool [P])."""
    import torch
    idx = (keys & 0x7FFFFFFF).to(torch.int32)
    cost = ((keys >> 31) & 0x7FFFFFFF).to(torch.int32).view(torch.float32)
    return cost, idx, ((keys >> 62) & 1) == 0


def gather_best_keys(keys, group=None):
    """THE collective of a restart sweep whose ranks hold different restarts of the same problems: one
    all-gather of the per-problem order keys (8 B per problem and rank), reduced locally with an elementwise
    minimum.  CUDA tensors over NCCL, CPU tensors over gloo.  Returns the winning keys [P] (on every rank)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return keys
    world = dist.get_world_size(group)
    out = torch.empty(world * keys.numel(), dtype=torch.int64, device=keys.device)
    dist.all_gather_into_tensor(out, keys.contiguous(), group=group)
    return out.view(world, keys.numel()).min(dim=0).values


def gather_best(cost, idx, n_total=None, group=None):
    """The collective of a sweep whose ranks hold whole problems (contiguous shards, ``shard_range``): ONE
    all-gather of the per-problem (best cost, global trajectory index) pairs.  The shard sizes follow from
    ``shard_range(n_total, rank, world)`` on every rank - nothing is exchanged to learn them.  Works on CUDA
    tensors over NCCL and on CPU tensors over gloo.  Returns (cost [n_total], idx [n_total]) by problem index."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return cost, idx
    world = dist.get_world_size(group)
    if n_total is None:
        n_total = cost.numel() * world              # equal shards
    sizes = [hi - lo for lo, hi in (shard_range(n_total, r, world) for r in range(world))]
    assert sizes[dist.get_rank(group)] == cost.numel(), "shard size does not follow shard_range(n_total, rank, world)"
    m = max(sizes)
    packed = torch.zeros(m, 2, dtype=torch.int32, device=cost.device)       # integer container: no float canonicalisation
    packed[: cost.numel(), 0] = cost.to(torch.float32).contiguous().view(torch.int32)
    packed[: cost.numel(), 1] = idx.to(torch.int32)
    out = torch.empty(world * m, 2, dtype=torch.int32, device=cost.device)
    dist.all_gather_into_tensor(out, packed, group=group)
    out = out.view(world, m, 2)
    costs = torch.cat([out[r, :s, 0] for r, s in enumerate(sizes)]).contiguous().view(torch.float32)
    idxs = torch.cat([out[r, :s, 1] for r, s in enumerate(sizes)])
    return costs, idxs
