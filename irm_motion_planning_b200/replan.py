"""Warm-start re-planning on a stream of scene updates.

The reference plans one trajectory from a cold start (``optimizer_BLS.py:58``:
``initTrajectory`` then ``jit_optimize``) and its blog names 50 Hz re-planning
as the goal (``DevBlog-Theme/blog-post.html:350-354``, ``README.md:25``).  Here
the device handle, the coefficient buffer and the loop-state tensors persist
between plans: every ``update(obstacles)`` uploads the new obstacle set with
``fgd_set_obstacles_async`` (no recompilation, no re-creation of the handle),
clears the loop state and re-runs the optimiser from the PREVIOUS solution --
one memset + one async copy + one kernel launch per plan, all on one stream.

Semantics of one plan = the reference's ``jit_optimize(alpha_prev, obstacles,
start, goal)``: penalty weights, step size and counters start fresh, only the
coefficients are carried over (that is what "warm start" means for a
penalty-method loop; the mirror oracle reproduces it in ``tests/``).
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from .batch import BatchedFGD, BatchResult


class WarmStartPlanner:
    """B persistent trajectories re-optimised whenever the scene changes."""

    def __init__(self, optimizer, start, goal, alpha0=None):
        """optimizer: a ``BacktrackingLineSearchOptimizer`` / ``GradientDescentOptimizer`` (its
        ``.trajectory`` owns the device handle); start/goal: [B,3] (or [3]) joint configurations;
        alpha0: optional [B,T,3] initial coefficients (default: the reference's straight-line fit)."""
        import torch
        self.trajectory = optimizer.trajectory
        self.engine: BatchedFGD = optimizer.engine
        start = np.ascontiguousarray(np.asarray(start, np.float32).reshape(-1, 3))
        goal = np.ascontiguousarray(np.asarray(goal, np.float32).reshape(-1, 3))
        if alpha0 is None:
            alpha0 = self.trajectory.initTrajectory(start, goal)
        alpha0 = np.asarray(alpha0, np.float32).reshape(-1, self.trajectory.N_timesteps, 3)
        self.B = alpha0.shape[0]
        dev = torch.device("cuda")
        self.alpha = torch.as_tensor(alpha0, device=dev).contiguous().clone()
        self.start = torch.as_tensor(np.broadcast_to(start, (self.B, 3)).copy(), device=dev)
        self.goal = torch.as_tensor(np.broadcast_to(goal, (self.B, 3)).copy(), device=dev)
        self.fstate, self.istate = self.engine.new_state(self.B)
        self.plans = 0

    def update(self, obstacles, start=None, goal=None) -> BatchResult:
        """New obstacle set (host array [O,2] -- pinned for a truly asynchronous copy -- or CUDA tensor) and
        optionally new start / goal configurations; returns the re-optimised batch (device tensors, valid
        after the stream is synchronised).  alpha is updated in place and seeds the next plan."""
        import torch
        self.trajectory.set_obstacles(obstacles)
        if start is not None:
            self.start.copy_(torch.as_tensor(np.asarray(start, np.float32).reshape(-1, 3)).expand(self.B, 3), non_blocking=True)
        if goal is not None:
            self.goal.copy_(torch.as_tensor(np.asarray(goal, np.float32).reshape(-1, 3)).expand(self.B, 3), non_blocking=True)
        self.fstate.zero_()
        self.istate.zero_()
        res = self.engine.optimize_device(self.alpha, self.start, self.goal, self.fstate, self.istate)
        self.plans += 1
        return res

    def trajectory_points(self, index: Optional[int] = None):
        """q(t) = K alpha J of the current plan(s): [T,3] for one index, [B,T,3] otherwise."""
        q = self.trajectory.evaluate(self.alpha, self.trajectory.km, self.trajectory.jac)
        return q[index] if index is not None else q
