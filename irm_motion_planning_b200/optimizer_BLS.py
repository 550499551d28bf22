"""Backtracking-line-search optimiser, GPU-batched.

Mirror of the reference's ``optimizer_BLS.py``: same class name, constructor
``(args)``, attributes ``.env`` / ``.trajectory`` and method ``.optimize()``.
The jitted triple loop (optimizer_BLS.py:126-213) is replaced by one launch of
the persistent sm_100a kernel (``fgd_optimize_bls``); ``optimize()`` keeps the
reference's single-trajectory behaviour, ``optimize_batch()`` is the batched
entry the reference does not have.
"""
from __future__ import annotations

import time

import numpy as np

from .batch import BatchedFGD, BatchResult
from .environment import Environment
from .trajectory import Trajectory

np.set_printoptions(precision=4)


class BacktrackingLineSearchOptimizer:
    MODE = "bls"

    def __init__(self, args, warmup: bool = True):
        self.jitLoop = args.jit_loop
        self.max_inner_iteration = args.max_inner_iteration
        self.max_outer_iteration = args.max_outer_iteration
        self.loop_loss_reduction = args.loop_loss_reduction
        self.lambda_constraint_increase = args.lambda_constraint_increase
        self.lambda_sg_constraint = args.lambda_sg_constraint
        self.lambda_jl_constraint = args.lambda_jl_constraint
        self.lambda_max_cost = args.lambda_max_cost
        self.lambda_reg = args.lambda_reg
        self.bls_max_iter = args.max_bls_iteration
        self.bls_lr_start = args.bls_lr_start
        self.bls_alpha = args.bls_alpha
        self.bls_beta_minus = args.bls_beta_minus
        self.bls_beta_plus = args.bls_beta_plus
        self.extendedVis = args.extended_vis

        self.env = Environment()
        self.trajectory = Trajectory(args, obstacle_capacity=int(getattr(args, "obstacle_capacity", 1024)),
                                     strict_math=bool(getattr(args, "strict_math", False)))
        self.engine = BatchedFGD(self.trajectory, self.MODE)

        # the reference compiles here by running optimize() once; this loads the CUDA module and warms it up
        if warmup:
            t1 = time.time()
            _ = self.optimize()
            t2 = time.time()
            print("setup object, jit-compile took", 1000 * (t2 - t1), "ms")

    # -- reference entry point: one trajectory on the default scene --------
    def optimize(self):
        init_alpha = self.trajectory.initTrajectory(self.env.start_config, self.env.goal_config)
        if self.extendedVis or not self.jitLoop:
            return self.plain_optimize(init_alpha)
        return self.jit_optimize(init_alpha, self.env.obstacles, self.env.start_config, self.env.goal_config)

    def jit_optimize(self, alpha, obstacles, start_config, goal_config):
        """Operator seam of the reference (optimizer_BLS.py:126-127); returns alpha as a CUDA tensor."""
        res = self.optimize_batch(np.asarray(alpha, np.float32)[None], np.asarray(start_config, np.float32)[None],
                                  np.asarray(goal_config, np.float32)[None], obstacles)
        return res.alpha[0]

    def plain_optimize(self, alpha):
        """--jit-loop false / --extended-vis true (optimizer_BLS.py:65-123): the loop advances one inner
        iteration per launch so that ``self.env.obstacles`` is re-read every iteration and the accepted
        iterates can be recorded for trajectory_series.txt."""
        import torch
        a = torch.as_tensor(np.asarray(alpha, np.float32)[None], device="cuda").contiguous()
        s = torch.as_tensor(self.env.start_config[None], device="cuda")
        g = torch.as_tensor(self.env.goal_config[None], device="cuda")
        fs, is_ = self.engine.new_state(1)
        p = [self.trajectory.evaluate(a[0], self.trajectory.km, self.trajectory.jac)] if self.extendedVis else None
        while True:
            self.trajectory.set_obstacles(self.env.obstacles)
            before = is_.cpu().numpy()[0].copy()
            self.engine.optimize_device(a, s, g, fs, is_, max_launch_iters=1)
            after = is_.cpu().numpy()[0]
            from . import backend as _b
            if self.extendedVis and after[_b.I_ACCEPTS] > before[_b.I_ACCEPTS] and after[_b.I_INNER] > 0 \
                    and after[_b.I_OUTER] == before[_b.I_OUTER] and after[_b.I_STATUS] != _b.ST_DONE:
                p.append(self.trajectory.evaluate(a[0], self.trajectory.km, self.trajectory.jac))
            if after[_b.I_STATUS] == _b.ST_DONE:
                break
        return (a[0], p) if self.extendedVis else a[0]

    # -- batched entry -----------------------------------------------------
    def optimize_batch(self, alpha, start, goal, obstacles=None, on_host: bool = False) -> BatchResult:
        """alpha [B,T,3], start/goal [B,3].  NumPy inputs are uploaded; CUDA tensors are used in place.
        ``on_host=True`` goes through the host-buffer C-ABI call and returns NumPy arrays."""
        import torch
        self.trajectory.set_obstacles(self.env.obstacles if obstacles is None else obstacles)
        if on_host:
            return self.engine.optimize_host(alpha, start, goal)
        dev = torch.device("cuda")
        a = torch.as_tensor(alpha, dtype=torch.float32, device=dev).contiguous()
        if isinstance(alpha, torch.Tensor) and a.data_ptr() == alpha.data_ptr():
            a = a.clone()
        s = torch.as_tensor(start, dtype=torch.float32, device=dev).reshape(-1, 3).contiguous()
        g = torch.as_tensor(goal, dtype=torch.float32, device=dev).reshape(-1, 3).contiguous()
        return self.engine.optimize_device(a, s, g)
