"""Fixed-step gradient-descent optimiser, GPU-batched.

Mirror of the reference's ``optimizer_GD.py`` (class ``GradientDescentOptimizer``):
single-level loop when ``--max-outer-iteration 1`` (optimizer_GD.py:68-97), the
dual penalty loop with the per-outer-iteration ``--gd-lr`` table otherwise
(optimizer_GD.py:172-232).  Both run in the persistent kernel (``fgd_optimize_gd``).
"""
from __future__ import annotations

import sys
import time

import numpy as np

from .batch import BatchedFGD
from .environment import Environment
from .optimizer_BLS import BacktrackingLineSearchOptimizer
from .trajectory import Trajectory

np.set_printoptions(precision=4)


class GradientDescentOptimizer(BacktrackingLineSearchOptimizer):
    MODE = "gd"

    def __init__(self, args, warmup: bool = True):
        self.jitLoop = args.jit_loop
        self.dualOptimization = args.max_outer_iteration > 1
        self.max_inner_iteration = args.max_inner_iteration
        self.max_outer_iteration = args.max_outer_iteration
        self.loop_loss_reduction = args.loop_loss_reduction
        self.lambda_constraint_increase = args.lambda_constraint_increase
        self.lambda_sg_constraint = args.lambda_sg_constraint
        self.lambda_jl_constraint = args.lambda_jl_constraint
        self.lambda_max_cost = args.lambda_max_cost
        self.lambda_reg = args.lambda_reg
        self.extendedVis = args.extended_vis

        if self.max_outer_iteration > len(args.gd_lr):
            print("FATAL: max_outer_iteration and dual_lr do not match")
            sys.exit(-1)
        self.dual_lr = np.asarray(args.gd_lr, dtype=np.float32)
        self.lr = self.dual_lr[0]

        self.env = Environment()
        self.trajectory = Trajectory(args, obstacle_capacity=int(getattr(args, "obstacle_capacity", 1024)),
                                     strict_math=bool(getattr(args, "strict_math", False)))
        self.engine = BatchedFGD(self.trajectory, self.MODE)

        if warmup:
            t1 = time.time()
            self.optimize()
            t2 = time.time()
            print("setup object, jit-compile took", 1000 * (t2 - t1), "ms")

    # reference-named aliases of the operator seam
    def jit_dual_optimize(self, alpha, obstacles, start_config, goal_config):
        return self.jit_optimize(alpha, obstacles, start_config, goal_config)

    def dual_optimize(self, alpha):
        return self.plain_optimize(alpha)
