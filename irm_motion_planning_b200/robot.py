"""Planar 3R arm parameters (mirror of ``robot.py``, class ``Robot``).

Forward kinematics, the position Jacobian and the four constraint predicates
(robot.py:29-36, 75-113) are evaluated per time sample inside the CUDA kernels.
The small NumPy ``fk``/``jacobian`` below exist only so plotting scripts written
against the reference keep working; the optimiser never calls them.
"""
from __future__ import annotations

import sys

import numpy as np


class Robot:
    def __init__(self, args):
        self.max_joint_velocity = args.max_joint_velocity
        self.min_joint_position = args.min_joint_position
        self.max_joint_position = args.max_joint_position
        self.N_joints = args.n_joints
        self.link_length = np.asarray(args.link_length, dtype=np.float32)
        if self.N_joints != len(self.link_length):
            print("FATAL: n_joints and link_length do not match")
            sys.exit(-1)
        if self.N_joints != 3:
            # the reference reshapes to (-1, 3) and builds a 3x3 J: only 3 joints ever worked
            print("FATAL: only n_joints == 3 is supported (as in the reference implementation)")
            sys.exit(-1)
        self.eps_velocity = args.eps_velocity
        self.eps_distance = args.eps_position

    # plotting helpers (host)
    def fk(self, config):
        ang = np.cumsum(np.asarray(config, np.float32).reshape(-1, 3), axis=1)
        return np.stack((np.cos(ang) @ self.link_length, np.sin(ang) @ self.link_length))

    def jacobian(self, config):
        ang = np.cumsum(np.asarray(config, np.float32).reshape(-1, 3), axis=1)
        sx = -self.link_length * np.sin(ang)
        sy = self.link_length * np.cos(ang)
        rev = lambda m: np.cumsum(m[:, ::-1], axis=1)[:, ::-1]
        return np.stack((rev(sx), rev(sy)))
