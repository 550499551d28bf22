"""Scene description: start / goal configuration and point obstacles.

Mirrors ``environment.py`` of the reference (class ``Environment``,
environment.py:11-29).  The obstacle *cost* functions of that file
(compute_cost / compute_cost_vg, environment.py:32-58) run on the GPU inside the
FGD kernels (csrc/fgd_device.cuh, cost_phase); this module only holds the data.
"""
from __future__ import annotations

import numpy as np


class Environment:
    def __init__(self, start_config=None, goal_config=None, obstacles=None):
        self.start_config = np.array([0.0, 0.0, 0.0] if start_config is None else start_config, dtype=np.float32)
        self.goal_config = np.array([1.2, 0.8, 0.3] if goal_config is None else goal_config, dtype=np.float32)
        if obstacles is None:
            # the reference's integer lattice scene, promoted to float32 on use
            obstacles = [(2, -3), (-2, 2), (3, 3), (-1, -2), (-2, 1), (-1, -1), (-2, -3), (-2, 0), (1, 3), (3, 2), (2, 3)]
        self.obstacles = np.asarray(obstacles, dtype=np.float32).reshape(-1, 2)


def random_obstacles(n: int, rng: np.random.Generator, extent: float = 4.0, keep_out: float = 0.5) -> np.ndarray:
    """n obstacles ~ U(-extent, extent)^2, none within `keep_out` of the arm base."""
    out = np.empty((0, 2), np.float32)
    while len(out) < n:
        c = rng.uniform(-extent, extent, size=(2 * n, 2)).astype(np.float32)
        out = np.concatenate([out, c[np.hypot(c[:, 0], c[:, 1]) >= keep_out]])
    return np.ascontiguousarray(out[:n])
