"""Synthetic workloads = the configurations of BASELINE.json.

Everything is generated on the host with ``numpy.random.default_rng(seed)``:
start/goal inside the joint limits, the reference's straight-line alpha fit per
trajectory, random obstacle sets and random-restart perturbations.
"""
from __future__ import annotations

from dataclasses import dataclass
from types import SimpleNamespace
from typing import Optional

import numpy as np

from .environment import Environment, random_obstacles


def default_args(**over):
    """The reference's argparse defaults as a namespace (main.py:17-98)."""
    from .main import build_parser
    ns = build_parser().parse_args([])
    for k, v in over.items():
        if not hasattr(ns, k):
            raise AttributeError(k)
        setattr(ns, k, v)
    return ns


def sample_start_goal(B: int, rng: np.random.Generator, lo: float = -0.9, hi: float = 1.9):
    start = rng.uniform(lo, hi, size=(B, 3)).astype(np.float32)
    goal = rng.uniform(lo, hi, size=(B, 3)).astype(np.float32)
    return start, goal


@dataclass
class Workload:
    name: str
    mode: str                 # "bls" | "gd"
    args: object              # argparse-like namespace
    obstacles: np.ndarray     # (O,2)
    start: np.ndarray         # (B,3)
    goal: np.ndarray          # (B,3)
    n_problems: int = 0       # C5: problems x restarts layout
    n_restarts: int = 1
    description: str = ""

    @property
    def B(self):
        """trajectories of the whole workload"""
        return self.start.shape[0] * (self.n_restarts if self.n_problems else 1)


def _restart_alpha(traj, start, goal, n_restarts, rng, amp: float = 0.3, restarts=None, problems=None):
    """C5: restart r bends the straight line by a random via-offset
    line + sin(pi c(t)) * delta, delta ~ N(0, amp^2)^3, then fits alpha like initTrajectory.
    restarts = (lo, hi) / problems = (lo, hi): generate only that block of the [P][R] sweep (a rank's shard);
    the via-offsets are drawn for the whole sweep first, so a shard equals the same block of the full sweep."""
    P = start.shape[0]
    T = traj.N_timesteps
    delta = (rng.standard_normal((P, n_restarts, 3)) * amp).astype(np.float32)
    delta[:, 0, :] = 0.0                     # restart 0 = the plain straight line
    r_lo, r_hi = restarts if restarts is not None else (0, n_restarts)
    p_lo, p_hi = problems if problems is not None else (0, P)
    delta = delta[p_lo:p_hi, r_lo:r_hi]
    start, goal = start[p_lo:p_hi], goal[p_lo:p_hi]
    n_r = r_hi - r_lo
    s = np.repeat(start[:, None, :], n_r, 1).reshape(-1, 3)
    g = np.repeat(goal[:, None, :], n_r, 1).reshape(-1, 3)
    bump = np.sin(np.float32(np.pi) * traj.c).astype(np.float32)
    jinv = np.linalg.inv(traj.jac).astype(np.float32)
    B = s.shape[0]
    out = np.empty((B, T, 3), np.float32)
    step = 65536                              # bounded temporaries for the 1 M sweep
    for lo in range(0, B, step):
        hi = min(B, lo + step)
        line = s[lo:hi, None, :] + (g[lo:hi] - s[lo:hi])[:, None, :] * traj.c[None, :, None] \
            + bump[None, :, None] * delta.reshape(-1, 1, 3)[lo:hi]
        rhs = line @ jinv
        sol = np.linalg.solve(traj.km, rhs.transpose(1, 0, 2).reshape(T, (hi - lo) * 3)).astype(np.float32)
        out[lo:hi] = sol.reshape(T, hi - lo, 3).transpose(1, 0, 2)
    return out, s, g


def make_workload(name: str, B: Optional[int] = None, seed: int = 0) -> Workload:
    rng = np.random.default_rng(seed)
    env = Environment()
    if name == "c1":      # main.py defaults, one trajectory, BLS
        return Workload("c1", "bls", default_args(), env.obstacles, env.start_config[None].copy(), env.goal_config[None].copy(),
                        description="main.py defaults: BLS, default arm and scene, 1 trajectory")
    if name == "c2":      # GD fixed step, default scene, 4096 random-init trajectories
        B = 4096 if B is None else B
        s, g = sample_start_goal(B, rng)
        return Workload("c2", "gd", default_args(optimizer_name="gd", max_outer_iteration=1), env.obstacles, s, g,
                        description=f"GD fixed step lr=2e-3 (--max-outer-iteration 1), default scene, {B} random start/goal")
    if name == "c3":      # BLS, T=256, 1024 obstacles
        B = 65536 if B is None else B
        s, g = sample_start_goal(B, rng)
        return Workload("c3", "bls", default_args(n_timesteps=256), random_obstacles(1024, rng), s, g,
                        description=f"BLS, T=256, 1024 random obstacles, {B} trajectories")
    if name == "c4":      # BLS, T=50, capacity 1024 / live 256, swaps every N iterations
        B = 262144 if B is None else B
        s, g = sample_start_goal(B, rng)
        return Workload("c4", "bls", default_args(), random_obstacles(256, rng), s, g,
                        description=f"BLS, T=50, 256 live obstacles swapped every 8 inner iterations, {B} trajectories")
    if name == "c5":      # problems x restarts
        n_restarts = 256
        P = (4096 if B is None else max(1, B // n_restarts))
        s, g = sample_start_goal(P, rng)
        return Workload("c5", "bls", default_args(), env.obstacles, s, g, n_problems=P, n_restarts=n_restarts,
                        description=f"BLS random-restart sweep: {P} problems x {n_restarts} restarts, default scene")
    raise ValueError(name)


def initial_alpha(wl: Workload, traj, seed: int = 0, restarts=None, problems=None):
    """alpha0 [B,T,3] plus per-trajectory start/goal (expanded for C5; `restarts` / `problems` select a
    block of the sweep, layout [problem][restart])."""
    if wl.n_problems:
        return _restart_alpha(traj, wl.start, wl.goal, wl.n_restarts, np.random.default_rng(seed + 1), restarts=restarts,
                              problems=problems)
    return traj.initTrajectory(wl.start, wl.goal), wl.start, wl.goal


def obstacle_swap(k: int, seed: int = 0, lo: int = 192, hi: int = 320) -> np.ndarray:
    """C4: the k-th replacement obstacle set (count varies in [lo, hi])."""
    rng = np.random.default_rng(seed + 1000 + k)
    return random_obstacles(int(rng.integers(lo, hi + 1)), rng)


# algorithmic FLOPs: FMA = 2, divide = 1, sin/cos = 1 each, D = 3
def flops_cost(T, O, joints=1):
    return 12 * T * T + 9 * T * O * joints + 156 * T


def flops_grad(T, O, joints=1):
    return 24 * T * T + 15 * T * O * joints + 272 * T


def flops_total(mode: str, T: int, O: int, inner_total, cand_evals, outer_bodies, joints: int = 1):
    """Reference-algorithm FLOPs consumed by a batch (arrays of per-trajectory counters).
    joints = 3 for the whole-arm obstacle cost (three joint positions per sample)."""
    inner_total = np.asarray(inner_total, np.float64)
    cand = np.asarray(cand_evals, np.float64)
    outer = np.asarray(outer_bodies, np.float64)
    E = 6 * T * T + 18 * T
    if mode == "bls":
        it = inner_total * (flops_grad(T, O, joints) + 15 * T) + cand * flops_cost(T, O, joints)
    else:
        it = inner_total * (flops_grad(T, O, joints) + 6 * T) + cand * flops_cost(T, O, joints) + outer * flops_cost(T, O, joints)
    return float(np.sum(it + outer * (2 * E + 30 * T)))
