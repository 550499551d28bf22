"""RKHS trajectory model: constants on the host, objective on the GPU.

Mirror of the reference's ``trajectory.py`` (class ``Trajectory``).  The
constructor builds ``t, c, km, dkm, jac`` exactly the way the reference does
(trajectory.py:23-48) with NumPy float32; they are inputs of the kernels.
``compute_trajectory_cost``, ``compute_trajectory_cost_g`` and
``constraintsFulfilled`` (trajectory.py:271-297, 129-137) call the CUDA
evaluation kernel through the C ABI (``fgd_eval_cost_grad``) -- for one
trajectory (reference signature) or a batch.
"""
from __future__ import annotations

import numpy as np

from . import backend
from ._prng import normal_key0
from .robot import Robot


def rbf_kernel(x_1, x_2, rbf_var):
    # "variance" is used as a standard deviation, as in the reference
    d = x_1 - x_2
    return np.exp(-(d * d) / (np.float32(2) * rbf_var * rbf_var)).astype(np.float32)


def d_rbf_kernel(x_1, x_2, rbf_var):
    d = x_1 - x_2
    return (d / (rbf_var * rbf_var) * np.exp(-(d * d) / (np.float32(2) * rbf_var * rbf_var))).astype(np.float32)


class Trajectory:
    def __init__(self, args, obstacle_capacity: int = 1024, strict_math: bool = False, jac_stream: str = "legacy",
                 create_handle: bool = True):
        self.robot = Robot(args)
        self.args = args
        self.rbf_var = np.float32(args.rbf_variance)
        self.constraint_violating_dependant_loss = args.constraint_violating_dependant_loss
        self.joint_safety_limit = args.joint_safety_limit
        self.mean_joint_position = 0.5 * (self.robot.max_joint_position + self.robot.min_joint_position)
        self.std_joint_position = 0.5 * (self.robot.max_joint_position - self.mean_joint_position)
        # the reference declares --n-timesteps as float (main.py:33); coerce so 256 works
        self.N_timesteps = int(args.n_timesteps)
        T = self.N_timesteps
        self.t = (np.arange(T, dtype=np.float32) / np.float32(T - 1)).astype(np.float32)
        t = self.t
        # c(0)=0, c(1)=1 with vanishing first and second derivatives at both ends
        self.c = (np.float32(6) * t**5 - np.float32(15) * t**4 + np.float32(10) * t**3).astype(np.float32)
        self.km = self.create_kernel_matrix(rbf_kernel, t, t)
        self.dkm = self.create_kernel_matrix(d_rbf_kernel, t, t)
        self.jac = (np.eye(3, dtype=np.float32)
                    + np.float32(args.jac_gaussian_mean) * normal_key0((3, 3), jac_stream)).astype(np.float32)
        self.lambda_max_cost = float(getattr(args, "lambda_max_cost", 0.5))
        self._handle = None
        self._obs_key = None
        self._capacity = int(obstacle_capacity)
        self._strict = bool(strict_math)
        if create_handle:
            self.handle  # noqa: B018  (fail early and loudly without CUDA / the library)

    # ------------------------------------------------------------------
    def create_kernel_matrix(self, kernel_f, x, x2):
        a, b = np.meshgrid(x, x2)          # 'xy' indexing: a[i, j] = x[j], b[i, j] = x2[i]
        return kernel_f(a, b, self.rbf_var)

    @property
    def handle(self) -> backend.Handle:
        if self._handle is None:
            hp = _HyperView(self.args, self.N_timesteps)
            cfg = backend.make_config(hp, self.km, self.dkm, self.jac, self._capacity, self._strict)
            self._handle = backend.Handle(cfg)
        return self._handle

    def set_obstacles(self, obstacles):
        """Upload the obstacle set (async memcpy into the double-buffered device array).  Host arrays are
        uploaded only when their CONTENT differs from the last upload made through this method (and nobody
        called ``handle.set_obstacles`` directly in between); tensors are always uploaded (<= 8 KB, async)."""
        import torch
        if isinstance(obstacles, torch.Tensor):
            self.handle.set_obstacles(obstacles)
            self._obs_key = None
            return
        arr = np.ascontiguousarray(np.asarray(obstacles, dtype=np.float32).reshape(-1, 2))
        key = (arr.tobytes(), self.handle.obstacle_generation)
        if key != self._obs_key:
            self.handle.set_obstacles(arr)
            self._obs_key = (key[0], self.handle.obstacle_generation)

    # ------------------------------------------------------------------
    def evaluate(self, alpha, kernel_matrix, jac):
        """K @ alpha @ J on the host (output formatting, main.py:145); the optimiser
        evaluates trajectories inside the kernels."""
        import torch
        if isinstance(alpha, torch.Tensor):
            alpha = alpha.detach().cpu().numpy()
        return (np.asarray(kernel_matrix, np.float32) @ np.asarray(alpha, np.float32)) @ np.asarray(jac, np.float32)

    def initTrajectory(self, start_config, goal_config):
        """Straight line in joint space fitted by an FP32 LU solve, trajectory.py:73-78.
        Accepts (3,) or batched (B,3) start/goal; returns (T,3) or (B,T,3)."""
        start = np.asarray(start_config, np.float32)
        goal = np.asarray(goal_config, np.float32)
        single = start.ndim == 1
        start, goal = start.reshape(-1, 3), goal.reshape(-1, 3)
        line = start[:, None, :] + (goal - start)[:, None, :] * self.c[None, :, None]      # (B,T,3)
        rhs = line @ np.linalg.inv(self.jac).astype(np.float32)
        B, T = rhs.shape[0], self.N_timesteps
        sol = np.linalg.solve(self.km, rhs.transpose(1, 0, 2).reshape(T, B * 3)).astype(np.float32)
        alpha = np.ascontiguousarray(sol.reshape(T, B, 3).transpose(1, 0, 2))
        return alpha[0] if single else alpha

    def init_basis(self):
        """u = K^-1 1, w = K^-1 c (one FP32 LU, the solver initTrajectory uses) and J^-1: the
        batch-shared pieces of the rank-2 form of trajectory.py:73-78."""
        rhs = np.stack([np.ones_like(self.c), self.c], axis=1).astype(np.float32)
        uw = np.linalg.solve(self.km, rhs).astype(np.float32)
        return np.ascontiguousarray(uw[:, 0]), np.ascontiguousarray(uw[:, 1]), np.linalg.inv(self.jac).astype(np.float32)

    def initTrajectoryDevice(self, start, goal):
        """initTrajectory for start/goal already on the GPU: (B,3) CUDA tensors -> alpha (B,T,3) CUDA
        tensor, no host round trip.  Rank-2 restatement (include/fgd_b200.h: fgd_init_trajectory);
        the host ``initTrajectory`` stays the reference-faithful per-trajectory LU solve."""
        import torch
        if not getattr(self, "_init_basis_set", False):
            self.handle.set_init_basis(*self.init_basis())
            self._init_basis_set = True
        start = start.to(torch.float32).reshape(-1, 3).contiguous()
        goal = goal.to(torch.float32).reshape(-1, 3).contiguous()
        alpha = torch.empty(start.shape[0], self.N_timesteps, 3, dtype=torch.float32, device=start.device)
        self.handle.init_trajectory(int(start.shape[0]), start, goal, alpha)
        return alpha

    # ------------------------------------------------------------------
    def _eval(self, alpha, obstacles, start_config, goal_config, lam_sg, lam_jl, lam_max, want):
        import torch
        dev = torch.device("cuda")
        a = torch.as_tensor(np.asarray(alpha, np.float32) if not isinstance(alpha, torch.Tensor) else alpha,
                            dtype=torch.float32, device=dev).reshape(-1, self.N_timesteps, 3).contiguous()
        B = a.shape[0]
        s = torch.as_tensor(np.asarray(start_config, np.float32) if not isinstance(start_config, torch.Tensor) else start_config,
                            dtype=torch.float32, device=dev).reshape(-1, 3).expand(B, 3).contiguous()
        g = torch.as_tensor(np.asarray(goal_config, np.float32) if not isinstance(goal_config, torch.Tensor) else goal_config,
                            dtype=torch.float32, device=dev).reshape(-1, 3).expand(B, 3).contiguous()
        if obstacles is not None:
            self.set_obstacles(obstacles)
        out = {}
        if "loss" in want:
            out["loss"] = torch.empty(B, device=dev)
        if "toc" in want:
            out["toc"] = torch.empty(B, device=dev)
        if "grad" in want:
            out["grad"] = torch.empty(B, self.N_timesteps, 3, device=dev)
        if "q" in want:
            out["q"] = torch.empty(B, self.N_timesteps, 3, device=dev)
        if "v" in want:
            out["v"] = torch.empty(B, self.N_timesteps, 3, device=dev)
        if "fulfilled" in want:
            out["fulfilled"] = torch.empty(B, dtype=torch.int32, device=dev)
        self.handle.eval(B, a, s, g, float(lam_sg), float(lam_jl), float(lam_max), **out)
        return out

    def compute_trajectory_cost(self, alpha, obstacles, start_config, goal_config, lambda_sg_constraint,
                                lambda_jl_constraint, lambda_max_cost):
        out = self._eval(alpha, obstacles, start_config, goal_config, lambda_sg_constraint, lambda_jl_constraint,
                         lambda_max_cost, ("loss",))["loss"]
        return out[0].item() if np.ndim(alpha) == 2 else out

    def compute_trajectory_cost_g(self, alpha, obstacles, start_config, goal_config, lambda_sg_constraint,
                                  lambda_jl_constraint, lambda_max_cost):
        out = self._eval(alpha, obstacles, start_config, goal_config, lambda_sg_constraint, lambda_jl_constraint,
                         lambda_max_cost, ("grad",))["grad"]
        return out[0] if np.ndim(alpha) == 2 else out

    def constraintsFulfilled(self, alpha, start_config, goal_config):
        out = self._eval(alpha, None, start_config, goal_config, 0.0, 0.0, -1.0, ("fulfilled",))["fulfilled"]
        return bool(out[0].item()) if np.ndim(alpha) == 2 else out.bool()

    def constraintsFulfilledVerbose(self, alpha, start_config, goal_config, verbose=True):
        """Same report lines as the reference (trajectory.py:140-180), from the device q and v."""
        out = self._eval(alpha, None, start_config, goal_config, 0.0, 0.0, -1.0, ("q", "v", "fulfilled"))
        q, v = out["q"][0].cpu().numpy(), out["v"][0].cpu().numpy()
        s, g = np.asarray(start_config, np.float32).reshape(3), np.asarray(goal_config, np.float32).reshape(3)
        r = self.robot
        n = np.linalg.norm
        checks = (
            (n(q[0] - s) < r.eps_distance and n(q[-1] - g) < r.eps_distance,
             "ok start goal position", "violated start goal position", (n(q[0] - s), n(q[-1] - g))),
            (n(v[0]) < r.eps_velocity and n(v[-1]) < r.eps_velocity,
             "ok start goal velocity", "violated start goal velocity", (n(v[0]), n(v[-1]))),
            (q.max() <= r.max_joint_position and q.min() >= r.min_joint_position,
             "ok joint limit with", "joint limit exceeded with", (q.max(), q.min())),
            (np.abs(v).max() <= r.max_joint_velocity,
             "ok velocity limit with", "joint velocity exceeded with", (np.abs(v).max(),)),
        )
        result = True
        for ok, msg_ok, msg_bad, vals in checks:
            if verbose:
                print(msg_ok if ok else msg_bad, *vals)
            result = result and bool(ok)
        return result


class _HyperView:
    """argparse namespace with n_timesteps coerced to int."""

    def __init__(self, args, T):
        self.__dict__.update(vars(args))
        self.n_timesteps = T
