"""NumPy restatement of the reference's functional-gradient-descent path.

TEST INFRASTRUCTURE ONLY.  Nothing in the shipped package may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do.  It is the *checker*, never the
thing measured as the product.

What it is: a line-by-line CPU restatement (NumPy, FP32 by default, FP64 on
request) of the arithmetic in the reference files

* ``trajectory.py``  (kernels :14-19, constants :23-42, evaluate :63-65,
  initTrajectory :73-78, obstacle cost :81-126, constraints :129-137,
  penalties :183-268, total cost :271-281, alpha-gradient :284-297)
* ``robot.py``       (fk :29-36, jacobian :75-87, predicates :90-113)
* ``environment.py`` (scene :14-29, compute_cost :32-43, compute_cost_vg :46-58)
* ``optimizer_BLS.py`` (plain loop :65-123, which is arithmetic-identical to
  the jitted loop :126-213)
* ``optimizer_GD.py`` (single level :68-119, dual :122-232)

Pinning status (see tests/test_oracle_golden.py and DESIGN.md):
the reference's executor (jax / jaxlib, unpinned in requirements.txt:1-4) is
NOT installed in this image, so the reference itself cannot be run here.  The
oracle is pinned against the only result artefacts the reference ships,
``visualization/trajectory_series.txt`` (146 iterates of one BLS run) and
``visualization/trajectory_result.txt``: every one of the 145 golden steps is
reproduced in direction (cosine) and the golden step lengths follow the BLS
learning-rate recurrence; final costs and constraint verdict match the blog's
table.  Anything tighter than that (ulp-level agreement with XLA:CPU) is
**parity unpinned** and is stated as such.

The matrix ``J`` (trajectory.py:42) comes from ``jax.random.normal(PRNGKey(0))``;
it is rebuilt here with a NumPy threefry2x32 (known-answer vectors in the tests).
"""
from __future__ import annotations

import dataclasses
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np

# --------------------------------------------------------------------------
# Hyper-parameters: the defaults of the reference CLI (main.py:17-98)
# --------------------------------------------------------------------------


@dataclass
class Hyper:
    n_timesteps: int = 50              # main.py:33 (float there; only int works)
    rbf_variance: float = 0.1          # main.py:35
    jac_gaussian_mean: float = 0.15    # main.py:37
    max_inner_iteration: int = 200     # main.py:41
    loop_loss_reduction: float = 1e-3  # main.py:43
    max_outer_iteration: int = 10      # main.py:47
    lambda_constraint_increase: float = 10  # main.py:49
    lambda_sg_constraint: float = 0.5  # main.py:52
    lambda_jl_constraint: float = 0.1  # main.py:54
    eps_position: float = 0.01         # main.py:57
    eps_velocity: float = 0.01         # main.py:59
    lambda_max_cost: float = 0.5       # main.py:63
    lambda_reg: float = 1e-4           # main.py:65
    constraint_violating_dependant_loss: bool = True  # main.py:67
    joint_safety_limit: float = 0.98   # main.py:69
    max_bls_iteration: int = 20        # main.py:73
    bls_lr_start: float = 0.2          # main.py:75
    bls_alpha: float = 0.01            # main.py:77
    bls_beta_plus: float = 1.2         # main.py:79
    bls_beta_minus: float = 0.5        # main.py:81
    gd_lr: List[float] = field(default_factory=lambda: [2e-3, 1e-4, 1e-5, 1e-6, 1e-7,
                                                        1e-8, 1e-8, 1e-8, 1e-8, 1e-8])  # main.py:85
    n_joints: int = 3                  # main.py:89
    link_length: List[float] = field(default_factory=lambda: [1.5, 1.0, 0.5])  # main.py:91
    max_joint_velocity: float = 7      # main.py:93
    max_joint_position: float = 2      # main.py:95
    min_joint_position: float = -1     # main.py:97
    # extension named by the reference's blog (DevBlog-Theme/blog-post.html:505-513), not in its code:
    # the obstacle cost of a sample is summed over ALL joint positions fk_j (robot.py:39-72)
    whole_arm_cost: bool = False


# --------------------------------------------------------------------------
# jax.random.normal(PRNGKey(0), (3, 3)) without jax            trajectory.py:42
# --------------------------------------------------------------------------

_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))


def threefry2x32(key: Tuple[int, int], x0, x1):
    """Threefry-2x32, 20 rounds (Salmon et al. 2011) as used by jax.random."""
    x0 = np.array(x0, dtype=np.uint32).copy()
    x1 = np.array(x1, dtype=np.uint32).copy()
    k0, k1 = np.uint32(key[0]), np.uint32(key[1])
    ks = (k0, k1, np.uint32(k0 ^ k1 ^ np.uint32(0x1BD11BDA)))
    with np.errstate(over="ignore"):
        x0 = x0 + ks[0]
        x1 = x1 + ks[1]
        for rnd in range(5):
            for r in _ROT[rnd % 2]:
                x0 = x0 + x1
                x1 = (x1 << np.uint32(r)) | (x1 >> np.uint32(32 - r))
                x1 = x1 ^ x0
            x0 = x0 + ks[(rnd + 1) % 3]
            x1 = x1 + ks[(rnd + 2) % 3] + np.uint32(rnd + 1)
    return x0, x1


def _bits_to_normal(bits: np.ndarray) -> np.ndarray:
    """uint32 -> N(0,1) float32 the way jax.random.normal does: uniform on
    (-1, 1) from the 23 mantissa bits, then sqrt(2) * erfinv."""
    from scipy.special import erfinv
    f = ((bits >> np.uint32(9)) | np.uint32(0x3F800000)).view(np.float32) - np.float32(1.0)
    lo = np.nextafter(np.float32(-1.0), np.float32(0.0), dtype=np.float32)
    hi = np.float32(1.0)
    u = np.maximum(lo, f * (hi - lo) + lo).astype(np.float32)
    return (np.float32(np.sqrt(2.0)) * erfinv(u.astype(np.float64)).astype(np.float32)).astype(np.float32)


def jax_normal_3x3(stream: str = "legacy") -> np.ndarray:
    """The (3,3) standard-normal draw of ``jax.random.normal(PRNGKey(0), (3,3))``.

    ``legacy``  = jax < 0.5 (jax_threefry_partitionable=False): 9 counters padded
                  to 10, split in halves, one threefry call, outputs concatenated.
                  The reference's golden files were produced with this stream.
    ``partitionable`` = jax >= 0.5 default: per element i, threefry(key,(0,i)),
                  bits = y0 ^ y1.
    """
    if stream == "legacy":
        cnt = np.arange(10, dtype=np.uint32)
        cnt[9] = 0  # odd size is padded with a zero counter
        y0, y1 = threefry2x32((0, 0), cnt[:5], cnt[5:])
        bits = np.concatenate([y0, y1])[:9]
    elif stream == "partitionable":
        y0, y1 = threefry2x32((0, 0), np.zeros(9, np.uint32), np.arange(9, dtype=np.uint32))
        bits = y0 ^ y1
    else:
        raise ValueError(stream)
    return _bits_to_normal(bits).reshape(3, 3)


def make_jac(mean: float, stream: str = "legacy", dtype=np.float32) -> np.ndarray:
    """trajectory.py:42  ``eye(3) + mean * normal``."""
    return (np.eye(3, dtype=np.float32) + np.float32(mean) * jax_normal_3x3(stream)).astype(dtype)


# --------------------------------------------------------------------------
# environment.py
# --------------------------------------------------------------------------

DEFAULT_OBSTACLES = np.array([[2, -3], [-2, 2], [3, 3], [-1, -2], [-2, 1], [-1, -1],
                              [-2, -3], [-2, 0], [1, 3], [3, 2], [2, 3]], dtype=np.int32)  # environment.py:17-29
DEFAULT_START = np.array([0.0, 0.0, 0.0])   # environment.py:14
DEFAULT_GOAL = np.array([1.2, 0.8, 0.3])    # environment.py:15


def compute_cost(f, obstacles, dt):
    """environment.py:32-43.  f (2,T), obstacles (O,2) -> (T,)"""
    fo = f[:, :, None] - obstacles.T[:, None, :].astype(dt)
    n = np.sum(np.square(fo), axis=0, dtype=dt)
    return np.sum(dt(0.8) / (dt(0.5) + dt(0.5) * n), axis=1, dtype=dt)


def compute_cost_vg(f, obstacles, dt):
    """environment.py:46-58 -> ((T,), (2,T))"""
    fo = f[:, :, None] - obstacles.T[:, None, :].astype(dt)
    n = np.sum(np.square(fo), axis=0, dtype=dt)
    den = dt(0.5) + dt(0.5) * n
    cost_v = np.sum(dt(0.8) / den, axis=1, dtype=dt)
    cost_g = np.sum(dt(-0.8) * fo / np.square(den)[None, :, :], axis=2, dtype=dt)
    return cost_v, cost_g


# --------------------------------------------------------------------------
# robot.py
# --------------------------------------------------------------------------


class RobotModel:
    def __init__(self, hp: Hyper, dt):
        if hp.n_joints != len(hp.link_length):
            raise ValueError("n_joints and link_length do not match")  # robot.py:21-23
        if hp.n_joints != 3:
            raise ValueError("the reference hard-wires 3 joints (robot.py:31,77; trajectory.py:42)")
        self.dt = dt
        self.ll = np.asarray(hp.link_length, dtype=dt)
        self.vmax = dt(hp.max_joint_velocity)
        self.qmax = dt(hp.max_joint_position)
        self.qmin = dt(hp.min_joint_position)
        self.eps_v = dt(hp.eps_velocity)
        self.eps_p = dt(hp.eps_position)

    def fk(self, q):                                    # robot.py:29-36
        c = np.cumsum(q.reshape(-1, 3), axis=1, dtype=self.dt)
        return np.stack((self.ll @ np.cos(c).T, self.ll @ np.sin(c).T))

    def jacobian(self, q):                              # robot.py:75-87
        c = np.cumsum(q.reshape(-1, 3), axis=1, dtype=self.dt)
        x = -(self.ll * np.sin(c))
        rx = x + np.sum(x, axis=1, dtype=self.dt)[:, None] - np.cumsum(x, axis=1, dtype=self.dt)
        y = self.ll * np.cos(c)
        ry = y + np.sum(y, axis=1, dtype=self.dt)[:, None] - np.cumsum(y, axis=1, dtype=self.dt)
        return np.stack((rx, ry))

    def fk_joint(self, q, j):                           # robot.py:39-72 (fk_joint_1/2/3), j = 1..3
        c = np.cumsum(q.reshape(-1, 3)[:, :j], axis=1, dtype=self.dt)
        return np.stack((self.ll[:j] @ np.cos(c).T, self.ll[:j] @ np.sin(c).T))

    def jacobian_joint(self, q, j):                     # d fk_joint_j / d q: reverse cumsum over the first j links
        c = np.cumsum(q.reshape(-1, 3), axis=1, dtype=self.dt)
        x = -(self.ll * np.sin(c))
        y = self.ll * np.cos(c)
        x[:, j:] = 0
        y[:, j:] = 0
        rx = x + np.sum(x, axis=1, dtype=self.dt)[:, None] - np.cumsum(x, axis=1, dtype=self.dt)
        ry = y + np.sum(y, axis=1, dtype=self.dt)[:, None] - np.cumsum(y, axis=1, dtype=self.dt)
        rx[:, j:] = 0
        ry[:, j:] = 0
        return np.stack((rx, ry))

    def _norm(self, x):
        return np.sqrt(np.sum(np.square(x), dtype=self.dt))

    def sg_pos_ok(self, s, g, start, goal):             # robot.py:90-94
        return bool(self._norm(s - start) < self.eps_p) and bool(self._norm(g - goal) < self.eps_p)

    def sg_vel_ok(self, vs, vg):                        # robot.py:97-101
        return bool(self._norm(vs) < self.eps_v) and bool(self._norm(vg) < self.eps_v)

    def pos_ok(self, q):                                # robot.py:104-108
        return bool(q.max() <= self.qmax) and bool(q.min() >= self.qmin)

    def vel_ok(self, v):                                # robot.py:111-113
        return bool(np.abs(v).max() <= self.vmax)


# --------------------------------------------------------------------------
# trajectory.py
# --------------------------------------------------------------------------


class TrajectoryModel:
    """Objective of the FGD loop; mirrors class Trajectory (trajectory.py:22)."""

    def __init__(self, hp: Hyper, dtype=np.float32, jac_stream: str = "legacy",
                 jac: Optional[np.ndarray] = None):
        dt = self.dt = np.dtype(dtype).type
        self.hp = hp
        self.robot = RobotModel(hp, dt)
        self.T = T = int(hp.n_timesteps)
        self.rbf = dt(hp.rbf_variance)
        self.safety = dt(hp.joint_safety_limit)
        self.mean_q = dt(0.5) * (self.robot.qmax + self.robot.qmin)        # trajectory.py:31
        self.std_q = dt(0.5) * (self.robot.qmax - self.mean_q)             # trajectory.py:32 (0.75, not 1.5)
        # jnp.linspace(0,1,T): i/(T-1) in working precision               trajectory.py:35
        self.t = (np.arange(T, dtype=dt) / dt(T - 1)).astype(dt)
        t = self.t
        self.c = (dt(6) * t**5 - dt(15) * t**4 + dt(10) * t**3).astype(dt)  # trajectory.py:38
        a, b = np.meshgrid(t, t)            # 'xy': a[i,j]=t[j], b[i,j]=t[i]  trajectory.py:45-48
        two_s2 = dt(2) * self.rbf**2
        e = np.exp(-(a - b) ** 2 / two_s2).astype(dt)
        self.km = e                                                        # trajectory.py:14-15
        self.dkm = ((a - b) / (self.rbf**2) * e).astype(dt)                # trajectory.py:18-19
        self.jac = (make_jac(hp.jac_gaussian_mean, jac_stream, dt) if jac is None
                    else np.asarray(jac, dtype=dt))

    # -- evaluation -------------------------------------------------------
    def evaluate(self, alpha, M):                        # trajectory.py:63-65
        return (M @ alpha) @ self.jac

    def init_trajectory(self, start, goal):              # trajectory.py:73-78
        dt = self.dt
        start = np.asarray(start, dt)
        goal = np.asarray(goal, dt)
        line = start + (goal - start) * self.c[:, None]
        return np.linalg.solve(self.km, line @ np.linalg.inv(self.jac)).astype(dt)

    # -- obstacle term ----------------------------------------------------
    def point_cost(self, f, obstacles, lam_max):         # trajectory.py:81-88
        dt = self.dt
        cv = compute_cost(f, obstacles, dt)
        return dt(lam_max) * cv.max() + (dt(1) - dt(lam_max)) * (np.sum(cv, dtype=dt) / dt(cv.shape[0]))

    def point_cost_g(self, f, obstacles, lam_max):       # trajectory.py:91-110
        dt = self.dt
        cv, cg = compute_cost_vg(f, obstacles, dt)
        T = cv.shape[0]
        w_max = np.zeros((1, T), dt)
        w_max[0, int(np.argmax(cv))] = 1
        return (dt(lam_max) * w_max + (dt(1) - dt(lam_max)) * (np.ones((1, T), dt) / dt(T))) * cg

    def obstacle_cost(self, q, obstacles, lam_max):      # trajectory.py:113-117
        if getattr(self.hp, "whole_arm_cost", False):
            return self._whole_arm_cost(q, obstacles, lam_max)
        return self.point_cost(self.robot.fk(q), obstacles, lam_max)

    def obstacle_cost_g(self, q, obstacles, lam_max):    # trajectory.py:120-126
        if getattr(self.hp, "whole_arm_cost", False):
            return self._whole_arm_cost_g(q, obstacles, lam_max)
        cg = self.point_cost_g(self.robot.fk(q), obstacles, lam_max)
        return np.einsum("ij,ijk->jk", cg, self.robot.jacobian(q))

    # whole-arm extension (blog-post.html:505-513): cost_v[t] = sum_j costmap(fk_j(q_t)), then the same
    # max/mean weighting (trajectory.py:81-110) on the summed per-sample cost
    def _whole_arm_cost(self, q, obstacles, lam_max):
        dt = self.dt
        cv = compute_cost(self.robot.fk_joint(q, 1), obstacles, dt)
        for j in (2, 3):
            cv = cv + compute_cost(self.robot.fk_joint(q, j), obstacles, dt)
        return dt(lam_max) * cv.max() + (dt(1) - dt(lam_max)) * (np.sum(cv, dtype=dt) / dt(cv.shape[0]))

    def _whole_arm_cost_g(self, q, obstacles, lam_max):
        dt = self.dt
        parts = [compute_cost_vg(self.robot.fk_joint(q, j), obstacles, dt) for j in (1, 2, 3)]
        cv = (parts[0][0] + parts[1][0]) + parts[2][0]
        T = cv.shape[0]
        w = (dt(1) - dt(lam_max)) * (np.ones((1, T), dt) / dt(T))
        w[0, int(np.argmax(cv))] += dt(lam_max)
        g = np.zeros_like(q)
        for j in (3, 2, 1):
            g = g + np.einsum("ij,ijk->jk", w * parts[j - 1][1], self.robot.jacobian_joint(q, j))
        return g

    # -- penalties --------------------------------------------------------
    def sg_cost(self, q, start, goal):                   # trajectory.py:183-188
        dt = self.dt
        return dt(0.5) * np.sum(np.square(q[0] - start), dtype=dt) + dt(0.5) * np.sum(np.square(q[self.T - 1] - goal), dtype=dt)

    def sg_cost_g(self, q, start, goal):                 # trajectory.py:191-198
        g = np.zeros_like(q)
        g[0] = q[0] - start
        g[-1] = q[self.T - 1] - goal
        return g

    def sgv_cost(self, v):                               # trajectory.py:201-204
        dt = self.dt
        return dt(0.5) * np.sum(np.square(v[0]), dtype=dt) + dt(0.5) * np.sum(np.square(v[-1]), dtype=dt)

    def sgv_cost_g(self, v):                             # trajectory.py:207-212
        g = np.zeros_like(v)
        g[0] = v[0]
        g[-1] = v[-1]
        return g

    def _q_mask(self, q):                                # trajectory.py:221-223
        return (q > self.safety * self.robot.qmax) | (q < self.safety * self.robot.qmin)

    def _v_mask(self, v):                                # trajectory.py:251
        return np.abs(v) > self.safety * self.robot.vmax

    def jl_cost(self, q):                                # trajectory.py:215-227
        dt = self.dt
        e = dt(0.5) * np.square((q - self.mean_q) / self.std_q)
        if self.hp.constraint_violating_dependant_loss:
            e = np.where(self._q_mask(q), e, dt(0))
        return np.sum(e, dtype=dt) / dt(self.T)

    def jl_cost_g(self, q):                              # trajectory.py:230-242
        dt = self.dt
        g = (q - self.mean_q) / np.square(self.std_q)
        if self.hp.constraint_violating_dependant_loss:
            g = np.where(self._q_mask(q), g, dt(0))
        return (g / dt(self.T)).astype(dt)

    def jv_cost(self, v):                                # trajectory.py:245-255
        dt = self.dt
        e = dt(0.5) * np.square(v / self.robot.vmax)
        if self.hp.constraint_violating_dependant_loss:
            e = np.where(self._v_mask(v), e, dt(0))
        return np.sum(e, dtype=dt) / dt(self.T)

    def jv_cost_g(self, v):                              # trajectory.py:258-268
        dt = self.dt
        g = v / np.square(self.robot.vmax)
        if self.hp.constraint_violating_dependant_loss:
            g = np.where(self._v_mask(v), g, dt(0))
        return (g / dt(self.T)).astype(dt)

    # -- total ------------------------------------------------------------
    def cost(self, alpha, obstacles, start, goal, lam_sg, lam_jl, lam_max):   # trajectory.py:271-281
        dt = self.dt
        q = self.evaluate(alpha, self.km)
        v = self.evaluate(alpha, self.dkm)
        toc = self.obstacle_cost(q, obstacles, lam_max)
        sg = self.sg_cost(q, start, goal) + self.sgv_cost(v)
        jl = self.jl_cost(q) + self.jv_cost(v)
        return dt(toc + dt(lam_sg) * sg + dt(lam_jl) * jl)

    def cost_parts(self, alpha, obstacles, start, goal, lam_max):
        q = self.evaluate(alpha, self.km)
        v = self.evaluate(alpha, self.dkm)
        return dict(q=q, v=v, toc=self.obstacle_cost(q, obstacles, lam_max),
                    sgp=self.sg_cost(q, start, goal), sgv=self.sgv_cost(v),
                    jp=self.jl_cost(q), jv=self.jv_cost(v))

    def cost_g(self, alpha, obstacles, start, goal, lam_sg, lam_jl, lam_max):  # trajectory.py:284-297
        dt = self.dt
        q = self.evaluate(alpha, self.km)
        v = self.evaluate(alpha, self.dkm)
        gq = self.obstacle_cost_g(q, obstacles, lam_max) + dt(lam_sg) * self.sg_cost_g(q, start, goal) \
            + dt(lam_jl) * self.jl_cost_g(q)
        gv = dt(lam_sg) * self.sgv_cost_g(v) + dt(lam_jl) * self.jv_cost_g(v)
        return ((self.km.T @ gq + self.dkm.T @ gv) @ self.jac.T).astype(dt)

    def constraints_fulfilled(self, alpha, start, goal):   # trajectory.py:129-137
        q = self.evaluate(alpha, self.km)
        v = self.evaluate(alpha, self.dkm)
        r = self.robot
        return r.sg_pos_ok(q[0], q[-1], start, goal) and r.sg_vel_ok(v[0], v[-1]) and r.pos_ok(q) and r.vel_ok(v)


# --------------------------------------------------------------------------
# optimizers
# --------------------------------------------------------------------------


@dataclass
class RunLog:
    outer_iters: int = 0
    inner_iters: int = 0          # inner-loop bodies executed
    accepts: int = 0              # accepted steps
    cost_evals: int = 0
    grad_evals: int = 0
    fulfilled: bool = False
    accept_j: List[int] = field(default_factory=list)   # BLS: index of accepting candidate (-1 = none)
    lrs: List[float] = field(default_factory=list)      # lr of every accepted step
    iterates: List[np.ndarray] = field(default_factory=list)


def bls_optimize(tm: TrajectoryModel, alpha, obstacles, start, goal, log: Optional[RunLog] = None,
                 keep_iterates: bool = False, obstacle_schedule=None):
    """optimizer_BLS.py:65-123 (plain loop).  ``obstacle_schedule(k)`` (optional)
    returns the obstacle array to use from global inner iteration k on, evaluated
    at iteration boundaries exactly where the reference re-reads self.env.obstacles
    (optimizer_BLS.py:79,82,90)."""
    hp, dt = tm.hp, tm.dt
    log = log if log is not None else RunLog()
    start = np.asarray(start, dt)
    goal = np.asarray(goal, dt)
    lam_sg, lam_jl = dt(hp.lambda_sg_constraint), dt(hp.lambda_jl_constraint)
    lam_max = hp.lambda_max_cost
    alpha = np.asarray(alpha, dt).copy()
    if keep_iterates:
        log.iterates.append(tm.evaluate(alpha, tm.km))
    k_global = 0
    for _outer in range(hp.max_outer_iteration):
        log.outer_iters += 1
        lr = dt(hp.bls_lr_start)
        for _inner in range(hp.max_inner_iteration):
            if obstacle_schedule is not None:
                obstacles = obstacle_schedule(k_global)
            k_global += 1
            log.inner_iters += 1
            loss = tm.cost(alpha, obstacles, start, goal, lam_sg, lam_jl, lam_max)
            g = tm.cost_g(alpha, obstacles, start, goal, lam_sg, lam_jl, lam_max)
            log.cost_evals += 1
            log.grad_evals += 1
            n = (g / np.sqrt(np.sum(np.square(g), dtype=dt))).astype(dt)       # :84
            alpha_norm = np.sum(g.T @ n, dtype=dt)                              # :86 (sum of all 9 entries)
            new_loss = loss
            accepted = -1
            for j in range(hp.max_bls_iteration):
                new_alpha = ((dt(1) - dt(hp.lambda_reg) * lr) * alpha - lr * n).astype(dt)   # :89
                new_loss = tm.cost(new_alpha, obstacles, start, goal, lam_sg, lam_jl, lam_max)
                log.cost_evals += 1
                required = loss - dt(hp.bls_alpha) * lr * alpha_norm            # :91
                if new_loss > required:                                         # :94 (NaN -> accept)
                    lr = dt(lr * dt(hp.bls_beta_minus))
                else:
                    alpha = new_alpha
                    log.lrs.append(float(lr))
                    lr = dt(lr * dt(hp.bls_beta_plus))
                    accepted = j
                    break
            log.accept_j.append(accepted)
            if accepted >= 0:
                log.accepts += 1
            if loss - new_loss < dt(hp.loop_loss_reduction):                    # :102
                break
            if keep_iterates:
                log.iterates.append(tm.evaluate(alpha, tm.km))
        if tm.constraints_fulfilled(alpha, start, goal):                        # :112
            log.fulfilled = True
            break
        lam_sg = dt(lam_sg * dt(hp.lambda_constraint_increase))
        lam_jl = dt(lam_jl * dt(hp.lambda_constraint_increase))
    return alpha, log


def gd_optimize(tm: TrajectoryModel, alpha, obstacles, start, goal, log: Optional[RunLog] = None):
    """optimizer_GD.py: single level (:68-119) when max_outer_iteration == 1,
    otherwise the dual loop (:122-232)."""
    hp, dt = tm.hp, tm.dt
    log = log if log is not None else RunLog()
    start = np.asarray(start, dt)
    goal = np.asarray(goal, dt)
    if hp.max_outer_iteration > len(hp.gd_lr):
        raise ValueError("max_outer_iteration and dual_lr do not match")      # optimizer_GD.py:34-36
    lam_sg, lam_jl = dt(hp.lambda_sg_constraint), dt(hp.lambda_jl_constraint)
    lam_max = hp.lambda_max_cost
    alpha = np.asarray(alpha, dt).copy()
    dual = hp.max_outer_iteration > 1
    for outer in range(hp.max_outer_iteration):
        log.outer_iters += 1
        lr = dt(hp.gd_lr[outer])
        last = tm.cost(alpha, obstacles, start, goal, lam_sg, lam_jl, lam_max)
        log.cost_evals += 1
        for _inner in range(hp.max_inner_iteration):
            log.inner_iters += 1
            g = tm.cost_g(alpha, obstacles, start, goal, lam_sg, lam_jl, lam_max)
            log.grad_evals += 1
            new_alpha = ((dt(1) - dt(hp.lambda_reg) * lr) * alpha - lr * g).astype(dt)
            new_loss = tm.cost(new_alpha, obstacles, start, goal, lam_sg, lam_jl, lam_max)
            log.cost_evals += 1
            if last - new_loss < dt(hp.loop_loss_reduction):
                break                                   # new_alpha discarded
            last = new_loss
            alpha = new_alpha
            log.accepts += 1
        if not dual:
            log.fulfilled = tm.constraints_fulfilled(alpha, start, goal)
            break
        if tm.constraints_fulfilled(alpha, start, goal):
            log.fulfilled = True
            break
        lam_sg = dt(lam_sg * dt(hp.lambda_constraint_increase))
        lam_jl = dt(lam_jl * dt(hp.lambda_constraint_increase))
    return alpha, log


def final_report(tm: TrajectoryModel, alpha, obstacles, start, goal):
    """main.py:141-143: avg (lambda_max=0) and max (lambda_max=1) obstacle cost."""
    avg = tm.cost(alpha, obstacles, start, goal, 0, 0, 0)
    mx = tm.cost(alpha, obstacles, start, goal, 0, 0, 1)
    return float(avg), float(mx), tm.constraints_fulfilled(alpha, np.asarray(start, tm.dt), np.asarray(goal, tm.dt))
