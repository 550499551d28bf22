"""ctypes front-end of the C mirror oracle (oracle/fgd_mirror.c).

TEST INFRASTRUCTURE ONLY (see the header of fgd_mirror.c).  Builds
``oracle/_ref/libfgd_mirror.so`` on demand with gcc (``make -C oracle``).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libfgd_mirror.so")

FS = 8
IS = 8
F_LAM_SG, F_LAM_JL, F_LR, F_LOSS, F_TOC, F_LAST_NEW_LOSS = range(6)
I_STATUS, I_OUTER, I_INNER, I_INNER_TOTAL, I_CAND_EVALS, I_ACCEPTS, I_FULFILLED, I_HASH = range(8)
ST_FRESH, ST_ACTIVE, ST_DONE = 0, 1, 2


class MirrorCfg(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("T", "n_obs", "max_inner", "max_outer", "max_bls", "cvdl", "mode", "strict")] + \
               [(n, C.c_float) for n in ("lam_sg0", "lam_jl0", "lam_inc", "lam_max", "lam_reg", "eps_loop", "eps_pos",
                                         "eps_vel", "bls_lr0", "bls_alpha", "bls_bp", "bls_bm", "safety", "qmax", "qmin",
                                         "vmax")] + \
               [("link", C.c_float * 3), ("J", C.c_float * 9), ("gd_lr", C.c_float * 16), ("whole_arm", C.c_int)]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "fgd_mirror.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        fp = C.POINTER(C.c_float)
        ip = C.POINTER(C.c_int)
        _lib.mirror_eval.argtypes = [C.POINTER(MirrorCfg), fp, fp, fp, C.c_int, fp, fp, fp, C.c_float, C.c_float,
                                     fp, fp, fp, fp, fp, ip]
        _lib.mirror_optimize.argtypes = [C.POINTER(MirrorCfg), fp, fp, fp, C.c_int, fp, fp, fp, fp, ip, C.c_int, C.c_int]
        _lib.mirror_init.argtypes = [C.c_int, C.c_int, fp, fp, fp, fp, fp, fp]
        _lib.mirror_max_threads.restype = C.c_int
    return _lib


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float)) if a is not None else None


def make_cfg(hp, jac: np.ndarray, n_obs: int, mode: str) -> MirrorCfg:
    """hp: any object with the reference's hyper-parameter attribute names."""
    c = MirrorCfg()
    c.T = int(hp.n_timesteps)
    c.n_obs = int(n_obs)
    c.max_inner = int(hp.max_inner_iteration)
    c.max_outer = int(hp.max_outer_iteration)
    c.max_bls = int(hp.max_bls_iteration)
    c.cvdl = int(bool(hp.constraint_violating_dependant_loss))
    c.mode = {"bls": 0, "gd": 1}[mode]
    c.strict = 1
    c.whole_arm = int(bool(getattr(hp, "whole_arm_cost", False)))
    c.lam_sg0, c.lam_jl0 = hp.lambda_sg_constraint, hp.lambda_jl_constraint
    c.lam_inc, c.lam_max, c.lam_reg = hp.lambda_constraint_increase, hp.lambda_max_cost, hp.lambda_reg
    c.eps_loop, c.eps_pos, c.eps_vel = hp.loop_loss_reduction, hp.eps_position, hp.eps_velocity
    c.bls_lr0, c.bls_alpha, c.bls_bp, c.bls_bm = hp.bls_lr_start, hp.bls_alpha, hp.bls_beta_plus, hp.bls_beta_minus
    c.safety, c.qmax, c.qmin, c.vmax = hp.joint_safety_limit, hp.max_joint_position, hp.min_joint_position, hp.max_joint_velocity
    for i in range(3):
        c.link[i] = hp.link_length[i]
    j = np.asarray(jac, np.float32).reshape(9)
    for i in range(9):
        c.J[i] = float(j[i])
    lrs = list(hp.gd_lr)[:16]
    for i in range(16):
        c.gd_lr[i] = lrs[i] if i < len(lrs) else lrs[-1]
    return c


class Mirror:
    """Batch evaluation / optimisation with the C mirror oracle."""

    def __init__(self, hp, km: np.ndarray, dkm: np.ndarray, jac: np.ndarray, obstacles: np.ndarray, mode: str = "bls"):
        self.T = int(hp.n_timesteps)
        self.K = np.ascontiguousarray(km, np.float32)
        self.dK = np.ascontiguousarray(dkm, np.float32)
        self.hp, self.jac, self.mode = hp, np.asarray(jac, np.float32), mode
        self.set_obstacles(obstacles)

    def set_obstacles(self, obstacles):
        self.obs = np.ascontiguousarray(obstacles, np.float32).reshape(-1, 2)
        self.cfg = make_cfg(self.hp, self.jac, len(self.obs), self.mode)

    @staticmethod
    def _prep(alpha, start, goal, T):
        alpha = np.ascontiguousarray(alpha, np.float32).reshape(-1, T, 3)
        B = alpha.shape[0]
        start = np.ascontiguousarray(np.broadcast_to(np.asarray(start, np.float32).reshape(-1, 3), (B, 3)))
        goal = np.ascontiguousarray(np.broadcast_to(np.asarray(goal, np.float32).reshape(-1, 3), (B, 3)))
        return alpha, start, goal, B

    def eval(self, alpha, start, goal, lam_sg: float, lam_jl: float):
        alpha, start, goal, B = self._prep(alpha, start, goal, self.T)
        T = self.T
        out = dict(loss=np.empty(B, np.float32), toc=np.empty(B, np.float32), grad=np.empty((B, T, 3), np.float32),
                   q=np.empty((B, T, 3), np.float32), v=np.empty((B, T, 3), np.float32),
                   fulfilled=np.empty(B, np.int32))
        rc = lib().mirror_eval(C.byref(self.cfg), _fp(self.K), _fp(self.dK), _fp(self.obs), B, _fp(alpha), _fp(start),
                               _fp(goal), lam_sg, lam_jl, _fp(out["loss"]), _fp(out["toc"]), _fp(out["grad"]),
                               _fp(out["q"]), _fp(out["v"]), out["fulfilled"].ctypes.data_as(C.POINTER(C.c_int)))
        if rc:
            raise RuntimeError(f"mirror_eval rc={rc}")
        return out

    def new_state(self, B):
        return np.zeros((B, FS), np.float32), np.zeros((B, IS), np.int32)

    def optimize(self, alpha, start, goal, fstate=None, istate=None, budget: int = -1, nthreads: int = 0):
        """Returns (alpha_out, fstate, istate); inputs are not modified."""
        alpha, start, goal, B = self._prep(alpha, start, goal, self.T)
        alpha = alpha.copy()
        if fstate is None:
            fstate, istate = self.new_state(B)
        rc = lib().mirror_optimize(C.byref(self.cfg), _fp(self.K), _fp(self.dK), _fp(self.obs), B, _fp(alpha),
                                   _fp(start), _fp(goal), _fp(fstate), istate.ctypes.data_as(C.POINTER(C.c_int)),
                                   budget, nthreads)
        if rc:
            raise RuntimeError(f"mirror_optimize rc={rc}")
        return alpha, fstate, istate


def init_trajectory(u, w, jinv, start, goal) -> np.ndarray:
    """Rank-2 initTrajectory (trajectory.py:73-78) in the device kernel's op order."""
    u, w = np.ascontiguousarray(u, np.float32), np.ascontiguousarray(w, np.float32)
    jinv = np.ascontiguousarray(jinv, np.float32).reshape(9)
    start = np.ascontiguousarray(start, np.float32).reshape(-1, 3)
    goal = np.ascontiguousarray(goal, np.float32).reshape(-1, 3)
    B, T = start.shape[0], u.shape[0]
    alpha = np.empty((B, T, 3), np.float32)
    rc = lib().mirror_init(T, B, _fp(u), _fp(w), _fp(jinv), _fp(start), _fp(goal), _fp(alpha))
    if rc:
        raise RuntimeError(f"mirror_init rc={rc}")
    return alpha


def max_threads() -> int:
    return int(lib().mirror_max_threads())
