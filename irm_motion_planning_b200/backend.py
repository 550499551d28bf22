"""ctypes binding of libfgd_b200.so (include/fgd_b200.h).

PyTorch supplies device memory and streams; every compute call goes through the
C ABI into the sm_100a kernels.  There is no CPU or PyTorch fallback: if the
library or a CUDA device is missing, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FGD_LIBRARY", os.path.join(_PKG, "libfgd_b200.so"))   # override: debugging builds only

FGD_ABI_VERSION = 3
FGD_MAX_T = 256
FGD_MAX_OUTER = 16
FGD_OBS_RING = 16
FGD_SWITCH_LOG = 64
FSTATE, ISTATE = 8, 8
F_LAM_SG, F_LAM_JL, F_LR, F_LOSS, F_TOC, F_LAST_NEW_LOSS = range(6)
I_STATUS, I_OUTER, I_INNER, I_INNER_TOTAL, I_CAND_EVALS, I_ACCEPTS, I_FULFILLED, I_HASH = range(8)
ST_FRESH, ST_ACTIVE, ST_DONE = 0, 1, 2


class FgdError(RuntimeError):
    def __init__(self, status: int, what: str, cuda_error: int = 0):
        self.status, self.cuda_error = status, cuda_error
        super().__init__(f"{what}: {_status_string(status)}" + (f" (cudaError {cuda_error})" if cuda_error else ""))


class FgdConfig(C.Structure):
    _fields_ = (
        [(n, C.c_int32) for n in ("abi_version", "n_timesteps", "n_joints", "obstacle_capacity", "strict_math",
                                  "max_inner_iteration", "max_outer_iteration", "max_bls_iteration",
                                  "constraint_violating_dependant_loss", "n_gd_lr", "whole_arm_cost")]
        + [(n, C.c_float) for n in ("lambda_sg_constraint", "lambda_jl_constraint", "lambda_constraint_increase",
                                    "lambda_max_cost", "lambda_reg", "loop_loss_reduction", "eps_position",
                                    "eps_velocity", "bls_lr_start", "bls_alpha", "bls_beta_plus", "bls_beta_minus",
                                    "joint_safety_limit", "max_joint_position", "min_joint_position",
                                    "max_joint_velocity")]
        + [("link_length", C.c_float * 3), ("jac", C.c_float * 9), ("gd_lr", C.c_float * FGD_MAX_OUTER),
           ("h_km", C.POINTER(C.c_float)), ("h_dkm", C.POINTER(C.c_float))]
    )


EXPORTED_SYMBOLS = (
    "fgd_create", "fgd_destroy", "fgd_status_string", "fgd_last_cuda_error", "fgd_set_obstacles_async",
    "fgd_obstacle_count", "fgd_obstacle_generation", "fgd_optimize_live", "fgd_eval_cost_grad", "fgd_optimize_bls", "fgd_optimize_gd", "fgd_optimize_host",
    "fgd_argmin_per_problem", "fgd_launch_geometry", "fgd_kernel_launches", "fgd_abi_version",
    "fgd_measure_fp32_peak", "fgd_measure_mufu_peak", "fgd_set_init_basis", "fgd_init_trajectory", "fgd_optimize_host_io", "fgd_zero_copy_calls", "fgd_speculative_launches",
    "fgd_sqrt_threshold",
)

_lib = None


def load_library(path: Optional[str] = None):
    """dlopen the in-tree CUDA library and declare the prototypes.  Loud on failure."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise FileNotFoundError(
            f"{p} is missing: build it with `python -m irm_motion_planning_b200.build` "
            "(nvcc, sm_100a).  There is no CPU fallback for the FGD hot path.")
    lib = C.CDLL(p)
    vp, fp, ip, i32, f32 = C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_float
    lib.fgd_abi_version.restype = C.c_int
    lib.fgd_status_string.restype = C.c_char_p
    lib.fgd_status_string.argtypes = [C.c_int]
    lib.fgd_create.argtypes = [C.POINTER(FgdConfig), C.POINTER(vp)]
    lib.fgd_destroy.argtypes = [vp]
    lib.fgd_last_cuda_error.argtypes = [vp]
    lib.fgd_set_obstacles_async.argtypes = [vp, fp, i32, i32, vp]
    lib.fgd_obstacle_count.argtypes = [vp]
    lib.fgd_obstacle_generation.argtypes = [vp]
    lib.fgd_optimize_live.argtypes = [vp, i32, i32, fp, fp, fp, fp, ip, i32, ip, vp]
    lib.fgd_eval_cost_grad.argtypes = [vp, i32, fp, fp, fp, f32, f32, f32, fp, fp, fp, fp, fp, ip, vp]
    lib.fgd_optimize_bls.argtypes = [vp, i32, fp, fp, fp, fp, ip, i32, vp]
    lib.fgd_optimize_gd.argtypes = [vp, i32, fp, fp, fp, fp, ip, i32, vp]
    lib.fgd_optimize_host.argtypes = [vp, i32, i32, fp, fp, fp, fp, ip, vp]
    lib.fgd_optimize_host_io.argtypes = [vp, i32, i32, fp, fp, fp, fp, fp, ip, vp]
    lib.fgd_argmin_per_problem.argtypes = [vp, i32, i32, fp, ip, i32, i32, fp, ip, vp, vp]
    lib.fgd_set_init_basis.argtypes = [vp, fp, fp, fp]
    lib.fgd_init_trajectory.argtypes = [vp, i32, fp, fp, fp, vp]
    lib.fgd_launch_geometry.argtypes = [vp, i32, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    lib.fgd_kernel_launches.argtypes = [vp]
    lib.fgd_kernel_launches.restype = C.c_int64
    lib.fgd_zero_copy_calls.argtypes = [vp]
    lib.fgd_zero_copy_calls.restype = C.c_int64
    lib.fgd_speculative_launches.argtypes = [vp]
    lib.fgd_speculative_launches.restype = C.c_int64
    lib.fgd_measure_fp32_peak.argtypes = [vp, C.POINTER(C.c_double), vp]
    lib.fgd_measure_mufu_peak.argtypes = [vp, C.POINTER(C.c_double), vp]
    if lib.fgd_abi_version() != FGD_ABI_VERSION:
        raise RuntimeError("libfgd_b200.so ABI version mismatch; rebuild")
    if path is None:
        _lib = lib
    return lib


def _status_string(status: int) -> str:
    try:
        return load_library().fgd_status_string(status).decode()
    except Exception:  # pragma: no cover
        return f"status {status}"


def make_config(hp, km: np.ndarray, dkm: np.ndarray, jac: np.ndarray, obstacle_capacity: int,
                strict_math: bool) -> FgdConfig:
    """hp: object with the reference's argparse attribute names (main.py:17-98)."""
    c = FgdConfig()
    c.abi_version = FGD_ABI_VERSION
    c.n_timesteps = int(hp.n_timesteps)
    c.n_joints = int(hp.n_joints)
    c.obstacle_capacity = int(obstacle_capacity)
    c.strict_math = int(bool(strict_math))
    c.max_inner_iteration = int(hp.max_inner_iteration)
    c.max_outer_iteration = int(hp.max_outer_iteration)
    c.max_bls_iteration = int(hp.max_bls_iteration)
    c.constraint_violating_dependant_loss = int(bool(hp.constraint_violating_dependant_loss))
    lrs = [float(x) for x in hp.gd_lr][:FGD_MAX_OUTER]
    c.n_gd_lr = len(lrs)
    c.whole_arm_cost = int(bool(getattr(hp, "whole_arm_cost", False)))
    for i, x in enumerate(lrs):
        c.gd_lr[i] = x
    c.lambda_sg_constraint = hp.lambda_sg_constraint
    c.lambda_jl_constraint = hp.lambda_jl_constraint
    c.lambda_constraint_increase = hp.lambda_constraint_increase
    c.lambda_max_cost = hp.lambda_max_cost
    c.lambda_reg = hp.lambda_reg
    c.loop_loss_reduction = hp.loop_loss_reduction
    c.eps_position, c.eps_velocity = hp.eps_position, hp.eps_velocity
    c.bls_lr_start, c.bls_alpha = hp.bls_lr_start, hp.bls_alpha
    c.bls_beta_plus, c.bls_beta_minus = hp.bls_beta_plus, hp.bls_beta_minus
    c.joint_safety_limit = hp.joint_safety_limit
    c.max_joint_position, c.min_joint_position = hp.max_joint_position, hp.min_joint_position
    c.max_joint_velocity = hp.max_joint_velocity
    for i in range(3):
        c.link_length[i] = float(hp.link_length[i])
    j = np.asarray(jac, np.float32).reshape(9)
    for i in range(9):
        c.jac[i] = float(j[i])
    c._km = np.ascontiguousarray(km, np.float32)       # keep alive
    c._dkm = np.ascontiguousarray(dkm, np.float32)
    c.h_km = c._km.ctypes.data_as(C.POINTER(C.c_float))
    c.h_dkm = c._dkm.ctypes.data_as(C.POINTER(C.c_float))
    return c


def _ptr(t):
    """Device (or pinned host) address of a torch tensor / numpy array, or NULL."""
    if t is None:
        return None
    if isinstance(t, np.ndarray):
        return t.ctypes.data
    return t.data_ptr()


class Handle:
    """Owns one FgdHandle* on the current CUDA device."""

    def __init__(self, cfg: FgdConfig):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("irm_motion_planning_b200 needs a CUDA device (B200, sm_100a); no CPU fallback exists")
        self._lib = load_library()
        torch.cuda.init()
        torch.zeros(1, device="cuda")          # make sure the primary context is current
        self._h = C.c_void_p()
        rc = self._lib.fgd_create(C.byref(cfg), C.byref(self._h))
        if rc:
            raise FgdError(rc, "fgd_create")
        self.T = int(cfg.n_timesteps)
        self._obs_keepalive = []               # device-resident sources of the most recent async uploads (copied device-to-device)

    def _check(self, rc: int, what: str):
        if rc:
            raise FgdError(rc, what, self._lib.fgd_last_cuda_error(self._h))

    @staticmethod
    def _stream():
        import torch
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.fgd_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    # -- obstacles ---------------------------------------------------------
    def set_obstacles(self, xy, stream=None):
        """xy: (O,2) float32 numpy array (host; pinned if it came from a pinned torch
        tensor) or CUDA torch tensor."""
        on_dev = 0
        if not isinstance(xy, np.ndarray):
            import torch
            xy = xy.to(torch.float32).contiguous()
            if xy.is_cuda:
                on_dev = 1
            else:
                xy = xy.numpy()
        if isinstance(xy, np.ndarray):
            xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
        if on_dev:
            self._obs_keepalive = (self._obs_keepalive + [xy])[-FGD_OBS_RING:]      # host sources are staged inside the call
        n = int(xy.shape[0])
        st = self._stream() if stream is None else C.c_void_p(stream)
        self._check(self._lib.fgd_set_obstacles_async(self._h, _ptr(xy), n, on_dev, st), "fgd_set_obstacles_async")

    def obstacle_count(self) -> int:
        return int(self._lib.fgd_obstacle_count(self._h))

    @property
    def obstacle_generation(self) -> int:
        """Number of obstacle sets published through this handle so far (callers that cache uploads compare it)."""
        return int(self._lib.fgd_obstacle_generation(self._h))

    # -- evaluation --------------------------------------------------------
    def eval(self, B, alpha, start, goal, lam_sg, lam_jl, lam_max, loss=None, toc=None, grad=None, q=None, v=None,
             fulfilled=None):
        self._check(self._lib.fgd_eval_cost_grad(self._h, B, _ptr(alpha), _ptr(start), _ptr(goal), lam_sg, lam_jl, lam_max,
                                                 _ptr(loss), _ptr(toc), _ptr(grad), _ptr(q), _ptr(v), _ptr(fulfilled),
                                                 self._stream()), "fgd_eval_cost_grad")

    # -- optimisation ------------------------------------------------------
    def optimize(self, mode: str, B, alpha, start, goal, fstate, istate, max_launch_iters: int = -1):
        fn = self._lib.fgd_optimize_bls if mode == "bls" else self._lib.fgd_optimize_gd
        self._check(fn(self._h, B, _ptr(alpha), _ptr(start), _ptr(goal), _ptr(fstate), _ptr(istate), max_launch_iters,
                       self._stream()), f"fgd_optimize_{mode}")

    def optimize_live(self, mode: str, B, alpha, start, goal, fstate, istate, poll_every: int, switch_log=None):
        """One persistent launch that follows obstacle sets published while it runs (fgd_optimize_live)."""
        self._check(self._lib.fgd_optimize_live(self._h, 1 if mode == "gd" else 0, B, _ptr(alpha), _ptr(start), _ptr(goal),
                                                _ptr(fstate), _ptr(istate), int(poll_every), _ptr(switch_log), self._stream()),
                    "fgd_optimize_live")

    def optimize_host(self, mode: str, B, alpha, start, goal, fstate, istate):
        self._check(self._lib.fgd_optimize_host(self._h, 1 if mode == "gd" else 0, B, _ptr(alpha), _ptr(start), _ptr(goal),
                                                _ptr(fstate), _ptr(istate), self._stream()), "fgd_optimize_host")

    def optimize_host_io(self, mode: str, B, alpha_in, alpha_out, start, goal, fstate_out, istate_out):
        self._check(self._lib.fgd_optimize_host_io(self._h, 1 if mode == "gd" else 0, B, _ptr(alpha_in), _ptr(alpha_out),
                                                   _ptr(start), _ptr(goal), _ptr(fstate_out), _ptr(istate_out),
                                                   self._stream()), "fgd_optimize_host_io")

    def argmin_per_problem(self, n_problems, n_restarts, fstate, istate, index_offset, best_cost=None, best_index=None,
                           best_key=None, problem_stride: int = 0):
        self._check(self._lib.fgd_argmin_per_problem(self._h, n_problems, n_restarts, _ptr(fstate), _ptr(istate),
                                                     index_offset, problem_stride, _ptr(best_cost), _ptr(best_index),
                                                     _ptr(best_key), self._stream()), "fgd_argmin_per_problem")

    # -- device-side initTrajectory ----------------------------------------
    def set_init_basis(self, u: np.ndarray, w: np.ndarray, jinv: np.ndarray):
        u, w = np.ascontiguousarray(u, np.float32), np.ascontiguousarray(w, np.float32)
        jinv = np.ascontiguousarray(jinv, np.float32).reshape(9)
        assert u.shape == (self.T,) and w.shape == (self.T,)
        self._check(self._lib.fgd_set_init_basis(self._h, _ptr(u), _ptr(w), _ptr(jinv)), "fgd_set_init_basis")

    def init_trajectory(self, B, start, goal, alpha):
        self._check(self._lib.fgd_init_trajectory(self._h, B, _ptr(start), _ptr(goal), _ptr(alpha), self._stream()),
                    "fgd_init_trajectory")

    def launch_geometry(self, B: int):
        g, b, s, t = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        self._check(self._lib.fgd_launch_geometry(self._h, B, C.byref(g), C.byref(b), C.byref(s), C.byref(t)),
                    "fgd_launch_geometry")
        return dict(grid=g.value, block=b.value, smem_bytes=s.value, warps_per_trajectory=t.value)

    def measure_fp32_peak(self) -> float:
        out = C.c_double()
        self._check(self._lib.fgd_measure_fp32_peak(self._h, C.byref(out), self._stream()), "fgd_measure_fp32_peak")
        return float(out.value)

    def measure_mufu_peak(self) -> float:
        """10^12 rcp.approx per second (the obstacle stage's second roofline)."""
        out = C.c_double()
        self._check(self._lib.fgd_measure_mufu_peak(self._h, C.byref(out), self._stream()), "fgd_measure_mufu_peak")
        return float(out.value)

    def kernel_launches(self) -> int:
        return int(self._lib.fgd_kernel_launches(self._h))

    def speculative_launches(self) -> int:
        """BLS launches that ran the speculative (parallel-candidate) line search kernel."""
        return int(self._lib.fgd_speculative_launches(self._h))

    def zero_copy_calls(self) -> int:
        """optimize_host_io calls that ran zero-copy on page-locked buffers (no staging copies)."""
        return int(self._lib.fgd_zero_copy_calls(self._h))
