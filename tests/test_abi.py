"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol
include/fgd_b200.h declares, and the Python host mirrors the reference's interface.
No compute calls (there is no GPU in the -m "not gpu" run)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "fgd_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fgd_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from irm_motion_planning_b200 import backend, build
    build.build()
    lib = ctypes.CDLL(backend.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in fgd_b200.h but not exported"
    assert sorted(backend.EXPORTED_SYMBOLS) == declared
    lib.fgd_abi_version.restype = ctypes.c_int
    assert lib.fgd_abi_version() == backend.FGD_ABI_VERSION


def test_library_contains_sm100a_code_only():
    import subprocess
    from irm_motion_planning_b200 import backend
    out = subprocess.run(["cuobjdump", "--list-elf", backend.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert not re.search(r"sm_(?!100a)\d+", out)


def test_config_struct_layout_matches_header():
    from irm_motion_planning_b200 import backend
    # 11 int32 + 16 float + 3 + 9 + 16 floats, then two pointers (8-byte aligned)
    off = 4 * (11 + 16 + 3 + 9 + 16)
    assert backend.FgdConfig.h_km.offset == (off + 7) // 8 * 8
    assert ctypes.sizeof(backend.FgdConfig) == backend.FgdConfig.h_km.offset + 16


def test_create_rejects_bad_arguments_without_a_gpu():
    """Argument validation happens before any CUDA call, so it is testable on the CPU."""
    from irm_motion_planning_b200 import backend
    from irm_motion_planning_b200.trajectory import Trajectory
    from irm_motion_planning_b200.workloads import default_args
    lib = backend.load_library()
    tr = Trajectory(default_args(), create_handle=False)
    h = ctypes.c_void_p()

    def create(**over):
        ns = default_args()
        km, dkm = tr.km, tr.dkm
        for k, v in over.items():
            if k == "km":
                km = v
            elif k == "dkm":
                dkm = v
            else:
                setattr(ns, k, v)
        cfg = backend.make_config(ns, km, dkm, tr.jac, 64, False)
        return lib.fgd_create(ctypes.byref(cfg), ctypes.byref(h))

    assert create(n_joints=4, link_length=[1, 1, 1, 1]) == 7            # FGD_ERR_JOINTS
    bad = tr.km.copy(); bad[3, 7] += 1e-3
    assert create(km=bad) == 3                                           # FGD_ERR_KERNEL_NOT_SYMMETRIC
    assert create(dkm=np.abs(tr.dkm)) == 3
    assert create(max_outer_iteration=17) == 1                           # FGD_ERR_INVALID_ARGUMENT
    tr300 = Trajectory(default_args(n_timesteps=300), create_handle=False)
    ns = default_args(n_timesteps=300)
    cfg = backend.make_config(ns, tr300.km, tr300.dkm, tr300.jac, 64, False)
    assert lib.fgd_create(ctypes.byref(cfg), ctypes.byref(h)) == 2       # FGD_ERR_UNSUPPORTED_T
    assert lib.fgd_status_string(3).decode().startswith("km must be symmetric")


def test_reference_flag_table_is_complete():
    """Every flag of the reference CLI (main.py:17-98) exists with the same default."""
    from irm_motion_planning_b200.main import parse_args
    a = parse_args([])
    want = dict(profiling=False, extended_vis=False, n_measurements=1, n_times=1, optimizer_name="bls", jit_loop=True,
                n_timesteps=50, rbf_variance=0.1, jac_gaussian_mean=0.15, max_inner_iteration=200,
                loop_loss_reduction=1e-3, max_outer_iteration=10, lambda_constraint_increase=10,
                lambda_sg_constraint=0.5, lambda_jl_constraint=0.1, eps_position=0.01, eps_velocity=0.01,
                lambda_max_cost=0.5, lambda_reg=1e-4, constraint_violating_dependant_loss=True, joint_safety_limit=0.98,
                max_bls_iteration=20, bls_lr_start=0.2, bls_alpha=0.01, bls_beta_plus=1.2, bls_beta_minus=0.5,
                gd_lr=[2e-3, 1e-4, 1e-5, 1e-6, 1e-7, 1e-8, 1e-8, 1e-8, 1e-8, 1e-8], n_joints=3,
                link_length=[1.5, 1.0, 0.5], max_joint_velocity=7, max_joint_position=2, min_joint_position=-1)
    for k, v in want.items():
        assert getattr(a, k) == v, k
    b = parse_args(["--n-timesteps", "256", "--jit-loop", "False", "--gd-lr", "1e-3", "2e-4", "--bls-beta_plus", "1.5"])
    assert b.n_timesteps == 256.0 and b.jit_loop is False and b.gd_lr == [1e-3, 2e-4] and b.bls_beta_plus == 1.5


def test_host_constants_match_oracle_bitwise():
    from oracle import fgd_numpy as O
    from irm_motion_planning_b200.trajectory import Trajectory
    from irm_motion_planning_b200.workloads import default_args
    for T in (50, 256):
        tr = Trajectory(default_args(n_timesteps=float(T)), create_handle=False)
        tm = O.TrajectoryModel(O.Hyper(n_timesteps=T))
        assert np.array_equal(tr.km, tm.km) and np.array_equal(tr.dkm, tm.dkm) and np.array_equal(tr.jac, tm.jac)
        assert np.array_equal(tr.c, tm.c)
    a_pkg = tr.initTrajectory(O.DEFAULT_START, O.DEFAULT_GOAL)
    assert a_pkg.shape == (256, 3)


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: no module of the shipped package may import or load it."""
    pkg = os.path.join(ROOT, "irm_motion_planning_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", text, flags=re.M), f
                assert "libfgd_mirror" not in text and "fgd_numpy" not in text, f


def test_missing_library_fails_loudly(tmp_path):
    from irm_motion_planning_b200 import backend
    with pytest.raises(FileNotFoundError, match="no CPU fallback"):
        backend.load_library(str(tmp_path / "nope.so"))


def test_shard_range_partitions_exactly():
    from irm_motion_planning_b200.batch import shard_range
    for n in (1, 7, 4096, 1048576):
        for w in (1, 2, 3, 8):
            parts = [shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in parts) - min(b - a for a, b in parts) <= 1


def test_sqrt_threshold_is_equivalent_to_the_square_root_test():
    """The kernels test the end-point predicates ||.|| < eps (robot.py:90-101) as  squared norm < fgd_sqrt_threshold(eps)
    (no square root on the device; the oracle takes the IEEE square root): equivalent for every float because the
    correctly rounded sqrt is monotonic - checked on the floats around the threshold and on random values."""
    from irm_motion_planning_b200 import backend, build
    build.build()
    lib = ctypes.CDLL(backend.LIB_PATH)
    lib.fgd_sqrt_threshold.restype = ctypes.c_float
    lib.fgd_sqrt_threshold.argtypes = [ctypes.c_float]
    f = np.float32
    rng = np.random.default_rng(0)
    for eps in [0.01, 0.05, 1e-3, 0.1, 1.0, 7.0, 1e-20, 1e15] + list(np.exp(rng.uniform(-20, 5, 100))):
        e = f(eps)
        t = f(lib.fgd_sqrt_threshold(e))
        lo, hi, xs = t, t, [t]
        for _ in range(200):
            lo, hi = np.nextafter(lo, f(0)), np.nextafter(hi, f(np.inf))
            xs += [lo, hi]
        xs = np.concatenate([np.array(xs, f), (np.abs(rng.standard_normal(5000)) * float(e) ** 2 * 4).astype(f), np.array([0.0, np.inf, np.nan], f)])
        assert np.array_equal(np.sqrt(xs) < e, xs < t), eps
    assert lib.fgd_sqrt_threshold(f(0.0)) == 0.0 and lib.fgd_sqrt_threshold(f(-1.0)) == 0.0
