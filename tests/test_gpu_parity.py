"""GPU parity tests: the CUDA path, called through the C ABI (ctypes -> libfgd_b200.so),
against the CPU mirror oracle on the same seeded inputs.

Bars (stated per test):
* strict-math mode: BIT-EXACT -- loss, gradient, q, v, final alpha, every loop counter and
  the decision-trace hash equal the oracle's, because both execute the same documented FP32
  operation sequence (IEEE add/mul/fma/div/sqrt only).
* fast-math mode (rcp.approx instead of the IEEE reciprocal, 1 ulp): per evaluation
  loss rel <= 1e-5 and gradient rel <= 1e-5 (north_star asks 1e-4); end to end the runs
  diverge chaotically, so the distribution is compared and the fraction of
  identical decision traces is reported.
"""
import os

import numpy as np
import pytest

from oracle import fgd_numpy as O
from oracle import mirror as M

pytestmark = pytest.mark.gpu


def _setup(T=50, n_obs=11, seed=0, B=64, strict=True, capacity=1024, **over):
    import torch
    from irm_motion_planning_b200.environment import Environment, random_obstacles
    from irm_motion_planning_b200.trajectory import Trajectory
    from irm_motion_planning_b200.workloads import default_args, sample_start_goal
    args = default_args(n_timesteps=float(T), **over)
    tr = Trajectory(args, obstacle_capacity=capacity, strict_math=strict)
    rng = np.random.default_rng(seed)
    obs = Environment().obstacles if n_obs == 11 else random_obstacles(n_obs, rng)
    start, goal = sample_start_goal(B, rng)
    alpha0 = tr.initTrajectory(start, goal)
    tr.set_obstacles(obs)
    return args, tr, obs, start, goal, alpha0


def _mirror(args, tr, obs, mode):
    hp = type("HP", (), dict(vars(args)))()
    hp.n_timesteps = tr.N_timesteps
    return M.Mirror(hp, tr.km, tr.dkm, tr.jac, obs, mode)


def _gpu_eval(tr, alpha, start, goal, lam_sg, lam_jl, lam_max=-1.0):
    out = tr._eval(alpha, None, start, goal, lam_sg, lam_jl, lam_max, ("loss", "toc", "grad", "q", "v", "fulfilled"))
    return {k: v.cpu().numpy() for k, v in out.items()}


def _gpu_optimize(tr, mode, alpha0, start, goal, budget=-1, state=None):
    import torch
    from irm_motion_planning_b200.batch import BatchedFGD
    eng = BatchedFGD(tr, mode)
    a = torch.as_tensor(alpha0, device="cuda").clone().contiguous()
    s = torch.as_tensor(start, device="cuda").contiguous()
    g = torch.as_tensor(goal, device="cuda").contiguous()
    fs, is_ = state if state is not None else eng.new_state(a.shape[0])
    eng.optimize_device(a, s, g, fs, is_, max_launch_iters=budget)
    torch.cuda.synchronize()
    return a, fs, is_


@pytest.mark.parametrize("T,n_obs", [(50, 11), (20, 11), (33, 37), (64, 5), (65, 9), (100, 64), (129, 20), (256, 300),
                                     (50, 256), (50, 67), (33, 100), (20, 64), (63, 130), (64, 200)])      # >= 64 obstacles, T < 64: two-chain obstacle sums (share_split)
def test_eval_bit_exact_strict(cuda_ready, T, n_obs):
    args, tr, obs, start, goal, alpha0 = _setup(T=T, n_obs=n_obs, B=24, seed=T)
    rng = np.random.default_rng(T + 1)
    smooth = (rng.standard_normal(alpha0.shape) * 0.05).astype(np.float32)
    m = _mirror(args, tr, obs, "bls")
    for alpha in (alpha0, smooth):
        for lam in ((0.5, 0.1), (500.0, 100.0)):
            g = _gpu_eval(tr, alpha, start, goal, *lam)
            c = m.eval(alpha, start, goal, *lam)
            for k in ("q", "v", "loss", "toc", "grad"):
                assert np.array_equal(g[k], c[k]), (T, n_obs, lam, k, np.abs(g[k] - c[k]).max())
            assert np.array_equal(g["fulfilled"], c["fulfilled"])


@pytest.mark.parametrize("T,n_obs", [(50, 13), (50, 80), (256, 20)])
def test_strict_reciprocal_slow_path_far_obstacles(cuda_ready, T, n_obs):
    """rcp_block (csrc/fgd_device.cuh): strict mode takes the packed Newton step only while every m = 1 + |f - o|^2 of a
    block is below 2^126; a block with an absurdly distant obstacle goes through __frcp_rn (denormal reciprocals, m = inf).
    Obstacles at 1e19 (m ~ 1e38 >= 2^126, 1/m denormal) and 1e20 (m overflows to inf, 1/m = 0) among ordinary ones:
    still bit-identical to the oracle's IEEE 1 / m, per evaluation and over a short optimisation."""
    args, tr, obs, start, goal, alpha0 = _setup(T=T, n_obs=n_obs, B=24, seed=3, max_inner_iteration=10, max_outer_iteration=2)
    obs = np.array(obs, np.float32).copy()
    obs[1] = (1e19, -2.0); obs[n_obs // 2] = (0.5, -1e20); obs[-1] = (-1e19, 1e19); obs[5] = (3e18, 4e18)
    tr.set_obstacles(obs)
    m = _mirror(args, tr, obs, "bls")
    for lam in ((0.5, 0.1), (500.0, 100.0)):
        g = _gpu_eval(tr, alpha0, start, goal, *lam)
        c = m.eval(alpha0, start, goal, *lam)
        for k in ("loss", "toc", "grad"):
            assert np.array_equal(g[k], c[k]), (T, n_obs, lam, k)
    a, fs, is_ = _gpu_optimize(tr, "bls", alpha0, start, goal)
    ca, cfs, cis = m.optimize(alpha0, start, goal)
    assert np.array_equal(is_.cpu().numpy(), cis) and np.array_equal(a.cpu().numpy(), ca)


def test_eval_fast_math_within_tolerance(cuda_ready, oracle_vectors):
    """Fast mode vs the committed NumPy-oracle vectors (FP64): loss rel <= 1e-5, grad rel <= 1e-5."""
    args, tr, obs, *_ = _setup(strict=False, B=1)
    v = oracle_vectors
    for li in (0, 1):
        lam = v["lam"][li]
        g = _gpu_eval(tr, v["alpha_smooth"], v["start"], v["goal"], float(lam[0]), float(lam[1]))
        k = f"smooth_l{li}_f64"
        np.testing.assert_allclose(g["loss"], v[k + "_loss"], rtol=1e-5)
        for b in range(len(g["loss"])):
            rel = np.linalg.norm(g["grad"][b] - v[k + "_grad"][b]) / np.linalg.norm(v[k + "_grad"][b])
            assert rel < 1e-5, rel
        assert np.abs(g["q"] - v[k + "_q"]).max() < 1e-5
        assert np.array_equal(g["fulfilled"].astype(bool), v[k + "_ful"])


def test_eval_lambda_max_override(cuda_ready):
    """main.py:141-142 evaluates with lambda_max_cost 0 (average) and 1 (max)."""
    args, tr, obs, start, goal, alpha0 = _setup(B=8)
    m = _mirror(args, tr, obs, "bls")
    for lm in (0.0, 1.0):
        m.cfg.lam_max = lm
        c = m.eval(alpha0, start, goal, 0.0, 0.0)
        g = _gpu_eval(tr, alpha0, start, goal, 0.0, 0.0, lm)
        assert np.array_equal(g["loss"], c["loss"])


@pytest.mark.parametrize("mode,wpt,T,n_obs,B,over", [
    ("bls", 1, 50, 11, 200, {}),
    ("bls", 1, 50, 11, 333, {}),
    ("bls", 1, 64, 11, 61, {}),
    ("gd", 1, 50, 11, 200, {"max_outer_iteration": 1}),
    ("gd", 1, 50, 11, 96, {}),
    ("bls", 1, 24, 30, 100, {}),
    ("bls", 1, 7, 3, 37, {}),
    ("bls", 1, 33, 11, 24, {}),
    ("bls", 2, 100, 40, 48, {"max_inner_iteration": 40, "max_outer_iteration": 3}),
    ("bls", 4, 256, 200, 16, {"max_inner_iteration": 12, "max_outer_iteration": 2}),
    ("bls", 2, 65, 11, 40, {"max_inner_iteration": 30, "max_outer_iteration": 2}),
    ("gd", 2, 128, 25, 32, {"max_inner_iteration": 25, "max_outer_iteration": 2}),
    ("bls", 4, 129, 17, 24, {"max_inner_iteration": 20, "max_outer_iteration": 2}),
    ("gd", 4, 200, 64, 20, {"max_inner_iteration": 15, "max_outer_iteration": 1}),
    ("bls", 1, 50, 11, 64, {"constraint_violating_dependant_loss": False, "lambda_max_cost": 0.25}),
    ("bls", 1, 50, 300, 40, {"max_inner_iteration": 30, "max_outer_iteration": 2}),      # two-chain obstacle sums (share_split)
    ("gd", 1, 33, 101, 30, {"max_inner_iteration": 30, "max_outer_iteration": 2}),
])
def test_optimize_bit_exact_strict(cuda_ready, mode, wpt, T, n_obs, B, over):
    """Whole optimisation (all outer / inner / line-search iterations) bit-identical to the oracle:
    final alpha, penalty weights, step size, loss, counters and decision hash of every trajectory."""
    args, tr, obs, start, goal, alpha0 = _setup(T=T, n_obs=n_obs, B=B, seed=B, **over)
    a, fs, is_ = _gpu_optimize(tr, mode, alpha0, start, goal)
    ca, cfs, cis = _mirror(args, tr, obs, mode).optimize(alpha0, start, goal)
    is_g, fs_g = is_.cpu().numpy(), fs.cpu().numpy()
    assert (is_g[:, M.I_STATUS] == M.ST_DONE).all()
    bad = np.where((is_g != cis).any(1))[0]
    assert len(bad) == 0, (len(bad), bad[:5], is_g[bad[:3]], cis[bad[:3]])
    assert np.array_equal(a.cpu().numpy(), ca)
    assert np.array_equal(fs_g[:, :6], cfs[:, :6])
    assert tr.handle.launch_geometry(B)["warps_per_trajectory"] == wpt


def test_budgeted_launches_and_resume_bit_exact(cuda_ready):
    args, tr, obs, start, goal, alpha0 = _setup(B=96, seed=5)
    ca, cfs, cis = _mirror(args, tr, obs, "bls").optimize(alpha0, start, goal)
    a, fs, is_ = _gpu_optimize(tr, "bls", alpha0, start, goal, budget=9)
    n = 1
    while not (is_.cpu().numpy()[:, M.I_STATUS] == M.ST_DONE).all():
        a, fs, is_ = _gpu_optimize(tr, "bls", a, start, goal, budget=9, state=(fs, is_))
        n += 1
        assert n < 1000
    assert n > 5
    assert np.array_equal(a.cpu().numpy(), ca) and np.array_equal(is_.cpu().numpy(), cis)


def test_dynamic_obstacles_track_the_oracle(cuda_ready):
    """Config 4 semantics: the live obstacle set is replaced (count changes too) every 8 inner
    iterations through fgd_set_obstacles_async -- same handle, no recompilation -- and the result
    equals the oracle run with the identical swap schedule, bit for bit."""
    from irm_motion_planning_b200.workloads import obstacle_swap
    args, tr, obs, start, goal, alpha0 = _setup(B=128, n_obs=256, seed=9)
    m = _mirror(args, tr, obs, "bls")
    ca, cfs, cis = alpha0.copy(), *m.new_state(128)
    a, fs, is_ = alpha0, None, None
    h_before = tr.handle
    for k in range(400):
        cur = obs if k == 0 else obstacle_swap(k, seed=9)
        tr.set_obstacles(cur)
        m.set_obstacles(cur)
        a, fs, is_ = _gpu_optimize(tr, "bls", a, start, goal, budget=8, state=None if fs is None else (fs, is_))
        ca, cfs, cis = m.optimize(ca, start, goal, cfs, cis, budget=8)
        assert tr.handle.obstacle_count() == len(cur)
        if (cis[:, M.I_STATUS] == M.ST_DONE).all():
            break
    assert k >= 3 and tr.handle is h_before
    assert np.array_equal(is_.cpu().numpy(), cis) and np.array_equal(a.cpu().numpy(), ca)


def test_fast_math_end_to_end_distribution(cuda_ready):
    """Product default (rcp.approx): same distribution of outcomes as the oracle on 512 problems."""
    args, tr, obs, start, goal, alpha0 = _setup(B=512, strict=False, seed=21)
    a, fs, is_ = _gpu_optimize(tr, "bls", alpha0, start, goal)
    ca, cfs, cis = _mirror(args, tr, obs, "bls").optimize(alpha0, start, goal)
    is_g, fs_g = is_.cpu().numpy(), fs.cpu().numpy()
    same_trace = (is_g[:, M.I_HASH] == cis[:, M.I_HASH]).mean()
    print(f"fast-math: identical decision traces for {100 * same_trace:.1f}% of 512 trajectories")
    assert abs(is_g[:, M.I_FULFILLED].mean() - cis[:, M.I_FULFILLED].mean()) < 0.06
    assert abs(is_g[:, M.I_INNER_TOTAL].mean() / cis[:, M.I_INNER_TOTAL].mean() - 1) < 0.08
    both = (is_g[:, M.I_FULFILLED] == 1) & (cis[:, M.I_FULFILLED] == 1)
    rel = np.abs(fs_g[both, M.F_TOC] - cfs[both, M.F_TOC]) / cfs[both, M.F_TOC]
    assert np.median(rel) < 1e-2
    dq = np.abs(np.einsum("ij,bjk->bik", tr.km, a.cpu().numpy() - ca) @ tr.jac).reshape(512, -1).max(1)
    print(f"fast-math: joint-angle error vs oracle median {np.median(dq):.2e} p90 {np.quantile(dq, 0.9):.2e}")
    assert np.median(dq[both]) < 5e-2


def test_default_problem_matches_reference_goldens(cuda_ready, reference_results):
    """Config 1 through the drop-in class: same verdict, costs within 1e-2 relative and joint
    angles within 5e-2 rad of the reference's shipped trajectory_result.txt."""
    from irm_motion_planning_b200.optimizer_BLS import BacktrackingLineSearchOptimizer
    from irm_motion_planning_b200.workloads import default_args
    opt = BacktrackingLineSearchOptimizer(default_args(), warmup=False)
    alpha = opt.optimize()
    tr, env = opt.trajectory, opt.env
    avg = tr.compute_trajectory_cost(alpha, env.obstacles, env.start_config, env.goal_config, 0, 0, 0)
    mx = tr.compute_trajectory_cost(alpha, env.obstacles, env.start_config, env.goal_config, 0, 0, 1)
    assert tr.constraintsFulfilledVerbose(alpha, env.start_config, env.goal_config, verbose=False)
    assert abs(avg - 1.685) / 1.685 < 1e-2 and abs(mx - 2.196) / 2.196 < 1e-2
    q = tr.evaluate(alpha, tr.km, tr.jac)
    assert np.abs(q - reference_results["trajectory_result"]).max() < 5e-2


def test_plain_loop_records_iterates_like_extended_vis(cuda_ready, reference_results):
    from irm_motion_planning_b200.optimizer_BLS import BacktrackingLineSearchOptimizer
    from irm_motion_planning_b200.workloads import default_args
    opt = BacktrackingLineSearchOptimizer(default_args(extended_vis=True, jit_loop=False, strict_math=True), warmup=False)
    alpha, p = opt.optimize()
    ser = reference_results["trajectory_series"].reshape(-1, 50, 3)
    p = np.array(p)
    assert 60 <= len(p) <= 200
    assert np.abs(p[:17] - ser[:17]).reshape(17, -1).max() < 3e-2     # same early iterates as the reference's run
    opt2 = BacktrackingLineSearchOptimizer(default_args(strict_math=True), warmup=False)
    assert np.array_equal(opt2.optimize().cpu().numpy(), alpha.cpu().numpy())   # per-iteration launches == one launch


def test_host_buffer_entry_equals_device_path(cuda_ready):
    from irm_motion_planning_b200.batch import BatchedFGD
    args, tr, obs, start, goal, alpha0 = _setup(B=70, seed=2)
    a, fs, is_ = _gpu_optimize(tr, "gd", alpha0, start, goal)
    res = BatchedFGD(tr, "gd").optimize_host(alpha0, start, goal)
    assert np.array_equal(res.alpha, a.cpu().numpy()) and np.array_equal(res.istate, is_.cpu().numpy())
    assert res.done.all()
    # the in-place entry (resumable state travels with the call) gives the same bits
    from irm_motion_planning_b200 import backend
    a2 = alpha0.copy()
    fs2, is2 = np.zeros((70, backend.FSTATE), np.float32), np.zeros((70, backend.ISTATE), np.int32)
    tr.handle.optimize_host("gd", 70, a2, start, goal, fs2, is2)
    assert np.array_equal(a2, res.alpha) and np.array_equal(is2, res.istate) and np.array_equal(fs2, res.fstate)
    assert np.array_equal(alpha0, _setup(B=70, seed=2)[5])          # optimize_host (io form) left its input untouched


@pytest.mark.parametrize("mode,T,n_obs,B", [("gd", 50, 11, 300), ("bls", 50, 11, 257), ("bls", 129, 20, 40), ("bls", 256, 64, 9)])
def test_zero_copy_host_io_equals_device_path(cuda_ready, mode, T, n_obs, B):
    """fgd_optimize_host_io on page-locked buffers: the kernel reads its inputs from and writes its results to HOST
    memory directly (no staging copies).  Same bits as the device-resident path and as the oracle; pageable buffers
    take the staged path and give the same bits again."""
    import torch
    from irm_motion_planning_b200 import backend
    from irm_motion_planning_b200.batch import BatchedFGD
    args, tr, obs, start, goal, alpha0 = _setup(T=T, n_obs=n_obs, B=B, seed=5)
    a, fs, is_ = _gpu_optimize(tr, mode, alpha0, start, goal)
    eng = BatchedFGD(tr, mode)
    pin = lambda x: torch.as_tensor(x).clone().pin_memory()
    a_in, s_pin, g_pin = pin(alpha0), pin(start), pin(goal)
    out_a = torch.full((B, T, 3), float("nan")).pin_memory()
    out_f = torch.full((B, backend.FSTATE), float("nan")).pin_memory()
    out_i = torch.full((B, backend.ISTATE), -7, dtype=torch.int32).pin_memory()
    z0, l0 = tr.handle.zero_copy_calls(), tr.handle.kernel_launches()
    eng.optimize_pinned(a_in, s_pin, g_pin, out_a, out_f, out_i)
    assert tr.handle.zero_copy_calls() == z0 + 1 and tr.handle.kernel_launches() == l0 + 1
    assert np.array_equal(out_a.numpy(), a.cpu().numpy())
    assert np.array_equal(out_i.numpy(), is_.cpu().numpy()) and np.array_equal(out_f.numpy(), fs.cpu().numpy())
    assert np.array_equal(a_in.numpy(), alpha0)                      # the input buffer is only read
    ca, cfs, cis = _mirror(args, tr, obs, mode).optimize(alpha0, start, goal)
    assert np.array_equal(out_a.numpy(), ca) and np.array_equal(out_i.numpy(), cis)
    res = eng.optimize_host(alpha0, start, goal)                      # pageable NumPy buffers: staged copies
    assert tr.handle.zero_copy_calls() == z0 + 1
    assert np.array_equal(res.alpha, ca) and np.array_equal(res.istate, cis) and np.array_equal(res.fstate, out_f.numpy())


def test_result_independent_of_batch_position(cuda_ready):
    """Batch-vs-loop consistency: a trajectory's result does not depend on where it sits in the
    batch, which trajectory shares its warp, or the batch size (fast-math mode, the product default)."""
    args, tr, obs, start, goal, alpha0 = _setup(B=300, strict=False, seed=4)
    a1, fs1, is1 = _gpu_optimize(tr, "bls", alpha0, start, goal)
    perm = np.random.default_rng(0).permutation(300)
    a4, fs4, is4 = _gpu_optimize(tr, "bls", alpha0[perm], start[perm], goal[perm])
    assert np.array_equal(a1.cpu().numpy()[perm], a4.cpu().numpy())
    assert np.array_equal(is1.cpu().numpy()[perm], is4.cpu().numpy())
    a7, _, is7 = _gpu_optimize(tr, "bls", alpha0[:7], start[:7], goal[:7])
    assert np.array_equal(a1.cpu().numpy()[:7], a7.cpu().numpy()) and np.array_equal(is1.cpu().numpy()[:7], is7.cpu().numpy())
    # identical problems -> identical rows
    rep = np.repeat(alpha0[:1], 64, 0)
    ar, _, isr = _gpu_optimize(tr, "bls", rep, np.repeat(start[:1], 64, 0), np.repeat(goal[:1], 64, 0))
    assert (ar.cpu().numpy() == ar.cpu().numpy()[0]).all() and (isr.cpu().numpy() == isr.cpu().numpy()[0]).all()


@pytest.mark.parametrize("T,B", [(50, 1), (50, 4097), (256, 300)])
def test_device_init_trajectory_bit_exact_and_optimisable(cuda_ready, T, B):
    """fgd_init_trajectory == oracle mirror_init bit for bit; optimising from the
    device-initialised alpha is bit-identical to the oracle optimising from the oracle's init."""
    import torch
    args, tr, obs, start, goal, _ = _setup(T=T, B=B, seed=5)
    s, g = torch.as_tensor(start, device="cuda"), torch.as_tensor(goal, device="cuda")
    a_dev = tr.initTrajectoryDevice(s, g)
    a_ref = M.init_trajectory(*tr.init_basis(), start, goal)
    assert np.array_equal(a_dev.cpu().numpy(), a_ref)
    if B <= 300 and T == 50:
        a, fs, is_ = _gpu_optimize(tr, "bls", a_dev, start, goal)
        ca, cfs, cis = _mirror(args, tr, obs, "bls").optimize(a_ref, start, goal)
        assert np.array_equal(is_.cpu().numpy(), cis) and np.array_equal(a.cpu().numpy(), ca)


@pytest.mark.parametrize("T,n_obs,mode", [(50, 11, "bls"), (100, 30, "gd"), (256, 64, "bls")])
def test_whole_arm_cost_bit_exact(cuda_ready, T, n_obs, mode):
    """obstacle cost over all joint positions (ARM kernels) -- per evaluation and whole
    optimisations bit-identical to the mirror oracle in strict mode."""
    over = {"whole_arm_cost": True, "max_inner_iteration": 25, "max_outer_iteration": 2}
    args, tr, obs, start, goal, alpha0 = _setup(T=T, n_obs=n_obs, B=20, seed=T + 3, **over)
    m = _mirror(args, tr, obs, mode)
    assert m.cfg.whole_arm == 1
    for lam in ((0.5, 0.1), (50.0, 10.0)):
        g = _gpu_eval(tr, alpha0, start, goal, *lam)
        c = m.eval(alpha0, start, goal, *lam)
        for k in ("q", "v", "loss", "toc", "grad"):
            assert np.array_equal(g[k], c[k]), (k, np.abs(g[k] - c[k]).max())
    a, fs, is_ = _gpu_optimize(tr, mode, alpha0, start, goal)
    ca, cfs, cis = m.optimize(alpha0, start, goal)
    assert np.array_equal(is_.cpu().numpy(), cis)
    assert np.array_equal(a.cpu().numpy(), ca)
    # and it is a different objective from the end-effector cost
    args0, tr0, *_ = _setup(T=T, n_obs=n_obs, B=20, seed=T + 3)
    tr0.set_obstacles(obs)
    g0 = _gpu_eval(tr0, alpha0, start, goal, 0.5, 0.1)
    assert (g["toc"] > g0["toc"]).all()


def test_warm_start_replanning_bit_exact(cuda_ready):
    """a stream of scene updates, every plan warm-started from the previous solution through
    the persistent handle (async obstacle upload, no re-creation); the oracle applies the same schedule."""
    from irm_motion_planning_b200.batch import BatchedFGD
    from irm_motion_planning_b200.replan import WarmStartPlanner
    args, tr, obs, start, goal, alpha0 = _setup(B=12, seed=21)
    opt = type("Opt", (), {"trajectory": tr, "engine": BatchedFGD(tr, "bls")})()
    planner = WarmStartPlanner(opt, start, goal, alpha0)
    rng = np.random.default_rng(3)
    h0 = tr.handle
    ca = alpha0
    for step in range(4):
        obs = (np.asarray(obs, np.float32) + rng.normal(0, 0.05, np.shape(obs))).astype(np.float32)
        res = planner.update(obs)
        ca, cfs, cis = _mirror(args, tr, obs, "bls").optimize(ca, start, goal)
        assert np.array_equal(res.istate.cpu().numpy(), cis), step
        assert np.array_equal(res.alpha.cpu().numpy(), ca), step
        assert np.array_equal(res.fstate.cpu().numpy()[:, :6], cfs[:, :6]), step
    assert tr.handle is h0 and planner.plans == 4
    assert planner.trajectory_points(0).shape == (50, 3)


def test_argmin_per_problem(cuda_ready):
    import torch
    from irm_motion_planning_b200.batch import BatchedFGD, BatchResult
    from irm_motion_planning_b200 import backend
    args, tr, *_ = _setup(B=1)
    rng = np.random.default_rng(0)
    P, R = 37, 50
    fs = np.zeros((P * R, 8), np.float32); is_ = np.zeros((P * R, 8), np.int32)
    fs[:, backend.F_TOC] = rng.uniform(1, 3, P * R)
    is_[:, backend.I_FULFILLED] = rng.uniform(size=P * R) < 0.3
    is_[5 * R:6 * R, backend.I_FULFILLED] = 0                       # a problem without any fulfilled restart
    fs[7 * R + 3, backend.F_TOC] = fs[7 * R + 9, backend.F_TOC] = 0.5   # tie -> lowest index
    is_[7 * R + 3, backend.I_FULFILLED] = is_[7 * R + 9, backend.I_FULFILLED] = 1
    res = BatchResult(None, torch.as_tensor(fs, device="cuda"), torch.as_tensor(is_, device="cuda"))
    cost, idx = BatchedFGD(tr, "bls").best_per_problem(res, P, R, index_offset=1000)
    cost, idx = cost.cpu().numpy(), idx.cpu().numpy()
    for p in range(P):
        c = fs[p * R:(p + 1) * R, backend.F_TOC].copy()
        ok = is_[p * R:(p + 1) * R, backend.I_FULFILLED] == 1
        if ok.any():
            c[~ok] = np.inf
        assert idx[p] == 1000 + p * R + int(np.argmin(c)) and cost[p] == c.min()


def test_edge_cases_and_errors(cuda_ready):
    import torch
    from irm_motion_planning_b200 import backend
    args, tr, obs, start, goal, alpha0 = _setup(B=3, capacity=16, seed=8)
    m = _mirror(args, tr, obs, "bls")
    # B = 1 and B smaller than one warp's slots
    for B in (1, 3):
        a, fs, is_ = _gpu_optimize(tr, "bls", alpha0[:B], start[:B], goal[:B])
        ca, cfs, cis = m.optimize(alpha0[:B], start[:B], goal[:B])
        assert np.array_equal(a.cpu().numpy(), ca) and np.array_equal(is_.cpu().numpy(), cis)
    # zero obstacles: the obstacle term vanishes
    tr.set_obstacles(np.zeros((0, 2), np.float32))
    assert tr.handle.obstacle_count() == 0
    g = _gpu_eval(tr, alpha0, start, goal, 0.5, 0.1)
    assert (g["toc"] == 0).all()
    # capacity overflow is an error code, not a crash
    with pytest.raises(backend.FgdError) as ei:
        tr.handle.set_obstacles(np.zeros((17, 2), np.float32))
    assert ei.value.status == 4
    # a capacity whose staging buffer cannot fit in shared memory is refused at creation, large-but-fitting ones work
    from irm_motion_planning_b200.trajectory import Trajectory as _T
    with pytest.raises(backend.FgdError) as ei:
        _T(args, obstacle_capacity=200000)
    assert ei.value.status == 4
    big = _T(args, obstacle_capacity=8192, strict_math=True)
    rng = np.random.default_rng(1)
    many = rng.uniform(-4, 4, (8000, 2)).astype(np.float32)
    big.set_obstacles(many)
    gb = {k: v.cpu().numpy() for k, v in big._eval(alpha0, None, start, goal, 0.5, 0.1, -1.0, ("loss", "grad")).items()}
    cb = _mirror(args, big, many, "bls").eval(alpha0, start, goal, 0.5, 0.1)
    assert np.array_equal(gb["loss"], cb["loss"]) and np.array_equal(gb["grad"], cb["grad"])
    # already finished trajectories are left untouched by another launch
    tr.set_obstacles(obs)
    a, fs, is_ = _gpu_optimize(tr, "bls", alpha0, start, goal)
    a2, fs2, is2 = _gpu_optimize(tr, "bls", a, start, goal, state=(fs.clone(), is_.clone()))
    assert torch.equal(a, a2) and torch.equal(is_, is2)
    # GD with more outer iterations than learning rates is rejected (optimizer_GD.py:34-36)
    from irm_motion_planning_b200.workloads import default_args
    from irm_motion_planning_b200.trajectory import Trajectory
    tr_bad = Trajectory(default_args(gd_lr=[1e-3, 1e-4], max_outer_iteration=5))
    with pytest.raises(backend.FgdError):
        _gpu_optimize(tr_bad, "gd", alpha0, start, goal)


def test_full_size_config2_properties(cuda_ready):
    """BASELINE config 2 at full size (GD, B=4096): every trajectory retires, loss decreased,
    S-independence holds at this size, and a random 256-subset equals the oracle bit for bit (strict)."""
    from irm_motion_planning_b200.workloads import make_workload, initial_alpha
    from irm_motion_planning_b200.trajectory import Trajectory
    wl = make_workload("c2")
    tr = Trajectory(wl.args, strict_math=True)
    tr.set_obstacles(wl.obstacles)
    alpha0, start, goal = initial_alpha(wl, tr)
    a, fs, is_ = _gpu_optimize(tr, "gd", alpha0, start, goal)
    is_g = is_.cpu().numpy()
    assert (is_g[:, M.I_STATUS] == M.ST_DONE).all() and (is_g[:, M.I_INNER_TOTAL] >= 1).all()
    l0 = _gpu_eval(tr, alpha0, start, goal, 0.5, 0.1)["loss"]
    assert (fs.cpu().numpy()[:, M.F_LOSS] <= l0 + 1e-6).all()
    sub = np.random.default_rng(0).choice(4096, 256, replace=False)
    ca, cfs, cis = _mirror(wl.args, tr, wl.obstacles, "gd").optimize(alpha0[sub], start[sub], goal[sub])
    assert np.array_equal(a.cpu().numpy()[sub], ca) and np.array_equal(is_g[sub], cis)


def test_config3_shape_subset(cuda_ready):
    """Config 3 shape (T=256, 1024 obstacles, BLS) on 96 trajectories with bounded iterations: bit-exact (strict)."""
    args, tr, obs, start, goal, alpha0 = _setup(T=256, n_obs=1024, B=96, seed=3, max_inner_iteration=6, max_outer_iteration=2)
    a, fs, is_ = _gpu_optimize(tr, "bls", alpha0, start, goal)
    ca, cfs, cis = _mirror(args, tr, obs, "bls").optimize(alpha0, start, goal)
    assert np.array_equal(is_.cpu().numpy(), cis) and np.array_equal(a.cpu().numpy(), ca)


def test_config3_shape_to_convergence(cuda_ready):
    """Config 3 shape (T=256, 1024 obstacles, BLS, 4 warps per trajectory), the reference's full iteration limits
    (200 inner x 10 outer), 64 trajectories run to convergence: bit-exact (strict)."""
    args, tr, obs, start, goal, alpha0 = _setup(T=256, n_obs=1024, B=64, seed=11)
    a, fs, is_ = _gpu_optimize(tr, "bls", alpha0, start, goal)
    ca, cfs, cis = _mirror(args, tr, obs, "bls").optimize(alpha0, start, goal)
    is_g = is_.cpu().numpy()
    assert (is_g[:, M.I_STATUS] == M.ST_DONE).all()
    assert is_g[:, M.I_INNER_TOTAL].mean() > 20 and is_g[:, M.I_OUTER].max() >= 2      # real runs, not a bounded prefix
    assert np.array_equal(is_g, cis) and np.array_equal(a.cpu().numpy(), ca)
    assert np.array_equal(fs.cpu().numpy()[:, :6], cfs[:, :6])


def _oracle_sweep_winners(cfs, cis, P, R):
    toc, ful = cfs[:, M.F_TOC].reshape(P, R), cis[:, M.I_FULFILLED].reshape(P, R).astype(bool)
    c = np.where(ful, toc, np.inf)
    c = np.where(np.isinf(c).all(1, keepdims=True), toc, c)
    r = c.argmin(1)
    return r, toc[np.arange(P), r], ful[np.arange(P), r]


def test_config5_restart_sweep_pipeline_equals_oracle(cuda_ready):
    """Config 5 end to end on a small sweep (12 problems x 32 restarts): restart inputs -> optimise (strict) ->
    fgd_argmin_per_problem -> winners, against mirror optimise -> NumPy argmin: identical (cost, index, fulfilled).
    Then the same sweep as two, four and eight restart-axis shards (what 2 / 4 / 8 ranks run), reduced by the elementwise min of
    the order keys (what the all-gather feeds): identical winners, independent of the number of shards."""
    import torch
    from irm_motion_planning_b200.batch import BatchedFGD, decode_keys, restart_shard
    from irm_motion_planning_b200.trajectory import Trajectory
    from irm_motion_planning_b200.workloads import initial_alpha, make_workload
    P, R = 12, 32
    wl = make_workload("c5", B=P * 256, seed=4)          # 12 problems ...
    wl.n_restarts = R                                   # ... same generator, fewer restarts per problem
    tr = Trajectory(wl.args, strict_math=True)
    tr.set_obstacles(wl.obstacles)
    alpha0, start, goal = initial_alpha(wl, tr, 4)
    assert alpha0.shape == (P * R, 50, 3)
    eng = BatchedFGD(tr, "bls")
    ca, cfs, cis = _mirror(wl.args, tr, wl.obstacles, "bls").optimize(alpha0, start, goal)
    r_ref, c_ref, f_ref = _oracle_sweep_winners(cfs, cis, P, R)
    assert 0 < f_ref.sum()                              # the sweep finds fulfilled winners

    def run(world):
        keys = []
        for rank in range(world):
            lo, hi = restart_shard(R, rank, world)
            # the rank's block of the sweep (sliced from the full generation: a rank that generates only its block gets the
            # same start / goal / via-offsets, but LAPACK's blocked solve may round a column differently for another nrhs)
            blk = lambda x: np.ascontiguousarray(x.reshape((P, R) + x.shape[1:])[:, lo:hi].reshape((-1,) + x.shape[1:]))
            a_s, s_s, g_s = blk(alpha0), blk(start), blk(goal)
            _, s_gen, g_gen = initial_alpha(wl, tr, 4, restarts=(lo, hi))
            assert np.array_equal(s_gen, s_s) and np.array_equal(g_gen, g_s)
            a, fs, is_ = _gpu_optimize(tr, "bls", a_s, s_s, g_s)
            keys.append(eng.best_keys(fs, is_, P, hi - lo, index_offset=lo, problem_stride=R))
        return decode_keys(torch.stack(keys).min(dim=0).values)

    for world in (1, 2, 4, 8):
        cost, idx, ful = run(world)
        assert np.array_equal(idx.cpu().numpy(), np.arange(P) * R + r_ref), world
        assert np.array_equal(cost.cpu().numpy(), c_ref), world
        assert np.array_equal(ful.cpu().numpy(), f_ref), world
    # (cost, index) outputs of the same kernel
    a, fs, is_ = _gpu_optimize(tr, "bls", alpha0, start, goal)
    c2, i2 = eng.best_per_problem(type("R", (), {"fstate": fs, "istate": is_})(), P, R)
    assert np.array_equal(i2.cpu().numpy(), np.arange(P) * R + r_ref) and np.array_equal(c2.cpu().numpy(), c_ref)


def test_fast_math_per_trajectory_tolerance(cuda_ready):
    """north_star: EACH final trajectory within a stated FP32 tolerance of the reference path.  Fast mode (the product
    default: rcp.approx, 1 ulp) against the oracle on 1024 trajectories:
      * trajectories whose decision trace equals the oracle's (>= 90 % of them): max-abs joint-angle error <= 2e-3 rad
        and relative obstacle-cost error <= 1e-3 for EVERY one of them;
      * all trajectories: p99 of the joint-angle error <= 1e-1 rad, and at most 3 % outside the stated end-to-end
        tolerance of 5e-2 rad - those are the chaotic divergences of SURVEY 0.3-3 (a 1-ulp perturbation
        flips an Armijo / stop decision; two CPU builds of the reference differ by 3.6e-2...6.5e-2 rad the same way);
      * same fulfilment verdict for >= 97 %; relative obstacle-cost error of the fulfilled ones: p99 <= 3e-2."""
    args, tr, obs, start, goal, alpha0 = _setup(B=1024, strict=False, seed=33)
    a, fs, is_ = _gpu_optimize(tr, "bls", alpha0, start, goal)
    ca, cfs, cis = _mirror(args, tr, obs, "bls").optimize(alpha0, start, goal)
    is_g, fs_g = is_.cpu().numpy(), fs.cpu().numpy()
    same = is_g[:, M.I_HASH] == cis[:, M.I_HASH]
    dq = np.abs(np.einsum("ij,bjk->bik", tr.km, a.cpu().numpy() - ca) @ tr.jac).reshape(1024, -1).max(1)
    rel = np.abs(fs_g[:, M.F_TOC] - cfs[:, M.F_TOC]) / cfs[:, M.F_TOC]
    both = (is_g[:, M.I_FULFILLED] == 1) & (cis[:, M.I_FULFILLED] == 1)
    print(f"fast-math per trajectory: same trace {same.mean():.3f}; same-trace max dq {dq[same].max():.2e} max rel cost {rel[same].max():.2e}; "
          f"all: dq p50 {np.median(dq):.2e} p99 {np.quantile(dq, 0.99):.2e} max {dq.max():.2e} outside 5e-2: {(dq > 5e-2).mean():.4f}; "
          f"fulfilled agree {(is_g[:, M.I_FULFILLED] == cis[:, M.I_FULFILLED]).mean():.4f}; rel cost p99 (both fulfilled) {np.quantile(rel[both], 0.99):.2e}")
    assert same.mean() >= 0.90
    assert dq[same].max() <= 2e-3 and rel[same].max() <= 1e-3
    assert np.quantile(dq, 0.99) <= 1e-1 and (dq > 5e-2).mean() <= 0.03
    assert (is_g[:, M.I_FULFILLED] == cis[:, M.I_FULFILLED]).mean() >= 0.97
    assert np.quantile(rel[both], 0.99) <= 3e-2


@pytest.mark.parametrize("extra,series", [([], False), (["--extended-vis", "true", "--jit-loop", "false", "--strict-math", "true"], True),
                                          (["--optimizer-name", "gd"], False), (["--batch", "40", "--n-measurements", "2"], False)])
def test_main_report_and_output_files(cuda_ready, tmp_path, monkeypatch, capsys, extra, series):
    """`python main.py [flags]` prints the reference's report lines (main.py:127,141-143) and writes
    trajectory_result.txt (50 x 3) / trajectory_series.txt (N x 150) in the np.savetxt format that
    visualization/visualization.py:91 loads with np.loadtxt."""
    from irm_motion_planning_b200 import main as fgd_main
    monkeypatch.chdir(tmp_path)
    fgd_main.main(extra)
    out = capsys.readouterr().out
    assert "setup object, jit-compile took" in out and "took" in out
    line = [l for l in out.splitlines() if l.startswith("result cost: ( avg")]
    assert len(line) == 1 and "constraint fulfiled" in line[0]
    avg, mx = float(line[0].split("avg")[1].split(",")[0]), float(line[0].split("max")[1].split(")")[0])
    assert 1.0 < avg < mx < 3.5
    if not extra:                                       # the default problem: the reference's own result (blog 1.69 / 2.19)
        assert abs(avg - 1.69) < 2e-2 and abs(mx - 2.19) < 2e-2 and line[0].rstrip().endswith("True")
        assert "ok start goal position" in out and "ok velocity limit with" in out
    res = np.loadtxt(tmp_path / "trajectory_result.txt")
    assert res.shape == (50, 3) and np.isfinite(res).all()
    if not extra or series:
        assert np.abs(res[0]).max() < 2e-2 and np.abs(res[-1] - np.array([1.2, 0.8, 0.3])).max() < 2e-2     # environment.py:14-15
    if series:
        ser = np.loadtxt(tmp_path / "trajectory_series.txt")
        assert ser.ndim == 2 and ser.shape[1] == 150 and 60 <= ser.shape[0] <= 200
        ref = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_results.npz"))["trajectory_series"]
        assert np.abs(ser[0] - ref[0]).max() < 2e-3    # row 0 = the fitted straight line, as in the reference's file
        assert f"({ser.shape[0]}, 50, 3)" in out        # main.py:152 prints the shape
    else:
        assert not (tmp_path / "trajectory_series.txt").exists()


def test_two_streams_on_one_handle(cuda_ready):
    """Two optimise launches of ONE handle enqueued on two streams may overlap on the device: each owns its work-queue
    counter, so both batches are fully processed and equal the single-stream results."""
    import torch
    from irm_motion_planning_b200.batch import BatchedFGD
    args, tr, obs, start, goal, alpha0 = _setup(B=3000, strict=False, seed=12)
    eng = BatchedFGD(tr, "bls")
    ref = [_gpu_optimize(tr, "bls", alpha0[i::2], start[i::2], goal[i::2]) for i in range(2)]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    ins = [(torch.as_tensor(alpha0[i::2], device="cuda").clone(), torch.as_tensor(start[i::2], device="cuda").contiguous(),
            torch.as_tensor(goal[i::2], device="cuda").contiguous(), *eng.new_state(1500)) for i in range(2)]
    torch.cuda.synchronize()
    for rep in range(3):
        for i, (st, x) in enumerate(zip(streams, ins)):
            with torch.cuda.stream(st):
                if rep:
                    x[3].zero_(); x[4].zero_()
                    x[0].copy_(torch.as_tensor(alpha0[i::2], device="cuda"))
                eng.optimize_device(*x)
        torch.cuda.synchronize()
        for i in range(2):
            assert (ins[i][4][:, M.I_STATUS] == M.ST_DONE).all()
            assert torch.equal(ins[i][0], ref[i][0]) and torch.equal(ins[i][4], ref[i][2])


@pytest.mark.parametrize("mode,strict,lo", [("bls", True, 192), ("gd", True, 192), ("bls", False, 192), ("bls", True, 20)])
def test_live_obstacle_updates_replay_bit_exact(cuda_ready, mode, strict, lo):
    """Config 4 without relaunches (fgd_optimize_live): ONE persistent launch while the host publishes new obstacle sets
    (count changes too) with fgd_set_obstacles_async on a side stream; every team polls the generation counter every 4
    inner iterations of its trajectory.  Which generation a trajectory saw at which iteration depends on timing, so the
    kernel records it (switch log) and the oracle REPLAYS that schedule per trajectory through its budgeted entry - the
    plain loop's semantics (optimizer_BLS.py:79,82,90): bit-identical alpha, counters and loss state (strict mode).
    The sets have 192-320 obstacles (20-320 in the last case), so the helper lanes of the shared obstacle loop are at work
    and, in the last case, switch on and off with the adopted set.
    Fast mode: the same replay through the budgeted CUDA entry points (relaunch path) gives identical bits."""
    import torch
    from irm_motion_planning_b200 import backend
    from irm_motion_planning_b200.batch import BatchedFGD
    from irm_motion_planning_b200.workloads import obstacle_swap
    B, POLL = 600, 4
    over = {} if mode == "bls" else {"max_outer_iteration": 3}
    args, tr, obs, start, goal, alpha0 = _setup(B=B, n_obs=256, seed=17, strict=strict, **over)
    # lo = 20: the obstacle count crosses the 64-obstacle threshold of the shared obstacle loop (share_split) from set to set
    sets = [np.asarray(obs, np.float32)] + [obstacle_swap(k, seed=17, lo=lo) for k in range(1, 40)]
    eng = BatchedFGD(tr, mode)
    a = torch.as_tensor(alpha0, device="cuda").clone()
    s, g = torch.as_tensor(start, device="cuda").contiguous(), torch.as_tensor(goal, device="cuda").contiguous()
    fs, is_ = eng.new_state(B)
    log = torch.zeros(B, backend.FGD_SWITCH_LOG, 2, dtype=torch.int32, device="cuda")
    gen0 = tr.handle.obstacle_generation                 # generation of `obs` (published by _setup)
    h_before = tr.handle
    n_pub = eng.optimize_live(a, s, g, fs, is_, sets, poll_every=POLL, period_us=100.0, switch_log=log, max_sets=45)
    torch.cuda.synchronize()
    assert tr.handle is h_before and tr.handle.obstacle_generation == gen0 + n_pub
    is_g, a_g, log = is_.cpu().numpy(), a.cpu().numpy(), log.cpu().numpy()
    assert (is_g[:, M.I_STATUS] == M.ST_DONE).all()
    n_sw = log[:, 0, 0]
    assert (n_sw >= 1).all() and n_sw.max() < backend.FGD_SWITCH_LOG
    assert n_pub >= 3 and (n_sw > 1).mean() > 0.2, (n_pub, n_sw.mean())      # the sets really changed under running trajectories
    set_of = lambda gen: sets[0] if gen <= gen0 else sets[(gen - gen0) % len(sets)]
    m = _mirror(args, tr, obs, mode)
    picks = np.random.default_rng(0).choice(B, 160, replace=False) if strict else np.arange(0)
    for b in picks:
        sched = log[b, 1:1 + n_sw[b]]
        assert sched[0, 0] == 0 and (np.diff(sched[:, 0]) > 0).all() and (sched[1:, 0] % POLL == 0).all()
        ca, cfs, cis = alpha0[b:b + 1].copy(), *m.new_state(1)
        for k, (it, gen) in enumerate(sched):
            m.set_obstacles(set_of(gen))
            budget = int(sched[k + 1, 0] - it) if k + 1 < len(sched) else -1
            ca, cfs, cis = m.optimize(ca, start[b:b + 1], goal[b:b + 1], cfs, cis, budget=budget)
        assert np.array_equal(cis[0], is_g[b]), (b, sched.tolist(), cis[0], is_g[b])
        assert np.array_equal(ca[0], a_g[b])
    if not strict:      # fast mode: replay through the relaunch path of the same library
        tr2 = type(tr)(args, strict_math=False)
        for b in np.random.default_rng(1).choice(B, 24, replace=False):
            sched = log[b, 1:1 + n_sw[b]]
            a2, st2 = alpha0[b:b + 1], None
            for k, (it, gen) in enumerate(sched):
                tr2.handle.set_obstacles(set_of(gen))
                budget = int(sched[k + 1, 0] - it) if k + 1 < len(sched) else -1
                a2, f2, i2 = _gpu_optimize(tr2, mode, a2, start[b:b + 1], goal[b:b + 1], budget=budget, state=st2)
                st2 = (f2, i2)
            assert np.array_equal(i2.cpu().numpy()[0], is_g[b]) and np.array_equal(a2.cpu().numpy()[0], a_g[b])
    # and a launch without updates equals the plain launch
    tr.set_obstacles(obs)
    a1, fs1, is1 = _gpu_optimize(tr, mode, alpha0[:64], start[:64], goal[:64])
    a3 = torch.as_tensor(alpha0[:64], device="cuda").clone()
    fs3, is3 = eng.new_state(64)
    tr.handle.optimize_live(mode, 64, a3, s[:64].contiguous(), g[:64].contiguous(), fs3, is3, POLL)
    torch.cuda.synchronize()
    assert torch.equal(a1, a3) and torch.equal(is1, is3)


def test_live_rejects_multi_warp_teams_and_oversized_capacity(cuda_ready):
    import torch
    from irm_motion_planning_b200 import backend
    from irm_motion_planning_b200.batch import BatchedFGD
    args, tr, obs, start, goal, alpha0 = _setup(T=100, n_obs=20, B=4, seed=1)
    eng = BatchedFGD(tr, "bls")
    fs, is_ = eng.new_state(4)
    with pytest.raises(backend.FgdError) as ei:
        tr.handle.optimize_live("bls", 4, torch.as_tensor(alpha0, device="cuda"), torch.as_tensor(start, device="cuda"),
                                torch.as_tensor(goal, device="cuda"), fs, is_, 8)
    assert ei.value.status == 2
    args, tr, obs, start, goal, alpha0 = _setup(T=50, n_obs=20, B=4, seed=1, capacity=8192)
    eng = BatchedFGD(tr, "bls")
    fs, is_ = eng.new_state(4)
    with pytest.raises(backend.FgdError) as ei:
        tr.handle.optimize_live("bls", 4, torch.as_tensor(alpha0, device="cuda"), torch.as_tensor(start, device="cuda"),
                                torch.as_tensor(goal, device="cuda"), fs, is_, 8)
    assert ei.value.status == 4


@pytest.mark.parametrize("T,B,over", [(50, 150, {}), (50, 296, {"max_bls_iteration": 6}), (33, 40, {"bls_alpha": 1e3, "max_outer_iteration": 2}),
                                      (64, 7, {"max_bls_iteration": 3, "bls_beta_minus": 0.7}), (50, 1, {})])
def test_speculative_line_search_equals_sequential(cuda_ready, monkeypatch, T, B, over):
    """north_star item 5: for batches smaller than the machine the Armijo candidates of a line search are evaluated IN
    PARALLEL (four warps per trajectory, four consecutive candidates per round) and the first accepting one in the
    reference's order is selected (optimizer_BLS.py:131-150).  The iterates, step sizes, counters and decision hash equal
    the sequential kernel's bit for bit - in fast mode too (same arithmetic) - and the oracle's in strict mode; candidate
    counts that are not a multiple of four and line searches that reject every candidate included."""
    from irm_motion_planning_b200.trajectory import Trajectory
    for strict in (True, False):
        args, tr, obs, start, goal, alpha0 = _setup(T=T, B=B, seed=B + 1, strict=strict, **over)
        s0 = tr.handle.speculative_launches()
        a, fs, is_ = _gpu_optimize(tr, "bls", alpha0, start, goal)
        assert tr.handle.speculative_launches() == s0 + 1
        monkeypatch.setenv("FGD_SPEC_MAX_BATCH", "0")
        tr_seq = Trajectory(args, strict_math=strict)
        monkeypatch.delenv("FGD_SPEC_MAX_BATCH")
        tr_seq.set_obstacles(obs)
        a2, fs2, is2 = _gpu_optimize(tr_seq, "bls", alpha0, start, goal)
        assert tr_seq.handle.speculative_launches() == 0
        import torch
        assert torch.equal(is_, is2) and torch.equal(a, a2) and torch.equal(fs[:, :6], fs2[:, :6])
        if strict:
            ca, cfs, cis = _mirror(args, tr, obs, "bls").optimize(alpha0, start, goal)
            assert np.array_equal(is_.cpu().numpy(), cis) and np.array_equal(a.cpu().numpy(), ca)
            if "bls_alpha" in over:
                assert (cis[:, M.I_ACCEPTS] == 0).all() and (cis[:, M.I_CAND_EVALS] >= 20).all()     # every candidate rejected
    # budgeted launches resume identically under speculation
    args, tr, obs, start, goal, alpha0 = _setup(T=50, B=20, seed=3)
    ca, cfs, cis = _mirror(args, tr, obs, "bls").optimize(alpha0, start, goal)
    a, fs, is_ = _gpu_optimize(tr, "bls", alpha0, start, goal, budget=5)
    for _ in range(1000):
        if (is_.cpu().numpy()[:, M.I_STATUS] == M.ST_DONE).all():
            break
        a, fs, is_ = _gpu_optimize(tr, "bls", a, start, goal, budget=5, state=(fs, is_))
    assert np.array_equal(a.cpu().numpy(), ca) and np.array_equal(is_.cpu().numpy(), cis)

