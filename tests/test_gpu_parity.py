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


@pytest.mark.parametrize("T,n_obs", [(50, 11), (20, 11), (33, 37), (64, 5), (65, 9), (100, 64), (129, 20), (256, 300)])
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
