"""Command-line entry point: the reference's ``main.py`` flags driving the CUDA backend.

All 33 flags of the reference (main.py:17-98) keep their names, types and
defaults, so ``python main.py [flags]`` behaves as before: build the optimiser,
time ``optimize()`` ``n_measurements x n_times`` times, print the average and
maximum obstacle cost with the verbose constraint report, write
``trajectory_result.txt`` (and ``trajectory_series.txt`` with --extended-vis).
Batch flags (new, all optional) run B trajectories in one launch.
"""
from __future__ import annotations

import argparse
import time

import numpy as np


def _flag(text):
    return str(text).lower() == "true"


# (flag, type, default, help) -- the reference's table, main.py:17-98
_REFERENCE_FLAGS = (
    ("--profiling", _flag, False, "wrap the timed loop in a CUDA profiler range (reference: jax profiler)"),
    ("--extended-vis", _flag, False, "record every accepted iterate (implies the per-iteration launch loop)"),
    ("--n-measurements", int, 1, "number of timing measurements"),
    ("--n-times", int, 1, "optimisations per measurement"),
    ("--jit-loop", _flag, True, "true: whole optimisation in one kernel launch; false: one launch per inner iteration"),
    ("--n-timesteps", float, 50, "time samples / RKHS support points (coerced to int)"),
    ("--rbf-variance", float, 0.1, "RBF kernel width"),
    ("--jac-gaussian-mean", float, 0.15, "scale of the Gaussian perturbation of the joint-mixing matrix J"),
    ("--max-inner-iteration", int, 200, "inner (descent) iterations per outer iteration"),
    ("--loop-loss-reduction", float, 1e-3, "stop the inner loop when the loss decreases by less than this"),
    ("--max-outer-iteration", int, 10, "penalty (outer) iterations"),
    ("--lambda-constraint-increase", int, 10, "penalty multiplier applied after every unfulfilled outer iteration"),
    ("--lambda-sg-constraint", float, 0.5, "initial start/goal penalty weight"),
    ("--lambda-jl-constraint", float, 0.1, "initial joint-limit penalty weight"),
    ("--eps-position", float, 0.01, "start/goal position tolerance"),
    ("--eps-velocity", float, 0.01, "start/goal velocity tolerance"),
    ("--lambda-max-cost", float, 0.5, "weight of the max term in the obstacle cost"),
    ("--lambda-reg", float, 1e-4, "weight decay of the update"),
    ("--constraint-violating-dependant-loss", _flag, True, "limit penalties only where the safety limit is violated"),
    ("--joint-safety-limit", float, 0.98, "fraction of the joint limits at which the limit penalty switches on"),
    ("--max-bls-iteration", int, 20, "candidates per backtracking line search"),
    ("--bls-lr-start", float, 0.2, "line-search step at the start of every outer iteration"),
    ("--bls-alpha", float, 0.01, "Armijo sufficient-decrease constant"),
    ("--bls-beta_plus", float, 1.2, "step growth after an accepted candidate"),
    ("--bls-beta_minus", float, 0.5, "step shrink after a rejected candidate"),
    ("--n-joints", int, 3, "number of joints (only 3 is supported, as in the reference)"),
    ("--max-joint-velocity", float, 7, "joint velocity limit"),
    ("--max-joint-position", float, 2, "upper joint position limit"),
    ("--min-joint-position", float, -1, "lower joint position limit"),
)


def build_parser() -> argparse.ArgumentParser:
    ap = argparse.ArgumentParser(description=__doc__.splitlines()[0])
    for name, typ, default, text in _REFERENCE_FLAGS:
        ap.add_argument(name, type=typ, default=default, help=f"{text} (default: {default})")
    ap.add_argument("--optimizer-name", choices=["gd", "bls"], default="bls", help="optimiser (default: bls)")
    ap.add_argument("--gd-lr", type=float, nargs="+",
                    default=[2e-3, 1e-4, 1e-5, 1e-6, 1e-7, 1e-8, 1e-8, 1e-8, 1e-8, 1e-8],
                    help="GD step per outer iteration")
    ap.add_argument("--link-length", type=float, nargs="+", default=[1.5, 1.0, 0.5], help="link lengths")
    # new, batch-related flags (no effect on the reference behaviour when left at their defaults)
    ap.add_argument("--batch", type=int, default=1, help="independent trajectories optimised in one launch (default: 1)")
    ap.add_argument("--seed", type=int, default=0, help="seed of the synthetic start/goal sampler for --batch > 1")
    ap.add_argument("--n-obstacles", type=int, default=0, help="0: the reference scene; N>0: N random obstacles")
    ap.add_argument("--obstacle-capacity", type=int, default=1024, help="size of the device obstacle buffer")
    ap.add_argument("--strict-math", type=_flag, default=False, help="IEEE reciprocal instead of rcp.approx")
    ap.add_argument("--whole-arm-cost", type=_flag, default=False,
                    help="obstacle cost summed over all joint positions instead of the end effector only (the reference blog's extension)")
    ap.add_argument("--replan", type=int, default=0,
                    help="after the timed runs: N warm-started re-plans on a drifting obstacle set (blog: 50 Hz re-planning)")
    return ap


def parse_args(argv=None):
    return build_parser().parse_args(argv)


def main(argv=None):
    import torch

    from .environment import random_obstacles
    from .optimizer_BLS import BacktrackingLineSearchOptimizer
    from .optimizer_GD import GradientDescentOptimizer

    args = parse_args(argv)
    if args.optimizer_name == "bls":
        optimizer = BacktrackingLineSearchOptimizer(args)
    elif args.optimizer_name == "gd":
        optimizer = GradientDescentOptimizer(args)
    else:
        print("FATAL: not defined optimizer", args.optimizer_name)
        raise SystemExit(-1)

    if args.n_obstacles > 0:
        optimizer.env.obstacles = random_obstacles(args.n_obstacles, np.random.default_rng(args.seed))

    batch_inputs = None
    if args.batch > 1:
        from .workloads import sample_start_goal
        start, goal = sample_start_goal(args.batch, np.random.default_rng(args.seed))
        start[0], goal[0] = optimizer.env.start_config, optimizer.env.goal_config      # trajectory 0 = the reference problem
        batch_inputs = (optimizer.trajectory.initTrajectory(start, goal), start, goal)

    def run_once():
        if batch_inputs is None:
            return optimizer.optimize()
        return optimizer.optimize_batch(*batch_inputs)

    def multiple_optimizations():
        runtimes, result = [], None
        for _ in range(args.n_measurements):
            st = time.time()
            for _ in range(args.n_times):
                result = run_once()
                torch.cuda.synchronize()
            et = time.time()
            runtimes.append(1000 * (et - st) / args.n_times)
            print("took", runtimes[-1], "ms")
        if args.n_measurements > 1:
            print("runtimes in ms: mean", np.mean(runtimes), "stddev", np.std(runtimes))
        return result

    if args.profiling:
        torch.cuda.profiler.start()
        result = multiple_optimizations()
        torch.cuda.profiler.stop()
    else:
        result = multiple_optimizations()

    p = None
    if batch_inputs is not None:
        res = result
        ok = res.fulfilled
        print(f"batch of {args.batch}: {int(ok.sum())} fulfil the constraints, mean inner iterations "
              f"{res.inner_iterations.mean():.1f}, mean obstacle cost {res.obstacle_cost.mean():.4f}")
        cand = np.where(ok)[0]
        best = int(cand[np.argmin(res.obstacle_cost[cand])]) if len(cand) else int(np.argmin(res.obstacle_cost))
        result_alpha, start_c, goal_c = res.alpha[best], batch_inputs[1][best], batch_inputs[2][best]
        print("reporting trajectory", best)
    else:
        if args.extended_vis:
            result_alpha, p = result
        else:
            result_alpha = result
        start_c, goal_c = optimizer.env.start_config, optimizer.env.goal_config

    tr, env = optimizer.trajectory, optimizer.env
    avg_result_cost = tr.compute_trajectory_cost(result_alpha, env.obstacles, start_c, goal_c, 0, 0, 0)
    max_result_cost = tr.compute_trajectory_cost(result_alpha, env.obstacles, start_c, goal_c, 0, 0, 1)
    print("result cost: ( avg", avg_result_cost, ", max", max_result_cost, "). constraint fulfiled",
          tr.constraintsFulfilledVerbose(result_alpha, start_c, goal_c, verbose=True))

    if args.replan > 0:
        _replan_demo(args, optimizer, batch_inputs)

    np.savetxt("trajectory_result.txt", np.array(tr.evaluate(result_alpha, tr.km, tr.jac)))
    if args.extended_vis and p is not None:
        p_np = np.array(p)
        print(p_np.shape)
        np.savetxt("trajectory_series.txt", p_np.reshape((-1, args.n_joints * int(args.n_timesteps))))


def _replan_demo(args, optimizer, batch_inputs):
    """--replan N: the obstacles drift a little every step; each step re-optimises from the previous
    solution (WarmStartPlanner) and, for comparison, from the reference's cold start."""
    import torch

    from .replan import WarmStartPlanner
    rng = np.random.default_rng(args.seed + 7)
    env = optimizer.env
    if batch_inputs is None:
        start, goal, alpha0 = env.start_config[None], env.goal_config[None], None
    else:
        alpha0, start, goal = batch_inputs
    obstacles = np.asarray(env.obstacles, np.float32).copy()
    planner = WarmStartPlanner(optimizer, start, goal, alpha0)
    planner.update(obstacles)                                  # first plan = the cold start
    torch.cuda.synchronize()
    warm_ms, warm_it, cold_it, ok = [], [], [], []
    for _ in range(args.replan):
        obstacles = (obstacles + rng.normal(0.0, 0.03, obstacles.shape)).astype(np.float32)
        t0 = time.time()
        res = planner.update(obstacles)
        torch.cuda.synchronize()
        warm_ms.append(1000 * (time.time() - t0))
        warm_it.append(float(res.inner_iterations.mean())); ok.append(float(res.fulfilled.mean()))
        cold = optimizer.optimize_batch(planner.trajectory.initTrajectory(start, goal).reshape(-1, planner.trajectory.N_timesteps, 3),
                                        start, goal, obstacles)
        cold_it.append(float(cold.inner_iterations.mean()))
    print(f"re-planning x{args.replan} ({planner.B} trajectories): {np.mean(warm_ms):.3f} ms per plan "
          f"({1000 / np.mean(warm_ms):.0f} Hz), inner iterations warm {np.mean(warm_it):.1f} vs cold {np.mean(cold_it):.1f}, "
          f"constraints fulfilled {100 * np.mean(ok):.0f} %")


if __name__ == "__main__":
    main()
