"""Small invocations of every kernel family (meant for compute-sanitizer --tool memcheck / racecheck / synccheck; the tool is
closed on this pool, so in round 2 the script only ran plain - the bit-exact parity tests are the bounds / race check here).
Covers: single-warp TMEM teams (T = 50 instance and runtime T), speculative replicas (shared-memory exchange), multi-warp
TMEM teams on named barriers (T = 100, 256), the LIVE instance with a publisher, zero-copy host I/O, the evaluation, argmin
and init kernels."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from irm_motion_planning_b200 import backend                                                  # noqa: E402
from irm_motion_planning_b200.batch import BatchedFGD                                         # noqa: E402
from irm_motion_planning_b200.environment import Environment, random_obstacles                # noqa: E402
from irm_motion_planning_b200.trajectory import Trajectory                                    # noqa: E402
from irm_motion_planning_b200.workloads import default_args, obstacle_swap, sample_start_goal  # noqa: E402


def run(T, B, mode, n_obs=11, live=False, **over):
    args = default_args(n_timesteps=float(T), **over)
    tr = Trajectory(args)
    rng = np.random.default_rng(T + B)
    obs = Environment().obstacles if n_obs == 11 else random_obstacles(n_obs, rng)
    tr.set_obstacles(obs)
    start, goal = sample_start_goal(B, rng)
    a0 = tr.initTrajectory(start, goal).reshape(B, T, 3)
    eng = BatchedFGD(tr, mode)
    a = torch.as_tensor(a0, device="cuda").clone()
    s, g = torch.as_tensor(start, device="cuda").contiguous(), torch.as_tensor(goal, device="cuda").contiguous()
    fs, is_ = eng.new_state(B)
    if live:
        sets = [np.asarray(obs, np.float32)] + [obstacle_swap(k, seed=1) for k in range(1, 6)]
        n = eng.optimize_live(a, s, g, fs, is_, sets, poll_every=2, period_us=200.0, max_sets=5)
    else:
        eng.optimize_device(a, s, g, fs, is_)
        n = 0
    torch.cuda.synchronize()
    assert (is_[:, backend.I_STATUS] == backend.ST_DONE).all()
    out = tr._eval(a, None, start, goal, 0.5, 0.1, -1.0, ("loss", "grad", "fulfilled"))
    keys = eng.best_keys(fs, is_, 1, B)
    torch.cuda.synchronize()
    print(f"T={T} B={B} {mode} live={live}: spec launches {tr.handle.speculative_launches()}, published {n}, "
          f"mean iterations {is_[:, backend.I_INNER_TOTAL].float().mean().item():.1f}, loss {out['loss'].mean().item():.3f}, key {int(keys[0]) & 0x7fffffff}")
    return tr, eng, a0, start, goal


small = {"max_inner_iteration": 6, "max_outer_iteration": 2}
run(50, 3, "bls", **small)                       # speculative replicas, T = 50 instance
run(33, 2, "bls", **small)                       # speculative replicas, runtime T
run(50, 600, "bls", **small)                     # 16 single-warp teams per CTA, queue
run(50, 40, "gd", **small)
run(100, 20, "bls", n_obs=40, **small)           # two-warp teams, K and dK in TMEM, named barriers
run(256, 10, "bls", n_obs=130, **small)          # four-warp teams, K in TMEM, dK from L2, pipelined obstacle loop
run(200, 9, "gd", n_obs=33, **small)
run(50, 300, "bls", n_obs=70, live=True, max_inner_iteration=8, max_outer_iteration=2)      # LIVE: helper lanes share the obstacle loop
run(50, 600, "bls", n_obs=300, **small)          # static many-obstacle scene, batch above the speculative limit: the HELP twin instance
run(33, 60, "gd", n_obs=100, **small)
tr, eng, a0, start, goal = run(50, 8, "gd", **small)
res = eng.optimize_host(a0, start, goal)         # staged host path
pin = lambda x: torch.as_tensor(x).clone().pin_memory()
out_a, out_f, out_i = torch.empty(8, 50, 3).pin_memory(), torch.empty(8, 8).pin_memory(), torch.empty(8, 8, dtype=torch.int32).pin_memory()
eng.optimize_pinned(pin(a0), pin(start), pin(goal), out_a, out_f, out_i)   # zero-copy host path
assert np.array_equal(out_a.numpy(), res.alpha)
dev = tr.initTrajectoryDevice(torch.as_tensor(start, device="cuda"), torch.as_tensor(goal, device="cuda"))
torch.cuda.synchronize()
print("sanitize.py: all kernels ran")
