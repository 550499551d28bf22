"""Latency of small BLS batches: speculative line search (four candidates in parallel per trajectory) against the
sequential kernel, same inputs, fast math.  Run on the GPU box from the repo root."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from irm_motion_planning_b200.batch import BatchedFGD                      # noqa: E402
from irm_motion_planning_b200.environment import Environment               # noqa: E402
from irm_motion_planning_b200.trajectory import Trajectory                 # noqa: E402
from irm_motion_planning_b200.workloads import default_args, sample_start_goal   # noqa: E402


def handle(spec):
    os.environ["FGD_SPEC_MAX_BATCH"] = str(spec)
    tr = Trajectory(default_args())
    tr.set_obstacles(Environment().obstacles)
    return tr


def timed(tr, a0, s, g, reps=30):
    eng = BatchedFGD(tr, "bls")
    B = a0.shape[0]
    bufs = [a0.clone() for _ in range(reps + 3)]
    ms = []
    for i, buf in enumerate(bufs):
        fs, is_ = eng.new_state(B)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.optimize_device(buf, s, g, fs, is_); e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            ms.append(e0.elapsed_time(e1))
    return float(np.median(ms)), is_.cpu().numpy(), bufs[-1]


env = Environment()
tr_spec, tr_seq = handle(100000), handle(0)
print("B      sequential ms   speculative ms   ratio   identical   mean inner iterations / candidate evaluations")
for B in (1, 4, 12, 37, 74, 148, 222, 296, 444, 592, 1184, 2368):
    if B == 1:
        start, goal = env.start_config[None].copy(), env.goal_config[None].copy()
    else:
        start, goal = sample_start_goal(B, np.random.default_rng(B))
    a0 = torch.as_tensor(tr_spec.initTrajectory(start, goal).reshape(B, 50, 3), device="cuda")
    s, g = torch.as_tensor(start, device="cuda").contiguous(), torch.as_tensor(goal, device="cuda").contiguous()
    t_seq, is_seq, a_seq = timed(tr_seq, a0, s, g)
    t_spec, is_spec, a_spec = timed(tr_spec, a0, s, g)
    same = bool(np.array_equal(is_seq, is_spec) and torch.equal(a_seq, a_spec))
    print(f"{B:5d}  {t_seq:12.4f}  {t_spec:14.4f}  {t_seq / t_spec:6.3f}   {same}   {is_seq[:, 3].mean():.1f} / {is_seq[:, 4].mean():.1f}")
assert tr_spec.handle.speculative_launches() > 0 and tr_seq.handle.speculative_launches() == 0
