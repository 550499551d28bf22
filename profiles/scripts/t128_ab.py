"""T = 128 (two-warp teams): K and dK in tensor memory (default) against the round-1 layout (one team per CTA, tables in L2:
FGD_VARIANT=1).  BLS, 64 random obstacles, 16 384 trajectories, fast math.  Run on the GPU box from the repo root."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from irm_motion_planning_b200.batch import BatchedFGD                           # noqa: E402
from irm_motion_planning_b200.environment import random_obstacles               # noqa: E402
from irm_motion_planning_b200.trajectory import Trajectory                      # noqa: E402
from irm_motion_planning_b200.workloads import default_args, flops_total, sample_start_goal   # noqa: E402

T, B, O = 128, 16384, 64
rng = np.random.default_rng(0)
obs = random_obstacles(O, rng)
start, goal = sample_start_goal(B, rng)
res = {}
for variant in ("0", "1"):
    os.environ["FGD_VARIANT"] = variant
    tr = Trajectory(default_args(n_timesteps=float(T)))
    tr.set_obstacles(obs)
    if variant == "0":
        a0 = torch.as_tensor(tr.initTrajectory(start, goal), device="cuda")
        s, g = torch.as_tensor(start, device="cuda").contiguous(), torch.as_tensor(goal, device="cuda").contiguous()
    eng = BatchedFGD(tr, "bls")
    ms = []
    for i in range(4):
        a = a0.clone(); fs, is_ = eng.new_state(B)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.optimize_device(a, s, g, fs, is_); e1.record(); torch.cuda.synchronize()
        if i:
            ms.append(e0.elapsed_time(e1))
    isn = is_.cpu().numpy()
    fl = flops_total("bls", T, O, isn[:, 3], isn[:, 4], np.maximum(1, isn[:, 1] + isn[:, 6]))
    peak = tr.handle.measure_fp32_peak()
    t = np.mean(ms) * 1e-3
    res[variant] = (a.cpu().numpy(), isn)
    print(f"FGD_VARIANT={variant} {tr.handle.launch_geometry(B)}: {1e3 * t:.2f} ms, {B / t:.0f} trajectories/s, {fl / t * 1e-12 / peak:.3f} of the FP32 peak (algorithmic)")
print("identical results:", bool(np.array_equal(res["0"][0], res["1"][0]) and np.array_equal(res["0"][1], res["1"][1])))
