"""Latency of small batches in a cluttered scene: T = 50, 256 static obstacles, BLS, B = 1 ... 296 (the speculative kernel)."""
import json, sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from irm_motion_planning_b200.batch import BatchedFGD
from irm_motion_planning_b200.trajectory import Trajectory
from irm_motion_planning_b200.workloads import initial_alpha, make_workload
out = {"lib": os.environ.get("FGD_LIBRARY", "tree")}
for B in (1, 16, 148, 296):
    wl = make_workload("c4", B=B, seed=0)
    tr = Trajectory(wl.args)
    tr.set_obstacles(wl.obstacles)
    alpha0, start, goal = initial_alpha(wl, tr, 0)
    eng = BatchedFGD(tr, wl.mode)
    a0 = torch.as_tensor(alpha0, device="cuda")
    s, g = torch.as_tensor(start, device="cuda").contiguous(), torch.as_tensor(goal, device="cuda").contiguous()
    ms = []
    for i in range(8):
        a = a0.clone(); fs, is_ = eng.new_state(B)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.optimize_device(a, s, g, fs, is_); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    out[f"B{B}"] = {"ms": round(float(np.median(ms[2:])), 4), "spec_launches": tr.handle.speculative_launches(),
                    "mean_inner": float(is_[:, 3].float().mean().item()), "hash0": int(is_[0, 7].item())}
print(json.dumps(out))
