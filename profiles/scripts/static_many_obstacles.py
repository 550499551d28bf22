"""T = 50, 256 static obstacles, BLS, one launch per batch: the config-4 scene without obstacle updates (non-LIVE kernels)."""
import json, sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from irm_motion_planning_b200.batch import BatchedFGD
from irm_motion_planning_b200.trajectory import Trajectory
from irm_motion_planning_b200.workloads import initial_alpha, make_workload
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
wl = make_workload("c4", B=B, seed=0)
tr = Trajectory(wl.args)
tr.set_obstacles(wl.obstacles)
alpha0, start, goal = initial_alpha(wl, tr, 0)
eng = BatchedFGD(tr, wl.mode)
a0 = torch.as_tensor(alpha0, device="cuda")
s, g = torch.as_tensor(start, device="cuda").contiguous(), torch.as_tensor(goal, device="cuda").contiguous()
ms = []
for i in range(4):
    a = a0.clone(); fs, is_ = eng.new_state(B)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.optimize_device(a, s, g, fs, is_); e1.record(); torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
print(json.dumps({"lib": os.environ.get("FGD_LIBRARY", "tree"), "B": B, "n_obs": int(len(wl.obstacles)), "ms": [round(m, 3) for m in ms[1:]],
                  "mean_inner": float(is_[:, 3].float().mean().item())}))
