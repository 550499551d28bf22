"""Per-iteration latency of a trajectory as a function of the co-resident warps per SM (c2 shape):
B copies of one 200-iteration trajectory, B = n x 148, so the launch time is 200 x latency(n)."""
import os, sys, numpy as np, torch
sys.path.insert(0, ".")
from irm_motion_planning_b200 import backend
from irm_motion_planning_b200.batch import BatchedFGD
from irm_motion_planning_b200.trajectory import Trajectory
from irm_motion_planning_b200.workloads import initial_alpha, make_workload
wl = make_workload("c2", B=4096)
traj = Trajectory(wl.args); traj.set_obstacles(wl.obstacles)
alpha0, start, goal = initial_alpha(wl, traj, 0)
eng = BatchedFGD(traj, wl.mode)
a = torch.as_tensor(alpha0, device="cuda"); s = torch.as_tensor(start, device="cuda"); g = torch.as_tensor(goal, device="cuda")
res = eng.optimize_device(a.clone(), s, g); torch.cuda.synchronize()
it = res.istate[:, backend.I_INNER_TOTAL].cpu().numpy()
j = int(np.argmax(it)); print("variant", os.environ.get("FGD_VARIANT", "0"), "trajectory", j, "iters", it[j], flush=True)
for n in (1, 148, 296, 592, 888, 1184, 1480, 1776, 2072, 2368, 4736):
    aa = a[j:j + 1].repeat(n, 1, 1).contiguous(); ss = s[j:j + 1].repeat(n, 1).contiguous(); gg = g[j:j + 1].repeat(n, 1).contiguous()
    ts = []
    for rep in range(4):
        x = aa.clone(); fs, is_ = eng.new_state(n)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.optimize_device(x, ss, gg, fs, is_); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = min(ts[1:])
    print(f"B={n:5d} ({n / 148:5.2f}/SM) geometry {traj.handle.launch_geometry(n)} ms {ms:.3f}  us/iter {1e3 * ms / it[j]:.2f}  SM-us per traj-iter {1e3 * ms / it[j] / max(n / 148, 1e-9):.3f}", flush=True)
