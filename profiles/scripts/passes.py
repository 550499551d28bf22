"""Does a budgeted first pass shorten the tail of a small batch?  (c2 shape)"""
import sys, numpy as np, torch
sys.path.insert(0, ".")
from irm_motion_planning_b200.batch import BatchedFGD
from irm_motion_planning_b200.trajectory import Trajectory
from irm_motion_planning_b200.workloads import initial_alpha, make_workload
for B in (4096, 8192, 16384):
    wl = make_workload("c2", B=B)
    traj = Trajectory(wl.args); traj.set_obstacles(wl.obstacles)
    alpha0, start, goal = initial_alpha(wl, traj, 0)
    eng = BatchedFGD(traj, wl.mode)
    a0 = torch.as_tensor(alpha0, device="cuda"); s = torch.as_tensor(start, device="cuda"); g = torch.as_tensor(goal, device="cuda")
    for sched in ([-1], [8, -1], [16, -1], [32, -1], [48, -1], [64, -1], [96, -1], [32, 64, -1], [24, 48, 96, -1]):
        ts = []
        for rep in range(6):
            a = a0.clone(); fs, is_ = eng.new_state(B)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            for b in sched:
                eng.optimize_device(a, s, g, fs, is_, max_launch_iters=b)
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print(B, sched, "ms %.3f" % min(ts[1:]), "iters", int(is_[:, 3].sum()), flush=True)
