import os, sys, ctypes as C, numpy as np, torch
sys.path.insert(0, ".")
os.environ["FGD_LIBRARY"] = os.path.abspath("profiles/scripts/libfgd_clk.so")
from irm_motion_planning_b200 import backend
from irm_motion_planning_b200.batch import BatchedFGD
from irm_motion_planning_b200.trajectory import Trajectory
from irm_motion_planning_b200.workloads import initial_alpha, make_workload
names = ["fwd contract", "cost_phase", "decide+grad", "tail(fwd)", "back contract", "back post", "tail(back)", "trips"]
for wlname, B in (("c2", 1), ("c1", 1), ("c3", 1)):
    wl = make_workload(wlname, B=B)
    traj = Trajectory(wl.args); traj.set_obstacles(wl.obstacles)
    alpha0, start, goal = initial_alpha(wl, traj, 0)
    eng = BatchedFGD(traj, wl.mode)
    lib = backend.load_library(); lib.fgd_debug_buffer.restype = C.POINTER(C.c_int); lib.fgd_debug_buffer.argtypes = [C.c_void_p]
    a = torch.as_tensor(alpha0[:B], device="cuda"); s = torch.as_tensor(start[:B], device="cuda"); g = torch.as_tensor(goal[:B], device="cuda")
    for rep in range(2):
        res = eng.optimize_device(a.clone(), s, g)
        torch.cuda.synchronize()
    buf = lib.fgd_debug_buffer(traj.handle._h)
    v = [buf[i] * 16 for i in range(8)]
    trips = v[7] // 16
    it = int(res.istate[0, 3]); ce = int(res.istate[0, 4])
    print(wlname, "iters", it, "cand evals", ce, "trips", trips)
    for n, x in zip(names[:7], v[:7]):
        print(f"   {n:14s} {x:10d} clk total   {x / max(it,1):9.1f} per iteration")
