"""Where does a lone trajectory spend its cycles?  Needs a -DFGD_PHASE_CLOCKS build of the library:
    nvcc <flags of irm_motion_planning_b200/build.py> -DFGD_PHASE_CLOCKS -o profiles/scripts/libfgd_clk.so irm_motion_planning_b200/csrc/fgd_api.cu
CTA 0 / warp 0 accumulates clock64() per phase of the loop; run on the GPU box from the repo root.
usage: python profiles/scripts/clk.py [spec_max_batch ...]     (one run per value: 0 = sequential line search, 592 = speculative)
       CLK_RUNS="c3:2368,c5:4736" python profiles/scripts/clk.py 0   (workload:batch pairs; the clocks are those of CTA 0 / team 0 under that load)"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, ".")
os.environ["FGD_LIBRARY"] = os.path.abspath("profiles/scripts/libfgd_clk.so")
from irm_motion_planning_b200 import backend                                  # noqa: E402
from irm_motion_planning_b200.batch import BatchedFGD                          # noqa: E402
from irm_motion_planning_b200.trajectory import Trajectory                     # noqa: E402
from irm_motion_planning_b200.workloads import initial_alpha, make_workload    # noqa: E402

names = ["fwd contract", "cost_phase", "decide+grad", "tail(fwd)", "back contract", "back post", "tail(back)", "trips"]
for spec in [int(x) for x in sys.argv[1:]] or [0]:
    os.environ["FGD_SPEC_MAX_BATCH"] = str(spec)
    runs = [(w.split(":")[0], int(w.split(":")[1])) for w in os.environ.get("CLK_RUNS", "c2:1,c1:1").split(",")]
    for wlname, B in runs:
        wl = make_workload(wlname, B=B)
        traj = Trajectory(wl.args)
        traj.set_obstacles(wl.obstacles)
        alpha0, start, goal = initial_alpha(wl, traj, 0)
        eng = BatchedFGD(traj, wl.mode)
        lib = backend.load_library()
        lib.fgd_debug_buffer.restype = C.POINTER(C.c_int)
        lib.fgd_debug_buffer.argtypes = [C.c_void_p]
        a = torch.as_tensor(alpha0[:B], device="cuda"); s = torch.as_tensor(start[:B], device="cuda"); g = torch.as_tensor(goal[:B], device="cuda")
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); res = eng.optimize_device(a.clone(), s, g); e1.record()
            torch.cuda.synchronize()
        buf = lib.fgd_debug_buffer(traj.handle._h)
        v = [buf[i] * 16 for i in range(8)]
        trips = v[7] // 16
        it = int(res.istate[0, 3]); ce = int(res.istate[0, 4])
        print(f"{wlname} spec_max_batch={spec} speculative launches {traj.handle.speculative_launches()}: iters {it} cand evals {ce} trips {trips} "
              f"kernel+launch {e0.elapsed_time(e1):.4f} ms, sum of phases {sum(v[:7])} clk")
        for n, x in zip(names[:7], v[:7]):
            print(f"   {n:14s} {x:10d} clk total   {x / max(it, 1):9.1f} per iteration")
