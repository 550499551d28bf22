"""What could a smarter trajectory scheduler gain on the headline batch (c2, B = 4096)?  CPU only.

Iteration counts come from the C mirror oracle; the per-iteration latency of a warp as a function of the warps
resident on its SM comes from profiles/scripts/latency.py (profiles/r01g_latency_vs_occupancy.txt).  Every SM is simulated
with W team slots; FIFO is what the persistent kernel does (global atomic queue).  Run from the repo root:
    python profiles/scripts/sim_schedule.py [scale]        # scale = latency scale relative to the measured curve (default 1)
"""
import sys
from collections import deque

import numpy as np

sys.path.insert(0, ".")
import bench                                                    # noqa: E402
from irm_motion_planning_b200.trajectory import Trajectory       # noqa: E402
from irm_motion_planning_b200.workloads import initial_alpha, make_workload   # noqa: E402
from oracle import mirror as M                                   # noqa: E402

SCALE = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
XS, YS = [0, 4, 8, 12, 16], [3.97, 3.97, 4.97, 5.6, 6.83]        # warps per SM -> us per iteration (round-1f kernel)


def lat(n):
    return SCALE * np.interp(n, XS, YS)


def iteration_counts():
    wl = make_workload("c2")
    traj = Trajectory(wl.args, create_handle=False)
    alpha0, start, goal = initial_alpha(wl, traj, 0)
    m = M.Mirror(bench.hp_view(wl.args, traj.N_timesteps), traj.km, traj.dkm, traj.jac, wl.obstacles, wl.mode)
    _, _, is_ = m.optimize(alpha0, start, goal)
    loss0 = m.eval(alpha0, start, goal, wl.args.lambda_sg_constraint, wl.args.lambda_jl_constraint)["loss"]
    return is_[:, M.I_INNER_TOTAL].astype(float), loss0, np.linalg.norm(goal - start, axis=1)


def simulate(it, order=None, quantum=None, policy="fifo", W=16, n_sm=148, resume_cost=0.6):
    """policy: fifo | first_quantum (yield after `quantum` iterations while fresh work waits) | round_robin."""
    fresh = deque((it if order is None else it[order]).tolist())
    requeued = deque()
    sms = [[] for _ in range(n_sm)]

    def fetch(s):
        if fresh:
            r = fresh.popleft()
            if policy != "fifo" and r > quantum:
                sms[s].append([quantum + 1.0, r - quantum])
            else:
                sms[s].append([r + 1.0, 0.0])                     # + the initial evaluation
        elif requeued:
            r = requeued.popleft()
            if policy == "round_robin" and r > quantum:
                sms[s].append([quantum + resume_cost, r - quantum])
            else:
                sms[s].append([r + resume_cost, 0.0])

    for s in range(n_sm):
        for _ in range(W):
            fetch(s)
    t = 0.0
    while True:
        dts = [min(x[0] for x in sm) * lat(len(sm)) for sm in sms if sm]
        if not dts:
            return t / 1e3
        dt = min(dts)
        t += dt
        for s in range(n_sm):
            if not sms[s]:
                continue
            l, keep, done = lat(len(sms[s])), [], 0
            for x in sms[s]:
                x[0] -= dt / l
                if x[0] > 1e-9:
                    keep.append(x)
                else:
                    done += 1
                    if x[1] > 0:
                        requeued.append(x[1])
            sms[s] = keep
            for _ in range(done):
                fetch(s)


if __name__ == "__main__":
    it, loss0, dist = iteration_counts()
    from scipy.stats import spearmanr
    print(f"mean iterations {it.mean():.1f}, at the cap of 200: {(it >= 200).mean():.1%}; rank correlation of the iteration count with "
          f"the initial loss {spearmanr(loss0, it)[0]:.2f}, with |goal - start| {spearmanr(dist, it)[0]:.2f}")
    print(f"FIFO (the kernel)          {simulate(it):.3f} ms")
    for W in (12, 8):
        print(f"FIFO, {W} slots per SM      {simulate(it, W=W):.3f} ms")
    print(f"longest first (oracle)     {simulate(it, order=np.argsort(-it)):.3f} ms")
    for q in (10, 20, 40):
        print(f"first quantum {q:3d}          {simulate(it, quantum=q, policy='first_quantum'):.3f} ms    "
              f"round robin {q:3d}: {simulate(it, quantum=q, policy='round_robin'):.3f} ms")
