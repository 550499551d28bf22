"""What limits the strong scaling of the restart sweep (config 5)?  CPU only (the C mirror oracle gives the per-trajectory
trip counts: inner iterations + candidate evaluations + outer bodies).  Run from the repo root:
    python profiles/scripts/c5_tail_sim.py [problems]      (default 128 problems x 256 restarts)
Prints the distribution of the trip count, how much of its variance sits between problems (whole-problem sharding makes
the ranks' work differ by that noise; restart-axis sharding does not), and a fluid simulation of the persistent kernel's
queue (2368 resident teams, per-trip latency depending on occupancy) for 1 / 2 / 4 / 8 ranks: FIFO (what the kernel does),
longest-first (oracle knowledge) and a two-phase order by the trips of the first k outer iterations."""
import heapq
import sys

import numpy as np
from scipy.stats import spearmanr

sys.path.insert(0, ".")
from irm_motion_planning_b200.trajectory import Trajectory                    # noqa: E402
from irm_motion_planning_b200.workloads import initial_alpha, make_workload   # noqa: E402
from oracle import mirror as M                                                # noqa: E402

P0 = int(sys.argv[1]) if len(sys.argv) > 1 else 128
wl = make_workload("c5", B=P0 * 256, seed=0)
tr = Trajectory(wl.args, create_handle=False)
a0, s, g = initial_alpha(wl, tr, 0)


def trips_with(max_outer):
    hp = type("HP", (), dict(vars(wl.args)))()
    hp.n_timesteps = 50
    hp.max_outer_iteration = max_outer
    _, _, is_ = M.Mirror(hp, tr.km, tr.dkm, tr.jac, wl.obstacles, "bls").optimize(a0, s, g)
    return (is_[:, M.I_INNER_TOTAL] + is_[:, M.I_CAND_EVALS] + is_[:, M.I_OUTER] + 1).astype(float), is_


trips, is_ = trips_with(10)
print(f"{len(trips)} trajectories: inner iterations mean {is_[:, M.I_INNER_TOTAL].mean():.1f} max {is_[:, M.I_INNER_TOTAL].max()}, "
      f"candidates per iteration {is_[:, M.I_CAND_EVALS].sum() / is_[:, M.I_INNER_TOTAL].sum():.2f}, fulfilled {is_[:, M.I_FULFILLED].mean():.3f}")
print("trips quantiles 50/90/98/99/99.9/100 %:", np.quantile(trips, [0.5, 0.9, 0.98, 0.99, 0.999, 1.0]))
print("escalations histogram:", np.bincount(is_[:, M.I_OUTER]))
tp = trips.reshape(P0, 256)
print(f"variance of the problem means / total variance = {tp.mean(1).var() / trips.var():.3f}")
ps = tp.sum(1)
print(f"per-problem work: cv {ps.std() / ps.mean():.3f} -> per-rank sum at 512 whole problems per rank: cv {ps.std() / ps.mean() / np.sqrt(512):.4f}; "
      f"at 131072 single trajectories per rank: cv {trips.std() / trips.mean() / np.sqrt(131072):.4f}")
first = {k: trips_with(k)[0] for k in (1, 2)}
for k, t in first.items():
    print(f"trips of the first {k} outer iteration(s): {t.sum() / trips.sum():.3f} of the work, rank correlation with the total {spearmanr(trips, t)[0]:.2f}")

XS, YS = [0, 4, 8, 12, 16], np.array([3.97, 3.97, 4.97, 5.6, 6.83]) / 2 * 0.87      # us per trip vs warps per SM (profiles/r01g_latency_vs_occupancy.txt, scaled)


def lat(n):
    return np.interp(n / 148.0, XS, YS)


def sim(L, M_=2368):
    L = list(L); n = len(L); i = 0; heap = []; V = 0.0; t = 0.0
    while i < n and len(heap) < M_:
        heapq.heappush(heap, V + L[i]); i += 1
    while heap:
        f = heapq.heappop(heap)
        t += (f - V) * lat(len(heap) + 1); V = f
        if i < n:
            heapq.heappush(heap, V + L[i]); i += 1
    return t / 1e3


rng = np.random.default_rng(1)
for N in (1, 2, 4, 8):
    Pn = 4096 // N
    probs = rng.integers(0, P0, Pn)
    idx = (probs[:, None] * 256 + np.arange(256)[None, :]).reshape(-1)
    L = trips[idx]
    ideal = L.sum() * lat(2368) / 2368 / 1e3
    fifo, lpt = sim(L), sim(np.sort(L)[::-1])
    a = first[2][idx]
    rem = L - a
    keep = rem > 0
    two_phase = sim(a) + sim(rem[keep][np.argsort(-a[keep], kind="stable")])
    print(f"N={N}: {len(L)} trajectories per rank, ideal {ideal:.2f} ms, FIFO {fifo:.2f} ms (efficiency {ideal / fifo:.3f}), "
          f"longest first {lpt:.2f} ({ideal / lpt:.3f}), two-phase ordered by the first 2 outer iterations {two_phase:.2f} ({ideal / two_phase:.3f})")
