"""Regenerates the fixtures in this directory.  Run from the repo root in the
build container (needs /root/reference; the GPU box does not have it):

    python tests/golden/make_golden.py

* reference_results.npz -- the two result files the reference ships
  (visualization/trajectory_result.txt: 50x3 final joint angles;
   visualization/trajectory_series.txt: 146 iterates x 150), stored as float32.
  They are the reference's own outputs and the only result-pinning artefacts
  it has.
* oracle_vectors.npz -- seeded inputs and the NumPy oracle's outputs for them
  (per-evaluation loss/gradient/q/v in FP32 and FP64), used to pin the C mirror
  oracle and, through it, the CUDA kernels without needing jax.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/visualization"


def main():
    from oracle import fgd_numpy as O

    res = np.loadtxt(os.path.join(REF, "trajectory_result.txt")).astype(np.float32)
    ser = np.loadtxt(os.path.join(REF, "trajectory_series.txt")).astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "reference_results.npz"), trajectory_result=res, trajectory_series=ser)

    rng = np.random.default_rng(1234)
    hp = O.Hyper()
    tm32, tm64 = O.TrajectoryModel(hp), O.TrajectoryModel(hp, dtype=np.float64)
    obs = O.DEFAULT_OBSTACLES
    n = 6
    start = rng.uniform(-0.9, 1.9, (n, 3)).astype(np.float32)
    goal = rng.uniform(-0.9, 1.9, (n, 3)).astype(np.float32)
    # well-conditioned alphas (smooth, small) and the reference's ill-conditioned LU fits
    smooth = (rng.standard_normal((n, 50, 3)) * 0.05).astype(np.float32)
    fitted = np.stack([tm32.init_trajectory(s, g) for s, g in zip(start, goal)])
    out = dict(start=start, goal=goal, alpha_smooth=smooth, alpha_fitted=fitted, lam=np.array([[0.5, 0.1], [500.0, 100.0]], np.float32))
    for tag, alphas in (("smooth", smooth), ("fitted", fitted)):
        for li, (lsg, ljl) in enumerate(out["lam"]):
            for name, tm in (("f32", tm32), ("f64", tm64)):
                dt = tm.dt
                loss, grad, q, v, ful = [], [], [], [], []
                for a, s, g in zip(alphas, start, goal):
                    a_, s_, g_ = a.astype(dt), s.astype(dt), g.astype(dt)
                    loss.append(tm.cost(a_, obs, s_, g_, lsg, ljl, 0.5))
                    grad.append(tm.cost_g(a_, obs, s_, g_, lsg, ljl, 0.5))
                    q.append(tm.evaluate(a_, tm.km)); v.append(tm.evaluate(a_, tm.dkm))
                    ful.append(tm.constraints_fulfilled(a_, s_, g_))
                k = f"{tag}_l{li}_{name}"
                out[k + "_loss"] = np.array(loss); out[k + "_grad"] = np.array(grad)
                out[k + "_q"] = np.array(q); out[k + "_v"] = np.array(v); out[k + "_ful"] = np.array(ful)
    # end-to-end: the default problem, FP32 oracle
    a0 = tm32.init_trajectory(O.DEFAULT_START, O.DEFAULT_GOAL)
    a, log = O.bls_optimize(tm32, a0, obs, O.DEFAULT_START, O.DEFAULT_GOAL)
    out["c1_alpha0"] = a0; out["c1_bls_alpha"] = a; out["c1_bls_q"] = tm32.evaluate(a, tm32.km)
    out["c1_bls_report"] = np.array(O.final_report(tm32, a, obs, O.DEFAULT_START, O.DEFAULT_GOAL)[:2])
    out["c1_bls_counts"] = np.array([log.outer_iters, log.inner_iters, log.accepts, log.cost_evals])
    np.savez_compressed(os.path.join(HERE, "oracle_vectors.npz"), **out)
    print("written", os.listdir(HERE))


if __name__ == "__main__":
    main()
