import sys, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
what, T, nobs, B, budget = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
import torch
from tests.test_gpu_parity import _setup, _gpu_eval, _gpu_optimize
args, tr, obs, start, goal, alpha0 = _setup(T=T, n_obs=nobs, B=B, seed=T)
if what == "eval":
    g = _gpu_eval(tr, alpha0, start, goal, 0.5, 0.1)
    print("eval ok", T, nobs, float(g["loss"][0]))
elif what == "evalnograd":
    out = tr._eval(alpha0, None, start, goal, 0.5, 0.1, -1.0, ("loss",)); torch.cuda.synchronize()
    print("evalnograd ok", T, nobs, float(out["loss"][0]))
else:
    a, fs, is_ = _gpu_optimize(tr, what, alpha0, start, goal, budget=budget)
    print(what, "ok", T, nobs, B, budget, is_.cpu().numpy()[:2])
