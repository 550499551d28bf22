import sys, os, time, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
what, T, nobs, B, budget = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
t0 = time.time()
import torch
from tests.test_gpu_parity import _setup, _gpu_eval, _gpu_optimize
print("imports", time.time() - t0, flush=True)
args, tr, obs, start, goal, alpha0 = _setup(T=T, n_obs=nobs, B=B, seed=T)
print("setup done", time.time() - t0, flush=True)
if what == "eval":
    g = _gpu_eval(tr, alpha0, start, goal, 0.5, 0.1)
    print("eval ok", T, nobs, float(g["loss"][0]), flush=True)
else:
    a, fs, is_ = _gpu_optimize(tr, what, alpha0, start, goal, budget=budget)
    print(what, "ok", T, nobs, B, budget, is_.cpu().numpy()[:2], flush=True)
