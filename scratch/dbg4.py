import sys, os, time, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
T, mode = int(sys.argv[1]), sys.argv[2]
import torch
from tests.test_gpu_parity import _setup
args, tr, obs, start, goal, alpha0 = _setup(T=T, n_obs=11, B=1, seed=T, strict=(mode != "fast"))
print("alpha0 finite", np.isfinite(alpha0).all(), np.abs(alpha0).max(), flush=True)
if mode == "zeros":
    alpha0 = np.zeros_like(alpha0)
want = ("loss",) if mode != "q" else ("q",)
out = tr._eval(alpha0, None, start, goal, 0.5, 0.1, -1.0, want); torch.cuda.synchronize()
print("ok", T, mode, {k: float(v.flatten()[0]) for k, v in out.items()}, flush=True)
