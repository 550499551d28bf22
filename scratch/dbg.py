import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
what, T, nobs, B = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from tests.test_gpu_parity import _setup, _mirror, _gpu_eval, _gpu_optimize
from oracle import mirror as M
args, tr, obs, start, goal, alpha0 = _setup(T=T, n_obs=nobs, B=B, seed=T)
if what == "eval":
    g = _gpu_eval(tr, alpha0, start, goal, 0.5, 0.1)
    c = _mirror(args, tr, obs, "bls").eval(alpha0, start, goal, 0.5, 0.1)
    print("eval", T, nobs, {k: bool(np.array_equal(g[k], c[k])) for k in ("q", "v", "loss", "toc", "grad")})
else:
    a, fs, is_ = _gpu_optimize(tr, what, alpha0, start, goal)
    ca, cfs, cis = _mirror(args, tr, obs, what).optimize(alpha0, start, goal)
    print(what, T, nobs, B, "alpha", bool(np.array_equal(a.cpu().numpy(), ca)), "istate", bool(np.array_equal(is_.cpu().numpy(), cis)))
