import sys, numpy as np
sys.path.insert(0, ".")
import tests.test_gpu_parity as P
over = {"whole_arm_cost": True, "max_inner_iteration": 25, "max_outer_iteration": 2}
for T, n_obs in ((50, 11), (50, 4), (50, 1), (50, 0), (100, 30)):
    args, tr, obs, start, goal, alpha0 = P._setup(T=T, n_obs=max(n_obs, 1), B=20, seed=T + 3, **over)
    if n_obs == 0:
        obs = np.zeros((0, 2), np.float32); tr.set_obstacles(obs)
    elif n_obs != 11:
        obs = obs[:n_obs]; tr.set_obstacles(obs)
    m = P._mirror(args, tr, obs, "bls")
    g = P._gpu_eval(tr, alpha0, start, goal, 0.5, 0.1)
    c = m.eval(alpha0, start, goal, 0.5, 0.1)
    print(T, n_obs, {k: (int((g[k] != c[k]).sum()), float(np.abs(g[k] - c[k]).max())) for k in ("q", "v", "loss", "toc", "grad")})
