import sys, os, time, ctypes, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
T = int(sys.argv[1])
import torch
from tests.test_gpu_parity import _setup
from irm_motion_planning_b200 import backend
args, tr, obs, start, goal, alpha0 = _setup(T=T, n_obs=11, B=1, seed=T)
lib = backend.load_library()
lib.fgd_debug_buffer.restype = ctypes.POINTER(ctypes.c_int); lib.fgd_debug_buffer.argtypes = [ctypes.c_void_p]
buf = lib.fgd_debug_buffer(tr.handle._h)
dev = torch.device("cuda")
a = torch.as_tensor(alpha0, device=dev).contiguous(); s = torch.as_tensor(start, device=dev).contiguous(); g = torch.as_tensor(goal, device=dev).contiguous()
loss = torch.empty(1, device=dev)
torch.cuda.synchronize()
tr.handle.eval(1, a, s, g, 0.5, 0.1, -1.0, loss=loss)
time.sleep(3)
arr = np.ctypeslib.as_array(buf, shape=(4096,))[:128].copy()
print("T", T, "markers per lane (128 threads):")
print(arr.reshape(4, 32), flush=True)
os._exit(0)
