import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gym_softrobot_b200 import _native as nat
from gym_softrobot_b200.envs.soft_pendulum import _make_handle, pendulum_init_params
h = _make_handle(8, 50, 1e-4, 0, nat.MATH_FAST, nat.DTYPE_F32)
h.reset_host(pendulum_init_params(np.linspace(0.1, 0.9, 8)))
obs, rew, term = h.step_host(np.zeros((8, 1), np.float32), 3)
torch.cuda.synchronize()
print(obs[:2], term)
