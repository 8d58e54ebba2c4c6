import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gym_softrobot_b200 import _native as nat
from gym_softrobot_b200.envs.soft_pendulum import _make_handle, pendulum_init_params
n_env, K = int(sys.argv[1]), int(sys.argv[2])
u = np.random.default_rng(0).random(n_env)
h64 = _make_handle(n_env, 50, 1e-4, 0, nat.MATH_FAST, nat.DTYPE_F64)
h32 = _make_handle(n_env, 50, 1e-4, 0, nat.MATH_FAST, nat.DTYPE_F32)
h64.reset_host(pendulum_init_params(u)); h32.reset_host(pendulum_init_params(u))
rng = np.random.default_rng(1)
for s in range(3):
    a = rng.uniform(-22, 22, size=(n_env, 1)).astype(np.float32)
    l0 = h32.launch_count
    h64.step_host(a, K); o, r, t = h32.step_host(a, K)
    f64 = {k: v.double().cpu().numpy() for k, v in h64.fields().items()}
    f32 = {k: v.double().cpu().numpy() for k, v in h32.fields().items()}
    print(s, "launches", h32.launch_count - l0, "term", int(t.sum()), {k.split("_")[0]: f"{np.abs(f32[k] - f64[k]).max() / max(np.abs(f64[k]).max(), 1e-30):.1e}" for k in f64}, flush=True)
