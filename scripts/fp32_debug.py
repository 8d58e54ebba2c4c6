import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gym_softrobot_b200 import _native as nat
from gym_softrobot_b200.envs.soft_pendulum import _make_handle, pendulum_init_params
n_env = int(sys.argv[1]) if len(sys.argv) > 1 else 4
u = np.linspace(0.1, 0.9, n_env)
h64 = _make_handle(n_env, 50, 1e-4, 0, nat.MATH_FAST, nat.DTYPE_F64)
h32 = _make_handle(n_env, 50, 1e-4, 0, nat.MATH_FAST, nat.DTYPE_F32)
h64.reset_host(pendulum_init_params(u)); h32.reset_host(pendulum_init_params(u))
a = np.full((n_env, 1), 7.5, np.float32)
done = 0
for K in (1, 1, 8, 40, 150, 200):
    h64.step_host(a, K); h32.step_host(a, K); done += K
    f64 = {k: v.double().cpu().numpy() for k, v in h64.fields().items()}
    f32 = {k: v.double().cpu().numpy() for k, v in h32.fields().items()}
    print(done, {k.split("_")[0]: f"{np.abs(f32[k] - f64[k]).max() / max(np.abs(f64[k]).max(), 1e-30):.1e}" for k in f64})
    if done == 2:
        i = 0
        print(" v64 node0..3", f64["velocity_collection"][i][:, :4].T.tolist())
        print(" v32 node0..3", f32["velocity_collection"][i][:, :4].T.tolist())
        print(" w64 el0..2", f64["omega_collection"][i][:, :3].T.tolist())
        print(" w32 el0..2", f32["omega_collection"][i][:, :3].T.tolist())
