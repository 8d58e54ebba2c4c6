"""A/B of the fast-only / fallback kernel pair against the single safe kernel, per instantiation (diagnostic).

Run it twice on the same box:   python scripts/bench_ab.py ;  SOFTROD_FASTPATH=0 python scripts/bench_ab.py
(the switch is read once per process).  These are the figures quoted in DESIGN.md 4.1 / 4.2."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import gym_softrobot_b200 as g
from gym_softrobot_b200 import _native as nat
from gym_softrobot_b200.envs.arm_single import arm_contact_params, _ROD
from gym_softrobot_b200.envs.octo_flat import OctoFlatVectorEnv

MODE = "pair" if os.environ.get("SOFTROD_FASTPATH", "1") != "0" else "safe kernel only"


def timed(fn, K, W):
    for _ in range(W): fn()            # W >= 10 lets the adaptive switch settle
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(K): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / K


def out(n_env):
    return (torch.empty((n_env, 6), dtype=torch.float32, device="cuda"), torch.empty(n_env, dtype=torch.float64, device="cuda"),
            torch.empty(n_env, dtype=torch.uint8, device="cuda"))


for dtype in ("float64", "float32"):
    for n_env in (4096, 65536):
        env = g.make_vec("SoftPendulum-v0", n_env, autoreset=False, dtype=dtype); env.reset(seed=42)
        a = (torch.rand((n_env, 1), device="cuda") * 44 - 22).float()
        ms = timed(lambda: env.handle.step(a, 400, env.obs, env.reward, env.terminated), 6 if n_env == 4096 else 3, 10 if n_env == 4096 else 3)
        print(f"[{MODE}] SoftPendulum-v0 {dtype} {n_env} envs: {ms:.3f} ms per env-step ({n_env / ms * 1e3:.4e} env-steps/s)")
        env.close()

n_env = 4096
env = g.make_vec("SoftPendulum3D-v0", n_env, autoreset=False); env.reset(seed=42)
a = (torch.rand((n_env, 2), device="cuda") * 2 - 1).float()
print(f"[{MODE}] SoftPendulum3D-v0 {n_env} envs: {timed(lambda: env.handle.step(a, 400, env.obs, env.reward, env.terminated), 6, 10):.3f} ms per env-step")
env.close()

h = nat.Handle(model=nat.MODEL_ROD, n_env=n_env, n_elem=50, dt=7e-5, gravity=(0, 0, -9.81), damping_constant=1e-2, bc_kind=nat.BC_FREE,
               contact=arm_contact_params(), **_ROD)
init = np.zeros((n_env, 9)); init[:, 3] = 1.0; init[:, 8] = 1.0
h.reset_host(init)
h.rest_kappa_tensor()[:, 0, :] = torch.as_tensor(np.random.default_rng(1).uniform(-5, 5, (n_env, 1)) * np.ones((1, 49)), device="cuda")
o = out(n_env)
print(f"[{MODE}] rod on frictional plane n=50, {n_env} envs: {timed(lambda: h.step(None, 400, *o), 6, 10):.3f} ms per 400 substeps")
h.close()

env = OctoFlatVectorEnv(16384, n_elems=10, time_step=7e-5, autoreset=False); env.reset(seed=42)
env.handle.rest_kappa_tensor()[:, 0, :] = torch.as_tensor(np.random.default_rng(2).uniform(-5, 5, (16384 * 8, 1)) * np.ones((1, 9)), device="cuda")
o6, rew, term = env._scratch
print(f"[{MODE}] 8-arm assembly n_elem=10, 16384 envs: {timed(lambda: env.handle.step(None, 400, o6, rew, term), 3, 10):.2f} ms per 400 substeps")
env.close()

env = g.make_vec("ContinuumSnake-v0", n_env, autoreset=False); env.reset()
mu = env.handle.muscle_tensor()
mu[:, 2:] = torch.as_tensor(np.random.default_rng(3).uniform(-4e-3, 4e-3, (n_env, 6)), device="cuda") @ env._W.T
mu[:, 1] = 2 * np.pi / 0.97
o6, rew, term = env._scratch
print(f"[{MODE}] ContinuumSnake-v0 callback segment (2083 substeps), {n_env} envs: {timed(lambda: env.handle.step(None, 2083, o6, rew, term), 3, 10):.2f} ms")
env.close()
