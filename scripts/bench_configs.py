"""Throughput of the BASELINE.json configs other than the headline one (diagnostic table for BASELINE.md)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import gym_softrobot_b200 as g
from gym_softrobot_b200 import _native as nat

def timed(step_fn, K=10, W=3):
    for _ in range(W): step_fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    for a, b in ev:
        a.record(); step_fn(); b.record()
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in ev) / K * 1e-3

peak = nat.measure_fp64_peak(0)
rows = []
def report(name, n_env, n_elem, K, sec, flop):
    es = n_env * n_elem * K / sec
    rows.append(dict(config=name, n_env=n_env, n_elem=n_elem, substeps=K, ms_per_step=sec * 1e3, env_steps_per_s=n_env / sec,
                     elem_substeps_per_s=es, fp64_frac=es * flop / 1e12 / peak, flop_per_elem_substep=flop))
    print(json.dumps(rows[-1]), flush=True)

# config 2 (headline) at larger batches
for n_env in (4096, 16384, 65536):
    env = g.make_vec("SoftPendulum-v0", n_env, autoreset=False); env.reset(seed=42)
    a = (torch.rand((n_env, 1), device="cuda") * 44 - 22).float()
    report("SoftPendulum-v0", n_env, 50, 400, timed(lambda: env.handle.step(a, 400, env.obs, env.reward, env.terminated)), 440)
    env.close()
# config 3: clamped rod n=100, gravity + damping, 8192 envs per GPU
n_env = 8192
h = nat.Handle(model=nat.MODEL_ROD, n_env=n_env, n_elem=100, dt=5e-5, base_length=1.0, base_radius=0.025, density=1000.0,
               youngs_modulus=1e6, gravity=(0, -9.80665, 0), damping_constant=2e-3, bc_kind=nat.BC_ONE_END_FIXED)
ang = np.deg2rad(np.random.default_rng(42).uniform(-5, 5, n_env)); init = np.zeros((n_env, 9))
init[:, 3], init[:, 4], init[:, 6], init[:, 7] = np.cos(ang), np.sin(ang), -np.sin(ang), np.cos(ang)
h.reset_host(init)
obs = torch.empty((n_env, 6), dtype=torch.float32, device="cuda"); rew = torch.empty(n_env, dtype=torch.float64, device="cuda"); term = torch.empty(n_env, dtype=torch.uint8, device="cuda")
report("rod n=100 clamped (config 3, per GPU)", n_env, 100, 400, timed(lambda: h.step(None, 400, obs, rew, term)), 440)
assert int(term.sum()) == 0
h.close()
# f1: SoftPendulum3D
n_env = 4096
env = g.make_vec("SoftPendulum3D-v0", n_env, autoreset=False); env.reset(seed=42)
a = (torch.rand((n_env, 2), device="cuda") * 2 - 1).float()
report("SoftPendulum3D-v0", n_env, 50, 400, timed(lambda: env.handle.step(a, 400, env.obs, env.reward, env.terminated)), 440 + 168)
env.close()
# config 5 topology: free rod on a frictional plane, rest-curvature actuation (OctoArmSingle-v0, n=50; and n=200)
from gym_softrobot_b200.envs.arm_single import arm_contact_params, _ROD
for n_elem, dt in ((50, 7e-5), (200, 2e-5)):
    n_env = 4096
    h = nat.Handle(model=nat.MODEL_ROD, n_env=n_env, n_elem=n_elem, dt=dt, gravity=(0, 0, -9.81), damping_constant=1e-2,
                   bc_kind=nat.BC_FREE, contact=arm_contact_params(), **_ROD)
    init = np.zeros((n_env, 9)); init[:, 3] = 1.0; init[:, 8] = 1.0
    h.reset_host(init)
    h.rest_kappa_tensor()[:, 0, :] = torch.as_tensor(np.random.default_rng(1).uniform(-5, 5, (n_env, 1)) * np.ones((1, n_elem - 1)), device="cuda")
    obs = torch.empty((n_env, 6), dtype=torch.float32, device="cuda"); rew = torch.empty(n_env, dtype=torch.float64, device="cuda"); term = torch.empty(n_env, dtype=torch.uint8, device="cuda")
    report(f"rod on frictional plane n={n_elem} (config 5 topology)", n_env, n_elem, 400, timed(lambda: h.step(None, 400, obs, rew, term), K=5), 440 + 280)
    assert int(term.sum()) == 0
    h.close()
# config 5 as specified: long slender rod n_elem = 512 on the frictional plane (one rod per CTA), FP64 and FP32 modes
n_env, n_elem = 4096, 512
for dtype, tag in ((nat.DTYPE_F64, "FP64"), (nat.DTYPE_F32, "FP32")):
    h = nat.Handle(model=nat.MODEL_ROD, n_env=n_env, n_elem=n_elem, dt=5e-6, gravity=(0, 0, -9.81), damping_constant=1e-2,
                   bc_kind=nat.BC_FREE, contact={**arm_contact_params(), "plane_origin": [0.0, 0.0, -0.005]}, base_length=1.0,
                   base_radius=0.005, density=1000.0, youngs_modulus=1e6, dtype=dtype)
    init = np.zeros((n_env, 9)); init[:, 3] = 1.0; init[:, 8] = 1.0
    h.reset_host(init)
    rk = h.rest_kappa_tensor()
    rk[:, 0, :] = torch.as_tensor(np.random.default_rng(1).uniform(-3, 3, (n_env, 1)) * np.ones((1, n_elem - 1)), device="cuda").to(rk.dtype)
    obs = torch.empty((n_env, 6), dtype=torch.float32, device="cuda"); rew = torch.empty(n_env, dtype=torch.float64, device="cuda"); term = torch.empty(n_env, dtype=torch.uint8, device="cuda")
    report(f"long slender rod n=512 on frictional plane (config 5 as specified), {tag}", n_env, n_elem, 50, timed(lambda: h.step(None, 50, obs, rew, term), K=5), 440 + 280)
    assert int(term.sum()) == 0
    h.close()
# config 4: 8-arm assembly with head + joints + contact (OctoFlat topology), n_elem = 10 (registered env) and 40 (BASELINE)
from gym_softrobot_b200.envs.octo_flat import OctoFlatVectorEnv
# n_elem = 40 needs dt <= 3e-5: the k = 1e6 joint spring on a 0.67 g end node has w dt = 2.7 at 7e-5
for n_elem, n_env, dt in ((10, 16384, 7e-5), (40, 16384, 3e-5)):
    env = OctoFlatVectorEnv(n_env, n_elems=n_elem, time_step=dt, autoreset=False); env.reset(seed=42)
    env.handle.rest_kappa_tensor()[:, 0, :] = torch.as_tensor(np.random.default_rng(2).uniform(-5, 5, (n_env * 8, 1)) * np.ones((1, n_elem - 1)), device="cuda")
    o6, rew, term = env._scratch
    report(f"8-arm assembly n_elem={n_elem} (config 4 topology)", n_env, 8 * n_elem, 400, timed(lambda: env.handle.step(None, 400, o6, rew, term), K=5), 440 + 280)
    assert int(term.sum()) == 0
    env.close()
# f2: ContinuumSnake-v0 (muscle torques + kinetic friction): one callback segment of 2083 substeps per launch
n_env = 4096
env = g.make_vec("ContinuumSnake-v0", n_env, autoreset=False); env.reset()
mu = env.handle.muscle_tensor()
mu[:, 2:] = torch.as_tensor(np.random.default_rng(3).uniform(-4e-3, 4e-3, (n_env, 6)), device="cuda") @ env._W.T
mu[:, 1] = 2 * np.pi / 0.97
o6, rew, term = env._scratch
report("ContinuumSnake-v0 (per 2083-substep callback segment; env-step = 12 of these)", n_env, 50, 2083,
       timed(lambda: env.handle.step(None, 2083, o6, rew, term), K=3, W=1), 440 + 280 + 120)
env.close()
# f4: SoftArmTracking-v0 (spline muscle torques, clamped arm n=40, 50 substeps per env-step)
n_env = 16384
env = g.make_vec("SoftArmTracking-v0", n_env, autoreset=False); env.reset(seed=1)
acts = torch.rand((n_env, 8), device="cuda", dtype=torch.float64) * 2 - 1
report("SoftArmTracking-v0 (whole env.step incl. host layer)", n_env, 40, 50, timed(lambda: env.step(acts), K=10), 440 + 20)
env.close()
# end to end through the public vector-env API: host (pinned) actions -> device, step(), observations + rewards -> host
def e2e(name, env_id, n_env, act_shape, lo, hi, K, dtype=torch.float32, W=18, **kw):   # W: past the adaptive switch's first checks
    env = g.make_vec(env_id, n_env, **kw)
    env.reset(seed=1)
    host_a = ((torch.rand((n_env,) + act_shape, dtype=dtype) * (hi - lo) + lo) if np.isscalar(lo) else
              (torch.rand((n_env,) + act_shape, dtype=dtype) * torch.as_tensor(hi - lo, dtype=dtype) + torch.as_tensor(lo, dtype=dtype))).pin_memory()
    def step():
        obs, rew, term, trunc, info = env.step(host_a.to("cuda", non_blocking=True))
        o = obs["individual"] if isinstance(obs, dict) else obs
        return o.cpu(), rew.cpu()
    sec = timed(step, K=K, W=W)
    rows.append(dict(config=name + " e2e (vector-env step, host in/out)", n_env=n_env, ms_per_step=sec * 1e3, env_steps_per_s=n_env / sec))
    print(json.dumps(rows[-1]), flush=True)
    env.close()

e2e("SoftPendulum-v0", "SoftPendulum-v0", 4096, (1,), -22.0, 22.0, 10)
e2e("SoftPendulum3D-v0", "SoftPendulum3D-v0", 4096, (2,), -1.0, 1.0, 10)
e2e("OctoArmSingle-v0", "OctoArmSingle-v0", 4096, (7,), -22.0, 22.0, 5)
e2e("OctoFlat-v0", "OctoFlat-v0", 4096, (24,), -22.0, 22.0, 3)
snake_lo = np.array([-4e-3] * 6 + [0.9], dtype=np.float32); snake_hi = np.array([4e-3] * 6 + [1.1], dtype=np.float32)
e2e("ContinuumSnake-v0", "ContinuumSnake-v0", 4096, (7,), snake_lo, snake_hi, 2, W=2)
e2e("SoftArmTracking-v0", "SoftArmTracking-v0", 16384, (8,), -0.3, 0.3, 10, dtype=torch.float64)
print("fp64 peak", peak)
