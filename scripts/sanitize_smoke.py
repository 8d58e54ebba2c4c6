"""Tiny run of every kernel variant, meant to be executed under compute-sanitizer (racecheck / memcheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gym_softrobot_b200 as g
from gym_softrobot_b200 import _native as nat
from gym_softrobot_b200.envs.arm_single import arm_contact_params, _ROD

K = 6
env = g.make_vec("SoftPendulum-v0", 7, autoreset=False); env.reset(seed=1)
env.handle.step(torch.ones((7, 1), device="cuda"), K, env.obs, env.reward, env.terminated); env.close()
env = g.make_vec("SoftPendulum-v0", 7, autoreset=False, dtype="float32"); env.reset(seed=1)
env.handle.step(torch.ones((7, 1), device="cuda"), K, env.obs, env.reward, env.terminated); env.close()
env = g.make_vec("SoftPendulum-v0", 3, autoreset=False, math=nat.MATH_FAITHFUL); env.reset(seed=1)
env.handle.step(torch.ones((3, 1), device="cuda"), K, env.obs, env.reward, env.terminated); env.close()
env = g.make_vec("SoftPendulum3D-v0", 6, autoreset=False); env.reset(seed=1)
env.handle.step(torch.full((6, 2), 0.5, device="cuda"), K, env.obs, env.reward, env.terminated); env.close()
h = nat.Handle(model=nat.MODEL_ROD, n_env=5, n_elem=100, dt=5e-5, base_length=1.0, base_radius=0.025, density=1000.0,
               youngs_modulus=1e6, gravity=(0, -9.80665, 0), damping_constant=2e-3, bc_kind=nat.BC_ONE_END_FIXED)
init = np.zeros((5, 9)); init[:, 3] = 1; init[:, 7] = 1; h.reset_host(init); h.step_host(None, K); h.close()
h = nat.Handle(model=nat.MODEL_ROD, n_env=6, n_elem=50, dt=7e-5, gravity=(0, 0, -9.81), damping_constant=1e-2,
               bc_kind=nat.BC_FREE, contact=arm_contact_params(), **_ROD)
init = np.zeros((6, 9)); init[:, 3] = 1; init[:, 8] = 1; h.reset_host(init); h.rest_kappa_tensor()[:, 0, :] = 3.0
h.step_host(None, K); h.close()
env = g.make_vec("OctoFlat-v0", 5, autoreset=False); env.reset(seed=1)
env.handle.rest_kappa_tensor()[:, 0, :] = 2.0
o6, rew, term = env._scratch
env.handle.step(None, K, o6, rew, term); torch.cuda.synchronize(); env.close()
# torque forcings: travelling-wave muscle torques + kinetic friction (ContinuumSnake), spline torques (SoftArmTracking)
env = g.make_vec("ContinuumSnake-v0", 6, autoreset=False); env.reset()
mu = env.handle.muscle_tensor(); mu[:, 2:] = 3e-3; mu[:, 1] = 6.4
o6, rew, term = env._scratch
env.handle.step(None, K, o6, rew, term); torch.cuda.synchronize(); env.close()
env = g.make_vec("SoftArmTracking-v0", 7, game_mode=2, autoreset=False); env.reset(seed=3)
for _ in range(2):
    env.step(torch.full((7, 8), 0.3, device="cuda", dtype=torch.float64))
torch.cuda.synchronize(); env.close()
# fast-only / fallback pair with flagged envs: two of five free rods spin outside the narrow rotation range
h = nat.Handle(model=nat.MODEL_ROD, n_env=5, n_elem=30, dt=1e-4, base_length=1.0, base_radius=0.05, density=1000.0,
               youngs_modulus=1e6, gravity=(0, -9.80665, 0), damping_constant=2e-3)
init = np.zeros((5, 9)); init[:, 3] = 1; init[:, 7] = 1; h.reset_host(init)
h.fields()["omega_collection"][1, 2, :] = 2.0e4; h.fields()["omega_collection"][3, 2, :] = -2.0e4
h.step_host(None, K); h.step_host(None, K); h.close()
# round 2: the split schedule (more env groups than resident CTA slots: items change CTAs through global scratch) of the
# lean kernel's plain, contact and assembly variants; a tapered rod with sucker and external loads (generic kernel, VARY)
sm = torch.cuda.get_device_properties(0).multi_processor_count
env = g.make_vec("SoftPendulum-v0", 10 * sm + 37, autoreset=False); env.reset(seed=1)
env.handle.step(torch.ones((10 * sm + 37, 1), device="cuda"), 5, env.obs, env.reward, env.terminated); torch.cuda.synchronize(); env.close()
n_big = 10 * sm + 23
h = nat.Handle(model=nat.MODEL_ROD, n_env=n_big, n_elem=50, dt=7e-5, gravity=(0, 0, -9.81), damping_constant=1e-2,
               bc_kind=nat.BC_FREE, contact=arm_contact_params(), **_ROD)
init = np.zeros((n_big, 9)); init[:, 3] = 1; init[:, 8] = 1; h.reset_host(init); h.rest_kappa_tensor()[:, 0, :] = 3.0
h.step_host(None, 5); h.close()
env = g.make_vec("OctoFlat-v0", 4 * sm + 9, autoreset=False); env.reset(seed=1)
env.handle.rest_kappa_tensor()[:, 0, :] = 2.0
o6, rew, term = env._scratch
env.handle.step(None, 5, o6, rew, term); torch.cuda.synchronize(); env.close()
env = g.make_vec("SoftPendulum3D-v0", 10 * sm + 11, autoreset=False); env.reset(seed=1)
env.handle.step(torch.full((10 * sm + 11, 2), 0.5, device="cuda"), 5, env.obs, env.reward, env.terminated); torch.cuda.synchronize(); env.close()
h = nat.Handle(model=nat.MODEL_ROD, n_env=5, n_elem=25, dt=2e-5, base_length=0.2, base_radius=0.012, density=1050.0, youngs_modulus=1e5,
               gravity=(0, 0, -9.81), damping_constant=0.05, bc_kind=nat.BC_FREE, tip_radius=0.002, sucker_index=0)
init = np.zeros((5, 9)); init[:, 3] = 1; init[:, 8] = 1; h.reset_host(init)
h.sucker_tensor()[:] = 0.5
f_t, c_t = h.ext_load_tensors(); f_t[:, 2, :] = 1e-4; c_t[:, 0, :] = 1e-6
h.step_host(None, K); h.close()
print("sanitize_smoke done")
