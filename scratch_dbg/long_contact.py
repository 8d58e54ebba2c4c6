import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
import rod_oracle as ro
from gym_softrobot_b200 import _native as nat
from gym_softrobot_b200.envs.arm_single import arm_contact_params
n, dt, L, r = 512, 5e-6, 1.0, 0.005
for bf in (True, False):
    c = arm_contact_params(before_forcing=bf); c["plane_origin"] = [0.0, 0.0, -r]
    rk = np.random.default_rng(9).uniform(-3, 3, size=(1, 1)) * np.sin(np.linspace(0, 3 * np.pi, n - 1))[None, :]
    o = ro.OracleRod(n, [0, 0, 0], [1.0, 0, 0], [0, 0, 1.0], L, r, 1000.0, 1e6, dt, gravity=(0.0, 0.0, -9.81), damping_constant=1e-2, contact=c)
    o.rest_kappa[0, :] = rk[0]
    h = nat.Handle(model=nat.MODEL_ROD, n_env=1, n_elem=n, dt=dt, gravity=(0.0, 0.0, -9.81), damping_constant=1e-2, bc_kind=nat.BC_FREE, contact=c, base_length=L, base_radius=r, density=1000.0, youngs_modulus=1e6)
    init = np.zeros((1, 9)); init[:, 3] = 1.0; init[:, 8] = 1.0
    h.reset_host(init)
    h.rest_kappa_tensor()[:, 0, :] = torch.as_tensor(rk, device="cuda")
    done = 0
    for tgt in (1, 2, 5, 10, 50, 100, 200, 400):
        o.substeps(tgt - done); h.step_host(None, tgt - done); done = tgt
        f = {k: v.double().cpu().numpy()[0] for k, v in h.fields().items()}
        dx = np.abs(f["position_collection"] - o.position_collection); dv = np.abs(f["velocity_collection"] - o.velocity_collection)
        print(f"bf={bf} sub={tgt:4d} max|dx|={dx.max():.3e} at {np.unravel_index(dx.argmax(), dx.shape)}  max|dv|={dv.max():.3e} at {np.unravel_index(dv.argmax(), dv.shape)}  |v|max={np.abs(o.velocity_collection).max():.3e}")
    h.close()
