import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
from gym_softrobot_b200 import _native as nat
from gym_softrobot_b200.envs.snake import snake_contact_params, beta_spline_matrix
g = np.load(os.path.join(os.path.dirname(__file__), "snake_variants.npz"))
for tag, sc in (("nofric", 0.0), ("fric", 1.0)):
    c = snake_contact_params(); c["kinetic_mu"] = c["kinetic_mu"] * sc
    L=0.35; E=1e6
    h = nat.Handle(model=nat.MODEL_ROD, n_env=1, n_elem=50, dt=8e-6, base_length=L, base_radius=L*0.011, density=1000.0, youngs_modulus=E,
                   shear_modulus=E/1.5, gravity=(0,-9.80665,0), damping_constant=1e-4, contact=c,
                   muscle=dict(period=2.0, ramp_up_time=2.0, phase_shift=0.0, direction=(0,1.0,0)))
    init = np.zeros((1,9)); init[0,5]=1; init[0,7]=1
    h.reset_host(init)
    b = np.array([5.4791206e-03, -1.2224312e-03, 7.1719582e-03, 3.9473604e-03, -8.1164530e-03, 9.5124468e-03], dtype=np.float32).astype(np.float64)
    mu = h.muscle_tensor(); mu[0,2:] = torch.as_tensor(beta_spline_matrix(6,50) @ b, device="cuda")
    mu[0,1] = float(np.float32(2*np.pi)/np.float32(2.4028492))
    done=0
    for tgt in (1, 10, 100, 2083, 6000):
        h.step_host(None, tgt-done); done=tgt
        f = {k: x.double().cpu().numpy()[0] for k, x in h.fields().items()}
        msg=[f"{tag:6s} sub={tgt:5d}"]
        for k in ("position_collection","velocity_collection","director_collection","omega_collection"):
            ref = g[f"{tag}/s{tgt}/{k}"]; d = np.abs(f[k]-ref)
            msg.append(f"{k[:3]} {d.max():.2e}/{np.abs(ref).max():.2e}@{np.unravel_index(d.argmax(), d.shape)}")
        print("  ".join(msg))
    h.close()
