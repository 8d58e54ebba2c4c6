import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import rod_oracle as ro
from gym_softrobot_b200 import _native as nat
n_elem, bc, damp_first = int(sys.argv[1]), int(sys.argv[2]), bool(int(sys.argv[3]))
L, r0, E, rho = 1.0, 0.05, 1e6, 2000.0
dt = float(0.03 * (L / n_elem) / np.sqrt(E / rho))
ang = np.deg2rad(25.0)
d = np.array([np.cos(ang), 0.0, np.sin(ang)]); nn = np.array([0.0, 1.0, 0.0])
for order in (7, 0):
    kw = dict(gravity=(0.0, 0.0, -9.80665), damping_constant=0.3, laplace_filter_order=order, bc_kind=bc, damping_before_constraints=damp_first)
    h = nat.Handle(model=nat.MODEL_ROD, n_env=3, n_elem=n_elem, dt=dt, base_length=L, base_radius=r0, density=rho, youngs_modulus=E, **kw)
    init = np.zeros((3, 9)); init[:, 3:6] = d; init[:, 6:9] = nn
    h.reset_host(init)
    o = ro.OracleRod(n_elem, [0, 0, 0], list(d), list(nn), L, r0, rho, E, dt, **kw)
    for chunk in (1, 9, 90, 200):
        h.step_host(None, chunk); o.substeps(chunk)
        f = {k: v.cpu().numpy() for k, v in h.fields().items()}
        out = []
        for name in ("position_collection", "velocity_collection", "director_collection", "omega_collection"):
            ref = getattr(o, name); err = np.abs(f[name][1] - ref)
            idx = np.unravel_index(err.argmax(), err.shape)
            out.append(f"{name[:3]} {err.max() / max(np.abs(ref).max(), 1e-300):.1e}@{idx[-1]}")
        print(f"order {order} n={n_elem} bc={bc} df={damp_first} after +{chunk}: " + "  ".join(out), "| fallback", h.fallback_count(), h.fallback_causes(), flush=True)
    h.close(); o.close()
