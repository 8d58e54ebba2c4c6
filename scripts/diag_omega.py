"""Where does the absolute 3-5e-10 rad/s omega offset of a clamped rod come from?  Abs errors over the first substeps,
with / without damper and gravity tilt, FAST and FAITHFUL math."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import rod_oracle as ro
from gym_softrobot_b200 import _native as nat
n_elem = 100
L, r0, E, rho = 1.0, 0.05, 1e6, 2000.0
dt = float(0.03 * (L / n_elem) / np.sqrt(E / rho))
ang = np.deg2rad(5.0)
d = np.array([np.cos(ang), 0.0, np.sin(ang)]); nn = np.array([0.0, 1.0, 0.0])
for label, damping, math in (("fast, damper 0.3", 0.3, nat.MATH_FAST), ("fast, no damper", -1.0, nat.MATH_FAST), ("faithful, damper 0.3", 0.3, nat.MATH_FAITHFUL)):
    kw = dict(gravity=(0.0, 0.0, -9.80665), damping_constant=damping, bc_kind=1)
    h = nat.Handle(model=nat.MODEL_ROD, n_env=3, n_elem=n_elem, dt=dt, base_length=L, base_radius=r0, density=rho, youngs_modulus=E, math=math, **kw)
    init = np.zeros((3, 9)); init[:, 3:6] = d; init[:, 6:9] = nn
    h.reset_host(init)
    o = ro.OracleRod(n_elem, [0, 0, 0], list(d), list(nn), L, r0, rho, E, dt, **kw)
    f = {k: v.cpu().numpy() for k, v in h.fields().items()}
    print(label, "initial: pos", np.abs(f["position_collection"][1] - o.position_collection).max(), "dir", np.abs(f["director_collection"][1] - o.director_collection).max())
    tot = 0
    for chunk in (1, 1, 3, 5, 10, 30, 50, 100, 300):
        h.step_host(None, chunk); o.substeps(chunk); tot += chunk
        f = {k: v.cpu().numpy() for k, v in h.fields().items()}
        ew = np.abs(f["omega_collection"][1] - o.omega_collection); ev = np.abs(f["velocity_collection"][1] - o.velocity_collection)
        print(f"  {label} after {tot:4d}: |w| {np.abs(o.omega_collection).max():.3e} abs err w {ew.max():.2e} @elem {int(ew.max(axis=0).argmax())} comp {int(ew.max(axis=1).argmax())} | |v| {np.abs(o.velocity_collection).max():.3e} abs err v {ev.max():.2e} | sigma err {np.abs(f['sigma'][1] - o.sigma).max():.1e} kappa err {np.abs(f['kappa'][1] - o.kappa).max():.1e}", flush=True)
    h.close(); o.close()
