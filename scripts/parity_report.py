"""Per-field error of the CUDA path vs the C oracle (diagnostic; prints a table)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import rod_oracle
from gym_softrobot_b200.envs.soft_pendulum import _make_handle, pendulum_init_params

def rel(a, b): return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
n_env = 8
rng = np.random.default_rng(0)
u = rng.random(n_env)
acts = rng.uniform(-22, 22, size=(6, n_env, 1)).astype(np.float32)
for math, dtype in ((0, 0), (1, 0), (0, 1)):
    h = _make_handle(n_env, 50, 1e-4, 0, math, dtype)
    h.reset_host(pendulum_init_params(u))
    orc = [rod_oracle.OracleSoftPendulum() for _ in range(n_env)]
    for i, o in enumerate(orc): o.reset(u01=u[i])
    for s in range(6):
        h.step_host(acts[s], 400)
        f = {k: v.double().cpu().numpy() for k, v in h.fields().items()}
        errs = {}
        for i, o in enumerate(orc):
            o.step(acts[s][i])
            for name in ("position_collection", "velocity_collection", "director_collection", "omega_collection", "tangents", "kappa"):
                errs[name] = max(errs.get(name, 0), rel(f[name][i], getattr(o.rod, name)))
        print(f"dtype={'f64' if dtype == 0 else 'f32'} math={math} substeps={400*(s+1)}: " + " ".join(f"{k.split('_')[0]}={v:.1e}" for k, v in errs.items()), flush=True)
    h.close()
