"""Which range do SoftPendulum3D / OctoArmSingle envs leave under random actions?  Per step: fallback delta and the
batch maxima of the three checked quantities reconstructed from the state (rotation per update, bend per element, z)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import gym_softrobot_b200 as g
which = sys.argv[1] if len(sys.argv) > 1 else "pend3d"
n_env = 4096
if which == "pend3d":
    env = g.make_vec("SoftPendulum3D-v0", n_env); lo, hi, shape, dt, dl = -1.0, 1.0, (2,), 1e-4, 1.0 / 50
elif which == "flat":
    env = g.make_vec("OctoFlat-v0", n_env); lo, hi, shape, dt, dl = -22.0, 22.0, (24,), 7e-5, None
else:
    env = g.make_vec("OctoArmSingle-v0", n_env); lo, hi, shape, dt, dl = -22.0, 22.0, (7,), 7e-5, None
env.reset(seed=1)
gen = torch.Generator().manual_seed(7)
prev = 0
cycle = [(torch.rand((n_env,) + shape, generator=gen) * (hi - lo) + lo).to("cuda") for _ in range(4)]   # bench_envs.py cycles four action sets
for s in range(int(sys.argv[2]) if len(sys.argv) > 2 else 24):
    a = cycle[(s + 1) % 4]
    env.step(a)
    f = env.fields()
    fb = env.handle.fallback_count()
    w = f["omega_collection"]; rot = (w.norm(dim=1) * dt)
    rot0 = rot[:, 0].clone()
    if which == "pend3d": rot[:, 0] = 0
    kap = f["kappa"].norm(dim=1)
    L = env.handle.cfg.base_length / env.handle.cfg.n_elem
    bend = kap * L
    em1 = (f["dilatation"] - 1).abs()
    print(f"step {s:2d} fallback +{fb - prev:5d} causes(rot,bend,stretch)={env.handle.fallback_causes()}  max rot/update {float(rot.max()):.3f} (lim 0.1)  max bend/elem {float(bend.max()):.3f} rad (lim 0.40)  "
          f"max |e-1| {float(em1.max()):.3f}  envs with bend>0.4: {int((bend.max(dim=1).values > 0.4).sum())}  rot>0.1: {int((rot.max(dim=1).values > 0.1).sum())}  "
          f"argmax bend elem {int(bend.max(dim=0).values.argmax())}  elem0 rot {float(rot0.max()):.3e}  envs elem0 rot>0.1: {int((rot0 > 0.1).sum())}  max|v| {float(f['velocity_collection'].abs().max()):.3f} finite {bool(torch.isfinite(f['velocity_collection']).all())}", flush=True)
    prev = fb
if which == "pend3d":
    f = env.fields()
    e = f["dilatation"] - 1
    worst = int(e.abs().max(dim=1).values.argmax())
    print("worst env", worst, "e-1 profile", [round(float(v), 4) for v in e[worst, :8]], "...", [round(float(v), 4) for v in e[worst, 44:50]])
    print("base pos", env.handle.aux_tensor()[worst].tolist())
    x = f["position_collection"][worst]
    print("node0", x[:, 0].tolist(), "node1", x[:, 1].tolist(), "tip", x[:, 50].tolist())
    # transient inside one env-step: 40 launches of 10 substeps with a zero action (the base stays where it is)
    z = torch.zeros((n_env, 2), device="cuda")
    peak = 0.0
    for k in range(40):
        env.handle.step(z, 10, env.obs, env.reward, env.terminated)
        peak = max(peak, float((env.fields()["dilatation"] - 1).abs().max()))
    print("peak |e-1| over the next 400 substeps sampled every 10 (zero action):", peak, "fallback total", env.handle.fallback_count(), env.handle.fallback_causes())
