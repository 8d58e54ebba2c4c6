import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
import gym_softrobot_b200 as gsb
g = np.load("tests/golden/continuum_snake_seed42.npz")
env = gsb.make("ContinuumSnake-v0")
obs0, _ = env.reset(seed=42)
print("obs0 err", np.abs(obs0 - g["obs0"]).max())
F = {"position": "position_collection", "velocity": "velocity_collection", "director": "director_collection", "omega": "omega_collection"}
t0 = time.time()
for i, a in enumerate(g["actions"]):
    obs, r, te, tr, info = env.step(a)
    msg = f"step {i+1:2d} t_err {abs(env.time - float(g['time'][i])):.1e} reward {r:+.6e} ref {float(g['reward'][i]):+.6e}"
    if f"state{i+1}/position" in g.files:
        st = env.rod_state()
        for gk, fk in F.items():
            ref = g[f"state{i+1}/{gk}"]
            msg += f"  {gk[:3]} {np.abs(st[fk]-ref).max():.2e}/{np.abs(ref).max():.2e}"
    print(msg, flush=True)
print("wall", time.time() - t0)
st = env.rod_state()
for gk, fk in F.items():
    ref = g[f"state_final/{gk}"]
    print("final", gk, f"{np.abs(st[fk]-ref).max():.2e}/{np.abs(ref).max():.2e}")
v = env._vec; S = len(g["cb_time"])
com, vel = v._com[0, :S].cpu().numpy(), v._vel[0, :S].cpu().numpy()
print("samples", len(v._times), S, "time err", np.abs(np.array(v._times) - g["cb_time"]).max())
dc = np.abs(com - g["cb_com"]).max(axis=1); dv = np.abs(vel - g["cb_avg_velocity"]).max(axis=1)
for k in (1, 12, 36, 120, 240, S - 1):
    print(f"sample {k}: com err {dc[k]:.2e} (|com| {np.abs(g['cb_com'][k]).max():.2e})  vel err {dv[k]:.2e} (|vel| {np.abs(g['cb_avg_velocity'][k]).max():.2e})")
