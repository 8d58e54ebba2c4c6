import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
import gym_softrobot_b200 as gsb
g = np.load(os.path.join(os.path.dirname(__file__), "snake_small.npz"))
v = gsb.make_vec("ContinuumSnake-v0", 2)
v.reset()
mu = v.handle.muscle_tensor()
a = torch.as_tensor(g["action"], device="cuda")[None].repeat(2, 1)
b = a[:, :6].double()
beta = sum(b[:, k:k + 1] * v._W[:, k] for k in range(6))
mu[:, 2:] = beta
mu[:, 1] = (torch.tensor(2.0 * np.pi, dtype=torch.float32, device="cuda") / a[:, 6]).double()
print("beta err", (beta[0].cpu().numpy() - g["beta"]).__abs__().max(), "kw", mu[0, 1].item(), float(g["kw"]))
done = 0
o6, r, t = v._scratch
for tgt in (1, 10, 100, 2083, 6000):
    v.handle.step(None, tgt - done, o6, r, t); done = tgt
    f = {k: x.double().cpu().numpy()[0] for k, x in v.handle.fields().items()}
    msg = [f"sub={tgt:5d} t_err={abs(mu[0,0].item()-float(g[f's{tgt}/time'])):.1e}"]
    for k in ("position_collection", "velocity_collection", "director_collection", "omega_collection"):
        ref = g[f"s{tgt}/{k}"]
        msg.append(f"{k[:3]} {np.abs(f[k]-ref).max():.2e}/{np.abs(ref).max():.2e}")
    print("  ".join(msg))
