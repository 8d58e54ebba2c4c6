"""Per-env-step kernel time and state extremes over an episode (diagnostic)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import gym_softrobot_b200 as g
n_env = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
env = g.make_vec("SoftPendulum-v0", n_env, autoreset=False)
env.reset(seed=42)
gen = torch.Generator(device="cuda").manual_seed(42)
for s in range(nsteps):
    a = (torch.rand((n_env, 1), generator=gen, device="cuda") * 44 - 22).float()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); env.handle.step(a, 400, env.obs, env.reward, env.terminated); e1.record()
    torch.cuda.synchronize()
    f = env.fields()
    kap = f["kappa"].abs().amax().item() * (1.0 / 50)
    w = f["omega_collection"].norm(dim=1).amax().item() * 1e-4
    dil = (f["dilatation"] - 1).abs().amax().item()
    print(f"step {s+1:3d}: {e0.elapsed_time(e1):7.3f} ms  max bend angle {kap:.4f} rad  max |w|dt {w:.2e}  max|e-1| {dil:.2e}  term {int(env.terminated.sum())}", flush=True)
