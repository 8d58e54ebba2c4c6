"""End-to-end throughput of the registered envs through the public vector-env API under RANDOM actions (host pinned
actions -> device, step(), observations + rewards -> host), with the number of env-steps the fast-only kernels handed
to the safe kernel: the figure that shows whether a polynomial range is wide enough for what a policy really does."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import gym_softrobot_b200 as g

def timed(step_fn, K, W):
    for _ in range(W): step_fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    for a, b in ev:
        a.record(); step_fn(); b.record()
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in ev) / K * 1e-3

def e2e(env_id, n_env, act_shape, lo, hi, K, dtype=torch.float32, W=18, **kw):   # W: past the adaptive switch's first checks
    env = g.make_vec(env_id, n_env, **kw)
    env.reset(seed=1)
    gen = torch.Generator().manual_seed(7)
    acts = [((torch.rand((n_env,) + act_shape, dtype=dtype, generator=gen) * torch.as_tensor(hi - lo, dtype=dtype) + torch.as_tensor(lo, dtype=dtype))).pin_memory()
            for _ in range(4)]
    it = [0]
    def step():
        it[0] += 1
        obs, rew, term, trunc, info = env.step(acts[it[0] % 4].to("cuda", non_blocking=True))
        o = obs["individual"] if isinstance(obs, dict) else obs
        return o.cpu(), rew.cpu()
    sec = timed(step, K=K, W=W)
    print(json.dumps(dict(env=env_id, n_env=n_env, ms_per_step=round(sec * 1e3, 3), env_steps_per_s=round(n_env / sec, 1),
                          fallback_env_steps=env.handle.fallback_count(), env_steps_run=(K + W) * n_env, launches=env.handle.launch_count,
                          knobs={k: v for k, v in os.environ.items() if k.startswith("SOFTROD_")})), flush=True)
    env.close()

which = sys.argv[1].split(",") if len(sys.argv) > 1 else ["pend", "pend3d", "arm", "flat", "softarm"]
if "pend" in which: e2e("SoftPendulum-v0", 4096, (1,), np.float32(-22.0), np.float32(22.0), 10)
if "pend3d" in which: e2e("SoftPendulum3D-v0", 4096, (2,), np.float32(-1.0), np.float32(1.0), 10)
if "arm" in which: e2e("OctoArmSingle-v0", 4096, (7,), np.float32(-22.0), np.float32(22.0), 5)
if "flat" in which: e2e("OctoFlat-v0", 4096, (24,), np.float32(-22.0), np.float32(22.0), 3)
if "softarm" in which: e2e("SoftArmTracking-v0", 16384, (8,), -0.3, 0.3, 10, dtype=torch.float64)
# the COOMM muscle envs (transverse muscle in the kernel, tapered-rod generic kernel): continuous actions in [0, 1]
if "push" in which: e2e("OctoArmPush-v1", 4096, (2,), np.float32(0.0), np.float32(1.0), 5)
if "pull" in which: e2e("OctoArmPullWeight-v0", 4096, (2,), np.float32(0.0), np.float32(1.0), 3)
if "crawl" in which: e2e("OctoCrawl-v0", 1024, (24,), np.float32(0.0), np.float32(1.0), 3)
# the muscle-layer kernel (two longitudinal + the transverse muscle, per-element activations)
if "reach" in which: e2e("OctoReach-v0", 1024, (480,), np.float32(0.0), np.float32(1.0), 3, W=2)
if "armtwo" in which: e2e("OctoArmTwo-v0", 4096, (18,), np.float32(0.0), np.float32(1.0), 3, W=2)
