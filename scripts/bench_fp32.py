"""FP32 vs FP64 mode: speed ratio on the headline workload (diagnostic)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import gym_softrobot_b200 as g
n_env = 4096
for dtype in ("float64", "float32"):
    env = g.make_vec("SoftPendulum-v0", n_env, autoreset=False, dtype=dtype); env.reset(seed=42)
    gen = torch.Generator(device="cuda").manual_seed(42)
    acts = (torch.rand((28, n_env, 1), generator=gen, device="cuda") * 44 - 22).float()
    for s in range(3): env.handle.step(acts[s], 400, env.obs, env.reward, env.terminated)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(25)]
    for s, (a, b) in enumerate(ev):
        a.record(); env.handle.step(acts[3 + s], 400, env.obs, env.reward, env.terminated); b.record()
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in ev) / 25
    print(json.dumps({"dtype": dtype, "ms_per_step": ms, "env_steps_per_s": n_env / ms * 1e3, "terminated": int(env.terminated.sum())}))
    env.close()
