"""compute-sanitizer workload for the muscle-layer instantiation of the generic kernel (OctoReach-v0: 8 arms + pinned
head, 2 env groups per 384-thread CTA with an odd tail; OctoArmTwo-v0: 2 arms + free head + fixed-index suckers, 8 env
groups per CTA) and, beside it, the transverse-muscle-only instantiation it shares the code with (OctoCrawl-v0)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import gym_softrobot_b200 as g

K = 12
gen = torch.Generator().manual_seed(3)
for env_id, n_env, n_act in (("OctoReach-v0", 3, 480), ("OctoArmTwo-v0", 11, 18), ("OctoCrawl-v0", 3, 24)):
    env = g.make_vec(env_id, n_env, autoreset=False)
    env.reset(seed=1)
    env.set_action(torch.rand((n_env, n_act), generator=gen).cuda())
    o6, rew, term = env._scratch
    for _ in range(2):
        env.handle.step(None, K, o6, rew, term)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(env.fields()["position_collection"]).all())
    env.close()
print("sanitize_muscle done")
