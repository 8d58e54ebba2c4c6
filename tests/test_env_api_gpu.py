"""GPU: Gymnasium-API conformance of every registered env, mirroring the reference's own tests
(`/root/reference/tests/envs/test_envs.py:18-68`, `test_determinism.py:7-58`), plus the batched
envs' autoreset and the state get/set entry points."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

IDS = ["SoftPendulum-v0", "SoftPendulum3D-v0", "OctoArmSingle-v0", "OctoFlat-v0", "OctoFlatLite-v0"]
FAST_KW = {"OctoArmSingle-v0": dict(recording_fps=100), "OctoFlat-v0": dict(recording_fps=100),
           "OctoFlatLite-v0": dict(recording_fps=100)}


def _contains(space_or_shapes, obs):
    if isinstance(obs, dict):
        return all(np.isfinite(v).all() and v.dtype == np.float32 for v in obs.values())
    return space_or_shapes.contains(obs)


@pytest.mark.parametrize("env_id", IDS)
def test_env_api(env_id):
    import gym_softrobot_b200 as gsb
    env = gsb.make(env_id, **FAST_KW.get(env_id, {}))
    ob, info = env.reset()
    assert isinstance(info, dict)
    if not isinstance(ob, dict):
        assert env.observation_space.contains(ob) and ob.dtype == env.observation_space.dtype
    a = env.action_space.sample()
    assert env.action_space.contains(a)
    ob, reward, terminated, truncated, info = env.step(a)
    assert _contains(getattr(env, "observation_space", None), ob)
    assert np.isscalar(reward) and isinstance(terminated, bool) and isinstance(truncated, bool)
    assert "time" in info
    # reset(seed) twice gives the same first observation (check_env's determinism requirement)
    o1, _ = env.reset(seed=3)
    o2, _ = env.reset(seed=3)
    if isinstance(o1, dict):
        assert all(np.array_equal(o1[k], o2[k]) for k in o1)
    else:
        assert np.array_equal(o1, o2)
    env.close()


@pytest.mark.parametrize("env_id", IDS)
def test_env_determinism_protocol(env_id):
    """Two fresh envs, reset(seed=0), action_space.seed(0), 3 steps: exact equality."""
    import gym_softrobot_b200 as gsb
    runs = []
    for _ in range(2):
        env = gsb.make(env_id, **FAST_KW.get(env_id, {}))
        o0, _ = env.reset(seed=0)
        env.action_space.seed(0)
        acts = [env.action_space.sample() for _ in range(3)]
        runs.append((o0, acts, [env.step(a) for a in acts]))
        env.close()
    (o1, a1, r1), (o2, a2, r2) = runs

    def eq(x, y):
        return all(np.array_equal(x[k], y[k]) for k in x) if isinstance(x, dict) else np.array_equal(x, y)
    assert eq(o1, o2) and all(np.array_equal(x, y) for x, y in zip(a1, a2))
    for (ob1, rw1, t1, x1, _), (ob2, rw2, t2, x2, _) in zip(r1, r2):
        assert eq(ob1, ob2) and rw1 == rw2 and t1 == t2 and x1 == x2


def test_vector_env_autoreset_and_truncation():
    """Batched SoftPendulum: truncation fires on env-step 126 for every env (time accumulated like the
    reference), finished envs are rebuilt inside the same step() and keep stepping."""
    import torch
    import gym_softrobot_b200 as gsb
    n_env = 32
    env = gsb.make_vec("SoftPendulum-v0", n_env, final_time=0.2)   # 5 env-steps to t = 0.2, truncated on the 6th? see below
    obs, _ = env.reset(seed=1)
    first = env._first_truncated
    assert env._time_table[first] > 0.2 and env._time_table[first - 1] <= 0.2
    gen = torch.Generator(device="cuda").manual_seed(0)
    for s in range(1, first + 3):
        a = (torch.rand((n_env, 1), generator=gen, device="cuda") * 44 - 22).float()
        obs, rew, term, trunc, info = env.step(a)
        assert bool(trunc.all()) == (s == first) or s > first
        if s == first:
            assert "final_obs" in info and info["final_obs"].shape == (n_env, 4)
            assert int(env.step_count.max()) == 0                     # all rebuilt
            assert torch.equal(obs[:, 1], torch.zeros(n_env, device="cuda"))   # fresh rods are at rest
            assert not torch.equal(obs, info["final_obs"])
    assert int(env.step_count.min()) == 2 and torch.isfinite(obs).all()
    env.close()


def test_state_get_set_roundtrip():
    """sr_get_state / sr_set_state: copying the SoA state between handles reproduces the trajectory bit for bit."""
    import torch
    from gym_softrobot_b200.envs.soft_pendulum import _make_handle, pendulum_init_params
    n_env = 9
    a = torch.full((n_env, 1), 5.0, device="cuda")
    bufs = lambda: (torch.empty((n_env, 4), dtype=torch.float32, device="cuda"),
                    torch.empty(n_env, dtype=torch.float64, device="cuda"),
                    torch.empty(n_env, dtype=torch.uint8, device="cuda"))
    h1, h2 = _make_handle(n_env, 50, 1e-4, 0, 0), _make_handle(n_env, 50, 1e-4, 0, 0)
    init = torch.as_tensor(pendulum_init_params(np.linspace(0.1, 0.9, n_env)), device="cuda").contiguous()
    h1.reset(init); h2.reset(init)
    o1, r1, t1 = bufs(); o2, r2, t2 = bufs()
    h1.step(a, 300, o1, r1, t1)
    h2.set_state_from(h1)
    h1.step(a, 300, o1, r1, t1); h2.step(a, 300, o2, r2, t2)
    torch.cuda.synchronize()
    v = h1.state_view()
    assert (v.n_env, v.n_fields, v.stride, v.elem_size) == (n_env, 32, 64, 8)
    assert torch.equal(h1.state_tensor(), h2.state_tensor()) and torch.equal(o1, o2) and torch.equal(r1, r2)
    assert h1.launch_count >= 3
    h1.close(); h2.close()
