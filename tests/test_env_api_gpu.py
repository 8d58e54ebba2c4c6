"""GPU: Gymnasium-API conformance of every registered env, mirroring the reference's own tests
(`/root/reference/tests/envs/test_envs.py:18-68`, `test_determinism.py:7-58`), plus the batched
envs' autoreset and the state get/set entry points."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

IDS = ["SoftPendulum-v0", "SoftPendulum3D-v0", "OctoArmSingle-v0", "OctoFlat-v0", "OctoFlatLite-v0",
       "ContinuumSnake-v0", "SoftArmTracking-v0", "OctoCrawl-v0", "OctoArmPush-v0", "OctoArmPush-v1",
       "OctoArmPullWeight-v0", "OctoReach-v0", "OctoArmTwo-v0"]
INFO_KEY = {"ContinuumSnake-v0": None, "SoftArmTracking-v0": "ctime"}     # what the reference env puts in info
# The reference's ArmTwoEnv keeps `_prev_kappa` (part of the observation) across reset(): the first reset after a step
# still shows the last step's curvature (arm_two_env.py:103-106,201-219); mirrored, so the pair compared here is taken
# from back-to-back resets like check_env's own, which runs on a fresh env.
CARRIES_KAPPA = {"OctoArmTwo-v0"}
FAST_KW = {"OctoArmSingle-v0": dict(recording_fps=100), "OctoFlat-v0": dict(recording_fps=100),
           "OctoFlatLite-v0": dict(recording_fps=100)}


def _contains(space_or_shapes, obs):
    if isinstance(obs, dict):
        return all(np.isfinite(v).all() and v.dtype == np.float32 for v in obs.values()) and space_or_shapes.contains(obs)
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
    key = INFO_KEY.get(env_id, "time")
    assert key is None or key in info
    # reset(seed) twice gives the same first observation (check_env's determinism requirement)
    if env_id in CARRIES_KAPPA:
        env.reset(seed=3)
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


def test_soft_arm_vector_env_truncation_and_autoreset():
    """Batched SoftArmTracking (moving targets): truncation on env-step 500 (tick * sim_dt >= 5 evaluated in
    float64 like soft_arm_tracking.py:254), rebuilt inside the same step(), fresh targets per episode."""
    import torch
    import gym_softrobot_b200 as gsb
    n_env = 5
    env = gsb.make_vec("SoftArmTracking-v0", n_env, game_mode=2)
    obs0, _ = env.reset(seed=7)
    assert env.n_updates == 500 and obs0.dtype == torch.float64 and obs0.shape == (n_env, 14)
    first_targets = env._targets.clone()
    assert not torch.equal(first_targets[0], first_targets[1])            # one stream per env
    gen = torch.Generator(device="cuda").manual_seed(0)
    for s in range(1, 503):
        # moderate actions: full-range white noise re-drawn every 10 ms can blow the arm up (the reference
        # anticipates it: "Episode blew up. Maybe try a smaller dt?"); that path is tested separately below
        a = 0.3 * (torch.rand((n_env, 8), generator=gen, device="cuda", dtype=torch.float64) * 2 - 1)
        obs, rew, term, trunc, info = env.step(a)
        assert not bool(term.any()) and bool(trunc.all()) == (s == 500)
        assert torch.isfinite(obs).all() and bool((rew <= 0).all())
        if s == 500:
            assert info["final_obs"].shape == (n_env, 14) and int(env.tick.max()) == 0
            assert torch.equal(obs[:, :8], torch.zeros((n_env, 8), dtype=torch.float64, device="cuda"))   # straight arm
            assert torch.allclose(obs[:, 8:11], torch.tensor([0.0, 1.0, 0.0], dtype=torch.float64, device="cuda").expand(n_env, 3))
            assert not torch.equal(env._targets, first_targets)           # new trajectories
            pts, mags = env.handle.spline_tensors()
            assert float(pts.abs().max()) == 0.0 and float(mags.abs().max()) == 0.0   # fresh forcing instances
    assert int(env.tick.min()) == 2
    env.close()


def test_soft_arm_blow_up_path():
    """soft_arm_tracking.py:247-252: a NaN state gives reward -100, a nan_to_num'ed observation and
    terminated=True; the batched env rebuilds that env only."""
    import torch
    import gym_softrobot_b200 as gsb
    env = gsb.make_vec("SoftArmTracking-v0", 4)
    env.reset(seed=0)
    a = torch.zeros((4, 8), dtype=torch.float64, device="cuda")
    env.step(a)
    env.fields()["velocity_collection"][2, 0, 5] = float("nan")
    obs, rew, term, trunc, info = env.step(a)
    assert term.tolist() == [False, False, True, False] and not bool(trunc.any())
    assert float(rew[2]) == -100.0 and bool((rew[[0, 1, 3]] > -1.0).all())
    assert torch.isfinite(info["final_obs"]).all() and info["reset_idx"].tolist() == [2]
    assert env.tick.tolist() == [2, 2, 0, 2] and torch.isfinite(obs).all()
    env.close()


def test_snake_vector_env_truncation_and_autoreset():
    """Batched ContinuumSnake: the in-kernel clock reaches final_time = 22.02 s on env-step 111
    (continuum_snake.py:279-287, 211-212), every env is rebuilt in that step and the clock restarts."""
    import torch
    import gym_softrobot_b200 as gsb
    n_env = 3
    env = gsb.make_vec("ContinuumSnake-v0", n_env)
    obs0, _ = env.reset()
    a = torch.tensor([[3.4e-3, 3.3e-3, 4.2e-3, 2.6e-3, 3.6e-3, 3.5e-3, 0.97]], device="cuda").repeat(n_env, 1)
    for s in range(1, 113):
        obs, rew, term, trunc, info = env.step(a)
        assert not bool(term.any()) and bool(trunc.all()) == (s == 111), s
        if s == 110:
            assert float(rew.min()) > 0.01                                # the published gait crawls forward
        if s == 111:
            assert torch.equal(obs, obs0) and float(env.handle.muscle_tensor()[:, 0].max()) == 0.0
            assert len(env._times) == 1 and info["final_obs"].shape == (n_env, 756)
    assert abs(float(info["time"][0]) - 0.2) < 1e-9 and torch.isfinite(obs).all()
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


def test_second_episode_reset_observation_matches_reference(golden_dir):
    """the reference never clears `_prev_action` in reset() for SoftPendulum / OctoArmSingle / OctoFlat
    (soft_pendulum.py:97, arm_single_env.py:100, flat_env.py:135-139), so the observation returned by the reset that
    opens the second episode still carries the last action; SoftPendulum3D does clear it (soft_pendulum_3d.py:68).
    Fixture: the reference envs themselves (oracle/gen_golden.py:gen_second_episode_reset_obs)."""
    import gym_softrobot_b200 as gsb
    g = np.load(os.path.join(golden_dir, "second_episode_reset_obs.npz"))
    for env_id, kw in (("SoftPendulum-v0", {}), ("SoftPendulum3D-v0", {}), ("OctoArmSingle-v0", {}),
                       ("OctoFlat-v0", {"recording_fps": 50})):
        tag = env_id.split("-")[0]
        env = gsb.make(env_id, **kw)
        env.reset(seed=1)
        env.step(g[f"{tag}/action"])
        obs2, _ = env.reset(seed=2)
        if isinstance(obs2, dict):
            for k, v in obs2.items():
                np.testing.assert_allclose(v, g[f"{tag}/obs2/{k}"], rtol=1e-5, atol=1e-6, err_msg=f"{env_id} {k}")
        else:
            np.testing.assert_allclose(obs2, g[f"{tag}/obs2"], rtol=1e-5, atol=1e-6, err_msg=env_id)
        env.close()


@pytest.mark.parametrize("env_id,kw", [("SoftPendulum3D-v0", {}), ("OctoFlat-v0", dict(recording_fps=100)),
                                       ("ContinuumSnake-v0", {}), ("SoftArmTracking-v0", {}),
                                       ("OctoReach-v0", {}), ("OctoArmTwo-v0", {})])
def test_clone_from_copies_everything_that_evolves(env_id, kw):
    """sr_copy_from: unlike sr_set_state (rod arrays only) it also carries BC anchors, the 3D pendulum's
    base controller, rigid heads, rest curvatures and the forcings' state (incl. the per-element muscle activations and
    the fixed-index sucker ratios of the muscle-layer envs) — a cloned handle continues bit for bit."""
    import torch
    import gym_softrobot_b200 as gsb
    n_env = 5
    e1, e2 = gsb.make_vec(env_id, n_env, autoreset=False, **kw), gsb.make_vec(env_id, n_env, autoreset=False, **kw)
    e1.reset(seed=3); e2.reset(seed=99)
    sp = e1.single_action_space
    rng = np.random.default_rng(0)
    act = lambda: torch.as_tensor(rng.uniform(sp.low, sp.high, size=(n_env,) + sp.shape).astype(sp.dtype), device="cuda")
    for _ in range(2):
        e1.step(act())
    e2.handle.clone_from(e1.handle)
    K, obs = 150, torch.empty((n_env, e1.handle.obs_dim), dtype=torch.float32, device="cuda")
    rew, term = torch.empty(n_env, dtype=torch.float64, device="cuda"), torch.empty(n_env, dtype=torch.uint8, device="cuda")
    a = torch.zeros((n_env, max(e1.handle.action_dim, 1)), dtype=torch.float32, device="cuda")
    for h in (e1.handle, e2.handle):
        h.step(a if h.action_dim else None, K, obs, rew, term)
    torch.cuda.synchronize()
    assert torch.equal(e1.handle.state_tensor(), e2.handle.state_tensor())
    assert torch.equal(e1.handle.aux_tensor(), e2.handle.aux_tensor())
    if env_id in ("OctoFlat-v0", "OctoReach-v0", "OctoArmTwo-v0"):
        assert torch.equal(e1.handle.head_tensor(), e2.handle.head_tensor())
    if env_id in ("OctoReach-v0", "OctoArmTwo-v0"):
        assert float(e2.handle.muscle_activation_tensor().abs().max()) > 0.0      # the activations travelled with the clone
    e1.close(); e2.close()
