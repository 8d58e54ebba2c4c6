"""CPU: host-side logic of the drop-in layer (no kernels): seeding, spaces, time accumulation,
env sharding (incl. a world_size-2 gloo run) and the registry."""
import os
import subprocess
import sys

import numpy as np
import pytest

import gym_softrobot_b200 as gsb
from gym_softrobot_b200 import distributed as D
from gym_softrobot_b200.compat import Box
from gym_softrobot_b200.envs import soft_pendulum as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_registry_keeps_reference_ids():
    """Every id the reference registers (/root/reference/gym_softrobot/__init__.py:6-80; SoftArmTracking-v1 is commented
    out there) has a drop-in facade and a batched vector env here, with the reference's registration kwargs."""
    assert "SoftPendulum-v0" in gsb.REGISTRY and "SoftPendulum-v0" in gsb.VECTOR_REGISTRY
    reference_ids = ["OctoFlat-v0", "OctoFlatLite-v0", "OctoCrawl-v0", "OctoReach-v0", "OctoArmSingle-v0", "OctoArmTwo-v0",
                     "OctoArmPush-v0", "OctoArmPush-v1", "OctoArmPullWeight-v0", "ContinuumSnake-v0", "SoftArmTracking-v0",
                     "SoftPendulum-v0", "SoftPendulum3D-v0"]
    assert sorted(gsb.REGISTRY) == sorted(reference_ids) == sorted(gsb.VECTOR_REGISTRY)
    assert gsb.REGISTRY["OctoFlatLite-v0"][1] == dict(n_arm=1, n_action=8)                 # __init__.py:11-15
    assert gsb.REGISTRY["OctoArmPush-v1"][1] == dict(mode="continuous")                    # :42-46
    assert gsb.REGISTRY["OctoArmPullWeight-v0"][1] == dict(mode="continuous")              # :48-52


def test_init_params_match_reference_build(golden_dir):
    """direction/normal from the env generator exactly as soft_pendulum/build.py:47-51."""
    g = np.load(os.path.join(golden_dir, "soft_pendulum_seed42_episode.npz"))
    u = np.random.Generator(np.random.PCG64(np.random.SeedSequence(42))).random()
    assert u == 0.7739560485559633
    init = sp.pendulum_init_params(u)[0]
    Q0 = g["state0/director"][:, :, 0]
    # d3 = normalised position difference of the linspace nodes, d1 = normal/|normal|: equal to
    # direction / normal up to one rounding (the reset kernel repeats the reference's arithmetic)
    np.testing.assert_allclose(init[3:6], Q0[2], rtol=0, atol=4e-16)
    np.testing.assert_allclose(init[6:9], Q0[0], rtol=0, atol=4e-16)
    x_tip = g["state0/position"][:, -1]
    np.testing.assert_allclose(init[3:6] * 1.0, x_tip, rtol=0, atol=1e-16)


def test_time_accumulation_and_truncation_index(golden_dir):
    g = np.load(os.path.join(golden_dir, "soft_pendulum_seed42_episode.npz"))
    t = np.float64(0.0)
    for i in range(int(g["n_steps"])):
        t = sp._advance_time(t, 1e-4, 400)
        assert t == g["time"][i]
        assert bool(t > 5.0) == bool(g["truncated"][i])
    assert 400 * 1e-4 * 125 == 5.0 and g["time"][124] < 5.0  # n*dt would truncate a step early


def test_box_sampling_matches_golden_actions(golden_dir):
    g = np.load(os.path.join(golden_dir, "softpendulum_v0_determinism_seed0.npz"))
    box = Box(np.ones(1) * -22, np.ones(1) * 22, shape=(1,), dtype=np.float32)
    box.seed(0)
    for a in g["actions"]:
        s = box.sample()
        assert s.dtype == np.float32 and np.array_equal(s, a)
    assert box.contains(np.array([3.0], dtype=np.float32)) and not box.contains(np.array([30.0], dtype=np.float32))


@pytest.mark.parametrize("n,world", [(4096, 1), (4096, 8), (65536, 8), (10, 3), (2, 4), (0, 2)])
def test_shard_envs_partitions_exactly(n, world):
    shards = [D.shard_envs(n, r, world) for r in range(world)]
    assert shards[0].start == 0 and shards[-1].stop == n
    for a, b in zip(shards, shards[1:]):
        assert a.stop == b.start
    counts = [s.count for s in shards]
    assert sum(counts) == n and max(counts) - min(counts) <= 1


def test_shard_envs_rejects_bad_rank():
    with pytest.raises(ValueError):
        D.shard_envs(8, 2, 2)


_GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, sys.argv[1])
import torch, torch.distributed as dist
from gym_softrobot_b200 import distributed as D
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
sh = D.shard_envs(10, rank, world)
st = D.EpisodeStats()
# every rank reports one episode per env it owns, return = global env index
for i in range(sh.start, sh.stop):
    st.add(episodes=1, return_sum=float(i), length_sum=126, nan_count=int(i == 3))
out = st.reduce()
assert out["episodes"] == 10 and out["return_sum"] == 45.0 and out["length_sum"] == 1260 and out["nan_count"] == 1, out
assert abs(out["mean_return"] - 4.5) < 1e-12
# ranks own disjoint contiguous ranges covering everything
lo = torch.tensor([sh.start, sh.stop]); all_lo = [torch.zeros(2, dtype=torch.long) for _ in range(world)]
dist.all_gather(all_lo, lo)
assert all_lo[0][0] == 0 and all_lo[-1][1] == 10 and all(all_lo[i][1] == all_lo[i+1][0] for i in range(world-1))
dist.destroy_process_group()
print("ok", rank)
"""


def test_two_rank_gloo_sharding_and_stats(tmp_path):
    """N>1 path on CPU: world_size 2, gloo, rendezvous on 127.0.0.1."""
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29517", str(script), ROOT]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env, timeout=180)
    assert res.returncode == 0, res.stdout[-2000:]
    assert res.stdout.count("ok") == 2


def test_curvature_interp_matrices_match_scipy_calls():
    """The actuation interpolation is linear in the action: the precomputed matrices must reproduce the
    reference's per-step scipy calls (arm_single_env.py:229-234, flat_env.py:296-309)."""
    from scipy.interpolate import interp1d
    from gym_softrobot_b200.envs.arm_single import curvature_interp_matrix
    from gym_softrobot_b200.envs.octo_flat import padded_curvature_interp_matrix
    rng = np.random.default_rng(0)
    a = rng.uniform(-22, 22, 7)
    ref = interp1d(np.linspace(0, 1, 7), a, kind="cubic", axis=-1)(np.linspace(0, 1, 49))
    np.testing.assert_allclose(curvature_interp_matrix(7, 49) @ a, ref, rtol=0, atol=1e-12)
    a = rng.uniform(-22, 22, (8, 3))
    padded = np.concatenate([np.zeros((8, 1)), a, np.zeros((8, 1))], axis=-1)
    ref = interp1d(np.linspace(0, 1, 5), padded, kind="cubic", axis=-1)(np.linspace(0, 1, 9))
    np.testing.assert_allclose(a @ padded_curvature_interp_matrix(3, 9).T, ref, rtol=0, atol=1e-12)


def test_count_crossings_known_answers():
    """Batched polyline-intersection count (utils/intersection.py semantics) on hand-made cases."""
    import torch
    from gym_softrobot_b200.envs.octo_flat import count_crossings
    t = np.linspace(0, 1, 11)
    line_a = np.stack([t, np.zeros_like(t)])                      # along x at y = 0
    cross1 = np.stack([np.full_like(t, 0.52), t - 0.47])         # vertical through it: 1 crossing (off-node:
    # a crossing exactly at a shared node counts for both segments, as in the reference's inclusive test)
    apart = np.stack([t, np.full_like(t, 0.7)])                  # parallel above: 0
    zig = np.stack([t, 0.3 * np.sin(3 * np.pi * t + 0.4)])       # sine: 3 crossings of y = 0 inside (0, 1)
    p1 = torch.as_tensor(np.stack([line_a, line_a, line_a]))
    p2 = torch.as_tensor(np.stack([cross1, apart, zig]))
    assert count_crossings(p1, p2).tolist() == [1, 0, 3]


def test_octopus_init_matches_reference_layout(golden_dir):
    """Arm start points / directions of build_octopus (build.py:66-90) vs the fixture's initial state."""
    from gym_softrobot_b200.envs.octo_flat import octopus_init_params
    g = np.load(os.path.join(golden_dir, "octo_flat_seed42.npz"))
    row = octopus_init_params(8)[0]
    for a in range(8):
        x0 = g[f"state0/arm{a}/position"][:, 0]
        np.testing.assert_allclose(row[9 * a:9 * a + 3], x0, rtol=0, atol=1e-15)
        d3 = g[f"state0/arm{a}/director"][2, :, 0]
        np.testing.assert_allclose(row[9 * a + 3:9 * a + 6], d3, rtol=0, atol=1e-15)
    np.testing.assert_allclose(g["state0/head/position"][:, 0], [0, 0, 0], atol=1e-18)


def test_snake_beta_spline_and_reward_match_reference(golden_dir):
    """ContinuumSnake-v0 host logic: the linearised B-spline amplitude equals the reference's
    `MuscleTorques.my_spline` (fixture, continuum_snake.py:186-198), and the batched
    `compute_projected_velocity` port reproduces every step's reward from the recorded callback
    samples (continuum_snake.py:40-101, 207-209)."""
    import torch
    from gym_softrobot_b200.envs.snake import beta_spline_matrix, projected_forward_velocity
    g = np.load(os.path.join(golden_dir, "continuum_snake_seed42.npz"))
    W = beta_spline_matrix(6, 50)
    for i in range(3):
        b = g["actions"][i, :6].astype(np.float64)
        np.testing.assert_allclose(W @ b, g[f"beta{i + 1}"], rtol=0, atol=1e-17)
    step_skip, every = int(g["step_skip"]), 2083
    com, vel, t = torch.as_tensor(g["cb_com"])[None], torch.as_tensor(g["cb_avg_velocity"])[None], g["cb_time"]
    assert np.array_equal(g["cb_step"], every * np.arange(len(t)))
    for i, r in enumerate(g["reward"]):
        S = 1 + (step_skip * (i + 1)) // every          # samples recorded when step i returns
        got = projected_forward_velocity(t[:S], com[:, :S], vel[:, :S], 2.0)[0].item()
        assert abs(got - float(r)) <= 1e-12 * max(1.0, abs(float(r))), (i, got, r)
    assert np.count_nonzero(g["reward"]) >= 2           # the non-trivial branch is exercised


def test_soft_arm_target_trajectory_matches_reference(golden_dir):
    """SoftArmTracking-v0 game_mode 2: the seeded target trajectory (`generate_trajectory`,
    soft_arm_tracking.py:44-98) bit for bit, from the env's own generator after reset(seed=42)."""
    from gym_softrobot_b200.envs.soft_arm_tracking import target_trajectory
    g = np.load(os.path.join(golden_dir, "soft_arm_tracking_mode2_seed42.npz"))
    rng = np.random.Generator(np.random.PCG64(np.random.SeedSequence(42)))
    w = target_trajectory(5.0, 2e-4, 0.1, rng)
    assert w.shape == (27500, 3)
    assert np.array_equal(w[::50], g["targets"])


def test_native_spline_basis_matches_scipy_not_a_knot():
    """sr_spline_basis (host-only entry point) == scipy's make_interp_spline(x, y) with zero end values,
    the spline MuscleTorquesWithVaryingBetaSplines fits (muscle_torques_with_bspline.py:146-148),
    including the extrapolation past base_length that a stretched arm evaluates."""
    from scipy.interpolate import make_interp_spline
    from gym_softrobot_b200 import _native as nat
    rng = np.random.default_rng(5)
    for P, L in ((4, 1000.0), (2, 0.35), (7, 3.0)):
        tab = nat.spline_basis(P, L)
        x = np.linspace(0, L, P + 2)
        y = np.zeros(P + 2); y[1:-1] = rng.uniform(-1, 1, P)
        ref = make_interp_spline(x, y)
        s = np.sort(rng.uniform(0, 1.08 * L, 300))
        m = np.clip(np.floor(s * (P + 1) / L).astype(int), 0, P)
        t = s - m * (L / (P + 1))
        val = sum(y[1 + i] * (tab[m, i, 0] + t * (tab[m, i, 1] + t * (tab[m, i, 2] + t * tab[m, i, 3])))
                  for i in range(P))
        assert np.abs(val - ref(s)).max() < 5e-15 * max(1.0, np.abs(ref(s)).max())


def test_arm_two_activation_map_and_layout_match_reference_fixture(golden_dir):
    """Host logic of OctoArmTwo-v0 without a GPU: the [n_elems, 3] matrix that replaces the three cubic `interp1d` calls of
    arm_two_env.py:236-247 reproduces the activations the unmodified reference env wrote into its muscles (fixture
    `muscle_activations`), the sucker locations follow arm_two_env.py:78-82, and the arm layout is build_two_arms'."""
    from gym_softrobot_b200.envs.arm_two import activation_interp_matrix, two_arm_init_params
    g = np.load(os.path.join(golden_dir, "octo_arm_two_seed42.npz"), allow_pickle=True)
    n = int(g["n_elems"])
    loc = [n // 6 * (2 * i + 1) for i in range(3)]
    assert loc == [int(v) for v in g["sucker_location"]]
    W = activation_interp_matrix([0] + loc + [n - 1], n)
    for i, a in enumerate(g["actions"]):
        a = a.reshape(2, 9)
        lm = a[:, 3:6] - np.float32(0.5)
        ctrl = np.stack([np.maximum(lm, 0).astype(np.float64), np.abs(np.minimum(lm, 0)).astype(np.float64),
                         a[:, 6:9].astype(np.float64)], axis=1)                      # [arm, muscle, control value]
        np.testing.assert_allclose(ctrl @ W.T, g["muscle_activations"][i], rtol=0, atol=1e-14)
    init, angles = two_arm_init_params()
    assert angles == [90.0, 270.0]
    for k in range(2):
        p0 = g[f"state0/arm{k}/position"]
        np.testing.assert_allclose(init[0, 9 * k:9 * k + 3], p0[:, 0], atol=1e-17)
        np.testing.assert_allclose(init[0, 9 * k + 3:9 * k + 6], (p0[:, -1] - p0[:, 0]) / 0.25, atol=1e-15)


def test_longitudinal_muscle_positions_match_the_layer_set():
    """create_es_muscle_layers (envs/octopus/build.py:303-328): ratio_muscle_position (0, -6/9, 0) rotated by
    muscle_init_angle = +-pi/2 puts the two longitudinal muscles at +-2/3 r on d1."""
    from gym_softrobot_b200.envs.octo_reach import es_longitudinal_positions
    (px0, py0), (px1, py1) = es_longitudinal_positions()
    assert abs(px0 - 2 / 3) < 1e-15 and abs(px1 + 2 / 3) < 1e-15 and abs(py0) < 1e-16 and abs(py1) < 1e-16
