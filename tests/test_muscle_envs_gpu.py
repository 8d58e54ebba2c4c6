"""GPU parity of the COOMM-driven envs: OctoArmPush-v0 / -v1, OctoArmPullWeight-v0,
OctoCrawl-v0 — CUDA path through the C-ABI vs (a) fixtures produced by the UNMODIFIED reference env code on the oracle
shims (oracle/shims/elastica + oracle/shims/coomm, oracle/gen_golden.py coomm) and (b) the C oracle.

Parity is UNPINNED twice over here: PyElastica and COOMM (coomm 0.1.1 @ d33fa034, /root/reference/uv.lock:172-179)
are both third-party packages outside the reference tree, restated from their published algorithms.

Tolerance rule (as tests/test_parity_gpu.py): max|mine - ref| <= 1e-9 * max(|ref|, field floor) per field; the one
way out is a MEASURED conditioning floor — 20 x the divergence of a C-oracle replica started 1e-13 (relative) away.
Observations are float32 casts; rewards are differences of FP64 norms of O(0.2 m) centre-of-mass positions and are
compared to 1e-9 of that scale.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-9
FIELDS = {"position": "position_collection", "velocity": "velocity_collection", "director": "director_collection",
          "omega": "omega_collection", "kappa": "kappa", "sigma": "sigma"}
# a field is compared relative to max|ref|, but not below the magnitude it has once the arm is actuated (the straight
# arm of the push envs keeps kappa / omega at exactly zero)
FLOOR = dict(position_collection=1e-2, velocity_collection=1e-3, director_collection=1.0, omega_collection=1e-2,
             kappa=1.0, sigma=1e-3)
HEAD = ((slice(0, 3), "position", "position_collection"), (slice(3, 6), "velocity", "velocity_collection"),
        (slice(6, 15), "director", "director_collection"), (slice(15, 18), "omega", "omega_collection"))


def _err(mine, ref, key):
    return float(np.abs(np.asarray(mine).reshape(np.shape(ref)) - ref).max() / max(np.abs(ref).max(), FLOOR[key]))


def _push_oracle(g, n, pull_weight, perturb=0.0, seed=0):
    """The arm of arm_push_env.py:156-215 (PullWeight: :520-620) on the C oracle, stepped through the fixture's actions;
    yields the oracle objects after every env-step."""
    import rod_oracle as ro
    dt = float(g["time_step"])
    arm = dict(n_elem=n, start=(0, 0, 0), direction=(1, 0, 0), normal=(0, 1, 0), base_length=0.2, base_radius=0.012,
               density=700.0, youngs_modulus=1e4, shear_modulus=1e4 / 1.5,
               damping_constant=0.05 * 2 * (5e2 if pull_weight else 1e2), tip_radius=0.001, taper_node_mean=True)
    if pull_weight:
        head = dict(start=(-0.015 * 0.9, 0, -0.024), direction=(0, 0, 1), normal=(0, 1, 0), length=0.024, radius=0.015,
                    density=700.0)
        asm = ro.OracleAssembly([arm], dt, head=head, joint=dict(k=1e6, nu=1e-2, kt=1e0, radius=0.015), angles_deg=[0.0])
        rod = asm.arms[0]
    else:
        asm, rod = None, ro.OracleRod(dt=dt, **arm)
    if perturb:
        rng = np.random.default_rng(seed)
        rod.position_collection[...] *= 1.0 + perturb * rng.standard_normal(rod.position_collection.shape)
    rod.set_tm_muscle(1.0, 0.012)
    for a in g["actions"]:
        if np.ndim(a) == 0:
            idx, act = (0, 0.5) if int(a) == 0 else (-1, 0.0)
        else:
            idx, act = int(np.clip(a[0] * n, 0, n - 1)), float(a[1])
        rod.set_sucker(0, idx, 0.9 if pull_weight else 1.0)
        rod.set_tm_activation(act)
        (asm or rod).substeps(int(g["step_skip"]))
        yield rod, asm


@pytest.mark.parametrize("env_id,tag", [("OctoArmPush-v0", "octo_arm_push_v0"), ("OctoArmPush-v1", "octo_arm_push_v1"),
                                        ("OctoArmPullWeight-v0", "octo_arm_pull_weight")])
def test_arm_push_env_golden(golden_dir, env_id, tag):
    """The three single-arm muscle envs through the Gymnasium facade vs the reference-env-on-shims fixture AND the C
    oracle: state (every field), observation, reward, flags, time; 500 (PullWeight: 1000) substeps per env-step, the
    stretch reaches 0.5 and the sucker hops between node 0, the last node (index -1) and interior nodes."""
    import gym_softrobot_b200 as gsb
    g = np.load(os.path.join(golden_dir, f"{tag}_seed42.npz"), allow_pickle=True)
    pull = env_id == "OctoArmPullWeight-v0"
    env = gsb.make(env_id)
    assert env.step_skip == int(g["step_skip"]) and env.time_step == float(g["time_step"])
    obs0, _ = env.reset(seed=42)
    assert obs0.dtype == np.float32 and obs0.shape == g["obs0"].shape
    np.testing.assert_allclose(obs0, g["obs0"], rtol=1e-6, atol=1e-7)
    oracle = _push_oracle(g, 40, pull)
    replica = _push_oracle(g, 40, pull, perturb=1e-13, seed=3)
    worst = 0.0
    for i, a in enumerate(g["actions"]):
        obs, r, te, tr, info = env.step(a)
        st = env.rod_state()
        (orod, oasm), (rrod, rasm) = next(oracle), next(replica)
        for gk, fk in FIELDS.items():
            ref = g[f"state{i + 1}/{gk}"]
            sens = float(np.abs(getattr(rrod, fk) - getattr(orod, fk)).max())
            scale = max(float(np.abs(ref).max()), FLOOR[fk])
            for what, r_ in (("fixture", ref), ("C oracle", getattr(orod, fk))):
                e = float(np.abs(st[fk] - r_).max())
                worst = max(worst, e / scale)
                assert e <= max(TOL * scale, 20 * sens), \
                    f"{env_id} step {i} {gk} vs {what}: {e / scale:.3e} (replica 1e-13 away: {sens / scale:.1e})"
        if pull:
            hd = env.head_state()
            for sl, gk, fk in HEAD:
                e = _err(hd[sl], g[f"state{i + 1}/head/{gk}"].reshape(-1), fk)
                worst = max(worst, e)
                assert e < TOL, f"step {i} head {gk}: {e:.3e}"
        np.testing.assert_allclose(obs, g["obs"][i], rtol=2e-6, atol=2e-7)
        assert abs(r - float(g["reward"][i])) < 1e-9 * 0.2, f"reward {r} vs {float(g['reward'][i])}"
        assert (te, tr) == (bool(g["terminated"][i]), bool(g["truncated"][i]))
        assert info["time"] == float(g["time"][i])          # bit-exact: accumulated as 2K additions of dt/2
    print(f"{env_id}: worst field error {worst:.2e}")
    env.close()


def _crawl_oracle(g, perturb=0.0, seed=0, variant=None):
    import rod_oracle as ro
    from gym_softrobot_b200.envs.octo_crawl import crawl_init_params
    n, dt = int(g["n_elems"]), float(g["time_step"])
    init, angles = crawl_init_params()
    hr, r0 = 0.04, 0.013
    arms = []
    for a in range(8):
        s0 = init[0, 9 * a:9 * a + 9]
        arms.append(dict(n_elem=n, start=s0[0:3], direction=s0[3:6], normal=s0[6:9], base_length=0.25, base_radius=r0,
                         density=1000.0, youngs_modulus=1.5e4, shear_modulus=1.5e4 / 1.5,
                         damping_constant=0.2 * 1e-2 * (7e-5 / dt), tip_radius=0.0042))
    head = dict(start=(0, 0, -2 * r0), direction=(0, 0, 1), normal=(0, 1, 0), length=2 * r0, radius=hr, density=50.0)
    mk = lambda: ro.OracleAssembly(arms, dt, head=head, joint=dict(k=1e6, nu=1e-3, kt=1e2, radius=hr), angles_deg=angles)
    if variant:
        with ro.variant(variant):
            asm = mk()
    else:
        asm = mk()
    rng = np.random.default_rng(seed)
    for rod in asm.arms:
        rod.set_tm_muscle(1.0, r0)
        if perturb:
            rod.position_collection[...] *= 1.0 + perturb * rng.standard_normal(rod.position_collection.shape)
    for a in g["actions"]:
        a = a.reshape(8, 3)
        for k, rod in enumerate(asm.arms):
            rod.set_sucker(0, int(np.clip(a[k, 0] * n, 0, n - 1)), float(a[k, 2]))
            rod.set_tm_activation(float(a[k, 1]))
        asm.substeps(int(g["step_skip"]))
        yield asm


def test_octo_crawl_env_golden(golden_dir):
    """OctoCrawl-v0 through the Gymnasium facade vs the reference-env-on-shims fixture (3 env-steps x 800 substeps:
    eight tapered arms, light head, joints, per-arm sucker index / ratio and transverse-muscle activation from the
    action).  Bound per field: max(1e-9 of scale, 20 x the divergence of C-oracle replicas started 1e-13 away / built
    with FMA contraction) — the light head on kt = 1e2 amplifies round-off (tests/test_parity_gpu.py:
    test_tapered_muscle_octopus_topology_vs_c_oracle measured kt dt / I ~ 3e3 per substep)."""
    import gym_softrobot_b200 as gsb
    g = np.load(os.path.join(golden_dir, "octo_crawl_seed42.npz"), allow_pickle=True)
    env = gsb.make("OctoCrawl-v0")
    assert env.step_skip == int(g["step_skip"]) and env.n_elems == int(g["n_elems"])
    obs0, _ = env.reset(seed=42)
    assert obs0.dtype == np.float32 and obs0.shape == g["obs0"].shape
    np.testing.assert_allclose(obs0, g["obs0"], rtol=1e-6, atol=1e-7)
    oracle = _crawl_oracle(g)
    reps = [_crawl_oracle(g, perturb=1e-13, seed=5), _crawl_oracle(g, variant="fma")]
    worst, worst_ratio = 0.0, 0.0
    for i, a in enumerate(g["actions"]):
        obs, r, te, tr, info = env.step(a)
        st, hd = env.arm_states(), env.head_state()
        oasm, rasm = next(oracle), [next(r_) for r_ in reps]
        for arm in range(8):
            for gk, fk in FIELDS.items():
                ref = g[f"state{i + 1}/arm{arm}/{gk}"]
                oref = getattr(oasm.arms[arm], fk)
                sens = max(float(np.abs(getattr(q.arms[arm], fk) - oref).max()) for q in rasm)
                scale = max(float(np.abs(ref).max()), FLOOR[fk])
                e = float(np.abs(st[fk][arm] - ref).max())
                worst, worst_ratio = max(worst, e / scale), max(worst_ratio, e / max(TOL * scale, 20 * sens))
                assert e <= max(TOL * scale, 20 * sens), \
                    f"step {i} arm {arm} {gk}: {e / scale:.3e} (oracle replicas: {sens / scale:.1e})"
        mine = {"position": hd[0:3], "velocity": hd[3:6], "director": hd[6:15], "omega": hd[15:18]}
        for (sl, gk, fk) in HEAD:
            ref = g[f"state{i + 1}/head/{gk}"].reshape(-1)
            oref = getattr(oasm, "head_" + gk).reshape(-1)
            sens = max(float(np.abs(getattr(q, "head_" + gk).reshape(-1) - oref).max()) for q in rasm)
            scale = max(float(np.abs(ref).max()), FLOOR[fk])
            e = float(np.abs(mine[gk] - ref).max())
            worst, worst_ratio = max(worst, e / scale), max(worst_ratio, e / max(TOL * scale, 20 * sens))
            assert e <= max(TOL * scale, 20 * sens), f"step {i} head {gk}: {e / scale:.3e} (oracle replicas: {sens / scale:.1e})"
        np.testing.assert_allclose(obs, g["obs"][i], rtol=1e-4, atol=1e-5)
        # reward = 1e2 x a difference of distances to the target 5 m away: 1e-9 of that scale
        assert abs(r - float(g["reward"][i])) < 1e-9 * 5e2, f"reward {r} vs {float(g['reward'][i])}"
        assert (te, tr) == (bool(g["terminated"][i]), bool(g["truncated"][i]))
    print(f"OctoCrawl-v0: worst field error {worst:.2e}, worst error / bound {worst_ratio:.2f}")
    env.close()


def test_transverse_muscle_on_a_bent_arm_vs_c_oracle():
    """The in-kernel transverse muscle away from the straight configuration: a tapered free arm bent and twisted by
    random external couples (so that directors, shear and the muscle's round-off couple all matter), random
    activations, sucker indices (interior, 0, -1) and ratios per env, 3 x 300 substeps, CUDA vs the C oracle at 1e-9
    on every field of every env."""
    import torch
    import rod_oracle as ro
    from gym_softrobot_b200 import _native as nat
    rng = np.random.default_rng(11)
    n_env, n, L, r_base, r_tip, E, rho, dt = 6, 24, 0.2, 0.012, 0.002, 2e4, 900.0, 2e-5
    kw = dict(base_length=L, base_radius=r_base, density=rho, youngs_modulus=E, shear_modulus=E / 1.5)
    h = nat.Handle(model=nat.MODEL_ROD, n_env=n_env, n_elem=n, dt=dt, gravity=(0.0, 0.0, 0.0), damping_constant=2.0,
                   bc_kind=nat.BC_FREE, tip_radius=r_tip, taper_node_mean=True, sucker_index=0,
                   tm_muscle=dict(max_stress=1.5, radius_ref=r_base), **kw)
    init = np.tile(np.array([[0.0, 0.0, 0.0, 0.6, 0.0, 0.8, 0.0, 1.0, 0.0]]), (n_env, 1))
    h.reset(torch.as_tensor(init, device="cuda").contiguous())
    s_e = (np.arange(n) + 0.5) / n
    taper = (1 + (r_tip / r_base - 1) * s_e) ** 3
    rods = []
    f_t, c_t = h.ext_load_tensors()
    for e in range(n_env):
        rod = ro.OracleRod(n, init[e, 0:3], init[e, 3:6], init[e, 6:9], L, r_base, rho, E, dt, shear_modulus=E / 1.5,
                           damping_constant=2.0, tip_radius=r_tip, taper_node_mean=True)
        rod.set_tm_muscle(1.5, r_base)
        cext = 3e-5 * taper * np.stack([rng.uniform(-1, 1) * np.sin(np.pi * s_e), rng.uniform(-1, 1) * np.sin(np.pi * s_e),
                                        0.3 * rng.uniform(-1, 1) * s_e])
        rod.user_torques[...] = cext
        c_t[e] = torch.as_tensor(cext, device="cuda")
        rods.append(rod)
    o6 = torch.empty((n_env, 6), dtype=torch.float32, device="cuda")
    rew = torch.empty(n_env, dtype=torch.float64, device="cuda")
    term = torch.empty(n_env, dtype=torch.uint8, device="cuda")
    worst = 0.0
    for chunk in range(3):
        idx = np.array([rng.integers(1, n - 1), 0, -1, rng.integers(1, n - 1), n - 1, -1])
        ratio = rng.uniform(0.2, 1.0, n_env)
        act = rng.uniform(0.0, 1.0, n_env)
        act[3] = 0.0
        h.sucker_index_tensor()[:] = torch.as_tensor(idx, dtype=torch.int32, device="cuda")
        h.sucker_tensor()[:] = torch.as_tensor(ratio, device="cuda")
        h.tm_activation_tensor()[:] = torch.as_tensor(act, device="cuda")
        h.step(None, 300, o6, rew, term)
        assert int(term.sum()) == 0
        f = {k: v.cpu().numpy() for k, v in h.fields().items()}
        for e, rod in enumerate(rods):
            rod.set_sucker(0, int(idx[e]), float(ratio[e]))
            rod.set_tm_activation(float(act[e]))
            rod.substeps(300)
            for gk, fk in FIELDS.items():
                err = _err(f[fk][e], getattr(rod, fk), fk)
                worst = max(worst, err)
                assert err < TOL, f"chunk {chunk} env {e} {gk}: {err:.3e}"
    # the muscle did something: the activated arms are longer than the passive one
    lengths = np.linalg.norm(np.diff(f["position_collection"], axis=2), axis=1).sum(axis=1)
    assert lengths[0] > lengths[3] * 1.005
    print(f"transverse muscle on a bent arm: worst field error {worst:.2e}")
    for rod in rods:
        rod.close()
    h.close()


def test_arm_push_early_termination_mode(golden_dir):
    """config_early_termination=True (arm_push_env.py:309-312,436-452) vs the reference env on the shims: reward -10
    whatever happens, terminated = truncated = (kinetic + shear + bending energy < 1e-7 J); the energy itself to 1e-9."""
    import gym_softrobot_b200 as gsb
    g = np.load(os.path.join(golden_dir, "octo_arm_push_early_seed1.npz"), allow_pickle=True)
    env = gsb.make("OctoArmPush-v1", config_early_termination=True)
    env.reset(seed=1)
    for i, a in enumerate(g["actions"]):
        obs, r, te, tr, info = env.step(a)
        assert r == float(g["reward"][i]) == -10.0
        assert (te, tr) == (bool(g["terminated"][i]), bool(g["truncated"][i]))
        ham, ref = env.cal_desired_Hamiltonian(), float(g["hamiltonian"][i])
        assert abs(ham - ref) <= 1e-9 * max(ref, 1e-3), (i, ham, ref)
    assert bool(g["terminated"][0]) and not bool(g["terminated"][1])
    env.close()


@pytest.mark.parametrize("env_id,n_act", [("OctoArmPush-v1", 2), ("OctoArmPullWeight-v0", 2), ("OctoCrawl-v0", 24)])
def test_muscle_vector_envs_are_batch_independent_and_autoreset(env_id, n_act):
    """An env's bits do not depend on the batch it runs in (1 env vs the same env among 7 others with other actions),
    and a finished env is rebuilt inside the same step() (fresh muscles / suckers) while the others keep stepping."""
    import torch
    import gym_softrobot_b200 as gsb
    kw = dict(final_time=0.06) if env_id != "OctoCrawl-v0" else dict(final_time=0.09)
    gen = torch.Generator().manual_seed(3)
    acts = torch.rand((4, 8, n_act), generator=gen).cuda()
    solo, batch = gsb.make_vec(env_id, 1, **kw), gsb.make_vec(env_id, 8, **kw)
    solo.reset(seed=0); batch.reset(seed=0)
    first = int((solo._time_table > kw["final_time"]).nonzero()[0][0]) if not hasattr(solo, "_first_truncated") else solo._first_truncated
    for s in range(4):
        o1, r1, te1, tr1, i1 = solo.step(acts[s, 5:6])
        o8, r8, te8, tr8, i8 = batch.step(acts[s])
        assert torch.equal(o1[0], o8[5]) and torch.equal(r1[0], r8[5]) and bool(te1[0]) == bool(te8[5])
        assert bool(tr8.all()) == (s + 1 == first), (s, first)
        if s + 1 == first:
            assert "final_obs" in i8 and int(batch.step_count.max()) == 0
            assert float(batch.handle.tm_activation_tensor().abs().max()) == 0.0      # fresh muscles
            assert int(batch.handle.sucker_index_tensor().abs().max()) == 0           # fresh SuckerController(index=0)
    assert torch.isfinite(o8).all()
    solo.close(); batch.close()
