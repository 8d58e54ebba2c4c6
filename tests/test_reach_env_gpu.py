"""GPU parity of OctoReach-v0 (second half): the muscle-layer kernel — two off-axis
longitudinal muscles + the transverse muscle with per-element activations, rigid head pinned by OneEndFixedBC —
through the C-ABI vs (a) the fixture the UNMODIFIED reference ReachEnv produced on the oracle shims
(oracle/gen_golden.py reach) and (b) the C oracle (oracle/rod_oracle.c:apply_muscle_layers).

Parity is UNPINNED (PyElastica and COOMM are third-party packages outside the reference tree, restated from their
published algorithms; see tests/test_muscle_envs_gpu.py).  Tolerance: max|mine - ref| <= 1e-9 * max(|ref|, field
floor) per field, or 20 x the measured divergence of C-oracle replicas (started 1e-13 away / FMA-contracted build).
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-9
FIELDS = {"position": "position_collection", "velocity": "velocity_collection", "director": "director_collection",
          "omega": "omega_collection", "kappa": "kappa", "sigma": "sigma"}
FLOOR = dict(position_collection=1e-2, velocity_collection=1e-3, director_collection=1.0, omega_collection=1e-2,
             kappa=1.0, sigma=1e-3)


def _reach_oracle(n, dt, actions, step_skip, perturb=0.0, seed=0, variant=None):
    import rod_oracle as ro
    from gym_softrobot_b200.envs.octo_crawl import crawl_init_params
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
    asm.set_head_fixed(True)
    acts = [rod.set_es_muscle_layers(r0) for rod in asm.arms]
    rng = np.random.default_rng(seed)
    if perturb:
        for rod in asm.arms:
            rod.position_collection[...] *= 1.0 + perturb * rng.standard_normal(rod.position_collection.shape)
    for a in actions:
        a = np.asarray(a, dtype=np.float32).reshape(8, 3, n).astype(np.float64)
        for k in range(8):
            acts[k][...] = a[k]
        asm.substeps(int(step_skip))
        yield asm


def _check_against(env, oasm, rasm, refs, tag):
    """refs(arm, golden key, field) -> reference array; returns (worst relative error, worst error / bound)."""
    st = env.arm_states()
    worst, worst_ratio = 0.0, 0.0
    for arm in range(8):
        for gk, fk in FIELDS.items():
            ref = refs(arm, gk, fk)
            oref = getattr(oasm.arms[arm], fk)
            sens = max(float(np.abs(getattr(q.arms[arm], fk) - oref).max()) for q in rasm)
            scale = max(float(np.abs(ref).max()), FLOOR[fk])
            e = float(np.abs(st[fk][arm] - ref).max())
            worst, worst_ratio = max(worst, e / scale), max(worst_ratio, e / max(TOL * scale, 20 * sens))
            assert e <= max(TOL * scale, 20 * sens), \
                f"{tag} arm {arm} {gk}: {e / scale:.3e} (oracle replicas: {sens / scale:.1e})"
    return worst, worst_ratio


def test_octo_reach_env_golden(golden_dir):
    """OctoReach-v0 through the Gymnasium facade vs the reference-env-on-shims fixture (2 env-steps x 800 substeps,
    random activations in [0, 1] on all 8 x 3 x 20 muscle elements); the head stays exactly at its reset pose."""
    import gym_softrobot_b200 as gsb
    g = np.load(os.path.join(golden_dir, "octo_reach_seed42.npz"), allow_pickle=True)
    env = gsb.make("OctoReach-v0")
    n, dt, skip = int(g["n_elems"]), float(g["time_step"]), int(g["step_skip"])
    assert env.step_skip == skip and env.n_elems == n
    obs0, _ = env.reset(seed=42)
    np.testing.assert_allclose(env._target, g["target"], rtol=0, atol=1e-15)   # same np_random stream, same rest lengths
    assert obs0.dtype == np.float32 and obs0.shape == g["obs0"].shape
    np.testing.assert_allclose(obs0, g["obs0"], rtol=1e-6, atol=1e-7)
    head0 = env.head_state().copy()
    oracle = _reach_oracle(n, dt, g["actions"], skip)
    reps = [_reach_oracle(n, dt, g["actions"], skip, perturb=1e-13, seed=5), _reach_oracle(n, dt, g["actions"], skip, variant="fma")]
    worst, worst_ratio = 0.0, 0.0
    for i, a in enumerate(g["actions"]):
        obs, r, te, tr, info = env.step(a)
        oasm, rasm = next(oracle), [next(r_) for r_ in reps]
        w, wr = _check_against(env, oasm, rasm, lambda arm, gk, fk: g[f"state{i + 1}/arm{arm}/{gk}"], f"step {i}")
        worst, worst_ratio = max(worst, w), max(worst_ratio, wr)
        hd = env.head_state()
        np.testing.assert_array_equal(hd[0:18], head0[0:18])                     # pinned: bit for bit
        np.testing.assert_array_equal(hd[0:3], g[f"state{i + 1}/head/position"].reshape(-1))
        np.testing.assert_array_equal(hd[6:15], g[f"state{i + 1}/head/director"].reshape(-1))
        np.testing.assert_allclose(obs, g["obs"][i], rtol=1e-4, atol=1e-5)
        assert abs(r - float(g["reward"][i])) < 1e-9 * max(1.0, abs(float(g["reward"][i]))), f"reward {r} vs {float(g['reward'][i])}"
        assert (te, tr) == (bool(g["terminated"][i]), bool(g["truncated"][i]))
    assert float(np.abs(env.arm_states()["kappa"]).max()) > 1.0                  # the longitudinal muscles bent the arms
    print(f"OctoReach-v0: worst field error {worst:.2e}, worst error / bound {worst_ratio:.2f}")
    env.close()


def test_octo_reach_one_sided_contraction_vs_c_oracle():
    """Longitudinal muscle 1 alone, fully activated with a ramp along the arm (a strong one-sided bend, plus the
    transverse muscle on the distal half), 3 env-steps live against the C oracle — a deterministic load away from the
    fixture's random activations."""
    import gym_softrobot_b200 as gsb
    env = gsb.make("OctoReach-v0")
    n, dt, skip = env.n_elems, env.time_step, env.step_skip
    env.reset(seed=3)
    a = np.zeros((8, 3, n), dtype=np.float32)
    a[:, 0, :] = np.linspace(1.0, 0.2, n)[None, :]
    a[::2, 0, :] = 0.0
    a[::2, 1, :] = np.linspace(0.3, 1.0, n)[None, :]          # every other arm bends the other way
    a[:, 2, n // 2:] = 0.7
    actions = [a.reshape(-1)] * 3
    oracle = _reach_oracle(n, dt, actions, skip)
    reps = [_reach_oracle(n, dt, actions, skip, perturb=1e-13, seed=7), _reach_oracle(n, dt, actions, skip, variant="fma")]
    worst, worst_ratio = 0.0, 0.0
    for i, act in enumerate(actions):
        env.step(act)
        oasm, rasm = next(oracle), [next(r_) for r_ in reps]
        w, wr = _check_against(env, oasm, rasm, lambda arm, gk, fk: getattr(oasm.arms[arm], fk), f"step {i}")
        worst, worst_ratio = max(worst, w), max(worst_ratio, wr)
    k = env.arm_states()["kappa"]
    assert float(np.abs(k[1]).max()) > 2.0 and float(np.abs(k[0]).max()) > 2.0
    print(f"one-sided contraction: worst field error {worst:.2e}, worst error / bound {worst_ratio:.2f}, max |kappa| {np.abs(k).max():.1f}")
    env.close()


def test_octo_reach_vector_env_batch_independent_and_autoreset():
    """Batched OctoReach-v0: an env's bits do not depend on its batch mates (same action alone / in a batch of 5 with
    other actions), spaces and autoreset plumbing."""
    import torch
    import gym_softrobot_b200 as gsb
    rng = np.random.default_rng(11)
    vec = gsb.make_vec("OctoReach-v0", 5, final_time=0.075)
    one = gsb.make_vec("OctoReach-v0", 1, final_time=0.075)
    tgt = rng.random((5, 3)) * 0.25
    o5, _ = vec.reset(target=tgt)
    o1, _ = one.reset(target=tgt[3:4])
    assert tuple(o5.shape) == (5, vec.single_observation_space.shape[0]) and o5.dtype == torch.float32
    assert torch.equal(o5[3], o1[0])
    for step in range(2):
        a = torch.as_tensor(rng.random((5, 8 * 3 * vec.n_elems)).astype(np.float32), device=vec.device)
        o5, r5, te5, tr5, info5 = vec.step(a)
        o1, r1, te1, tr1, info1 = one.step(a[3:4])
        if step == 0:
            assert torch.equal(o5[3], o1[0]) and torch.equal(r5[3], r1[0])
            assert not bool(tr5.any()) and bool(torch.isfinite(o5).all())
    # second step passed final_time = 0.075 s (2 x 0.04 s): truncated, autoreset zeroes the activations and re-draws targets
    assert bool(tr5.all()) and "final_obs" in info5
    assert float(vec.handle.muscle_activation_tensor().abs().max()) == 0.0
    assert int(vec.step_count.max()) == 0
    vec.close(); one.close()


HEAD = ((slice(0, 3), "position", "position_collection"), (slice(3, 6), "velocity", "velocity_collection"),
        (slice(6, 15), "director", "director_collection"), (slice(15, 18), "omega", "omega_collection"))


def _arm_two_oracle(g, perturb=0.0, seed=0, variant=None):
    import rod_oracle as ro
    from gym_softrobot_b200.envs.arm_two import two_arm_init_params
    n, dt = int(g["n_elems"]), float(g["time_step"])
    init, angles = two_arm_init_params()
    hr, r0 = 0.04, 0.013
    arms = []
    for a in range(2):
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
    acts = [rod.set_es_muscle_layers(r0) for rod in asm.arms]
    rng = np.random.default_rng(seed)
    if perturb:
        for rod in asm.arms:
            rod.position_collection[...] *= 1.0 + perturb * rng.standard_normal(rod.position_collection.shape)
    loc = [int(v) for v in g["sucker_location"]]
    for i, a in enumerate(g["actions"]):
        a = a.reshape(2, 9)
        for k, rod in enumerate(asm.arms):
            for s_ in range(3):
                rod.set_sucker(s_, loc[s_], float(a[k, s_]))
            acts[k][...] = g["muscle_activations"][i, k]
        asm.substeps(int(g["step_skip"]))
        yield asm


def test_octo_arm_two_env_golden(golden_dir):
    """OctoArmTwo-v0 through the Gymnasium facade vs the reference-env-on-shims fixture (3 env-steps x 800 substeps: two
    tapered arms on a free light head, three fixed-index suckers per arm, cubic-interpolated activations of all three
    muscles).  Bound per field as in test_octo_crawl_env_golden: max(1e-9 of scale, 20 x the divergence of C-oracle
    replicas) — the light head on kt = 1e2 amplifies round-off."""
    import torch
    import gym_softrobot_b200 as gsb
    g = np.load(os.path.join(golden_dir, "octo_arm_two_seed42.npz"), allow_pickle=True)
    env = gsb.make("OctoArmTwo-v0")
    assert env.step_skip == int(g["step_skip"]) and env.n_elems == int(g["n_elems"])
    assert list(env.sucker_location) == [int(v) for v in g["sucker_location"]]
    obs0, _ = env.reset(seed=42)
    assert obs0.dtype == np.float32 and obs0.shape == g["obs0"].shape
    np.testing.assert_allclose(obs0, g["obs0"], rtol=1e-6, atol=1e-7)
    oracle = _arm_two_oracle(g)
    reps = [_arm_two_oracle(g, perturb=1e-13, seed=5), _arm_two_oracle(g, variant="fma")]
    worst, worst_ratio = 0.0, 0.0
    for i, a in enumerate(g["actions"]):
        # the host-side activation map (float32 action -> clamp -> cubic interpolation matrix) vs scipy's interp1d
        mine_act = env._vec.muscle_activations(torch.as_tensor(a.reshape(1, -1), device=env._vec.device))[0].cpu().numpy()
        np.testing.assert_allclose(mine_act, g["muscle_activations"][i], rtol=0, atol=1e-14)
        obs, r, te, tr, info = env.step(a)
        st, hd = env.arm_states(), env.head_state()
        oasm, rasm = next(oracle), [next(r_) for r_ in reps]
        for arm in range(2):
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
    print(f"OctoArmTwo-v0: worst field error {worst:.2e}, worst error / bound {worst_ratio:.2f}")
    env.close()


@pytest.mark.parametrize("n_elems,time_step,fps", [(11, 5e-5, 100), (40, 2e-5, 250)], ids=["n11-three-envs-per-cta", "n40-one-env-per-cta"])
def test_octo_reach_other_arm_resolutions_vs_c_oracle(n_elems, time_step, fps):
    """The muscle-layer kernel away from the default 20 elements per arm: 11 (three env groups per 384-thread CTA, warps
    shared between groups) and 40 (one group of 329 threads per CTA: CTA-wide barriers), random per-element activations,
    live against the C oracle.  Two envs with different actions, so that a mix-up between env groups would show."""
    import torch
    import gym_softrobot_b200 as gsb
    rng = np.random.default_rng(n_elems)
    vec = gsb.make_vec("OctoReach-v0", 2, n_elems=n_elems, time_step=time_step, recording_fps=fps, autoreset=False)
    vec.reset(seed=1)
    skip = vec.step_skip
    acts = rng.random((2, 2, 8 * 3 * n_elems)).astype(np.float32)          # [step, env, action]
    worst = 0.0
    for env_i in range(2):
        seq = [acts[s, env_i] for s in range(2)]
        oracle = _reach_oracle(n_elems, time_step, seq, skip)
        reps = [_reach_oracle(n_elems, time_step, seq, skip, perturb=1e-13, seed=5), _reach_oracle(n_elems, time_step, seq, skip, variant="fma")]
        if env_i == 0:
            states = []
            for s in range(2):
                vec.step(torch.as_tensor(acts[s], device=vec.device))
                states.append({k: v.cpu().numpy() for k, v in vec.fields().items()})
        for s in range(2):
            oasm, rasm = next(oracle), [next(r_) for r_ in reps]
            for arm in range(8):
                for gk, fk in FIELDS.items():
                    ref = getattr(oasm.arms[arm], fk)
                    sens = max(float(np.abs(getattr(q.arms[arm], fk) - ref).max()) for q in rasm)
                    scale = max(float(np.abs(ref).max()), FLOOR[fk])
                    e = float(np.abs(states[s][fk][env_i, arm] - ref).max())
                    worst = max(worst, e / scale)
                    assert e <= max(TOL * scale, 20 * sens), \
                        f"env {env_i} step {s} arm {arm} {gk}: {e / scale:.3e} (oracle replicas: {sens / scale:.1e})"
    print(f"OctoReach-v0 n_elems={n_elems}: worst field error {worst:.2e}")
    vec.close()


def test_octo_arm_two_other_resolution_vs_c_oracle():
    """OctoArmTwo-v0 at 30 elements per arm (suckers at elements 5 / 15 / 25; 384 // 63 = 6 env groups per CTA), three envs
    with different actions, 2 env-steps live against the C oracle (activations taken from the env's own host-side map,
    which test_octo_arm_two_env_golden checks against scipy)."""
    import torch
    import rod_oracle as ro
    import gym_softrobot_b200 as gsb
    from gym_softrobot_b200.envs.arm_two import two_arm_init_params
    n, dt, fps = 30, 3e-5, 200
    rng = np.random.default_rng(30)
    vec = gsb.make_vec("OctoArmTwo-v0", 3, n_elems=n, time_step=dt, recording_fps=fps, autoreset=False)
    assert vec.sucker_location == [5, 15, 25]
    vec.reset()
    acts = rng.random((2, 3, 18)).astype(np.float32)
    init, angles = two_arm_init_params()
    hr, r0 = 0.04, 0.013
    arms = [dict(n_elem=n, start=init[0, 9 * a:9 * a + 3], direction=init[0, 9 * a + 3:9 * a + 6], normal=init[0, 9 * a + 6:9 * a + 9],
                 base_length=0.25, base_radius=r0, density=1000.0, youngs_modulus=1.5e4, shear_modulus=1.5e4 / 1.5,
                 damping_constant=0.2 * 1e-2 * (7e-5 / dt), tip_radius=0.0042) for a in range(2)]
    head = dict(start=(0, 0, -2 * r0), direction=(0, 0, 1), normal=(0, 1, 0), length=2 * r0, radius=hr, density=50.0)

    def make(perturb=0.0, variant=None):
        mk = lambda: ro.OracleAssembly(arms, dt, head=head, joint=dict(k=1e6, nu=1e-3, kt=1e2, radius=hr), angles_deg=angles)
        if variant:
            with ro.variant(variant):
                asm = mk()
        else:
            asm = mk()
        views = [rod.set_es_muscle_layers(r0) for rod in asm.arms]
        if perturb:
            prng = np.random.default_rng(5)
            for rod in asm.arms:
                rod.position_collection[...] *= 1.0 + perturb * prng.standard_normal(rod.position_collection.shape)
        return asm, views

    systems = [[make(), make(1e-13), make(variant="fma")] for _ in range(3)]
    worst = 0.0
    for s in range(2):
        a = torch.as_tensor(acts[s], device=vec.device)
        mus = vec.muscle_activations(a).cpu().numpy()                       # [env, arm, 3, n]
        vec.step(a)
        st = {k: v.cpu().numpy() for k, v in vec.fields().items()}
        hd = vec.handle.head_tensor().cpu().numpy()
        for e_i in range(3):
            for asm, views in systems[e_i]:
                for k, rod in enumerate(asm.arms):
                    for s_ in range(3):
                        rod.set_sucker(s_, vec.sucker_location[s_], float(acts[s, e_i].reshape(2, 9)[k, s_]))
                    views[k][...] = mus[e_i, k]
                asm.substeps(vec.step_skip)
            (oasm, _), reps = systems[e_i][0], [q for q, _ in systems[e_i][1:]]
            for arm in range(2):
                for gk, fk in FIELDS.items():
                    ref = getattr(oasm.arms[arm], fk)
                    sens = max(float(np.abs(getattr(q.arms[arm], fk) - ref).max()) for q in reps)
                    scale = max(float(np.abs(ref).max()), FLOOR[fk])
                    err = float(np.abs(st[fk][e_i, arm] - ref).max())
                    worst = max(worst, err / scale)
                    assert err <= max(TOL * scale, 20 * sens), \
                        f"env {e_i} step {s} arm {arm} {gk}: {err / scale:.3e} (oracle replicas: {sens / scale:.1e})"
            for (sl, gk, fk) in HEAD:
                ref = getattr(oasm, "head_" + gk).reshape(-1)
                sens = max(float(np.abs(getattr(q, "head_" + gk).reshape(-1) - ref).max()) for q in reps)
                scale = max(float(np.abs(ref).max()), FLOOR[fk])
                err = float(np.abs(hd[e_i, sl] - ref).max())
                assert err <= max(TOL * scale, 20 * sens), f"env {e_i} step {s} head {gk}: {err / scale:.3e}"
    print(f"OctoArmTwo-v0 n_elems={n}: worst field error {worst:.2e}")
    vec.close()


def test_octo_arm_two_vector_env_autoreset_and_prev_kappa():
    """Batched OctoArmTwo-v0: truncation past final_time rebuilds the envs inside step() (suckers back to ratio 1,
    muscles at rest), the truncated step's reward carries the reference's `- distance` term (arm_two_env.py:312-314), and
    the observation's prev_kappa block shows the previous observation's curvature (arm_two_env.py:201-219)."""
    import torch
    import gym_softrobot_b200 as gsb
    rng = np.random.default_rng(2)
    vec = gsb.make_vec("OctoArmTwo-v0", 4, final_time=0.075)
    o0, _ = vec.reset()
    n_seg = vec.n_seg
    assert tuple(o0.shape) == (4, vec.single_observation_space.shape[0])
    a = torch.as_tensor(rng.random((4, 18)).astype(np.float32), device=vec.device)
    o1, r1, te1, tr1, _ = vec.step(a)
    assert not bool(tr1.any()) and not bool(te1.any())
    k1 = o1.reshape(4, 2, -1)[:, :, :n_seg].clone()
    assert float(k1.abs().max()) > 0.0
    o2, r2, te2, tr2, info = vec.step(a)
    assert bool(tr2.all()) and "final_obs" in info and int(vec.step_count.max()) == 0
    # the pre-reset observation of the truncated step carries the previous step's curvature in its prev_kappa block
    assert torch.equal(info["final_obs"].reshape(4, 2, -1)[:, :, n_seg:2 * n_seg], k1)
    assert bool((r2 < -4.0).all())                        # forward_reward -= |target - head| ~ 5 m
    assert float(vec.handle.muscle_activation_tensor().abs().max()) == 0.0
    assert float((vec.handle.fixed_sucker_tensor() - 1.0).abs().max()) == 0.0
    assert float(o2.reshape(4, 2, -1)[:, :, :n_seg].abs().max()) < 1e-12      # fresh arms are straight
    vec.close()
