"""TEST INFRASTRUCTURE (oracle) — generates tests/golden/*.npz.

Runs the UNMODIFIED reference env classes (`/root/reference/gym_softrobot/...`)
on top of the oracle's PyElastica/gymnasium shims (oracle/shims) and records
golden input/output vectors.  The gym-softrobot layer (env logic, plugins,
seeding, reward, truncation) is therefore the reference's own code; the
PyElastica layer underneath is the NumPy restatement (pyelastica==1.0.0 is not
installable here) — fixtures are labelled  parity="unpinned (restated PyElastica)".

Run only in the build container (needs /root/reference):
    python oracle/gen_golden.py
"""
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, _HERE)
import ref_loader  # noqa: E402

OUT = os.path.join(os.path.dirname(_HERE), "tests", "golden")
LABEL = "unpinned (reference gym-softrobot code on restated PyElastica, oracle/shims)"


def rod_state(rod):
    return dict(
        position=rod.position_collection.copy(),
        velocity=rod.velocity_collection.copy(),
        director=rod.director_collection.copy(),
        omega=rod.omega_collection.copy(),
        tangents=rod.tangents.copy(),
        kappa=rod.kappa.copy(),
        sigma=rod.sigma.copy(),
        dilatation=rod.dilatation.copy(),
    )


def pack(prefix, d, out):
    for k, v in d.items():
        out[f"{prefix}/{k}"] = v


def gen_soft_pendulum_episode(seed=42, n_state_steps=4):
    """Config 1 of BASELINE.json: SoftPendulum-v0, seed 42, random actions, full episode."""
    env = ref_loader.load_reference_env("SoftPendulum-v0")
    obs0, _ = env.reset(seed=seed)
    env.action_space.seed(seed)
    rod = env.unwrapped.shearable_rod
    out = {"label": LABEL, "seed": seed, "obs0": obs0}
    pack("state0", rod_state(rod), out)
    actions, obs, rew, term, trunc, times = [], [], [], [], [], []
    step = 0
    while True:
        a = env.action_space.sample()
        o, r, te, tr, info = env.step(a)
        actions.append(a); obs.append(o); rew.append(r); term.append(te); trunc.append(tr)
        times.append(info["time"])
        step += 1
        if step <= n_state_steps:
            pack(f"state{step}", rod_state(rod), out)
        if te or tr:
            break
    pack("state_final", rod_state(rod), out)
    out.update(actions=np.array(actions, dtype=np.float32), obs=np.array(obs, dtype=np.float32),
               reward=np.array(rew, dtype=np.float64), terminated=np.array(term), truncated=np.array(trunc),
               time=np.array(times, dtype=np.float64), n_steps=step)
    np.savez_compressed(os.path.join(OUT, f"soft_pendulum_seed{seed}_episode.npz"), **out)
    print("soft_pendulum episode:", step, "steps; last obs", obs[-1], "reward", rew[-1])


def gen_soft_pendulum_substeps(seed=42):
    """State after 1, 10, 100, 400, 1000 raw PositionVerlet substeps (constant action 7.5)."""
    env = ref_loader.load_reference_env("SoftPendulum-v0")
    env.reset(seed=seed)
    e = env.unwrapped
    rod = e.shearable_rod
    e.set_action(np.array([7.5], dtype=np.float32))
    out = {"label": LABEL, "seed": seed, "action": 7.5}
    done = 0
    for target in (1, 10, 100, 400, 1000):
        for _ in range(target - done):
            e.time = e.do_step(e.simulator, e.time, e.time_step)
        done = target
        pack(f"sub{target}", rod_state(rod), out)
        out[f"sub{target}/time"] = e.time
    np.savez_compressed(os.path.join(OUT, f"soft_pendulum_seed{seed}_substeps.npz"), **out)
    print("soft_pendulum substeps: x_tip", rod.position_collection[:, -1])


def gen_determinism(env_id, seed=0, n=3):
    """The reference's own determinism protocol (tests/envs/test_determinism.py:7-58)."""
    env = ref_loader.load_reference_env(env_id)
    obs0, _ = env.reset(seed=seed)
    env.action_space.seed(seed)
    acts = [env.action_space.sample() for _ in range(n)]
    resp = [env.step(a) for a in acts]
    out = {"label": LABEL, "seed": seed, "obs0": obs0, "actions": np.array(acts),
           "obs": np.array([r[0] for r in resp]), "reward": np.array([r[1] for r in resp], dtype=np.float64),
           "terminated": np.array([r[2] for r in resp]), "truncated": np.array([r[3] for r in resp])}
    pack("state_final", rod_state(env.unwrapped.shearable_rod), out)
    np.savez_compressed(os.path.join(OUT, f"{env_id.replace('-', '_').lower()}_determinism_seed{seed}.npz"), **out)
    print(env_id, "determinism: obs", resp[-1][0], "reward", resp[-1][1])


def gen_soft_pendulum_3d(seed=42, n=6):
    env = ref_loader.load_reference_env("SoftPendulum3D-v0")
    obs0, _ = env.reset(seed=seed)
    env.action_space.seed(seed)
    rod = env.unwrapped.shearable_rod
    out = {"label": LABEL, "seed": seed, "obs0": obs0}
    pack("state0", rod_state(rod), out)
    acts, obs, rew, term, trunc, tilt = [], [], [], [], [], []
    for i in range(n):
        a = env.action_space.sample()
        o, r, te, tr, info = env.step(a)
        acts.append(a); obs.append(o); rew.append(r); term.append(te); trunc.append(tr); tilt.append(info["tilt"])
        pack(f"state{i + 1}", rod_state(rod), out)
    out.update(actions=np.array(acts, dtype=np.float32), obs=np.array(obs, dtype=np.float32),
               reward=np.array(rew), terminated=np.array(term), truncated=np.array(trunc), tilt=np.array(tilt))
    np.savez_compressed(os.path.join(OUT, f"soft_pendulum_3d_seed{seed}.npz"), **out)
    print("soft_pendulum_3d:", obs[-1], rew[-1])


def gen_arm_single(seed=42, n=5):
    """OctoArmSingle-v0 (free rod on a frictional plane, rest-curvature actuation): 5 env-steps."""
    env = ref_loader.load_reference_env("OctoArmSingle-v0")
    obs0, _ = env.reset(seed=seed)
    env.action_space.seed(seed)
    rod = env.unwrapped.shearable_rod
    out = {"label": LABEL + "; operator order = build-code call order (OperatorGroupFIFO): synchronize = [forcing..., contact]",
           "seed": seed, "obs0": obs0}
    pack("state0", rod_state(rod), out)
    acts, obs, rew, term, trunc = [], [], [], [], []
    for i in range(n):
        a = env.action_space.sample()
        o, r, te, tr, info = env.step(a)
        acts.append(a); obs.append(o); rew.append(r); term.append(te); trunc.append(tr)
        pack(f"state{i + 1}", rod_state(rod), out)
    out.update(actions=np.array(acts, dtype=np.float32), obs=np.array(obs, dtype=np.float32),
               reward=np.array(rew, dtype=np.float64), terminated=np.array(term), truncated=np.array(trunc))
    np.savez_compressed(os.path.join(OUT, f"octo_arm_single_seed{seed}.npz"), **out)
    print("arm_single:", rew, term)


def gen_octo_flat(seed=42, n=3, recording_fps=50):
    """OctoFlat-v0 (8 arms + rigid head + FixedJoint2Rigid joints + plane contact, rest-curvature
    actuation).  recording_fps=50 (285 substeps per env-step instead of 2857) keeps the NumPy run short;
    every other parameter is the registered default."""
    env = ref_loader.load_reference_env("OctoFlat-v0", recording_fps=recording_fps)
    obs0, _ = env.reset(seed=seed)
    env.action_space.seed(seed)
    e = env.unwrapped
    out = {"label": LABEL + "; operator order = build-code call order (OperatorGroupFIFO): synchronize = [connections, forcing, contact]",
           "seed": seed, "recording_fps": recording_fps, "step_skip": e.step_skip, "target": e._target.copy(),
           "obs0/individual": obs0["individual"], "obs0/shared": obs0["shared"]}

    def snap(tag):
        for a, rod in enumerate(e.shearable_rods):
            pack(f"{tag}/arm{a}", rod_state(rod), out)
        h = e.rigid_rod
        out[f"{tag}/head/position"] = h.position_collection.copy()
        out[f"{tag}/head/velocity"] = h.velocity_collection.copy()
        out[f"{tag}/head/director"] = h.director_collection.copy()
        out[f"{tag}/head/omega"] = h.omega_collection.copy()

    snap("state0")
    acts, rew, term, trunc = [], [], [], []
    for i in range(n):
        a = env.action_space.sample()
        o, r, te, tr, info = env.step(a)
        acts.append(a); rew.append(r); term.append(te); trunc.append(tr)
        out[f"obs{i + 1}/individual"], out[f"obs{i + 1}/shared"] = o["individual"], o["shared"]
        snap(f"state{i + 1}")
    out.update(actions=np.array(acts, dtype=np.float32), reward=np.array(rew, dtype=np.float64),
               terminated=np.array(term), truncated=np.array(trunc))
    np.savez_compressed(os.path.join(OUT, f"octo_flat_seed{seed}.npz"), **out)
    print("octo_flat:", rew, term, "head", e.rigid_rod.position_collection[:, 0])


def gen_octo_flat_decentralized(seed=42, recording_fps=50):
    """OctoFlat-v0, policy_mode="decentralized": same physics, per-arm action space, one-hot arm id in the
    individual observation (flat_env.py:111-145,248-260).  One env-step is enough to pin the layout."""
    env = ref_loader.load_reference_env("OctoFlat-v0", recording_fps=recording_fps, policy_mode="decentralized")
    obs0, _ = env.reset(seed=seed)
    e = env.unwrapped
    rng = np.random.default_rng(seed)
    a = rng.uniform(-22, 22, size=(e.n_arm, e.n_action)).astype(np.float32)     # one action per arm
    o, r, te, tr, info = e.step(a)
    np.savez_compressed(os.path.join(OUT, f"octo_flat_decentralized_seed{seed}.npz"), label=LABEL, seed=seed,
                        recording_fps=recording_fps, action=a, action_shape=np.array(env.action_space.shape),
                        **{"obs0/individual": obs0["individual"], "obs0/shared": obs0["shared"],
                           "obs1/individual": o["individual"], "obs1/shared": o["shared"]}, reward=np.float64(r))
    print("octo_flat decentralized:", o["individual"].shape, r)


def gen_octo_cfg4(seed=42, n=3, n_elems=40, time_step=3e-5, recording_fps=100):
    """BASELINE config 4 as specified: build_octopus(n_arm=8, n_elem=40) topology (envs/octopus/build.py:52-217)
    under FlatEnv's rest-curvature actuation.  dt = 3e-5 because the registered 7e-5 is unstable for 40 elements
    (joint spring k = 1e6 on a 0.67 g end node: dt < 2 / sqrt(k / m) = 5.2e-5); 3 x 333 = 999 substeps.
    The rest curvatures the env wrote are stored so that checkers need no scipy."""
    env = ref_loader.load_reference_env("OctoFlat-v0", n_elems=n_elems, time_step=time_step, recording_fps=recording_fps)
    obs0, _ = env.reset(seed=seed)
    env.action_space.seed(seed)
    e = env.unwrapped
    out = {"label": LABEL + "; operator order = build-code call order (OperatorGroupFIFO): synchronize = [connections, forcing, contact]",
           "seed": seed, "n_elems": n_elems, "time_step": time_step, "recording_fps": recording_fps,
           "step_skip": e.step_skip, "target": e._target.copy()}

    def snap(tag):
        for a, rod in enumerate(e.shearable_rods):
            pack(f"{tag}/arm{a}", rod_state(rod), out)
        h = e.rigid_rod
        out[f"{tag}/head/position"] = h.position_collection.copy()
        out[f"{tag}/head/velocity"] = h.velocity_collection.copy()
        out[f"{tag}/head/director"] = h.director_collection.copy()
        out[f"{tag}/head/omega"] = h.omega_collection.copy()

    snap("state0")
    acts, rew, term, rk = [], [], [], []
    for i in range(n):
        a = env.action_space.sample()
        o, r, te, tr, info = env.step(a)
        acts.append(a); rew.append(r); term.append(te)
        rk.append(np.stack([rod.rest_kappa.copy() for rod in e.shearable_rods]))
        snap(f"state{i + 1}")
    out.update(actions=np.array(acts, dtype=np.float32), reward=np.array(rew, dtype=np.float64),
               terminated=np.array(term), rest_kappa=np.array(rk))
    np.savez_compressed(os.path.join(OUT, f"octo_cfg4_8x{n_elems}_seed{seed}.npz"), **out)
    print("octo cfg4:", rew, term, "head", e.rigid_rod.position_collection[:, 0])


def gen_second_episode_reset_obs():
    """Which observation does reset() return in the SECOND episode?  SoftPendulum / OctoArmSingle / OctoFlat never clear
    `_prev_action` in reset (soft_pendulum.py:97, arm_single_env.py:100, flat_env.py:135-139), so it still carries the
    last action; SoftPendulum3D clears it (soft_pendulum_3d.py:68)."""
    out = {"label": LABEL}
    for env_id, kw in (("SoftPendulum-v0", {}), ("SoftPendulum3D-v0", {}), ("OctoArmSingle-v0", {}),
                       ("OctoFlat-v0", {"recording_fps": 50})):
        env = ref_loader.load_reference_env(env_id, **kw)
        env.reset(seed=1)
        env.action_space.seed(1)
        a = env.action_space.sample()
        env.step(a)
        obs2, _ = env.reset(seed=2)
        tag = env_id.split("-")[0]
        out[f"{tag}/action"] = np.asarray(a)
        if isinstance(obs2, dict):
            for k, v in obs2.items():
                out[f"{tag}/obs2/{k}"] = v
        else:
            out[f"{tag}/obs2"] = obs2
    np.savez_compressed(os.path.join(OUT, "second_episode_reset_obs.npz"), **out)
    print("second-episode reset observations:", out["SoftPendulum/obs2"], out["SoftPendulum3D/obs2"][6:8])


def gen_spline_forcing(seed=5):
    """The reference's `MuscleTorquesWithVaryingBetaSplines` (muscle_torques_with_bspline.py) driven directly,
    outside any env, to cover what SoftArmTracking-v0 does not: a finite max_rate_of_change_of_activation
    (the spline is re-fitted at the current lengths on every substep until the cached control values reach
    the targets), the tangent (twist) direction, 3 control points, targets changed every 50 substeps."""
    ref_loader.install_shims()
    import elastica as ea
    from gym_softrobot.utils.custom_elastica.muscle_torque import MuscleTorquesWithVaryingBetaSplines

    class Sim(ea.BaseSystemCollection, ea.Constraints, ea.Forcing, ea.Damping):
        pass

    n, L, r, E, dt = 24, 0.5, 0.02, 1e6, 5e-5
    sim = Sim()
    rod = ea.CosseratRod.straight_rod(n, np.zeros(3), np.array([0.0, 0.0, 1.0]), np.array([1.0, 0.0, 0.0]), L, r, 1000.0,
                                      youngs_modulus=E)
    sim.append(rod)
    sim.constrain(rod).using(ea.OneEndFixedBC, constrained_position_idx=(0,), constrained_director_idx=(0,))
    sim.dampen(rod).using(ea.AnalyticalLinearDamper, damping_constant=0.5, time_step=dt)
    pts = {"normal": [], "tangent": []}
    for d in ("normal", "tangent"):
        sim.add_forcing_to(rod).using(MuscleTorquesWithVaryingBetaSplines, base_length=L, number_of_control_points=3,
                                      points_func_array=pts[d], muscle_torque_scale=1.0, direction=d, step_skip=10 ** 9,
                                      max_rate_of_change_of_activation=0.04)
    sim.finalize()
    stepper, t = ea.PositionVerlet(), np.float64(0.0)
    rng = np.random.default_rng(seed)
    out = {"label": LABEL, "n_elem": n, "base_length": L, "base_radius": r, "youngs_modulus": E, "dt": dt,
           "damping_constant": 0.5, "scale": 1.0, "max_rate": 0.04, "n_ctrl": 3, "segment": 50}
    targets = []
    for seg in range(8):
        tgt = rng.uniform(-1, 1, (2, 3))
        pts["normal"][:] = tgt[0]; pts["tangent"][:] = tgt[1]
        targets.append(tgt)
        for _ in range(50):
            t = stepper.step(sim, t, dt)
        pack(f"seg{seg + 1}", rod_state(rod), out)
    out["targets"] = np.array(targets)
    np.savez_compressed(os.path.join(OUT, f"spline_forcing_seed{seed}.npz"), **out)
    print("spline forcing: tip", rod.position_collection[:, -1], "twist", rod.kappa[2].mean())


def gen_muscle_torques(seed=9):
    """PyElastica `MuscleTorques` (shim restatement) on the snake rod WITHOUT friction, so that the
    comparison is round-off limited (with kinetic friction the reference dynamics chatter):
    case A = plane with zero friction coefficients + gravity (normal response only), case B = no plane, no
    gravity, oblique torque direction and a phase shift.  Both rebuild the forcing (new beta spline / wave
    number) half way, like `set_action` does."""
    ref_loader.install_shims()
    import elastica as ea

    class Sim(ea.BaseSystemCollection, ea.Constraints, ea.Forcing, ea.Damping, ea.Contact):
        pass

    n, L, E, dt, period = 50, 0.35, 1e6, 8e-6, 2.0
    r = L * 0.011
    rng = np.random.default_rng(seed)
    out = {"label": LABEL, "n_elem": n, "dt": dt, "period": period, "segment": 1500}
    for case, (plane, direction, phase) in {"A": (True, np.array([0.0, 1.0, 0.0]), 0.0),
                                            "B": (False, np.array([0.6, 0.8, 0.0]), 0.7)}.items():
        sim = Sim()
        rod = ea.CosseratRod.straight_rod(n, np.zeros(3), np.array([0.0, 0.0, 1.0]), np.array([0.0, 1.0, 0.0]), L, r, 1000.0,
                                          youngs_modulus=E, shear_modulus=E / 1.5)
        sim.append(rod)
        sim.dampen(rod).using(ea.AnalyticalLinearDamper, damping_constant=1e-4, time_step=dt)
        if plane:
            sim.add_forcing_to(rod).using(ea.GravityForces, acc_gravity=np.array([0.0, -9.80665, 0.0]))
        ref = {}

        class Probe(ea.MuscleTorques):
            def __init__(self, *a, **k):
                super().__init__(*a, **k)
                ref["f"] = self

        kw = dict(base_length=L, period=period, phase_shift=phase, rest_lengths=rod.rest_lengths, ramp_up_time=period,
                  direction=direction, with_spline=True)
        sim.add_forcing_to(rod).using(Probe, b_coeff=np.zeros(6), wave_number=2 * np.pi, **kw)
        if plane:
            pl = ea.Plane(plane_origin=np.array([0.0, -r, 0.0]), plane_normal=np.array([0.0, 1.0, 0.0]))
            sim.append(pl)
            sim.detect_contact_between(rod, pl).using(ea.RodPlaneContactWithAnisotropicFriction, k=1.0, nu=1e-6,
                                                      slip_velocity_tol=1e-8, static_mu_array=np.zeros(3),
                                                      kinetic_mu_array=np.zeros(3))
        sim.finalize()
        stepper, t = ea.PositionVerlet(), np.float64(0.0)
        for seg in range(2):
            b = rng.uniform(-5e-3, 5e-3, 6).astype(np.float32)
            wl = np.float32(rng.uniform(0.7, 1.5))
            ref["f"].__init__(b_coeff=b, wave_number=2.0 * np.pi / wl, **kw)
            out[f"{case}/b{seg}"], out[f"{case}/wave_number{seg}"] = b, np.float64(ref["f"].wave_number)
            out[f"{case}/beta{seg}"] = ref["f"].my_spline.copy()
            for _ in range(1500):
                t = stepper.step(sim, t, dt)
            pack(f"{case}/seg{seg + 1}", rod_state(rod), out)
        out[f"{case}/time"] = t
        out[f"{case}/direction"], out[f"{case}/phase"] = direction, phase
        print("muscle case", case, "max |omega|", np.abs(rod.omega_collection).max())
    np.savez_compressed(os.path.join(OUT, f"muscle_torques_seed{seed}.npz"), **out)


def gen_snake(seed=42, n_state=3, n=33):
    """ContinuumSnake-v0 (n=50 rod, travelling-wave MuscleTorques rebuilt per action, anisotropic plane
    friction, dt=8e-6, 25 000 substeps per env-step).  33 env-steps so that the reward
    (`compute_projected_velocity`, non-zero from the 31st step on) is exercised; full states are kept
    for the first `n_state` steps, the callback samples (time, centre of mass, its velocity) for all."""
    env = ref_loader.load_reference_env("ContinuumSnake-v0")
    obs0, _ = env.reset(seed=seed)
    env.action_space.seed(seed)
    e = env.unwrapped
    rod = e.shearable_rod
    out = {"label": LABEL + "; operator order = build-code call order (OperatorGroupFIFO): synchronize = [forcing..., contact]",
           "seed": seed, "obs0": obs0, "step_skip": e.step_skip}
    pack("state0", rod_state(rod), out)
    acts, rew, term, trunc, times = [], [], [], [], []
    # full-range random actions (|b| up to 1e-2, re-drawn every 0.2 s) blow the NumPy simulation up after ~4 s,
    # so the fixture perturbs the published gait of the PyElastica snake case instead: still a new,
    # seed-determined action in the Box every step
    b_gait = np.array([3.4e-3, 3.3e-3, 4.2e-3, 2.6e-3, 3.6e-3, 3.5e-3])
    for i in range(n):
        s = env.action_space.sample().astype(np.float64)
        a = np.concatenate([b_gait + 0.1 * s[:6], [0.97 + 0.04 * (s[6] - 1.75)]]).astype(np.float32)
        assert env.action_space.contains(a)
        o, r, te, tr, info = env.step(a)
        assert np.isfinite(o).all(), f"simulation diverged at step {i + 1}"
        acts.append(a); rew.append(r); term.append(te); trunc.append(tr); times.append(float(e.time))
        if i < n_state:
            out[f"obs{i + 1}"] = o
            pack(f"state{i + 1}", rod_state(rod), out)
            out[f"beta{i + 1}"] = e.muscle_torque.my_spline.copy()
        print("snake step", i + 1, "reward", r, flush=True)
    out[f"obs{n}"] = o
    pack("state_final", rod_state(rod), out)
    out.update(actions=np.array(acts, dtype=np.float32), reward=np.array(rew, dtype=np.float64),
               terminated=np.array(term), truncated=np.array(trunc), time=np.array(times),
               cb_time=np.array(e.data["time"]), cb_step=np.array(e.data["step"]),
               cb_com=np.array(e.data["center_of_mass"]), cb_avg_velocity=np.array(e.data["avg_velocity"]))
    np.savez_compressed(os.path.join(OUT, f"continuum_snake_seed{seed}.npz"), **out)
    print("snake:", rew[-3:], "com", e.data["center_of_mass"][-1])


def gen_snake_perturbed(seed=42):
    """The same ContinuumSnake-v0 episode as gen_snake (same seed, same actions) with the initial node positions
    perturbed by 1e-15 m: how far two runs of the REFERENCE drift apart on their own (kinetic-friction chatter,
    DESIGN.md 5).  ~25 min of NumPy stepping."""
    g = np.load(os.path.join(OUT, f"continuum_snake_seed{seed}.npz"))
    env = ref_loader.load_reference_env("ContinuumSnake-v0")
    env.reset(seed=seed)
    e = env.unwrapped
    rng = np.random.default_rng(1)
    e.shearable_rod.position_collection += 1e-15 * rng.standard_normal(e.shearable_rod.position_collection.shape)
    rew = [env.step(a)[1] for a in g["actions"]]
    np.savez_compressed(os.path.join(OUT, f"continuum_snake_seed{seed}_perturbed.npz"),
                        label=LABEL + "; same seed and actions as continuum_snake_seed42.npz, initial node positions "
                        "perturbed by 1e-15 * N(0,1) (default_rng(1))",
                        reward=np.array(rew), cb_com=np.array(e.data["center_of_mass"]),
                        cb_avg_velocity=np.array(e.data["avg_velocity"]))
    print("snake perturbed: rewards", rew[-3:], "vs", g["reward"][-3:])


def gen_soft_arm(seed=42, n=40, game_mode=1):
    """SoftArmTracking-v0 (clamped n=40 arm, two spline muscle-torque forcings re-fitted at the current
    lengths, fixed or moving target): `n` env-steps of 50 substeps, random actions from the Box."""
    env = ref_loader.load_reference_env("SoftArmTracking-v0", game_mode=game_mode)
    obs0, _ = env.reset(seed=seed)
    env.action_space.seed(seed)
    e = env.unwrapped
    rod = e.shearable_rod
    out = {"label": LABEL + "; operator order = build-code call order (OperatorGroupFIFO)",
           "seed": seed, "game_mode": game_mode, "obs0": obs0, "targets": e.wsol[::int(e.num_steps_per_update)].copy()}
    pack("state0", rod_state(rod), out)
    acts, obs, rew, term, trunc, ctime, mags = [], [], [], [], [], [], []
    for i in range(n):
        a = env.action_space.sample()
        if i in (3, 4):
            a = acts[-1].copy()      # a repeated action: the forcing keeps its cached torque profile
        o, r, te, tr, info = env.step(a)
        acts.append(a); obs.append(o); rew.append(r); term.append(te); trunc.append(tr); ctime.append(info["ctime"])
        if i < 6 or i == n - 1:
            pack(f"state{i + 1}", rod_state(rod), out)
    out.update(actions=np.array(acts), obs=np.array(obs), reward=np.array(rew, dtype=np.float64),
               terminated=np.array(term), truncated=np.array(trunc), ctime=np.array(ctime))
    np.savez_compressed(os.path.join(OUT, f"soft_arm_tracking_mode{game_mode}_seed{seed}.npz"), **out)
    print("soft_arm mode", game_mode, ":", rew[-1], obs[-1][8:11])


COOMM_LABEL = ("unpinned (reference gym-softrobot code on restated PyElastica AND restated COOMM muscle model, "
               "oracle/shims/elastica + oracle/shims/coomm; coomm 0.1.1 @ d33fa034 is not obtainable offline)")


def gen_arm_push(env_id="OctoArmPush-v0", tag="octo_arm_push_v0", seed=42, n=6):
    """OctoArmPush-v0 (discrete: 0 = hold the base + contract the transverse muscle, 1 = hold the tip + release),
    OctoArmPush-v1 (continuous: sucker location, activation) and OctoArmPullWeight-v0 (the same arm dragging a rigid
    cylinder through a FixedJoint2Rigid): tapered free arm, ControllableFixConstraint, COOMM ApplyMuscles
    (/root/reference/gym_softrobot/envs/octopus/arm_push_env.py)."""
    env = ref_loader.load_reference_env(env_id)
    obs0, _ = env.reset(seed=seed)
    env.action_space.seed(seed)
    e = env.unwrapped
    rod = e.shearable_rod
    out = {"label": COOMM_LABEL, "seed": seed, "obs0": obs0, "step_skip": e.step_skip, "time_step": e.time_step,
           "radius0": rod.radius.copy(), "mass": rod.mass.copy()}
    head = getattr(e, "rigid_rod", None)

    def snap(tagk):
        pack(tagk, rod_state(rod), out)
        if head is not None:
            out[f"{tagk}/head/position"] = head.position_collection.copy()
            out[f"{tagk}/head/velocity"] = head.velocity_collection.copy()
            out[f"{tagk}/head/director"] = head.director_collection.copy()
            out[f"{tagk}/head/omega"] = head.omega_collection.copy()

    snap("state0")
    acts, obs, rew, term, trunc, times = [], [], [], [], [], []
    for i in range(n):
        a = env.action_space.sample()
        if e.mode == 0:
            a = i % 2 if i < 4 else int(a)        # the crawling gait first, then whatever the space samples
        o, r, te, tr, info = env.step(a)
        acts.append(a); obs.append(o); rew.append(r); term.append(te); trunc.append(tr); times.append(info["time"])
        snap(f"state{i + 1}")
    out.update(actions=np.array(acts), obs=np.array(obs, dtype=np.float32), reward=np.array(rew, dtype=np.float64),
               terminated=np.array(term), truncated=np.array(trunc), time=np.array(times))
    np.savez_compressed(os.path.join(OUT, f"{tag}_seed{seed}.npz"), **out)
    print(env_id, "rewards", rew, "tip x", rod.position_collection[0, -1])


def gen_arm_push_early(seed=1):
    """OctoArmPush-v1 with config_early_termination=True (arm_push_env.py:309-312,436-452): every reward is -10,
    the episode ends when kinetic + shear + bending energy < 1e-7 J — true on a first step without activation."""
    env = ref_loader.load_reference_env("OctoArmPush-v1", config_early_termination=True)
    env.reset(seed=seed)
    e = env.unwrapped
    acts = np.array([[0.3, 0.0], [0.3, 0.6], [0.3, 0.0], [0.5, 0.2]], dtype=np.float32)
    rew, term, trunc, ham = [], [], [], []
    for a in acts:
        o, r, te, tr, info = env.step(a)
        rew.append(r); term.append(te); trunc.append(tr); ham.append(e.cal_desired_Hamiltonian())
    np.savez_compressed(os.path.join(OUT, f"octo_arm_push_early_seed{seed}.npz"), label=COOMM_LABEL, actions=acts,
                        reward=np.array(rew), terminated=np.array(term), truncated=np.array(trunc), hamiltonian=np.array(ham))
    print("arm push early termination:", term, trunc, ham)


def gen_octo_crawl(seed=42, n=3):
    """OctoCrawl-v0: build_octopus_muscles (eight tapered arms, light head, joints, BodyBoundaryCondition), one
    ControllableFixConstraint per arm whose index / ratio the action moves, transverse-muscle activation per arm
    (/root/reference/gym_softrobot/envs/octopus/crawl_env.py, build_muscle_octopus.py:70-179)."""
    env = ref_loader.load_reference_env("OctoCrawl-v0")
    obs0, _ = env.reset(seed=seed)
    env.action_space.seed(seed)
    e = env.unwrapped
    out = {"label": COOMM_LABEL, "seed": seed, "obs0": obs0, "step_skip": e.step_skip, "time_step": e.time_step,
           "n_elems": e.n_elems}

    def snap(tag):
        for a, rod in enumerate(e.shearable_rods):
            pack(f"{tag}/arm{a}", rod_state(rod), out)
        h = e.rigid_rod
        out[f"{tag}/head/position"] = h.position_collection.copy()
        out[f"{tag}/head/velocity"] = h.velocity_collection.copy()
        out[f"{tag}/head/director"] = h.director_collection.copy()
        out[f"{tag}/head/omega"] = h.omega_collection.copy()

    snap("state0")
    acts, obs, rew, term, trunc = [], [], [], [], []
    for i in range(n):
        a = env.action_space.sample()
        o, r, te, tr, info = env.step(a)
        acts.append(a); obs.append(o); rew.append(r); term.append(te); trunc.append(tr)
        snap(f"state{i + 1}")
    out.update(actions=np.array(acts, dtype=np.float32), obs=np.array(obs, dtype=np.float32),
               reward=np.array(rew, dtype=np.float64), terminated=np.array(term), truncated=np.array(trunc))
    np.savez_compressed(os.path.join(OUT, f"octo_crawl_seed{seed}.npz"), **out)
    print("octo crawl:", rew, term, "head", e.rigid_rod.position_collection[:, 0])


def gen_octo_reach(seed=42, n=2):
    """OctoReach-v0: build_octopus_muscles with all three muscles driven by per-element activations and the head
    pinned by OneEndFixedBC on top of BodyBoundaryCondition (/root/reference/gym_softrobot/envs/octopus/reach_env.py)."""
    env = ref_loader.load_reference_env("OctoReach-v0")
    obs0, _ = env.reset(seed=seed)
    env.action_space.seed(seed)
    e = env.unwrapped
    out = {"label": COOMM_LABEL, "seed": seed, "obs0": obs0, "step_skip": e.step_skip, "time_step": e.time_step,
           "n_elems": e.n_elems, "target": np.asarray(e._target, dtype=np.float64)}

    def snap(tag):
        for a, rod in enumerate(e.shearable_rods):
            pack(f"{tag}/arm{a}", rod_state(rod), out)
        h = e.rigid_rod
        out[f"{tag}/head/position"] = h.position_collection.copy()
        out[f"{tag}/head/velocity"] = h.velocity_collection.copy()
        out[f"{tag}/head/director"] = h.director_collection.copy()
        out[f"{tag}/head/omega"] = h.omega_collection.copy()

    snap("state0")
    acts, obs, rew, term, trunc = [], [], [], [], []
    for i in range(n):
        a = env.action_space.sample()
        o, r, te, tr, info = env.step(a)
        acts.append(a); obs.append(o); rew.append(r); term.append(te); trunc.append(tr)
        snap(f"state{i + 1}")
    out.update(actions=np.array(acts, dtype=np.float32), obs=np.array(obs, dtype=np.float32),
               reward=np.array(rew, dtype=np.float64), terminated=np.array(term), truncated=np.array(trunc))
    np.savez_compressed(os.path.join(OUT, f"octo_reach_seed{seed}.npz"), **out)
    print("octo reach:", rew, term, trunc, "tip0", e.shearable_rods[0].position_collection[:, -1])


def gen_octo_arm_two(seed=42, n=3):
    """OctoArmTwo-v0: build_two_arms (two tapered arms at 90 / 270 degrees, free head), three fixed-index
    ControllableFixConstraints per arm, cubic-interpolated per-element activations of all three muscles
    (/root/reference/gym_softrobot/envs/octopus/arm_two_env.py)."""
    env = ref_loader.load_reference_env("OctoArmTwo-v0")
    obs0, _ = env.reset(seed=seed)
    env.action_space.seed(seed)
    e = env.unwrapped
    out = {"label": COOMM_LABEL, "seed": seed, "obs0": obs0, "step_skip": e.step_skip, "time_step": e.time_step,
           "n_elems": e.n_elems, "sucker_location": np.array(e.sucker_location)}

    def snap(tag):
        for a, rod in enumerate(e.shearable_rods):
            pack(f"{tag}/arm{a}", rod_state(rod), out)
        h = e.rigid_rod
        out[f"{tag}/head/position"] = h.position_collection.copy()
        out[f"{tag}/head/velocity"] = h.velocity_collection.copy()
        out[f"{tag}/head/director"] = h.director_collection.copy()
        out[f"{tag}/head/omega"] = h.omega_collection.copy()

    snap("state0")
    acts, obs, rew, term, trunc, musc = [], [], [], [], [], []
    for i in range(n):
        a = env.action_space.sample()
        o, r, te, tr, info = env.step(a)
        acts.append(a); obs.append(o); rew.append(r); term.append(te); trunc.append(tr)
        musc.append(np.array([[m.activation.copy() for m in arm] for arm in e.muscle_activations]))
        snap(f"state{i + 1}")
    out.update(actions=np.array(acts, dtype=np.float32), obs=np.array(obs, dtype=np.float32),
               reward=np.array(rew, dtype=np.float64), terminated=np.array(term), truncated=np.array(trunc),
               muscle_activations=np.array(musc))
    np.savez_compressed(os.path.join(OUT, f"octo_arm_two_seed{seed}.npz"), **out)
    print("octo arm two:", rew, term, trunc, "head", e.rigid_rod.position_collection[:, 0])


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "snake":
        gen_snake()
        gen_snake_perturbed()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "muscle":
        gen_muscle_torques()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "spline":
        gen_spline_forcing()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "decentralized":
        gen_octo_flat_decentralized()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "episode2":
        gen_second_episode_reset_obs()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "cfg4":
        gen_octo_cfg4()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "reach":
        gen_octo_reach()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "arm_two":
        gen_octo_arm_two()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "coomm":
        gen_arm_push("OctoArmPush-v0", "octo_arm_push_v0")
        gen_arm_push("OctoArmPush-v1", "octo_arm_push_v1")
        gen_arm_push("OctoArmPullWeight-v0", "octo_arm_pull_weight", n=3)
        gen_arm_push_early()
        gen_octo_crawl()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "soft_arm":
        gen_soft_arm(game_mode=1)
        gen_soft_arm(game_mode=2)
        sys.exit(0)
    gen_soft_pendulum_substeps()
    gen_determinism("SoftPendulum-v0")
    gen_determinism("SoftPendulum3D-v0")
    gen_soft_pendulum_3d()
    gen_soft_pendulum_episode()
    gen_arm_single()
    gen_octo_flat()
    gen_octo_flat_decentralized()
    gen_octo_cfg4()
    gen_second_episode_reset_obs()
    gen_spline_forcing()
    gen_muscle_torques()
    gen_soft_arm(game_mode=1)
    gen_soft_arm(game_mode=2)
    gen_arm_push("OctoArmPush-v0", "octo_arm_push_v0")
    gen_arm_push("OctoArmPush-v1", "octo_arm_push_v1")
    gen_arm_push("OctoArmPullWeight-v0", "octo_arm_pull_weight", n=3)
    gen_arm_push_early()
    gen_octo_crawl()
    gen_octo_reach()
    gen_octo_arm_two()
    gen_snake()   # ~25 min of NumPy stepping
    gen_snake_perturbed()   # another ~25 min
