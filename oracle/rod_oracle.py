"""TEST INFRASTRUCTURE (oracle) — not product code.

ctypes wrapper over oracle/rod_oracle.c (the plain-C FP64 restatement of the
reference physics step; PARITY UNPINNED against real PyElastica — see
oracle/README.md).  Allowed importers: tests/, __graft_entry__.smoke(),
bench.py's cpu_baseline / `--impl reference` leg.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "librod_oracle.so")

BC_FREE, BC_ONE_END_FIXED, BC_PENDULUM_SLIDER, BC_MOVING_BASE = 0, 1, 2, 3


class ROConfig(C.Structure):
    _fields_ = [
        ("n_elem", C.c_int),
        ("start", C.c_double * 3),
        ("direction", C.c_double * 3),
        ("normal", C.c_double * 3),
        ("base_length", C.c_double),
        ("base_radius", C.c_double),
        ("density", C.c_double),
        ("youngs_modulus", C.c_double),
        ("shear_modulus", C.c_double),
        ("shear_convention", C.c_int),
        ("dt", C.c_double),
        ("gravity", C.c_double * 3),
        ("damping_constant", C.c_double),
        ("laplace_filter_order", C.c_int),
        ("bc_kind", C.c_int),
        ("point_force_on_base", C.c_int),
        ("damping_before_constraints", C.c_int),
        ("contact_on", C.c_int),
        ("contact_before_forcing", C.c_int),
        ("plane_origin", C.c_double * 3),
        ("plane_normal", C.c_double * 3),
        ("contact_k", C.c_double),
        ("contact_nu", C.c_double),
        ("slip_velocity_tol", C.c_double),
        ("surface_tol", C.c_double),
        ("static_mu", C.c_double * 3),
        ("kinetic_mu", C.c_double * 3),
        ("muscle_on", C.c_int),
        ("muscle_period", C.c_double),
        ("muscle_ramp_up_time", C.c_double),
        ("muscle_phase_shift", C.c_double),
        ("muscle_direction", C.c_double * 3),
        ("spline_dir_mask", C.c_int),
        ("spline_n_ctrl", C.c_int),
        ("spline_scale", C.c_double),
        ("spline_max_rate", C.c_double),
        ("tip_radius", C.c_double),
        ("taper_node_mean", C.c_int),
    ]


RO_MAX_ARMS = 16


class ROAsmConfig(C.Structure):
    _fields_ = [
        ("n_arm", C.c_int),
        ("has_head", C.c_int),
        ("dt", C.c_double),
        ("head_start", C.c_double * 3),
        ("head_direction", C.c_double * 3),
        ("head_normal", C.c_double * 3),
        ("head_length", C.c_double),
        ("head_radius", C.c_double),
        ("head_density", C.c_double),
        ("joint_k", C.c_double),
        ("joint_nu", C.c_double),
        ("joint_kt", C.c_double),
        ("joint_radius", C.c_double),
        ("joint_angle_deg", C.c_double * RO_MAX_ARMS),
    ]


_VARIANTS = {None: "librod_oracle.so", "fma": "librod_oracle_fma.so"}


def build(force: bool = False, variant=None) -> str:
    """Compile the C oracle (gcc, -ffp-contract=off; variant "fma": -O3 with FMA contraction — the same source built
    the other legitimate way, used to MEASURE how far two correct builds drift apart).  Building the checker is not using it."""
    src = os.path.join(_HERE, "rod_oracle.c")
    path = os.path.join(_HERE, "_build", _VARIANTS[variant])
    if force or not os.path.exists(path) or os.path.getmtime(path) < max(os.path.getmtime(src), os.path.getmtime(src[:-1] + "h")):
        subprocess.run(["make", "-C", _HERE, "-B", "_build/" + _VARIANTS[variant]], check=True, stdout=subprocess.DEVNULL)
    return path


_libs = {}
_variant = None


class variant:
    """`with rod_oracle.variant("fma"): ...` — oracle objects created inside use that build of the library."""

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        global _variant
        self.prev, _variant = _variant, self.name

    def __exit__(self, *a):
        global _variant
        _variant = self.prev


def lib(which="current"):
    v = _variant if which == "current" else which
    if v not in _libs:
        L = C.CDLL(build(variant=v))
        L.ro_create.restype = C.c_void_p
        L.ro_create.argtypes = [C.POINTER(ROConfig)]
        L.ro_destroy.argtypes = [C.c_void_p]
        L.ro_substeps.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p]
        L.ro_time.restype = C.c_double
        L.ro_time.argtypes = [C.c_void_p]
        for name in ("position", "velocity", "director", "omega", "tangents", "kappa", "sigma",
                     "dilatation", "rest_kappa", "external_forces", "mass", "internal_forces",
                     "internal_torques", "radius", "muscle", "spline_points", "spline_magnitude",
                     "external_torques"):
            f = getattr(L, "ro_" + name)
            f.restype = C.POINTER(C.c_double)
            f.argtypes = [C.c_void_p]
        L.ro_softpendulum_obs.argtypes = [C.c_void_p, C.c_float, C.POINTER(C.c_float)]
        L.ro_softpendulum_step.argtypes = [C.c_void_p, C.c_float, C.c_int, C.c_double,
                                           C.POINTER(C.c_float), C.POINTER(C.c_double),
                                           C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.ro_softpendulum_step_batch.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_int,
                                                 C.c_double, C.c_void_p, C.c_void_p, C.c_void_p,
                                                 C.c_void_p, C.c_int]
        L.ro_max_threads.restype = C.c_int
        L.ro_set_sucker.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double]
        L.ro_set_tm_muscle.argtypes = [C.c_void_p, C.c_double, C.c_double]
        L.ro_set_tm_activation.argtypes = [C.c_void_p, C.c_double]
        L.ro_set_muscle_layer.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double]
        L.ro_muscle_activation.restype = C.POINTER(C.c_double)
        L.ro_muscle_activation.argtypes = [C.c_void_p]
        L.ro_asm_create.restype = C.c_void_p
        L.ro_asm_create.argtypes = [C.POINTER(ROConfig), C.POINTER(ROAsmConfig)]
        L.ro_asm_destroy.argtypes = [C.c_void_p]
        L.ro_asm_set_head_fixed.argtypes = [C.c_void_p, C.c_int]
        L.ro_asm_substeps.argtypes = [C.c_void_p, C.c_int]
        L.ro_asm_arm.restype = C.c_void_p
        L.ro_asm_arm.argtypes = [C.c_void_p, C.c_int]
        L.ro_asm_time.restype = C.c_double
        L.ro_asm_time.argtypes = [C.c_void_p]
        for name in ("position", "velocity", "director", "omega"):
            f = getattr(L, "ro_asm_head_" + name)
            f.restype = C.POINTER(C.c_double)
            f.argtypes = [C.c_void_p]
        L.ro_substeps_batch.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int]
        L.ro_asm_substeps_batch.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int]
        _libs[v] = L
    return _libs[v]


class OracleRod:
    """One Cosserat rod stepped by the C oracle; arrays are NumPy views (reference layout)."""

    @staticmethod
    def _make_config(n_elem, start, direction, normal, base_length, base_radius, density,
                     youngs_modulus, dt, shear_modulus=0.0, shear_convention=0,
                     gravity=(0.0, 0.0, 0.0), damping_constant=-1.0, laplace_filter_order=0,
                     bc_kind=BC_FREE, point_force_on_base=False, damping_before_constraints=False,
                     contact=None, muscle=None, spline=None, tip_radius=0.0, taper_node_mean=False):
        cfg = ROConfig()
        cfg.tip_radius = float(tip_radius)
        cfg.taper_node_mean = int(taper_node_mean)
        cfg.n_elem = n_elem
        cfg.start[:] = list(map(float, start))
        cfg.direction[:] = list(map(float, direction))
        cfg.normal[:] = list(map(float, normal))
        cfg.base_length, cfg.base_radius, cfg.density = base_length, base_radius, density
        cfg.youngs_modulus, cfg.shear_modulus, cfg.shear_convention = youngs_modulus, shear_modulus, shear_convention
        cfg.dt = dt
        cfg.gravity[:] = list(map(float, gravity))
        cfg.damping_constant = damping_constant
        cfg.laplace_filter_order = laplace_filter_order
        cfg.bc_kind = bc_kind
        cfg.point_force_on_base = int(point_force_on_base)
        cfg.damping_before_constraints = int(damping_before_constraints)
        if contact is not None:   # dict: plane_origin, plane_normal, k, nu, slip_velocity_tol, static_mu, kinetic_mu
            cfg.contact_on = 1
            cfg.contact_before_forcing = int(contact.get("before_forcing", False))
            cfg.plane_origin[:] = list(map(float, contact["plane_origin"]))
            cfg.plane_normal[:] = list(map(float, contact["plane_normal"]))
            cfg.contact_k, cfg.contact_nu = contact["k"], contact["nu"]
            cfg.slip_velocity_tol, cfg.surface_tol = contact["slip_velocity_tol"], contact.get("surface_tol", 1e-4)
            cfg.static_mu[:] = list(map(float, contact["static_mu"]))
            cfg.kinetic_mu[:] = list(map(float, contact["kinetic_mu"]))
        if muscle is not None:    # dict: period, ramp_up_time, phase_shift, direction (MuscleTorques kwargs)
            cfg.muscle_on = 1
            cfg.muscle_period, cfg.muscle_ramp_up_time = muscle["period"], muscle["ramp_up_time"]
            cfg.muscle_phase_shift = muscle.get("phase_shift", 0.0)
            cfg.muscle_direction[:] = list(map(float, muscle["direction"]))
        if spline is not None:    # dict: directions (subset of 0,1,2), n_ctrl, scale, max_rate
            cfg.spline_dir_mask = sum(1 << int(d) for d in spline["directions"])
            cfg.spline_n_ctrl = spline["n_ctrl"]
            cfg.spline_scale, cfg.spline_max_rate = spline["scale"], spline.get("max_rate", float("inf"))
        return cfg

    def __init__(self, n_elem, *args, _handle=None, **kwargs):
        cfg = self._make_config(n_elem, *args, **kwargs)
        self.cfg = cfg
        self.n = n_elem
        self._owned = _handle is None
        self._L = lib()
        self._h = C.c_void_p(self._L.ro_create(C.byref(cfg))) if _handle is None else C.c_void_p(_handle)
        n = n_elem
        self.position_collection = self._view("position", (3, n + 1))
        self.velocity_collection = self._view("velocity", (3, n + 1))
        self.director_collection = self._view("director", (3, 3, n))
        self.omega_collection = self._view("omega", (3, n))
        self.tangents = self._view("tangents", (3, n))
        self.kappa = self._view("kappa", (3, n - 1))
        self.sigma = self._view("sigma", (3, n))
        self.dilatation = self._view("dilatation", (n,))
        self.rest_kappa = self._view("rest_kappa", (3, n - 1))
        self.user_forces = self._view("external_forces", (3, n + 1))
        self.user_torques = self._view("external_torques", (3, n))   # material frame, added every substep
        self.mass = self._view("mass", (n + 1,))
        self.internal_forces = self._view("internal_forces", (3, n + 1))
        self.internal_torques = self._view("internal_torques", (3, n))
        self.radius = self._view("radius", (n,))
        self.muscle = self._view("muscle", (n + 1,))       # wave number, beta(s_k)
        P = max(int(cfg.spline_n_ctrl), 0)
        self.spline_points = self._view("spline_points", (3, 2 * P + 1))     # per direction: P targets, P cached, flag
        self.spline_magnitude = self._view("spline_magnitude", (3, n))

    def _view(self, name, shape):
        p = getattr(self._L, "ro_" + name)(self._h)
        return np.ctypeslib.as_array(p, shape=(int(np.prod(shape)),)).reshape(shape)

    @property
    def time(self):
        return self._L.ro_time(self._h)

    def substeps(self, n, action=0.0, base_pos=None, base_vel=None):
        bp = None if base_pos is None else np.ascontiguousarray(base_pos, dtype=np.float64)
        bv = None if base_vel is None else np.ascontiguousarray(base_vel, dtype=np.float64)
        self._L.ro_substeps(self._h, int(n), float(action),
                          None if bp is None else bp.ctypes.data, None if bv is None else bv.ctypes.data)

    def set_sucker(self, slot, index, ratio):
        """ControllableFixConstraint(index, reduction_ratio) in slot `slot` (ratio 0 = released)."""
        self._L.ro_set_sucker(self._h, int(slot), int(index), float(ratio))

    def set_tm_muscle(self, max_stress, radius_ref):
        """COOMM TransverseMuscle(rest_muscle_area=(radius / radius_ref)**2, max_muscle_stress) under ApplyMuscles."""
        self._L.ro_set_tm_muscle(self._h, float(max_stress), float(radius_ref))

    def set_es_muscle_layers(self, radius_ref):
        """create_es_muscle_layers (envs/octopus/build.py:292-338): LM at (0, -2/3, 0) rotated by +-pi/2 about the axis
        (max stress 0.5), TM (max stress 1.0, sign flipped by the class); returns the (3, n) activation view."""
        for slot, ang in enumerate((np.pi / 2, -np.pi / 2)):
            c, s = np.cos(ang), np.sin(ang)
            px, py = c * 0.0 - s * (-6 / 9), s * 0.0 + c * (-6 / 9)
            self._L.ro_set_muscle_layer(self._h, slot, 1, 0.5, float(radius_ref), float(px), float(py))
        self._L.ro_set_muscle_layer(self._h, 2, 2, -1.0, float(radius_ref), 0.0, 0.0)
        p = self._L.ro_muscle_activation(self._h)
        return np.ctypeslib.as_array(p, shape=(3 * self.n,)).reshape(3, self.n)

    def set_tm_activation(self, activation):
        self._L.ro_set_tm_activation(self._h, float(activation))

    def close(self):
        if self._h:
            if self._owned:
                self._L.ro_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class OracleAssembly:
    """n_arm rods + rigid Cylinder head + FixedJoint2Rigid joints + BodyBoundaryCondition on the C oracle
    (reference envs/octopus/build.py:52-217 topology).  `arm_kwargs` are OracleRod keyword sets (one per arm)."""

    def __init__(self, arm_kwargs, dt, head=None, joint=None, angles_deg=None):
        n_arm = len(arm_kwargs)
        assert 1 <= n_arm <= RO_MAX_ARMS
        self._cfg_holders = []
        cfgs = (ROConfig * n_arm)()
        for i, kw in enumerate(arm_kwargs):
            cfgs[i] = OracleRod._make_config(dt=dt, **kw)
        ac = ROAsmConfig()
        ac.n_arm, ac.dt = n_arm, dt
        ac.has_head = int(head is not None)
        if head is not None:   # dict: start, direction, normal, length, radius, density
            ac.head_start[:] = list(map(float, head["start"]))
            ac.head_direction[:] = list(map(float, head["direction"]))
            ac.head_normal[:] = list(map(float, head["normal"]))
            ac.head_length, ac.head_radius, ac.head_density = head["length"], head["radius"], head["density"]
            ac.joint_k, ac.joint_nu, ac.joint_kt, ac.joint_radius = joint["k"], joint["nu"], joint["kt"], joint["radius"]
            for i, a in enumerate(angles_deg):
                ac.joint_angle_deg[i] = float(a)
        else:   # keep the (unused) rigid body well defined
            ac.head_direction[:] = [0.0, 0.0, 1.0]
            ac.head_normal[:] = [0.0, 1.0, 0.0]
            ac.head_length = ac.head_radius = ac.head_density = 1.0
        self._L = lib()
        self._h = C.c_void_p(self._L.ro_asm_create(cfgs, C.byref(ac)))
        assert self._h
        self.arms = [OracleRod(dt=dt, _handle=self._L.ro_asm_arm(self._h, i), **kw) for i, kw in enumerate(arm_kwargs)]
        hv = lambda name, shape: np.ctypeslib.as_array(getattr(self._L, "ro_asm_head_" + name)(self._h),
                                                       shape=(int(np.prod(shape)),)).reshape(shape)
        self.head_position, self.head_velocity = hv("position", (3,)), hv("velocity", (3,))
        self.head_director, self.head_omega = hv("director", (3, 3)), hv("omega", (3,))

    @property
    def time(self):
        return self._L.ro_asm_time(self._h)

    def substeps(self, n):
        self._L.ro_asm_substeps(self._h, int(n))

    def set_head_fixed(self, on=True):
        """OneEndFixedBC on the rigid head (reach_env.py:128-132)."""
        self._L.ro_asm_set_head_fixed(self._h, int(on))

    def close(self):
        if self._h:
            for a in self.arms:
                a.close()
            self._L.ro_asm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def substeps_batch(systems, n_substeps, n_threads=0):
    """Advance a list of OracleRod or OracleAssembly objects by n_substeps each on n_threads pthreads."""
    arr = (C.c_void_p * len(systems))(*[s._h for s in systems])
    if isinstance(systems[0], OracleAssembly):
        lib().ro_asm_substeps_batch(arr, len(systems), int(n_substeps), int(n_threads))
    else:
        lib().ro_substeps_batch(arr, len(systems), int(n_substeps), int(n_threads))


def pendulum_direction_normal(u01: float):
    """theta / direction / normal exactly as reference soft_pendulum/build.py:47-51."""
    theta = np.deg2rad(90 + (u01 - 0.5) * 10)
    direction = np.array([1.0 * np.cos(theta), 1.0 * np.sin(theta), 0.0])
    normal = np.array([1.0 * np.sin(theta), -1.0 * np.cos(theta), 0.0])
    return direction, normal


class OracleSoftPendulum:
    """SoftPendulum-v0 on the C oracle (reference soft_pendulum.py:59-251, build.py:29-115)."""

    def __init__(self, final_time=5.0, time_step=1.0e-4, recording_fps=25, n_elems=50,
                 shear_convention=0):
        self.final_time, self.time_step, self.n_elems = final_time, time_step, n_elems
        self.step_skip = int(1.0 / (recording_fps * time_step))
        self.shear_convention = shear_convention
        self.rod = None
        self.prev_action = np.float32(0.0)

    def reset(self, seed=None, u01=None):
        if u01 is None:
            u01 = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed))).random()
        direction, normal = pendulum_direction_normal(u01)
        if self.rod is not None:
            self.rod.close()
        self.rod = OracleRod(self.n_elems, np.zeros(3), direction, normal, 1.0, 0.05, 1000.0, 1e6,
                             self.time_step, shear_convention=self.shear_convention,
                             gravity=(0.0, -9.80665, 0.0), damping_constant=2e-3,
                             bc_kind=BC_PENDULUM_SLIDER, point_force_on_base=True)
        self.prev_action = np.float32(0.0)
        return self.obs(), {}

    def obs(self):
        out = (C.c_float * 4)()
        lib().ro_softpendulum_obs(self.rod._h, C.c_float(float(self.prev_action)), out)
        return np.array(out[:], dtype=np.float32)

    def step(self, action):
        a = np.float32(np.asarray(action, dtype=np.float32).reshape(-1)[0])
        out = (C.c_float * 4)()
        rew, term, trunc = C.c_double(), C.c_int(), C.c_int()
        lib().ro_softpendulum_step(self.rod._h, C.c_float(float(a)), self.step_skip, self.final_time,
                                   out, C.byref(rew), C.byref(term), C.byref(trunc))
        self.prev_action = a
        return (np.array(out[:], dtype=np.float32), rew.value, bool(term.value), bool(trunc.value),
                {"time": np.float64(self.rod.time), "TimeLimit.truncated": bool(trunc.value)})


class OracleSoftPendulum3D:
    """SoftPendulum3D-v0 on the C oracle: env logic restated from
    /root/reference/gym_softrobot/envs/soft_pendulum_3d/soft_pendulum_3d.py:60-158 and build.py:43-86."""

    def __init__(self, final_time=5.0, time_step=1.0e-4, recording_fps=25, n_elems=50):
        self.final_time, self.time_step, self.n_elems = final_time, time_step, n_elems
        self.step_skip = int(1.0 / (recording_fps * time_step))
        self.base_step, self.base_limit = 1e-3, 0.5
        self.rod = None

    def reset(self, seed=None, tilt_deg=None):
        if tilt_deg is None:
            tilt_deg = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed))).uniform(-1.0, 1.0)
        tilt = np.deg2rad(tilt_deg)
        direction = np.array([np.sin(tilt), 0.0, np.cos(tilt)])
        normal = np.array([0.0, 1.0, 0.0])
        if self.rod is not None:
            self.rod.close()
        self.rod = OracleRod(self.n_elems, np.zeros(3), direction, normal, 1.0, 0.1, 4000.0, 1e6, self.time_step,
                             gravity=(0.0, 0.0, -9.80665), damping_constant=1.0, laplace_filter_order=7,
                             bc_kind=BC_MOVING_BASE)
        self.position, self.velocity = np.zeros(3), np.zeros(3)
        self.prev_action = np.zeros(2, dtype=np.float32)
        return self.get_state(), {}

    def _tilt_angle(self):
        tangent = np.mean(self.rod.tangents, axis=1)
        tangent /= np.linalg.norm(tangent)
        return float(np.arccos(np.clip(tangent[2], -1.0, 1.0)))

    def get_state(self):
        return np.hstack([self.rod.position_collection[:, 0], self.rod.velocity_collection[:, 0],
                          self.prev_action, self._tilt_angle()]).astype(np.float32)

    def step(self, action):
        action = np.asarray(action, dtype=np.float32)
        displacement = self.base_step * action
        next_position = self.position.copy()
        next_position[:2] = np.clip(next_position[:2] + displacement, -self.base_limit, self.base_limit)
        actual = next_position - self.position
        self.position[:] = next_position
        self.velocity[:] = actual / (self.step_skip * self.time_step)
        self.prev_action[:] = action
        self.rod.substeps(self.step_skip, base_pos=self.position, base_vel=self.velocity)
        invalid = bool(np.isnan(self.rod.position_collection).any() or np.isnan(self.rod.velocity_collection).any())
        tilt = self._tilt_angle()
        base_distance = np.linalg.norm(self.position[:2])
        reward = -float(tilt ** 2 + 0.1 * base_distance ** 2 + 1e-3 * np.dot(action, action))
        if invalid:
            reward = -50.0
        t = self.rod.time
        return self.get_state(), reward, invalid, bool(t >= self.final_time), {"time": np.float64(t), "tilt": tilt}


def octopus_assembly(n_arm=8, n_elem=10, time_step=7e-5, head_radius=0.04, head_density=700.0,
                     body_arm_k=1e6, body_arm_kt=1.0, body_arm_nu=1e-3, friction_multiplier=1.0,
                     base_length=0.35, base_radius=0.35 * 0.02, youngs_modulus=1e6, density=1000.0,
                     tip_radius=0.0, plane=True, gravity=-9.81, damping_constant=1e-2, angle_offset=0.0):
    """The systems `build_octopus` assembles (reference envs/octopus/build.py:52-217), on the C oracle:
    n_arm arms at 360/n_arm degrees around a rigid Cylinder head, FixedJoint2Rigid joints, gravity and
    AnalyticalLinearDamper(1e-2) on the arms, anisotropic plane friction under each arm."""
    L0, r0 = base_length, base_radius
    g = gravity
    mu = L0 / (2.0 * 2.0 * 9.81 * 0.1)
    kinetic = np.array([mu, 1.5 * mu, 2.0 * mu]) * friction_multiplier
    contact = dict(plane_origin=(0.0, 0.0, -r0), plane_normal=(0.0, 0.0, 1.0), k=1e2, nu=1e1,
                   slip_velocity_tol=1e-8, static_mu=2 * kinetic, kinetic_mu=kinetic)
    angles = [angle_offset + 360 / n_arm * i for i in range(n_arm)]
    arms = []
    for ang in angles:
        # scipy Rotation.from_euler("z", ang, degrees=True).apply(v) for v along x
        c, s = np.cos(np.deg2rad(ang)), np.sin(np.deg2rad(ang))
        arms.append(dict(n_elem=n_elem, start=(c * head_radius, s * head_radius, 0.0), direction=(c, s, 0.0),
                         normal=(0.0, 0.0, 1.0), base_length=L0, base_radius=r0, density=density,
                         youngs_modulus=youngs_modulus, gravity=(0.0, 0.0, g), damping_constant=damping_constant,
                         contact=contact if plane else None, tip_radius=tip_radius))
    head = dict(start=(0.0, 0.0, -r0), direction=(0.0, 0.0, 1.0), normal=(0.0, 1.0, 0.0), length=2 * r0,
                radius=head_radius, density=head_density)
    joint = dict(k=body_arm_k, nu=body_arm_nu, kt=body_arm_kt, radius=head_radius)
    return OracleAssembly(arms, time_step, head=head, joint=joint, angles_deg=angles)
