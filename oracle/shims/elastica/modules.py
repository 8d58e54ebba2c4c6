"""TEST INFRASTRUCTURE (oracle) — not product code.

Restatement of PyElastica 1.0.0's system-collection mixins (``elastica/modules/*.py``,
[PE-recall], parity unpinned): BaseSystemCollection, Constraints, Connections,
Forcing, Damping, Contact, CallBacks.  The reference composes them at
`/root/reference/gym_softrobot/envs/soft_pendulum/soft_pendulum.py:34-42`.

Ordering rule (``elastica/modules/operator_group.py``, `OperatorGroupFIFO`): every user-facing
call (`add_forcing_to`, `connect`, `detect_contact_between` -> synchronize; `constrain`, `dampen`
-> constrain_rates; `constrain` -> constrain_values) appends its feature id to the group at call
time, and `finalize()` attaches the operators to that id; a group runs its operators in the order
the ids were appended, i.e. in the order the env's build code made the calls.  Consequences used
by the fixtures (DESIGN.md "conventions"):
  * every plane-contact env registers gravity (and the joints / muscle torques) BEFORE
    `detect_contact_between`, so the contact sees the weight it has to cancel — without that the
    friction magnitude (norm of the cancelling response only) would be zero and PyElastica's own
    ContinuumSnake case could not move;
  * SoftPendulum3D-v0 runs  [constrain_rates, AnalyticalLinearDamper, LaplaceDissipationFilter]
    (`soft_pendulum_3d/build.py:67-85`); SoftPendulum-v0's BC only zeroes components, so it is
    insensitive to the order.
There is no memory block here: systems are stepped one by one (B-8).
"""
import numpy as np


class OperatorGroupFIFO:
    """feature id -> operators, iterated in the order the ids were appended."""

    def __init__(self):
        self._ids = []
        self._ops = {}

    def append_id(self, feature):
        self._ids.append(id(feature))
        self._ops[id(feature)] = []

    def add_operators(self, feature, operators):
        self._ops[id(feature)].extend(operators)

    def __iter__(self):
        for i in self._ids:
            yield from self._ops[i]


class _Using:
    def __init__(self, system_idx):
        self._idx = system_idx
        self._cls = None
        self._args = ()
        self._kwargs = {}

    def using(self, cls, *args, **kwargs):
        self._cls, self._args, self._kwargs = cls, args, kwargs
        return self

    def id(self):
        return self._idx


class BaseSystemCollection:
    def __init__(self):
        self._feature_group_synchronize = OperatorGroupFIFO()
        self._feature_group_constrain_values = OperatorGroupFIFO()
        self._feature_group_constrain_rates = OperatorGroupFIFO()
        self._feature_group_callback = OperatorGroupFIFO()
        self._feature_group_finalize = []
        self._systems = []
        self._finalize_flag = False
        super().__init__()

    # sequence protocol used by the reference (`simulator.append(rod)`)
    def append(self, system):
        self._systems.append(system)

    def __len__(self):
        return len(self._systems)

    def __getitem__(self, i):
        return self._systems[i]

    def _get_sys_idx_if_valid(self, sys_to_be_added):
        if isinstance(sys_to_be_added, (int, np.integer)):
            return int(sys_to_be_added)
        for i, s in enumerate(self._systems):
            if s is sys_to_be_added:
                return i
        raise ValueError("system was not appended to the simulator")

    def block_systems(self):
        # only rods / rigid bodies are time-stepped; static surfaces (Plane) are not
        return [s for s in self._systems if hasattr(s, "velocity_collection")]

    def finalize(self):
        assert not self._finalize_flag, "The finalize cannot be called twice."
        for fn in self._feature_group_finalize:
            fn()
        self._feature_group_finalize.clear()
        self._finalize_flag = True

    def synchronize(self, time):
        for fn in self._feature_group_synchronize:
            fn(time)

    def constrain_values(self, time):
        for fn in self._feature_group_constrain_values:
            fn(time)

    def constrain_rates(self, time):
        for fn in self._feature_group_constrain_rates:
            fn(time)

    def apply_callbacks(self, time, current_step):
        for fn in self._feature_group_callback:
            fn(time, current_step)


class Constraints:
    def __init__(self):
        self._constraints = []
        super().__init__()
        self._feature_group_finalize.append(self._finalize_constraints)

    def constrain(self, system):
        u = _Using(self._get_sys_idx_if_valid(system))
        self._constraints.append(u)
        self._feature_group_constrain_values.append_id(u)
        self._feature_group_constrain_rates.append_id(u)
        return u

    def _finalize_constraints(self):
        built = []
        for u in self._constraints:
            rod = self._systems[u.id()]
            pos_idx = u._kwargs.get("constrained_position_idx", None)
            dir_idx = u._kwargs.get("constrained_director_idx", None)
            # B-7: copies of the state taken at finalize
            positions = [rod.position_collection[..., i].copy() for i in pos_idx] if pos_idx else []
            directors = [rod.director_collection[..., i].copy() for i in dir_idx] if dir_idx else []
            c = u._cls(*positions, *directors, *u._args, _system=rod, **u._kwargs)
            built.append((rod, c))
            self._feature_group_constrain_values.add_operators(
                u, [lambda time, c=c, rod=rod: c.constrain_values(rod, time)])
            self._feature_group_constrain_rates.add_operators(
                u, [lambda time, c=c, rod=rod: c.constrain_rates(rod, time)])
        # at t=0 constrain everything for compatibility with initial conditions
        for rod, c in built:
            c.constrain_values(rod, 0.0)
        for rod, c in built:
            c.constrain_rates(rod, 0.0)


class Forcing:
    def __init__(self):
        self._ext_forces_torques = []
        super().__init__()
        self._feature_group_finalize.append(self._finalize_forcing)

    def add_forcing_to(self, system):
        u = _Using(self._get_sys_idx_if_valid(system))
        self._ext_forces_torques.append(u)
        self._feature_group_synchronize.append_id(u)
        return u

    def _finalize_forcing(self):
        for u in self._ext_forces_torques:
            f, rod = u._cls(*u._args, **u._kwargs), self._systems[u.id()]
            self._feature_group_synchronize.add_operators(
                u, [lambda time, f=f, rod=rod: f.apply_forces(rod, time),
                    lambda time, f=f, rod=rod: f.apply_torques(rod, time)])


class Connections:
    def __init__(self):
        self._connections = []
        super().__init__()
        self._feature_group_finalize.append(self._finalize_connections)

    def connect(self, first_rod, second_rod, first_connect_idx=None, second_connect_idx=None):
        u = _Using((self._get_sys_idx_if_valid(first_rod), self._get_sys_idx_if_valid(second_rod)))
        u._connect_idx = (first_connect_idx, second_connect_idx)
        self._connections.append(u)
        self._feature_group_synchronize.append_id(u)
        return u

    def _finalize_connections(self):
        for u in self._connections:
            conn = u._cls(*u._args, **u._kwargs)
            s1, s2 = self._systems[u.id()[0]], self._systems[u.id()[1]]
            c1, c2 = u._connect_idx
            self._feature_group_synchronize.add_operators(
                u, [lambda time, conn=conn, s1=s1, c1=c1, s2=s2, c2=c2: conn.apply_forces(s1, c1, s2, c2),
                    lambda time, conn=conn, s1=s1, c1=c1, s2=s2, c2=c2: conn.apply_torques(s1, c1, s2, c2)])


class Damping:
    def __init__(self):
        self._dampers = []
        super().__init__()
        self._feature_group_finalize.append(self._finalize_dampers)

    def dampen(self, system):
        u = _Using(self._get_sys_idx_if_valid(system))
        self._dampers.append(u)
        self._feature_group_constrain_rates.append_id(u)
        return u

    def _finalize_dampers(self):
        for u in self._dampers:
            rod = self._systems[u.id()]
            d = u._cls(*u._args, _system=rod, **u._kwargs)
            self._feature_group_constrain_rates.add_operators(
                u, [lambda time, d=d, rod=rod: d.dampen_rates(rod, time)])


class Contact:
    def __init__(self):
        self._contacts = []
        super().__init__()
        self._feature_group_finalize.append(self._finalize_contact)

    def detect_contact_between(self, first_system, second_system):
        u = _Using((self._get_sys_idx_if_valid(first_system), self._get_sys_idx_if_valid(second_system)))
        self._contacts.append(u)
        self._feature_group_synchronize.append_id(u)
        return u

    def _finalize_contact(self):
        for u in self._contacts:
            c = u._cls(*u._args, **u._kwargs)
            s1, s2 = self._systems[u.id()[0]], self._systems[u.id()[1]]
            self._feature_group_synchronize.add_operators(u, [lambda time, c=c, s1=s1, s2=s2: c.apply_contact(s1, s2)])


class CallBacks:
    def __init__(self):
        self._callback_list = []
        super().__init__()
        self._feature_group_finalize.append(self._finalize_callback)

    def collect_diagnostics(self, system):
        u = _Using(self._get_sys_idx_if_valid(system))
        self._callback_list.append(u)
        self._feature_group_callback.append_id(u)
        return u

    def _finalize_callback(self):
        for u in self._callback_list:
            cb, rod = u._cls(*u._args, **u._kwargs), self._systems[u.id()]
            self._feature_group_callback.add_operators(
                u, [lambda time, current_step, cb=cb, rod=rod: cb.make_callback(rod, time, current_step)])
            cb.make_callback(rod, 0.0, 0)
