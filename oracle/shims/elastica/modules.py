"""TEST INFRASTRUCTURE (oracle) — not product code.

Restatement of PyElastica's system-collection mixins (``elastica/modules/*.py``,
[PE-recall], parity unpinned): BaseSystemCollection, Constraints, Connections,
Forcing, Damping, Contact, CallBacks.  The reference composes them at
`/root/reference/gym_softrobot/envs/soft_pendulum/soft_pendulum.py:34-42`.

Ordering rule reproduced from the recalled source: every mixin calls
``super().__init__()`` *before* appending its operators to the feature groups,
so operators register in reverse-MRO order.  For
SoftPendulum-v0 (Constraints, Connections, Forcing, Damping, CallBacks) this
gives  synchronize = [forcing, connections]  and
constrain_rates = [dampen_rates, constrain_rates];  SoftPendulum-v0 is
insensitive to either order (its BC only zeroes components).
There is no memory block here: systems are stepped one by one (B-8).
"""
import numpy as np


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
        self._feature_group_synchronize = []
        self._feature_group_constrain_values = []
        self._feature_group_constrain_rates = []
        self._feature_group_callback = []
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
        self._feature_group_constrain_values.append(self._constrain_values)
        self._feature_group_constrain_rates.append(self._constrain_rates)
        self._feature_group_finalize.append(self._finalize_constraints)

    def constrain(self, system):
        u = _Using(self._get_sys_idx_if_valid(system))
        self._constraints.append(u)
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
            built.append((u.id(), u._cls(*positions, *directors, *u._args, _system=rod, **u._kwargs)))
        built.sort(key=lambda t: t[0])  # stable: registration order within a system
        self._constraints = built
        # at t=0 constrain everything for compatibility with initial conditions
        self._constrain_values(0.0)
        self._constrain_rates(0.0)

    def _constrain_values(self, time):
        for idx, c in self._constraints:
            c.constrain_values(self._systems[idx], time)

    def _constrain_rates(self, time):
        for idx, c in self._constraints:
            c.constrain_rates(self._systems[idx], time)


class Forcing:
    def __init__(self):
        self._ext_forces_torques = []
        super().__init__()
        self._feature_group_synchronize.append(self._call_ext_forces_torques)
        self._feature_group_finalize.append(self._finalize_forcing)

    def add_forcing_to(self, system):
        u = _Using(self._get_sys_idx_if_valid(system))
        self._ext_forces_torques.append(u)
        return u

    def _finalize_forcing(self):
        built = [(u.id(), u._cls(*u._args, **u._kwargs)) for u in self._ext_forces_torques]
        built.sort(key=lambda t: t[0])
        self._ext_forces_torques = built

    def _call_ext_forces_torques(self, time):
        for idx, f in self._ext_forces_torques:
            f.apply_forces(self._systems[idx], time)
            f.apply_torques(self._systems[idx], time)


class Connections:
    def __init__(self):
        self._connections = []
        super().__init__()
        self._feature_group_synchronize.append(self._call_connections)
        self._feature_group_finalize.append(self._finalize_connections)

    def connect(self, first_rod, second_rod, first_connect_idx=None, second_connect_idx=None):
        u = _Using((self._get_sys_idx_if_valid(first_rod), self._get_sys_idx_if_valid(second_rod)))
        u._connect_idx = (first_connect_idx, second_connect_idx)
        self._connections.append(u)
        return u

    def _finalize_connections(self):
        built = []
        for u in self._connections:
            built.append((u.id()[0], u.id()[1], u._connect_idx[0], u._connect_idx[1],
                          u._cls(*u._args, **u._kwargs)))
        self._connections = built

    def _call_connections(self, time):
        for i1, i2, c1, c2, conn in self._connections:
            conn.apply_forces(self._systems[i1], c1, self._systems[i2], c2)
            conn.apply_torques(self._systems[i1], c1, self._systems[i2], c2)


class Damping:
    def __init__(self):
        self._dampers = []
        super().__init__()
        self._feature_group_constrain_rates.append(self._dampen_rates)
        self._feature_group_finalize.append(self._finalize_dampers)

    def dampen(self, system):
        u = _Using(self._get_sys_idx_if_valid(system))
        self._dampers.append(u)
        return u

    def _finalize_dampers(self):
        built = [(u.id(), u._cls(*u._args, _system=self._systems[u.id()], **u._kwargs))
                 for u in self._dampers]
        built.sort(key=lambda t: t[0])
        self._dampers = built

    def _dampen_rates(self, time):
        for idx, d in self._dampers:
            d.dampen_rates(self._systems[idx], time)


class Contact:
    def __init__(self):
        self._contacts = []
        super().__init__()
        self._feature_group_synchronize.append(self._call_contacts)
        self._feature_group_finalize.append(self._finalize_contact)

    def detect_contact_between(self, first_system, second_system):
        u = _Using((self._get_sys_idx_if_valid(first_system), self._get_sys_idx_if_valid(second_system)))
        self._contacts.append(u)
        return u

    def _finalize_contact(self):
        self._contacts = [(u.id()[0], u.id()[1], u._cls(*u._args, **u._kwargs)) for u in self._contacts]

    def _call_contacts(self, time):
        for i1, i2, c in self._contacts:
            c.apply_contact(self._systems[i1], self._systems[i2])


class CallBacks:
    def __init__(self):
        self._callback_list = []
        super().__init__()
        self._feature_group_callback.append(self._callback_execution)
        self._feature_group_finalize.append(self._finalize_callback)

    def collect_diagnostics(self, system):
        u = _Using(self._get_sys_idx_if_valid(system))
        self._callback_list.append(u)
        return u

    def _finalize_callback(self):
        built = [(u.id(), u._cls(*u._args, **u._kwargs)) for u in self._callback_list]
        built.sort(key=lambda t: t[0])
        self._callback_list = built
        for idx, cb in self._callback_list:
            cb.make_callback(self._systems[idx], 0.0, 0)

    def _callback_execution(self, time, current_step):
        for idx, cb in self._callback_list:
            cb.make_callback(self._systems[idx], time, current_step)
