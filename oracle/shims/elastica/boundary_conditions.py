"""TEST INFRASTRUCTURE (oracle) — not product code.

ConstraintBase / FreeBC / OneEndFixedBC of PyElastica
(``elastica/boundary_conditions.py``, [PE-recall]).
Subclassed by the reference at `/root/reference/gym_softrobot/envs/soft_pendulum/build.py:65-85`.
"""
import numpy as np


class ConstraintBase:
    def __init__(self, *args, **kwargs):
        self._system = kwargs["_system"]
        self._constrained_position_idx = np.array(kwargs.get("constrained_position_idx", []), dtype=int)
        self._constrained_director_idx = np.array(kwargs.get("constrained_director_idx", []), dtype=int)

    @property
    def system(self):
        return self._system

    @property
    def constrained_position_idx(self):
        return self._constrained_position_idx

    @property
    def constrained_director_idx(self):
        return self._constrained_director_idx

    def constrain_values(self, system, time):
        pass

    def constrain_rates(self, system, time):
        pass


class FreeBC(ConstraintBase):
    pass


class OneEndFixedBC(ConstraintBase):
    def __init__(self, fixed_position, fixed_directors, **kwargs):
        super().__init__(**kwargs)
        self.fixed_position = fixed_position
        self.fixed_directors = fixed_directors

    def constrain_values(self, system, time):
        system.position_collection[..., 0] = self.fixed_position
        system.director_collection[..., 0] = self.fixed_directors

    def constrain_rates(self, system, time):
        system.velocity_collection[..., 0] = 0.0
        system.omega_collection[..., 0] = 0.0
