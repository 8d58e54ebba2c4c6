"""TEST INFRASTRUCTURE (oracle) — not product code.

NoForces / GravityForces / EndpointForces of PyElastica
(``elastica/external_forces.py``, [PE-recall]); reference call sites
`/root/reference/gym_softrobot/envs/soft_pendulum/build.py:88-105`.
"""
import numpy as np
from ._linalg import _batch_product_i_k_to_ik


class NoForces:
    def __init__(self):
        pass

    def apply_forces(self, system, time=0.0):
        pass

    def apply_torques(self, system, time=0.0):
        pass


class GravityForces(NoForces):
    def __init__(self, acc_gravity=np.array([0.0, -9.80665, 0.0])):
        super().__init__()
        self.acc_gravity = np.asarray(acc_gravity, dtype=np.float64)

    def apply_forces(self, system, time=0.0):
        system.external_forces += _batch_product_i_k_to_ik(self.acc_gravity, system.mass)


class EndpointForces(NoForces):
    def __init__(self, start_force, end_force, ramp_up_time):
        super().__init__()
        self.start_force = np.asarray(start_force, dtype=np.float64)
        self.end_force = np.asarray(end_force, dtype=np.float64)
        assert ramp_up_time > 0.0
        self.ramp_up_time = ramp_up_time

    def apply_forces(self, system, time=0.0):
        factor = min(1.0, time / self.ramp_up_time)
        system.external_forces[..., 0] += self.start_force * factor
        system.external_forces[..., -1] += self.end_force * factor
