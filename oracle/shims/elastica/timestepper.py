"""TEST INFRASTRUCTURE (oracle) — not product code.

PositionVerlet of PyElastica (``elastica/timestepper/symplectic_steppers.py``,
[PE-recall], parity unpinned) following SURVEY.md Appendix A.2 line by line.
Boundary in the reference: `time = PositionVerlet().step(simulator, time, dt)`
(`/root/reference/gym_softrobot/envs/soft_pendulum/soft_pendulum.py:137-139,184`).
"""
import numpy as np
from ._rotations import _get_rotation_matrix
from ._linalg import _batch_matmul


def _kinematic_step(system, prefac):
    # x += prefac*v ; Q <- R(prefac*omega) Q          (A.2 line 1 / 8, A.2.1)
    system.position_collection += prefac * system.velocity_collection
    rot = _get_rotation_matrix(1.0, prefac * system.omega_collection)
    system.director_collection[:] = _batch_matmul(rot, system.director_collection)


def _dynamic_step(system, time, dt):
    # v += dt*a ; omega += dt*alpha                   (A.2 line 5)
    system.update_accelerations(time)
    system.velocity_collection += dt * system.acceleration_collection
    system.omega_collection += dt * system.alpha_collection


class PositionVerlet:
    def step(self, SystemCollection, time, dt):
        return self.do_step(SystemCollection, time, dt)

    @staticmethod
    def do_step(SystemCollection, time, dt):
        prefac = 0.5 * dt
        for system in SystemCollection.block_systems():
            _kinematic_step(system, prefac)
        time += prefac
        SystemCollection.constrain_values(time)
        for system in SystemCollection.block_systems():
            system.compute_internal_forces_and_torques(time)
        SystemCollection.synchronize(time)
        for system in SystemCollection.block_systems():
            _dynamic_step(system, time, dt)
        SystemCollection.constrain_rates(time)  # contains dampen_rates, see modules.py
        for system in SystemCollection.block_systems():
            _kinematic_step(system, prefac)
        time += prefac
        SystemCollection.constrain_values(time)
        SystemCollection.apply_callbacks(time, round(time / dt))
        for system in SystemCollection.block_systems():
            system.zeroed_out_external_forces_and_torques(time)
        return time


def extend_stepper_interface(stepper, system_collection):
    return stepper.do_step, None


def integrate(stepper, systems, final_time, n_steps=1000, **kwargs):
    dt = np.float64(float(final_time) / n_steps)
    time = np.float64(0.0)
    for _ in range(n_steps):
        time = stepper.step(systems, time, dt)
    return time
