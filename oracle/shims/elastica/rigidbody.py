"""TEST INFRASTRUCTURE (oracle) — not product code.

`Cylinder` rigid body of PyElastica (``elastica/rigidbody/cylinder.py`` + ``rigid_body.py``,
[PE-recall], parity unpinned).  One node integrated by the same
PositionVerlet as the rods.  Used by `/root/reference/gym_softrobot/envs/octopus/build.py:103-106`.

Inertia: recalled as the rod-style thin-section form  diag(I0 * density * length)  with
I0 = (A^2/4pi, A^2/4pi, A^2/2pi)  (J1 = J2 = m r^2/4, J3 = m r^2/2); SURVEY D.1 recalls
m(3r^2+L^2)/12 for J1,J2.  Every reference env pins w_x = w_y = 0 on its cylinder
(`utils/custom_elastica/constraint.py:83-84`), so only J3 (identical in both) matters there.
"""
import numpy as np

from ._linalg import _batch_matvec, _batch_cross


class RigidBodyBase:
    pass


class _RigidOps:
    """dynamics shared by the rigid bodies: a = F/m, alpha = J^-1 ((J w) x w + T)"""

    def compute_internal_forces_and_torques(self, time=0.0):
        pass

    update_internal_forces_and_torques = compute_internal_forces_and_torques

    def update_accelerations(self, time=0.0):
        np.copyto(self.acceleration_collection, self.external_forces / self.mass)
        J_omega = _batch_matvec(self.mass_second_moment_of_inertia, self.omega_collection)
        lagrangian_transport = _batch_cross(J_omega, self.omega_collection)
        np.copyto(self.alpha_collection,
                  _batch_matvec(self.inv_mass_second_moment_of_inertia,
                                lagrangian_transport + self.external_torques))

    def zeroed_out_external_forces_and_torques(self, time=0.0):
        self.external_forces[:] = 0.0
        self.external_torques[:] = 0.0

    def compute_position_center_of_mass(self):
        return self.position_collection[..., 0].copy()


class Sphere(_RigidOps, RigidBodyBase):
    """``elastica/rigidbody/sphere.py`` ([PE-recall]): J = 2/5 m r^2 I, identity directors.  In the
    reference it is a load-free target marker whose position / velocity the env overwrites every
    substep (`envs/soft_arm/soft_arm_tracking.py:223-224,437-447`)."""

    def __init__(self, center, base_radius, density):
        self.n_elems = 1
        self.n_nodes = 1
        self.radius = base_radius
        self.density = density
        self.length = 2 * base_radius
        self.volume = 4.0 / 3.0 * np.pi * base_radius ** 3
        self.mass = np.array([self.volume * density])
        J = np.zeros((3, 3))
        np.fill_diagonal(J, 2.0 / 5.0 * self.mass[0] * base_radius ** 2)
        self.mass_second_moment_of_inertia = J.reshape(3, 3, 1)
        self.inv_mass_second_moment_of_inertia = np.linalg.inv(J).reshape(3, 3, 1)
        self.position_collection = np.asarray(center, dtype=np.float64).reshape(3, 1).copy()
        self.velocity_collection = np.zeros((3, 1))
        self.omega_collection = np.zeros((3, 1))
        self.acceleration_collection = np.zeros((3, 1))
        self.alpha_collection = np.zeros((3, 1))
        self.director_collection = np.eye(3).reshape(3, 3, 1).copy()
        self.external_forces = np.zeros((3, 1))
        self.external_torques = np.zeros((3, 1))


class Cylinder(RigidBodyBase):
    def __init__(self, start, direction, normal, base_length, base_radius, density):
        start = np.asarray(start, dtype=np.float64)
        direction = np.asarray(direction, dtype=np.float64)
        normal = np.asarray(normal, dtype=np.float64)
        self.n_elems = 1
        self.n_nodes = 1
        self.length = base_length
        self.radius = base_radius
        self.density = density
        A0 = np.pi * base_radius * base_radius
        I0_1 = A0 * A0 / (4.0 * np.pi)
        I0 = np.array([I0_1, I0_1, 2.0 * I0_1])
        self.volume = np.pi * base_radius * base_radius * base_length
        self.mass = np.array([self.volume * density])
        J = np.zeros((3, 3))
        np.fill_diagonal(J, I0 * density * base_length)
        self.mass_second_moment_of_inertia = J.reshape(3, 3, 1)
        self.inv_mass_second_moment_of_inertia = np.linalg.inv(J).reshape(3, 3, 1)
        self.position_collection = np.zeros((3, 1))
        self.position_collection[:, 0] = start + direction * base_length / 2
        self.velocity_collection = np.zeros((3, 1))
        self.omega_collection = np.zeros((3, 1))
        self.acceleration_collection = np.zeros((3, 1))
        self.alpha_collection = np.zeros((3, 1))
        binormal = np.cross(direction, normal)
        self.director_collection = np.zeros((3, 3, 1))
        self.director_collection[0, :, 0] = normal
        self.director_collection[1, :, 0] = binormal
        self.director_collection[2, :, 0] = direction
        self.external_forces = np.zeros((3, 1))
        self.external_torques = np.zeros((3, 1))

    def compute_internal_forces_and_torques(self, time=0.0):
        pass

    update_internal_forces_and_torques = compute_internal_forces_and_torques

    def update_accelerations(self, time=0.0):
        np.copyto(self.acceleration_collection, self.external_forces / self.mass)
        J_omega = _batch_matvec(self.mass_second_moment_of_inertia, self.omega_collection)
        lagrangian_transport = _batch_cross(J_omega, self.omega_collection)
        np.copyto(self.alpha_collection,
                  _batch_matvec(self.inv_mass_second_moment_of_inertia,
                                lagrangian_transport + self.external_torques))

    def zeroed_out_external_forces_and_torques(self, time=0.0):
        self.external_forces[:] = 0.0
        self.external_torques[:] = 0.0

    def compute_position_center_of_mass(self):
        return self.position_collection[..., 0].copy()
