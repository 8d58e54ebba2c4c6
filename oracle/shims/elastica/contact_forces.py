"""TEST INFRASTRUCTURE (oracle) — not product code.

Plane surface and rod-plane contact with anisotropic friction, restated from PyElastica
(``elastica/surface/plane.py``, ``elastica/contact_forces.py``, ``elastica/_contact_functions.py``:
`_calculate_contact_forces_rod_plane`, `_calculate_contact_forces_rod_plane_with_anisotropic_friction`,
``elastica/contact_utils.py``/``interaction.py`` helpers) — [PE-recall], parity unpinned; SURVEY.md A.5.
Reference call sites: `/root/reference/gym_softrobot/envs/octopus/build.py:192-200,276-283`.
"""
import numpy as np

from ._linalg import _batch_matvec, _batch_cross, _batch_dot, _batch_norm


class SurfaceBase:
    pass


class Plane(SurfaceBase):
    def __init__(self, plane_origin, plane_normal):
        self.origin = np.asarray(plane_origin, dtype=np.float64).reshape(3, 1)
        self.normal = np.asarray(plane_normal, dtype=np.float64).reshape(3)
        self.normal = self.normal / np.linalg.norm(self.normal)


def node_to_element_mass_or_force(inp):
    out = np.zeros((3, inp.shape[1] - 1))
    out += 0.5 * (inp[:, :-1] + inp[:, 1:])
    out[:, 0] += 0.5 * inp[:, 0]
    out[:, -1] += 0.5 * inp[:, -1]
    return out


def node_to_element_position(x):
    return 0.5 * (x[:, :-1] + x[:, 1:])


def node_to_element_velocity(mass, v):
    out = mass[1:] * v[:, 1:] + mass[:-1] * v[:, :-1]
    out /= mass[1:] + mass[:-1]
    return out


def elements_to_nodes_inplace(vec_elem, vec_node):
    vec_node[:, :-1] += 0.5 * vec_elem
    vec_node[:, 1:] += 0.5 * vec_elem


def find_slipping_elements(velocity_slip, velocity_threshold):
    abs_velocity_slip = _batch_norm(velocity_slip)
    slip_points = np.where(np.fabs(abs_velocity_slip) > velocity_threshold)
    slip_function = np.ones(velocity_slip.shape[1])
    slip_function[slip_points] = np.fabs(
        1.0 - np.minimum(1.0, abs_velocity_slip[slip_points] / velocity_threshold - 1.0)
    )
    return slip_function


def _calculate_contact_forces_rod_plane(plane_origin, plane_normal, surface_tol, k, nu, radius, mass,
                                        position_collection, velocity_collection, internal_forces,
                                        external_forces):
    nodal_total_forces = internal_forces + external_forces
    element_total_forces = node_to_element_mass_or_force(nodal_total_forces)
    force_component_along_normal_direction = np.einsum("i,ik->k", plane_normal, element_total_forces)
    forces_along_normal_direction = np.einsum("i,k->ik", plane_normal, force_component_along_normal_direction)
    forces_along_normal_direction[..., np.where(force_component_along_normal_direction > 0)[0]] = 0.0
    plane_response_force = -forces_along_normal_direction
    element_position = node_to_element_position(position_collection)
    distance_from_plane = np.einsum("i,ik->k", plane_normal, (element_position - plane_origin))
    plane_penetration = np.minimum(distance_from_plane - radius, 0.0)
    elastic_force = -k * np.einsum("i,k->ik", plane_normal, plane_penetration)
    element_velocity = node_to_element_velocity(mass, velocity_collection)
    normal_component_of_element_velocity = np.einsum("i,ik->k", plane_normal, element_velocity)
    damping_force = -nu * np.einsum("i,k->ik", plane_normal, normal_component_of_element_velocity)
    plane_response_force_total = plane_response_force + elastic_force + damping_force
    no_contact_point_idx = np.where((distance_from_plane - radius) > surface_tol)[0]
    plane_response_force[..., no_contact_point_idx] = 0.0
    plane_response_force_total[..., no_contact_point_idx] = 0.0
    elements_to_nodes_inplace(plane_response_force_total, external_forces)
    return _batch_norm(plane_response_force), no_contact_point_idx


def _calculate_contact_forces_rod_plane_with_anisotropic_friction(
        plane_origin, plane_normal, surface_tol, slip_velocity_tol, k, nu,
        kinetic_mu_forward, kinetic_mu_backward, kinetic_mu_sideways,
        static_mu_forward, static_mu_backward, static_mu_sideways,
        radius, mass, tangents, position_collection, director_collection, velocity_collection,
        omega_collection, internal_forces, external_forces, internal_torques, external_torques):
    plane_response_force_mag, no_contact_point_idx = _calculate_contact_forces_rod_plane(
        plane_origin, plane_normal, surface_tol, k, nu, radius, mass, position_collection,
        velocity_collection, internal_forces, external_forces)

    tangent_along_normal_direction = np.einsum("i,ik->k", plane_normal, tangents)
    tangent_perpendicular_to_normal_direction = tangents - np.einsum(
        "i,k->ik", plane_normal, tangent_along_normal_direction)
    tangent_perpendicular_to_normal_direction_mag = _batch_norm(tangent_perpendicular_to_normal_direction)
    axial_direction = (1 / (tangent_perpendicular_to_normal_direction_mag + 1e-14)) * \
        tangent_perpendicular_to_normal_direction
    element_velocity = node_to_element_velocity(mass, velocity_collection)
    # axial kinetic friction
    velocity_mag_along_axial_direction = _batch_dot(element_velocity, axial_direction)
    velocity_along_axial_direction = velocity_mag_along_axial_direction * axial_direction
    velocity_sign_along_axial_direction = np.sign(velocity_mag_along_axial_direction)
    kinetic_mu = 0.5 * (kinetic_mu_forward * (1 + velocity_sign_along_axial_direction)
                        + kinetic_mu_backward * (1 - velocity_sign_along_axial_direction))
    slip_function_along_axial_direction = find_slipping_elements(velocity_along_axial_direction,
                                                                 slip_velocity_tol)
    # rolling kinetic friction
    rolling_direction = np.cross(axial_direction.T, plane_normal).T
    torque_arm = np.einsum("i,k->ik", -plane_normal, radius)
    velocity_along_rolling_direction = _batch_dot(element_velocity, rolling_direction)
    directors_transpose = np.transpose(director_collection, (1, 0, 2))
    rotation_velocity = _batch_matvec(
        directors_transpose, _batch_cross(omega_collection, _batch_matvec(director_collection, torque_arm)))
    rotation_velocity_along_rolling_direction = _batch_dot(rotation_velocity, rolling_direction)
    slip_velocity_mag_along_rolling_direction = (velocity_along_rolling_direction
                                                 + rotation_velocity_along_rolling_direction)
    slip_velocity_along_rolling_direction = slip_velocity_mag_along_rolling_direction * rolling_direction
    slip_function_along_rolling_direction = find_slipping_elements(slip_velocity_along_rolling_direction,
                                                                   slip_velocity_tol)
    unitized_total_velocity = slip_velocity_along_rolling_direction + velocity_along_axial_direction
    unitized_total_velocity /= _batch_norm(unitized_total_velocity + 1e-14)
    kinetic_friction_force_along_axial_direction = -(
        (1.0 - slip_function_along_axial_direction) * kinetic_mu * plane_response_force_mag
        * _batch_dot(unitized_total_velocity, axial_direction) * axial_direction)
    kinetic_friction_force_along_axial_direction[..., no_contact_point_idx] = 0.0
    elements_to_nodes_inplace(kinetic_friction_force_along_axial_direction, external_forces)
    kinetic_friction_force_along_rolling_direction = -(
        (1.0 - slip_function_along_rolling_direction) * kinetic_mu_sideways * plane_response_force_mag
        * _batch_dot(unitized_total_velocity, rolling_direction) * rolling_direction)
    kinetic_friction_force_along_rolling_direction[..., no_contact_point_idx] = 0.0
    elements_to_nodes_inplace(kinetic_friction_force_along_rolling_direction, external_forces)
    external_torques += _batch_matvec(
        director_collection, _batch_cross(torque_arm, kinetic_friction_force_along_rolling_direction))

    # axial static friction (forces re-collected, now including the responses added above)
    nodal_total_forces = internal_forces + external_forces
    element_total_forces = node_to_element_mass_or_force(nodal_total_forces)
    force_component_along_axial_direction = _batch_dot(element_total_forces, axial_direction)
    force_component_sign_along_axial_direction = np.sign(force_component_along_axial_direction)
    static_mu = 0.5 * (static_mu_forward * (1 + force_component_sign_along_axial_direction)
                       + static_mu_backward * (1 - force_component_sign_along_axial_direction))
    max_friction_force = slip_function_along_axial_direction * static_mu * plane_response_force_mag
    static_friction_force_along_axial_direction = -(
        np.minimum(np.fabs(force_component_along_axial_direction), max_friction_force)
        * force_component_sign_along_axial_direction * axial_direction)
    static_friction_force_along_axial_direction[..., no_contact_point_idx] = 0.0
    elements_to_nodes_inplace(static_friction_force_along_axial_direction, external_forces)

    # rolling static friction
    total_torques = _batch_matvec(directors_transpose, (internal_torques + external_torques))
    total_torques_along_axial_direction = _batch_dot(total_torques, axial_direction)
    force_component_along_rolling_direction = _batch_dot(element_total_forces, rolling_direction)
    noslip_force = -((radius * force_component_along_rolling_direction
                      - 2.0 * total_torques_along_axial_direction) / 3.0 / radius)
    max_friction_force = slip_function_along_rolling_direction * static_mu_sideways * plane_response_force_mag
    noslip_force_sign = np.sign(noslip_force)
    static_friction_force_along_rolling_direction = (
        np.minimum(np.fabs(noslip_force), max_friction_force) * noslip_force_sign * rolling_direction)
    static_friction_force_along_rolling_direction[..., no_contact_point_idx] = 0.0
    elements_to_nodes_inplace(static_friction_force_along_rolling_direction, external_forces)
    external_torques += _batch_matvec(
        director_collection, _batch_cross(torque_arm, static_friction_force_along_rolling_direction))


class NoContact:
    def __init__(self):
        pass

    def apply_contact(self, system_one, system_two):
        pass


class RodPlaneContactWithAnisotropicFriction(NoContact):
    def __init__(self, k, nu, slip_velocity_tol, static_mu_array, kinetic_mu_array):
        super().__init__()
        self.k, self.nu = k, nu
        self.surface_tol = 1e-4
        self.slip_velocity_tol = slip_velocity_tol
        (self.static_mu_forward, self.static_mu_backward, self.static_mu_sideways) = static_mu_array
        (self.kinetic_mu_forward, self.kinetic_mu_backward, self.kinetic_mu_sideways) = kinetic_mu_array

    def apply_contact(self, system_one, system_two):
        rod, plane = system_one, system_two
        _calculate_contact_forces_rod_plane_with_anisotropic_friction(
            plane.origin, plane.normal, self.surface_tol, self.slip_velocity_tol, self.k, self.nu,
            self.kinetic_mu_forward, self.kinetic_mu_backward, self.kinetic_mu_sideways,
            self.static_mu_forward, self.static_mu_backward, self.static_mu_sideways,
            rod.radius, rod.mass, rod.tangents, rod.position_collection, rod.director_collection,
            rod.velocity_collection, rod.omega_collection, rod.internal_forces, rod.external_forces,
            rod.internal_torques, rod.external_torques)
