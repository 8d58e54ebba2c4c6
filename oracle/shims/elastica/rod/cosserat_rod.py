"""TEST INFRASTRUCTURE (oracle) — not product code.

NumPy restatement of PyElastica's ``CosseratRod`` (``elastica/rod/cosserat_rod.py``
and ``elastica/rod/factory_function.py`` of pyelastica==1.0.0, which is NOT
installable in this environment — [PE-recall], "parity unpinned").
Follows SURVEY.md Appendix A.1 (construction) and A.3 (internal loads) term by
term; call sites in the reference:
`/root/reference/gym_softrobot/envs/soft_pendulum/build.py:54-61`,
`/root/reference/gym_softrobot/envs/soft_pendulum_3d/build.py:55-64`.

Arrays use the reference layout: vectors (3, n), frames (3, 3, n), rows of the
director matrix are d1, d2, d3.
"""
import numpy as np

from .._linalg import _batch_matvec, _batch_cross, _batch_dot, _batch_norm
from .._calculus import (
    position_difference_kernel,
    position_average,
    _difference,
    _trapezoidal,
)
from .._rotations import _inv_rotate


class RodBase:
    """Marker base (reference: `gym_softrobot/utils/render/base_renderer.py:6`)."""

    REQUISITE_MODULES = []


# Convention B-4: default shear modulus when none is passed.
# PyElastica >=0.3 factory: G = E / (2 (1 + nu)) with nu = 0.5 (recalled from the
# factory source text "shear_modulus = youngs_modulus / (2.0 * (1.0 + 0.5))").
# SURVEY.md A.1 recalls E / (1 + nu); both are kept selectable, see DESIGN.md.
DEFAULT_SHEAR_CONVENTION = "E/(2(1+nu))"


def default_shear_modulus(youngs_modulus, convention=None):
    convention = convention or DEFAULT_SHEAR_CONVENTION
    if convention == "E/(2(1+nu))":
        return youngs_modulus / (2.0 * (1.0 + 0.5))
    if convention == "E/(1+nu)":
        return youngs_modulus / (0.5 + 1.0)
    raise ValueError(convention)


class CosseratRod(RodBase):
    def __init__(self, n_elements, position, directors, radius, density, rest_lengths,
                 youngs_modulus, shear_modulus):
        n = n_elements
        self.n_elems = n
        self.n_nodes = n + 1
        self.ring_rod_flag = False
        self.position_collection = position
        self.director_collection = directors
        self.velocity_collection = np.zeros((3, n + 1))
        self.omega_collection = np.zeros((3, n))
        self.acceleration_collection = np.zeros((3, n + 1))
        self.alpha_collection = np.zeros((3, n))
        self.radius = radius
        self.density = density
        self.rest_lengths = rest_lengths
        self.lengths = rest_lengths.copy()
        self.rest_voronoi_lengths = 0.5 * (rest_lengths[1:] + rest_lengths[:-1])

        # A.1: second moments, mass, stiffness
        A0 = np.pi * radius * radius
        I0_1 = A0 * A0 / (4.0 * np.pi)
        I0_2 = I0_1
        I0_3 = 2.0 * I0_2
        I0 = np.array([I0_1, I0_2, I0_3]).transpose()  # (n, 3)
        msmoi = np.einsum("ij,i->ij", I0, density * rest_lengths)
        self.mass_second_moment_of_inertia = np.zeros((3, 3, n))
        self.inv_mass_second_moment_of_inertia = np.zeros((3, 3, n))
        for k in range(n):
            np.fill_diagonal(self.mass_second_moment_of_inertia[..., k], msmoi[k, :])
            self.inv_mass_second_moment_of_inertia[..., k] = np.linalg.inv(
                self.mass_second_moment_of_inertia[..., k]
            )
        alpha_c = 27.0 / 28.0
        self.shear_matrix = np.zeros((3, 3, n))
        bend = np.zeros((3, 3, n))
        for k in range(n):
            np.fill_diagonal(
                self.shear_matrix[..., k],
                [alpha_c * shear_modulus * A0[k], alpha_c * shear_modulus * A0[k],
                 youngs_modulus * A0[k]],
            )
            np.fill_diagonal(
                bend[..., k],
                [youngs_modulus * I0_1[k], youngs_modulus * I0_2[k], shear_modulus * I0_3[k]],
            )
        self.bend_matrix = (bend[..., 1:] * rest_lengths[1:] + bend[..., :-1] * rest_lengths[:-1]) / (
            rest_lengths[1:] + rest_lengths[:-1]
        )
        self.volume = np.pi * radius ** 2 * rest_lengths
        self.mass = np.zeros(n + 1)
        self.mass[:-1] += 0.5 * density * self.volume
        self.mass[1:] += 0.5 * density * self.volume

        self.internal_forces = np.zeros((3, n + 1))
        self.internal_torques = np.zeros((3, n))
        self.external_forces = np.zeros((3, n + 1))
        self.external_torques = np.zeros((3, n))
        self.tangents = np.zeros((3, n))
        self.dilatation = np.zeros(n)
        self.voronoi_dilatation = np.zeros(n - 1)
        self.dilatation_rate = np.zeros(n)
        self.sigma = np.zeros((3, n))
        self.kappa = np.zeros((3, n - 1))
        self.rest_sigma = np.zeros((3, n))
        self.rest_kappa = np.zeros((3, n - 1))
        self.internal_stress = np.zeros((3, n))
        self.internal_couple = np.zeros((3, n - 1))

        # strains are evaluated once at construction, so `rod.tangents` is valid at reset (A.1/A.6)
        self._compute_shear_stretch_strains()
        self._compute_bending_twist_strains()

    # ------------------------------------------------------------------ factory
    @classmethod
    def straight_rod(cls, n_elements, start, direction, normal, base_length, base_radius,
                     density, *, youngs_modulus, shear_modulus=None, **kwargs):
        if "poisson_ratio" in kwargs:
            raise NameError("Poisson's ratio is deprecated for Cosserat Rod; give shear_modulus")
        n = int(n_elements)
        start = np.asarray(start, dtype=np.float64)
        direction = np.asarray(direction, dtype=np.float64)
        normal = np.asarray(normal, dtype=np.float64).copy()
        end = start + direction * base_length
        position = np.zeros((3, n + 1))
        for i in range(3):
            position[i, ...] = np.linspace(start[i], end[i], n + 1)
        position_diff = position[..., 1:] - position[..., :-1]
        rest_lengths = _batch_norm(position_diff)
        tangents = position_diff / rest_lengths
        normal /= np.linalg.norm(normal)
        directors = np.zeros((3, 3, n))
        normal_collection = np.repeat(normal[:, np.newaxis], n, axis=1)
        assert np.allclose(_batch_dot(normal_collection, tangents), 0.0, atol=1e-8), (
            " Rod normal and tangent are not perpendicular to each other!"
        )
        directors[0, ...] = normal_collection
        directors[1, ...] = _batch_cross(tangents, normal_collection)
        directors[2, ...] = tangents
        radius = np.zeros(n)
        radius[:] = np.array(base_radius)
        density_array = np.zeros(n)
        density_array[:] = np.array(density)
        if not shear_modulus:
            shear_modulus = default_shear_modulus(youngs_modulus)
        return cls(n, position, directors, radius, density_array, rest_lengths,
                   youngs_modulus, shear_modulus)

    # ------------------------------------------------------------------ A.3
    def _compute_geometry_from_state(self):
        position_diff = position_difference_kernel(self.position_collection)
        self.lengths[:] = _batch_norm(position_diff) + 1e-14
        self.tangents[:] = position_diff / self.lengths
        self.radius[:] = np.sqrt(self.volume / self.lengths / np.pi)

    def _compute_all_dilatations(self):
        self._compute_geometry_from_state()
        self.dilatation[:] = self.lengths / self.rest_lengths
        voronoi_lengths = position_average(self.lengths)
        self.voronoi_dilatation[:] = voronoi_lengths / self.rest_voronoi_lengths

    def _compute_dilatation_rate(self):
        x, v = self.position_collection, self.velocity_collection
        r_dot_v = _batch_dot(x, v)
        r_plus_one_dot_v = _batch_dot(x[..., 1:], v[..., :-1])
        r_dot_v_plus_one = _batch_dot(x[..., :-1], v[..., 1:])
        self.dilatation_rate[:] = (
            (r_dot_v[:-1] + r_dot_v[1:] - r_dot_v_plus_one - r_plus_one_dot_v)
            / self.lengths / self.rest_lengths
        )

    def _compute_shear_stretch_strains(self):
        self._compute_all_dilatations()
        z_vector = np.array([0.0, 0.0, 1.0]).reshape(3, -1)
        self.sigma[:] = self.dilatation * _batch_matvec(self.director_collection, self.tangents) - z_vector

    def _compute_internal_shear_stretch_stresses_from_model(self):
        self._compute_shear_stretch_strains()
        self.internal_stress[:] = _batch_matvec(self.shear_matrix, self.sigma - self.rest_sigma)

    def _compute_internal_forces(self):
        self._compute_internal_shear_stretch_stresses_from_model()
        Q = self.director_collection
        cosserat_internal_stress = np.zeros((3, self.n_elems))
        for i in range(3):
            for j in range(3):
                cosserat_internal_stress[i] += Q[j, i] * self.internal_stress[j]
        cosserat_internal_stress /= self.dilatation
        self.internal_forces[:] = _difference(cosserat_internal_stress)

    def _compute_bending_twist_strains(self):
        temp = _inv_rotate(self.director_collection)
        self.kappa[:] = temp / self.rest_voronoi_lengths

    def _compute_internal_bending_twist_stresses_from_model(self):
        self._compute_bending_twist_strains()
        self.internal_couple[:] = _batch_matvec(self.bend_matrix, self.kappa - self.rest_kappa)

    def _compute_internal_torques(self):
        self._compute_internal_bending_twist_stresses_from_model()
        self._compute_dilatation_rate()
        voronoi_dilatation_inv_cube_cached = 1.0 / self.voronoi_dilatation ** 3
        bend_twist_couple_2D = _difference(self.internal_couple * voronoi_dilatation_inv_cube_cached)
        bend_twist_couple_3D = _trapezoidal(
            _batch_cross(self.kappa, self.internal_couple)
            * self.rest_voronoi_lengths
            * voronoi_dilatation_inv_cube_cached
        )
        shear_stretch_couple = (
            _batch_cross(_batch_matvec(self.director_collection, self.tangents), self.internal_stress)
            * self.rest_lengths
        )
        J_omega_upon_e = (
            _batch_matvec(self.mass_second_moment_of_inertia, self.omega_collection) / self.dilatation
        )
        lagrangian_transport = _batch_cross(J_omega_upon_e, self.omega_collection)
        unsteady_dilatation = J_omega_upon_e * self.dilatation_rate / self.dilatation
        self.internal_torques[:] = (
            bend_twist_couple_2D
            + bend_twist_couple_3D
            + shear_stretch_couple
            + lagrangian_transport
            + unsteady_dilatation
        )

    def compute_internal_forces_and_torques(self, time=0.0):
        self._compute_internal_forces()
        self._compute_internal_torques()

    update_internal_forces_and_torques = compute_internal_forces_and_torques

    def update_accelerations(self, time=0.0):
        self.acceleration_collection[:] = (self.internal_forces + self.external_forces) / self.mass
        self.alpha_collection[:] = (
            _batch_matvec(self.inv_mass_second_moment_of_inertia,
                          self.internal_torques + self.external_torques)
            * self.dilatation
        )

    def zeroed_out_external_forces_and_torques(self, time=0.0):
        self.external_forces[:] = 0.0
        self.external_torques[:] = 0.0

    def compute_position_center_of_mass(self):
        mass_times_position = self.mass * self.position_collection
        return mass_times_position.sum(axis=1) / self.mass.sum()

    def compute_velocity_center_of_mass(self):
        return (self.mass * self.velocity_collection).sum(axis=1) / self.mass.sum()

    # -- energies (PyElastica `CosseratRod.compute_*_energy`, [PE-recall]; used by the reference's
    #    `ArmPushEnv.cal_desired_Hamiltonian`, /root/reference/gym_softrobot/envs/octopus/arm_push_env.py:442-452)
    def compute_translational_energy(self):
        v = self.velocity_collection
        return (0.5 * (self.mass * np.einsum("ij,ij->j", v, v)).sum())

    def compute_rotational_energy(self):
        j_omega_upon_e = _batch_matvec(self.mass_second_moment_of_inertia, self.omega_collection) / self.dilatation
        return 0.5 * np.einsum("ik,ik->k", self.omega_collection, j_omega_upon_e).sum()

    def compute_bending_energy(self):
        kappa_diff = self.kappa - self.rest_kappa
        bending_internal_torques = _batch_matvec(self.bend_matrix, kappa_diff)
        return 0.5 * (np.einsum("ik,ik->k", kappa_diff, bending_internal_torques) * self.rest_voronoi_lengths).sum()

    def compute_shear_energy(self):
        strain_diff = self.sigma - self.rest_sigma
        shear_internal_forces = _batch_matvec(self.shear_matrix, strain_diff)
        return 0.5 * (np.einsum("ik,ik->k", strain_diff, shear_internal_forces) * self.rest_lengths).sum()
