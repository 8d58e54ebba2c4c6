"""TEST INFRASTRUCTURE (oracle) — not product code.  PARITY UNPINNED (see oracle/shims/coomm/__init__.py).

`coomm/actuations/muscles/muscle.py` as recalled (published model, Chang et al. 2023 section 2): a muscle m sits
at material-frame offset x_m(s) from the centre line; with nu = sigma + e3 and kappa averaged onto the elements

    muscle strain    nu_m = nu + kappa x x_m + d x_m / ds
    muscle length    l_m  = |nu_m|            (normalised by the rest value 1)
    muscle tangent   t_m  = nu_m / |nu_m|
    muscle area      A_m  = rest_area / e     (incompressible)
    force            n_m  = u * sigma_max * A_m * h(l_m) * t_m,   couple  m_m = x_m x n_m
    h(l) = max(3.06 l^3 - 13.64 l^2 + 18.01 l - 6.44, 0)

The reference only ever calls `apply_activation(scalar)` (crawl_env.py:242, arm_push_env.py:257-271,
arm_two_env.py:243-245, reach_env.py:223); the scalar is broadcast over the activation array.
"""
import numpy as np

from ..actuation import ContinuousActuation, ApplyActuations, _cross

F_L_COEFFICIENTS = np.array([-6.44, 18.01, -13.64, 3.06])   # ascending powers of the normalised length


def force_length_weight_poly(muscle_length, f_l_coefficients=F_L_COEFFICIENTS):
    w = np.zeros_like(muscle_length)
    for p in range(len(f_l_coefficients) - 1, -1, -1):     # Horner, highest power first
        w = w * muscle_length + f_l_coefficients[p]
    return np.maximum(w, 0.0)


class MuscleForce(ContinuousActuation):
    def __init__(self, ratio_muscle_position, rest_muscle_area, max_muscle_stress, type_name="muscle",
                 index=0, force_length_weight=force_length_weight_poly, **kwargs):
        n_elements = np.asarray(rest_muscle_area).shape[0]
        super().__init__(n_elements)
        self.type_name, self.index = type_name, index
        self.ratio_muscle_position = np.array(ratio_muscle_position, dtype=float)
        self.rest_muscle_area = np.array(rest_muscle_area, dtype=float)
        self.max_muscle_stress = float(max_muscle_stress)
        self.force_length_weight = force_length_weight
        self.activation = np.zeros(n_elements)
        self.muscle_area = self.rest_muscle_area.copy()
        self.muscle_position = np.zeros((3, n_elements))
        self.muscle_strain = np.zeros((3, n_elements))
        self.muscle_tangent = np.zeros((3, n_elements))
        self.muscle_length = np.ones(n_elements)
        self.muscle_force = np.zeros((3, n_elements))

    # -- activation ---------------------------------------------------------------------------------
    def apply_activation(self, activation):
        self.set_activation(activation)

    def set_activation(self, activation):
        self.activation[...] = activation

    def get_activation(self):
        return self.activation

    # -- geometry -----------------------------------------------------------------------------------
    def calculate_muscle_length(self):
        self.muscle_length[...] = np.sqrt(np.einsum("ik,ik->k", self.muscle_strain, self.muscle_strain))

    def __call__(self, system):
        self.reset_actuation()
        self.muscle_area[...] = self.rest_muscle_area / system.dilatation
        self.muscle_position[...] = self.ratio_muscle_position * system.radius
        nu = system.sigma.copy()
        nu[2] += 1.0
        n = self.n_elements
        kappa_e = np.zeros((3, n))                         # A_h without the half weights at the ends? no: plain trapezoid
        kappa_e[:, 0] = 0.5 * system.kappa[:, 0]
        kappa_e[:, 1:-1] = 0.5 * (system.kappa[:, 1:] + system.kappa[:, :-1])
        kappa_e[:, -1] = 0.5 * system.kappa[:, -1]
        dpos = np.zeros((3, n))                            # d x_m / ds on the elements (central, one-sided at the ends)
        if n > 2:
            s = np.cumsum(system.rest_lengths) - 0.5 * system.rest_lengths
            dpos[:, 1:-1] = (self.muscle_position[:, 2:] - self.muscle_position[:, :-2]) / (s[2:] - s[:-2])
            dpos[:, 0] = (self.muscle_position[:, 1] - self.muscle_position[:, 0]) / (s[1] - s[0])
            dpos[:, -1] = (self.muscle_position[:, -1] - self.muscle_position[:, -2]) / (s[-1] - s[-2])
        self.muscle_strain[...] = nu + _cross(kappa_e, self.muscle_position) + dpos
        self.calculate_muscle_length()
        norm = np.sqrt(np.einsum("ik,ik->k", self.muscle_strain, self.muscle_strain))
        self.muscle_tangent[...] = self.muscle_strain / norm
        weight = self.force_length_weight(self.muscle_length)
        magnitude = self.activation * self.max_muscle_stress * self.muscle_area * weight
        self.muscle_force[...] = magnitude * self.muscle_tangent
        self.internal_forces[...] = self.muscle_force
        couple_e = _cross(self.muscle_position, self.muscle_force)
        self.internal_couples[...] = 0.5 * (couple_e[:, 1:] + couple_e[:, :-1])
        self.internal_to_external(system)


class ApplyMuscles(ApplyActuations):
    def __init__(self, muscles, step_skip, callback_params_list):
        super().__init__(muscles, step_skip, callback_params_list)
        for m, muscle in enumerate(self.actuations):
            muscle.index = m
