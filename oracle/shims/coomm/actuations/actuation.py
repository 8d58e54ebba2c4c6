"""TEST INFRASTRUCTURE (oracle) — not product code.  PARITY UNPINNED (see oracle/shims/coomm/__init__.py).

`coomm/actuations/actuation.py` as recalled: a continuous actuation owns an internal force (elements, material
frame) and an internal couple (Voronoi points, material frame) and turns them into the loads a Cosserat rod
accepts exactly the way the rod turns its own stress resultants into internal forces / torques, but WITHOUT the
rod's dilatation factors (the muscle model already works with current areas and lengths):

    external_forces  (nodes, lab frame)       = Delta_h( Q^T n_m )
    external_couples (elements, material)     = Delta_h( m_m ) + A_h( kappa x m_m * rest_voronoi_length )
                                                + (Q t e) x n_m * rest_length

Call sites in the reference: `ApplyMuscles(muscles=..., step_skip=..., callback_params_list=...)` registered as
a forcing (/root/reference/gym_softrobot/envs/octopus/build_muscle_octopus.py:165-177,
arm_push_env.py:198-209,601-606).
"""
import numpy as np

from elastica.external_forces import NoForces


def _difference_kernel(a):
    """Delta_h of PyElastica: (3, m) -> (3, m + 1) with ghost zeros at both ends."""
    out = np.zeros((a.shape[0], a.shape[1] + 1))
    out[:, 0] = a[:, 0]
    out[:, 1:-1] = a[:, 1:] - a[:, :-1]
    out[:, -1] = -a[:, -1]
    return out


def _quadrature_kernel(a):
    """A_h of PyElastica (trapezoid with ghost zeros): (3, m) -> (3, m + 1)."""
    out = np.zeros((a.shape[0], a.shape[1] + 1))
    out[:, 0] = 0.5 * a[:, 0]
    out[:, 1:-1] = 0.5 * (a[:, 1:] + a[:, :-1])
    out[:, -1] = 0.5 * a[:, -1]
    return out


def _matvec(Q, v):      # (3,3,n) x (3,n)
    return np.einsum("ijk,jk->ik", Q, v)


def _matTvec(Q, v):
    return np.einsum("jik,jk->ik", Q, v)


def _cross(a, b):
    return np.cross(a, b, axis=0)


class ContinuousActuation:
    def __init__(self, n_elements):
        self.n_elements = n_elements
        self.internal_forces = np.zeros((3, n_elements))
        self.external_forces = np.zeros((3, n_elements + 1))
        self.internal_couples = np.zeros((3, n_elements - 1))
        self.external_couples = np.zeros((3, n_elements))

    def reset_actuation(self):
        self.internal_forces[...] = 0.0
        self.external_forces[...] = 0.0
        self.internal_couples[...] = 0.0
        self.external_couples[...] = 0.0

    def internal_to_external(self, system):
        n_lab = _matTvec(system.director_collection, self.internal_forces)
        self.external_forces[...] = _difference_kernel(n_lab)
        qt = _matvec(system.director_collection, system.tangents * system.dilatation)
        self.external_couples[...] = (
            _difference_kernel(self.internal_couples)
            + _quadrature_kernel(_cross(system.kappa, self.internal_couples) * system.rest_voronoi_lengths)
            + _cross(qt, self.internal_forces) * system.rest_lengths
        )

    def apply(self, system):
        system.external_forces += self.external_forces
        system.external_torques += self.external_couples


class ApplyActuations(NoForces):
    def __init__(self, actuations, step_skip, callback_params_list):
        super().__init__()
        self.actuations = actuations
        self.step_skip = step_skip
        self.callback_params_list = callback_params_list
        self.counter = 0

    def apply_torques(self, system, time=0.0):
        # forces and couples are evaluated together, on the state the rod's own force evaluation just used
        for actuation in self.actuations:
            actuation(system)
            actuation.apply(system)
        self.counter += 1
