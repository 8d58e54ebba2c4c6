"""TEST INFRASTRUCTURE (oracle) — not product code.

Difference / quadrature stencils of PyElastica ``elastica/_calculus.py``
([PE-recall]) and the NaN guard the reference envs call
(`/root/reference/gym_softrobot/envs/soft_pendulum/soft_pendulum.py:9,196`).
"""
import numpy as np


def _isnan_check(array) -> bool:
    # reference semantic (SURVEY B-11): any NaN anywhere
    return bool(np.isnan(array).any())


def position_difference_kernel(vector):
    return vector[..., 1:] - vector[..., :-1]


def position_average(vector):
    return 0.5 * (vector[..., 1:] + vector[..., :-1])


def _difference(vector):
    """Delta_h: out_0 = in_0 ; out_j = in_j - in_{j-1} ; out_n = -in_{n-1}."""
    blocksize = vector.shape[1]
    out = np.zeros((3, blocksize + 1))
    out[:, 0] = vector[:, 0]
    out[:, -1] = -vector[:, -1]
    out[:, 1:-1] = vector[:, 1:] - vector[:, :-1]
    return out


def _trapezoidal(vector):
    """A_h: out_0 = in_0/2 ; out_j = (in_j + in_{j-1})/2 ; out_n = in_{n-1}/2."""
    blocksize = vector.shape[1]
    out = np.empty((3, blocksize + 1))
    out[:, 0] = 0.5 * vector[:, 0]
    out[:, -1] = 0.5 * vector[:, -1]
    out[:, 1:-1] = 0.5 * (vector[:, 1:] + vector[:, :-1])
    return out


# the reference's block-structure variants reduce to the plain ones for a rod
# that is not adjacent to ghost elements (SURVEY B-8)
difference_kernel_for_block_structure = lambda v, ghost=None: _difference(v)
trapezoidal_for_block_structure = lambda v, ghost=None: _trapezoidal(v)
quadrature_kernel = _trapezoidal
difference_kernel = _difference
