"""TEST INFRASTRUCTURE (oracle) — not product code.

Rodrigues exponential and inverse-rotate (log map) of PyElastica
``elastica/_rotations.py`` ([PE-recall]),
including the hard-coded guards 1e-14 / 1e-10 (SURVEY B-3).
"""
import numpy as np
from ._linalg import _batch_matmul


def _get_rotation_matrix(scale, axis_collection):
    v0 = axis_collection[0].copy()
    v1 = axis_collection[1].copy()
    v2 = axis_collection[2].copy()
    theta = np.sqrt(v0 * v0 + v1 * v1 + v2 * v2)
    v0 /= theta + 1e-14
    v1 /= theta + 1e-14
    v2 /= theta + 1e-14
    theta = theta * scale
    u_prefix = np.sin(theta)
    u_sq_prefix = 1.0 - np.cos(theta)
    rot = np.empty((3, 3, axis_collection.shape[1]))
    rot[0, 0] = 1.0 - u_sq_prefix * (v1 * v1 + v2 * v2)
    rot[1, 1] = 1.0 - u_sq_prefix * (v0 * v0 + v2 * v2)
    rot[2, 2] = 1.0 - u_sq_prefix * (v0 * v0 + v1 * v1)
    rot[0, 1] = u_prefix * v2 + u_sq_prefix * v0 * v1
    rot[1, 0] = -u_prefix * v2 + u_sq_prefix * v0 * v1
    rot[0, 2] = -u_prefix * v1 + u_sq_prefix * v0 * v2
    rot[2, 0] = u_prefix * v1 + u_sq_prefix * v0 * v2
    rot[1, 2] = u_prefix * v0 + u_sq_prefix * v1 * v2
    rot[2, 1] = -u_prefix * v0 + u_sq_prefix * v1 * v2
    return rot


def _rotate(director_collection, scale, axis_collection):
    return _batch_matmul(_get_rotation_matrix(scale, axis_collection), director_collection)


def _inv_rotate(director_collection):
    """log(Q_{k+1} Q_k^T) as a vector, per Voronoi point (not yet divided by D-hat)."""
    Q = director_collection
    a = Q[:, :, 1:]   # Q_{k+1}
    b = Q[:, :, :-1]  # Q_k

    def row_dot(i, j):  # (Q_{k+1} Q_k^T)[i, j] = sum_m a[i,m] b[j,m]
        return a[i, 0] * b[j, 0] + a[i, 1] * b[j, 1] + a[i, 2] * b[j, 2]

    vec = np.empty((3, Q.shape[2] - 1))
    vec[0] = row_dot(2, 1) - row_dot(1, 2)
    vec[1] = row_dot(0, 2) - row_dot(2, 0)
    vec[2] = row_dot(1, 0) - row_dot(0, 1)
    trace = row_dot(0, 0) + row_dot(1, 1) + row_dot(2, 2)
    theta = np.arccos(0.5 * trace - 0.5 - 1e-10)
    vec *= -0.5 * theta / np.sin(theta + 1e-14)
    return vec
