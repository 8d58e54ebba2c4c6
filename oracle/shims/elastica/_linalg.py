"""TEST INFRASTRUCTURE (oracle) — not product code.

Batched 3-vector / 3x3 helpers with the accumulation order of PyElastica's
``elastica/_linalg.py`` Numba loops ([PE-recall], pyelastica==1.0.0 is not
installable here — "parity unpinned", see oracle/README.md).

Every helper accumulates ``out += a*b`` over the contracted index in increasing
order starting from 0.0, exactly like the reference's triple loops, so the
floating-point result is independent of NumPy's vectorisation.
"""
import numpy as np


def _batch_matvec(matrix_collection, vector_collection):
    # out[i,k] = sum_j M[i,j,k] v[j,k], j ascending, starting from 0.0
    out = np.zeros((3, vector_collection.shape[1]))
    for i in range(3):
        for j in range(3):
            out[i] += matrix_collection[i, j] * vector_collection[j]
    return out


def _batch_matmul(first, second):
    # out[i,m,k] = sum_j A[i,j,k] B[j,m,k]
    out = np.zeros(first.shape)
    for i in range(3):
        for j in range(3):
            for m in range(3):
                out[i, m] += first[i, j] * second[j, m]
    return out


def _batch_cross(a, b):
    out = np.empty(a.shape)
    out[0] = a[1] * b[2] - a[2] * b[1]
    out[1] = a[2] * b[0] - a[0] * b[2]
    out[2] = a[0] * b[1] - a[1] * b[0]
    return out


def _batch_dot(a, b):
    out = np.zeros(a.shape[1])
    for i in range(3):
        out += a[i] * b[i]
    return out


def _batch_norm(v):
    return np.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])


def _batch_product_i_k_to_ik(vector1, vector2):
    return vector1[:, None] * vector2[None, :]
