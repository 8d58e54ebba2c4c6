"""TEST INFRASTRUCTURE (oracle) — not product code.

NoForces / GravityForces / EndpointForces of PyElastica
(``elastica/external_forces.py``, [PE-recall]); reference call sites
`/root/reference/gym_softrobot/envs/soft_pendulum/build.py:88-105`.
"""
import numpy as np
from ._linalg import _batch_product_i_k_to_ik, _batch_matvec


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


def _bspline(t_coeff, l_centerline=1.0):
    """``elastica.utils._bspline`` ([PE-recall]): clamped cubic B-spline through
    ``len(t_coeff)`` equidistant control points on [0, l_centerline], with one zero coefficient
    prepended and appended (no torque at either rod end).  The knot multiplicity is fixed by
    scipy's rule len(knots) == len(coeffs) + degree + 1."""
    from scipy.interpolate import BSpline

    t_coeff = np.asarray(t_coeff)
    control_pts = l_centerline * np.linspace(0.0, 1.0, t_coeff.shape[0])
    degree = 3
    knots = np.hstack((np.full(degree, control_pts[0]), control_pts, np.full(degree, control_pts[-1])))
    coeffs = np.hstack((0.0, t_coeff, 0.0))
    return BSpline(knots, coeffs, degree, extrapolate=False), control_pts, coeffs


class MuscleTorques(NoForces):
    """Travelling-wave muscle torque of PyElastica (``elastica/external_forces.py``, [PE-recall]):
    A(s,t) = min(1, t/ramp) * beta(s) * sin(2 pi t / T - wave_number * s + phase), applied as an
    equal-and-opposite couple on consecutive elements, iterated tail-to-head.  Reference call
    site `/root/reference/gym_softrobot/envs/snake/continuum_snake.py:186-198,325-337`."""

    def __init__(self, base_length, b_coeff, period, wave_number, phase_shift, direction, rest_lengths,
                 ramp_up_time, with_spline=False):
        super().__init__()
        self.direction = direction
        self.angular_frequency = 2.0 * np.pi / period
        self.wave_number = wave_number
        self.phase_shift = phase_shift
        assert ramp_up_time > 0.0
        self.ramp_up_time = ramp_up_time
        self.s = np.cumsum(rest_lengths)
        self.s /= self.s[-1]
        if with_spline:
            assert b_coeff.size != 0, "Beta spline coefficient array (t_coeff) is empty"
            my_spline, _, _ = _bspline(b_coeff)
            self.my_spline = my_spline(self.s)
        else:
            self.my_spline = np.full_like(self.s, fill_value=1.0)

    def apply_torques(self, system, time=0.0):
        factor = min(1.0, time / self.ramp_up_time)
        torque_mag = factor * self.my_spline * np.sin(
            self.angular_frequency * time - self.wave_number * self.s + self.phase_shift)
        torque = _batch_product_i_k_to_ik(self.direction, torque_mag[::-1])
        Q = system.director_collection
        system.external_torques[..., 1:] += _batch_matvec(Q, torque)[..., 1:]
        system.external_torques[..., :-1] -= _batch_matvec(Q[..., :-1], torque[..., 1:])
