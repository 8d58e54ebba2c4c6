"""TEST INFRASTRUCTURE (oracle) — not product code.

AnalyticalLinearDamper / LaplaceDissipationFilter of PyElastica
(``elastica/dissipation.py``, [PE-recall]).
Reference call sites: `/root/reference/gym_softrobot/envs/soft_pendulum/build.py:108-113`,
`/root/reference/gym_softrobot/envs/soft_pendulum_3d/build.py:77-85`.
"""
import numpy as np


class DamperBase:
    def __init__(self, *args, **kwargs):
        self._system = kwargs["_system"]

    @property
    def system(self):
        return self._system

    def dampen_rates(self, system, time):
        pass


class AnalyticalLinearDamper(DamperBase):
    def __init__(self, damping_constant=None, time_step=None, **kwargs):
        super().__init__(**kwargs)
        nodal_mass = self._system.mass
        self.translational_damping_coefficient = np.exp(-damping_constant * time_step)
        element_mass = 0.5 * (nodal_mass[1:] + nodal_mass[:-1])
        element_mass[0] += 0.5 * nodal_mass[0]
        element_mass[-1] += 0.5 * nodal_mass[-1]
        self.rotational_damping_coefficient = np.exp(
            -damping_constant
            * time_step
            * element_mass
            * np.diagonal(self._system.inv_mass_second_moment_of_inertia).T
        )

    def dampen_rates(self, rod, time):
        rod.velocity_collection[:] = rod.velocity_collection * self.translational_damping_coefficient
        rod.omega_collection[:] = rod.omega_collection * np.power(
            self.rotational_damping_coefficient, rod.dilatation
        )


class LaplaceDissipationFilter(DamperBase):
    def __init__(self, filter_order, **kwargs):
        super().__init__(**kwargs)
        if not (filter_order > 0 and isinstance(filter_order, int)):
            raise ValueError("filter_order must be a positive integer")
        self.filter_order = filter_order
        self.velocity_filter_term = np.zeros_like(self._system.velocity_collection)
        self.omega_filter_term = np.zeros_like(self._system.omega_collection)

    @staticmethod
    def _filter_rate(rate_collection, filter_term, filter_order):
        filter_term[...] = rate_collection
        for _ in range(filter_order):
            filter_term[..., 1:-1] = (
                -filter_term[..., 2:] - filter_term[..., :-2] + 2.0 * filter_term[..., 1:-1]
            ) / 4.0
            filter_term[..., 0] = 0.0
            filter_term[..., -1] = 0.0
        rate_collection[...] = rate_collection - filter_term

    def dampen_rates(self, system, time):
        self._filter_rate(system.velocity_collection, self.velocity_filter_term, self.filter_order)
        self._filter_rate(system.omega_collection, self.omega_filter_term, self.filter_order)
