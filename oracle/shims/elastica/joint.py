"""TEST INFRASTRUCTURE (oracle) — FreeJoint of PyElastica ([PE-recall]): stores k, nu and
applies a spring-damper between two nodes; gym-softrobot subclasses it
(`/root/reference/gym_softrobot/utils/custom_elastica/joint.py:20`)."""
import numpy as np


class FreeJoint:
    def __init__(self, k, nu):
        self.k = k
        self.nu = nu

    def apply_forces(self, system_one, index_one, system_two, index_two):
        end_distance_vector = (system_two.position_collection[..., index_two]
                               - system_one.position_collection[..., index_one])
        elastic_force = self.k * end_distance_vector
        relative_velocity = (system_two.velocity_collection[..., index_two]
                             - system_one.velocity_collection[..., index_one])
        damping_force = self.nu * relative_velocity
        contact_force = elastic_force + damping_force
        system_one.external_forces[..., index_one] += contact_force
        system_two.external_forces[..., index_two] -= contact_force

    def apply_torques(self, system_one, index_one, system_two, index_two):
        pass
