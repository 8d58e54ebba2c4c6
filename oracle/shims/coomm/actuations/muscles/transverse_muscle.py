"""TEST INFRASTRUCTURE (oracle) — not product code.  PARITY UNPINNED (see oracle/shims/coomm/__init__.py).

TransverseMuscle (call site /root/reference/gym_softrobot/envs/octopus/build.py:329-333): radial fibres on the
centre line.  Contracting them thins the (incompressible) arm and LENGTHENS it, so the axial force has the
opposite sign of a longitudinal muscle's (recalled: the class hands -max_muscle_stress to the base class), and the
normalised fibre length is the radius ratio r / r0 = 1 / sqrt(stretch).
"""
import numpy as np

from .muscle import MuscleForce


class TransverseMuscle(MuscleForce):
    def __init__(self, rest_muscle_area, max_muscle_stress, **kwargs):
        n_elements = np.asarray(rest_muscle_area).shape[0]
        super().__init__(np.zeros((3, n_elements)), rest_muscle_area, -max_muscle_stress, type_name="TM", **kwargs)

    def calculate_muscle_length(self):
        stretch = np.sqrt(np.einsum("ik,ik->k", self.muscle_strain, self.muscle_strain))
        self.muscle_length[...] = 1.0 / np.sqrt(stretch)
