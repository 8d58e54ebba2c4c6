"""TEST INFRASTRUCTURE (oracle) — not product code.  PARITY UNPINNED (see oracle/shims/coomm/__init__.py).

LongitudinalMuscle as the reference constructs it (/root/reference/gym_softrobot/envs/octopus/build.py:303-328):
`muscle_init_angle` rotates the given `ratio_muscle_position` about the arm axis (recalled: the fork's keyword; with
the reference's (0, -2/3, 0) and +-pi/2 the two muscles sit at -+2/3 r on d1).  In every env this repo builds the
longitudinal activations stay zero, so only their exact-zero contribution is exercised.
"""
import numpy as np

from .muscle import MuscleForce


class LongitudinalMuscle(MuscleForce):
    def __init__(self, ratio_muscle_position, rest_muscle_area, max_muscle_stress, muscle_init_angle=0.0, **kwargs):
        c, s = np.cos(muscle_init_angle), np.sin(muscle_init_angle)
        p = np.array(ratio_muscle_position, dtype=float)
        rotated = np.stack((c * p[0] - s * p[1], s * p[0] + c * p[1], p[2]), axis=0)
        super().__init__(rotated, rest_muscle_area, max_muscle_stress, type_name="LM", **kwargs)
        self.muscle_init_angle = muscle_init_angle
