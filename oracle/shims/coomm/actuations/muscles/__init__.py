"""TEST INFRASTRUCTURE (oracle) — see oracle/shims/coomm/__init__.py."""
from .muscle import MuscleForce, ApplyMuscles, force_length_weight_poly  # noqa: F401
from .longitudinal_muscle import LongitudinalMuscle  # noqa: F401
from .transverse_muscle import TransverseMuscle  # noqa: F401
