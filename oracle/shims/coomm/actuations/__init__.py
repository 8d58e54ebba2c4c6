"""TEST INFRASTRUCTURE (oracle) — see oracle/shims/coomm/__init__.py."""
from .actuation import ContinuousActuation, ApplyActuations  # noqa: F401
