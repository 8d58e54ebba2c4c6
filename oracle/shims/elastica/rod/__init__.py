"""TEST INFRASTRUCTURE (oracle) — see cosserat_rod.py."""
from .cosserat_rod import RodBase, CosseratRod
