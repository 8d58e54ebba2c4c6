"""TEST INFRASTRUCTURE (oracle) — not product code; never imported by gym_softrobot_b200.

A minimal NumPy restatement of the part of the `pyelastica==1.0.0` API that
gym-softrobot's physics step touches, exposed
under the import name ``elastica`` so that the *unmodified* reference env code in
`/root/reference/gym_softrobot` can be executed on top of it to produce golden
vectors (oracle/gen_golden.py).  pyelastica itself is a third-party, un-vendored
dependency (pinned 1.0.0 in `/root/reference/uv.lock:845-857`) that cannot be
installed here: this restatement is recalled from its public source and the
published algorithm (Gazzola et al. 2018) — **parity unpinned**.
"""
from .rod import CosseratRod, RodBase
from .modules import (
    BaseSystemCollection,
    Constraints,
    Connections,
    Forcing,
    Damping,
    Contact,
    CallBacks,
)
from .timestepper import PositionVerlet, extend_stepper_interface, integrate
from .external_forces import NoForces, GravityForces, EndpointForces, MuscleTorques
from .dissipation import DamperBase, AnalyticalLinearDamper, LaplaceDissipationFilter
from .boundary_conditions import ConstraintBase, FreeBC, OneEndFixedBC
from .callback_functions import CallBackBaseClass
from .contact_forces import Plane, RodPlaneContactWithAnisotropicFriction, NoContact, SurfaceBase
from .joint import FreeJoint
from .rigidbody import Cylinder, Sphere, RigidBodyBase


def __getattr__(name):
    """Names other reference envs import at module load but this oracle does not restate
    (Sphere, MuscleTorques, ...): importable placeholders that fail loudly when used."""
    if name.startswith("__"):
        raise AttributeError(name)

    class _NotRestated:
        def __init__(self, *a, **k):
            raise NotImplementedError(f"elastica.{name} is not restated by the oracle shim")

    _NotRestated.__name__ = name
    return _NotRestated
