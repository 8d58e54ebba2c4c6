"""TEST INFRASTRUCTURE (oracle) — not product code.

Loads the UNMODIFIED reference env classes from `/root/reference/gym_softrobot`
on top of the oracle shims (oracle/shims/elastica, oracle/shims/gymnasium).
Only usable where `/root/reference` exists (the build container); never used by
`-m gpu` tests, smoke() or bench.py (those use committed golden fixtures and the
C oracle, which travel).
"""
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = "/root/reference"


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "gym_softrobot"))


def _stub(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


def install_shims():
    """Put the shims and the reference on sys.path; stub render-only deps."""
    shims = os.path.join(_HERE, "shims")
    for p in (REFERENCE_ROOT, shims):
        if p not in sys.path:
            sys.path.insert(0, p)
    # rendering-only imports of the reference (`utils/render/post_processing.py:1-5`)
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except ImportError:
            mpl = _stub("matplotlib")
            mpl.pyplot = _stub("matplotlib.pyplot")
            mpl.colors = _stub("matplotlib.colors", to_rgb=lambda c: c)
            mpl.animation = _stub("matplotlib.animation")


class _StubFinder:
    """Fabricates importable placeholder modules for third-party packages that cannot be obtained
    offline (COOMM muscle models: SURVEY §2 row 22).  Only import-time names are satisfied; using
    them raises."""

    PREFIXES = ("coomm",)

    def find_spec(self, fullname, path=None, target=None):
        import importlib.machinery
        if fullname.split(".")[0] in self.PREFIXES:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        mod = types.ModuleType(spec.name)
        mod.__path__ = []

        def _getattr(name, _m=spec.name):
            if name.startswith("__"):
                raise AttributeError(name)
            return type(name, (), {"__init__": lambda self, *a, **k: (_ for _ in ()).throw(
                NotImplementedError(f"{_m}.{name} is not available offline"))})

        mod.__getattr__ = _getattr
        return mod

    def exec_module(self, module):
        pass


def load_reference_env(env_id: str, **kwargs):
    """`gym.make(env_id)` against the real reference code + oracle shims."""
    if not reference_available():
        raise RuntimeError("reference tree not present (only exists in the build container)")
    install_shims()
    if not any(isinstance(f, _StubFinder) for f in sys.meta_path):
        sys.meta_path.append(_StubFinder())
    import gymnasium  # the shim
    import gym_softrobot  # noqa: F401  the real reference package (registers ids)

    return gymnasium.make(env_id, **kwargs)
