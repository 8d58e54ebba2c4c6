"""TEST INFRASTRUCTURE (oracle) — `gymnasium.envs.registration` stand-in."""
import importlib

registry = {}


class EnvSpec:
    def __init__(self, id, entry_point, kwargs=None):
        self.id = id
        self.entry_point = entry_point
        self.kwargs = dict(kwargs or {})
        self.nondeterministic = False

    def make(self, **kwargs):
        mod_name, attr = self.entry_point.split(":")
        cls = getattr(importlib.import_module(mod_name), attr)
        kw = dict(self.kwargs)
        kw.update(kwargs)
        env = cls(**kw)
        env.spec = self
        return env


def register(id, entry_point, kwargs=None, **_ignored):
    registry[id] = EnvSpec(id, entry_point, kwargs)


def spec(id):
    return registry[id]


def make(id, **kwargs):
    return registry[id].make(**kwargs)
