"""TEST INFRASTRUCTURE (oracle) — not product code.

Minimal stand-in for the parts of `gymnasium==1.0.0` the reference envs use
(`Env`, `spaces.Box`, `envs.registration.register`), recalled from its public
source.  gymnasium is not installable here; this exists
only so the unmodified reference env classes can be executed by
oracle/gen_golden.py.
"""
import numpy as np
from . import spaces
from .envs.registration import register, registry, make, spec


class Env:
    metadata = {"render_modes": []}
    render_mode = None
    spec = None
    _np_random = None
    _np_random_seed = None

    def reset(self, *, seed=None, options=None):
        if seed is not None:
            self._np_random, self._np_random_seed = np_random(seed)
        return None

    @property
    def unwrapped(self):
        return self

    @property
    def np_random(self):
        if self._np_random is None:
            self._np_random, self._np_random_seed = np_random()
        return self._np_random

    @np_random.setter
    def np_random(self, value):
        self._np_random = value

    def close(self):
        pass


def np_random(seed=None):
    # gymnasium.utils.seeding.np_random
    seed_seq = np.random.SeedSequence(seed)
    np_seed = seed_seq.entropy
    rng = np.random.Generator(np.random.PCG64(seed_seq))
    return rng, np_seed
