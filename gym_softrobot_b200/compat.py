"""Gymnasium compatibility: use the real package when importable, otherwise a
minimal stand-in with the same seeding / Box semantics.

gymnasium is not installed in the build image, so the host layer must not
require it; when it is present the envs subclass `gymnasium.Env`, use
`gymnasium.spaces.Box`, and the ids are registered at import time exactly like
`/root/reference/gym_softrobot/__init__.py:6-80`.
"""
import numpy as np

try:  # pragma: no cover - depends on the image
    import gymnasium as _gym
    from gymnasium import spaces as _spaces

    HAVE_GYMNASIUM = True
    Env = _gym.Env
    Box = _spaces.Box
    Dict = _spaces.Dict
    Discrete = _spaces.Discrete
except ImportError:
    HAVE_GYMNASIUM = False

    def np_random(seed=None):
        """gymnasium.utils.seeding.np_random: Generator(PCG64(SeedSequence(seed)))."""
        seed_seq = np.random.SeedSequence(seed)
        return np.random.Generator(np.random.PCG64(seed_seq)), seed_seq.entropy

    class Env:
        metadata = {"render_modes": []}
        render_mode = None
        spec = None
        _np_random = None
        _np_random_seed = None

        def reset(self, *, seed=None, options=None):
            if seed is not None:
                self._np_random, self._np_random_seed = np_random(seed)

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

    class Box:
        """Fully/partially bounded box with gymnasium 1.0 `sample()` semantics (Appendix E, B-12)."""

        def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
            self.dtype = np.dtype(dtype)
            if shape is None:
                shape = np.broadcast(np.asarray(low), np.asarray(high)).shape
            self._shape = tuple(int(s) for s in shape)
            self.low = np.broadcast_to(np.asarray(low, dtype=np.float64), self._shape).astype(self.dtype)
            self.high = np.broadcast_to(np.asarray(high, dtype=np.float64), self._shape).astype(self.dtype)
            self.bounded_below = -np.inf < self.low
            self.bounded_above = np.inf > self.high
            self._np_random = None
            if seed is not None:
                self.seed(seed)

        @property
        def shape(self):
            return self._shape

        @property
        def np_random(self):
            if self._np_random is None:
                self.seed()
            return self._np_random

        def seed(self, seed=None):
            self._np_random, s = np_random(seed)
            return s

        def sample(self):
            high = self.high
            sample = np.empty(self.shape)
            unbounded = ~self.bounded_below & ~self.bounded_above
            upp = ~self.bounded_below & self.bounded_above
            low = self.bounded_below & ~self.bounded_above
            bounded = self.bounded_below & self.bounded_above
            sample[unbounded] = self.np_random.normal(size=unbounded[unbounded].shape)
            sample[low] = self.np_random.exponential(size=low[low].shape) + self.low[low]
            sample[upp] = -self.np_random.exponential(size=upp[upp].shape) + high[upp]
            sample[bounded] = self.np_random.uniform(low=self.low[bounded], high=high[bounded],
                                                     size=bounded[bounded].shape)
            return sample.astype(self.dtype)

        def contains(self, x):
            if not isinstance(x, np.ndarray):
                try:
                    x = np.asarray(x, dtype=self.dtype)
                except (ValueError, TypeError):
                    return False
            return bool(np.can_cast(x.dtype, self.dtype) and x.shape == self.shape
                        and np.all(x >= self.low) and np.all(x <= self.high))

        def __repr__(self):
            return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"

    class Discrete:
        """`gymnasium.spaces.Discrete` stand-in ({start, ..., start + n - 1}; OctoArmPush-v0's action space,
        arm_push_env.py:101)."""

        def __init__(self, n, seed=None, start=0):
            self.n, self.start = int(n), int(start)
            self.dtype, self._shape = np.dtype(np.int64), ()
            self._np_random = None
            if seed is not None:
                self.seed(seed)

        @property
        def shape(self):
            return self._shape

        @property
        def np_random(self):
            if self._np_random is None:
                self.seed()
            return self._np_random

        def seed(self, seed=None):
            self._np_random, s = np_random(seed)
            return s

        def sample(self):
            return np.int64(self.start + self.np_random.integers(self.n))

        def contains(self, x):
            try:
                xi = int(x)
            except (TypeError, ValueError):
                return False
            return xi == x and self.start <= xi < self.start + self.n

        def __repr__(self):
            return f"Discrete({self.n})"

    class Dict:
        """`gymnasium.spaces.Dict` stand-in: an ordered mapping of sub-spaces (FlatEnv's observation space,
        flat_env.py:99-130)."""

        def __init__(self, spaces=None, seed=None, **kwargs):
            self.spaces = dict(spaces or {}, **kwargs)
            if seed is not None:
                self.seed(seed)

        def __getitem__(self, key):
            return self.spaces[key]

        def keys(self):
            return self.spaces.keys()

        def items(self):
            return self.spaces.items()

        def seed(self, seed=None):
            return {k: s.seed(None if seed is None else seed + i) for i, (k, s) in enumerate(self.spaces.items())}

        def sample(self):
            return {k: s.sample() for k, s in self.spaces.items()}

        def contains(self, x):
            return (isinstance(x, dict) and x.keys() == self.spaces.keys()
                    and all(self.spaces[k].contains(v) for k, v in x.items()))

        def __repr__(self):
            return "Dict(" + ", ".join(f"{k!r}: {s!r}" for k, s in self.spaces.items()) + ")"
