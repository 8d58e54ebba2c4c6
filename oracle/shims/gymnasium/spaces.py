"""TEST INFRASTRUCTURE (oracle) — `gymnasium.spaces.Box` stand-in."""
import numpy as np


class Space:
    def __init__(self, shape=None, dtype=None, seed=None):
        self._shape = None if shape is None else tuple(shape)
        self.dtype = None if dtype is None else np.dtype(dtype)
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
        seed_seq = np.random.SeedSequence(seed)
        self._np_random = np.random.Generator(np.random.PCG64(seed_seq))
        return seed_seq.entropy


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
        dtype = np.dtype(dtype)
        if shape is None:
            shape = np.broadcast(np.asarray(low), np.asarray(high)).shape
        shape = tuple(int(s) for s in shape)
        self.low = np.broadcast_to(np.asarray(low, dtype=np.float64), shape).astype(dtype)
        self.high = np.broadcast_to(np.asarray(high, dtype=np.float64), shape).astype(dtype)
        self.bounded_below = -np.inf < self.low
        self.bounded_above = np.inf > self.high
        super().__init__(shape, dtype, seed)

    def sample(self):
        high = self.high if self.dtype.kind == "f" else self.high.astype("int64") + 1
        sample = np.empty(self.shape)
        unbounded = ~self.bounded_below & ~self.bounded_above
        upp_bounded = ~self.bounded_below & self.bounded_above
        low_bounded = self.bounded_below & ~self.bounded_above
        bounded = self.bounded_below & self.bounded_above
        sample[unbounded] = self.np_random.normal(size=unbounded[unbounded].shape)
        sample[low_bounded] = (
            self.np_random.exponential(size=low_bounded[low_bounded].shape) + self.low[low_bounded]
        )
        sample[upp_bounded] = (
            -self.np_random.exponential(size=upp_bounded[upp_bounded].shape) + high[upp_bounded]
        )
        sample[bounded] = self.np_random.uniform(
            low=self.low[bounded], high=high[bounded], size=bounded[bounded].shape
        )
        if self.dtype.kind in ["i", "u", "b"]:
            sample = np.floor(sample)
        return sample.astype(self.dtype)

    def contains(self, x):
        if not isinstance(x, np.ndarray):
            try:
                x = np.asarray(x, dtype=self.dtype)
            except (ValueError, TypeError):
                return False
        return bool(
            np.can_cast(x.dtype, self.dtype)
            and x.shape == self.shape
            and np.all(x >= self.low)
            and np.all(x <= self.high)
        )

    def __repr__(self):
        return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"


class Dict(Space):
    """`gymnasium.spaces.Dict` stand-in: an ordered mapping of sub-spaces."""

    def __init__(self, spaces=None, seed=None, **kw):
        self.spaces = dict(spaces or {}, **kw)
        super().__init__(None, None, seed)

    def __getitem__(self, k):
        return self.spaces[k]

    def sample(self):
        return {k: s.sample() for k, s in self.spaces.items()}

    def contains(self, x):
        return isinstance(x, dict) and all(k in x and s.contains(x[k]) for k, s in self.spaces.items())


class Discrete(Space):
    def __init__(self, n, seed=None, start=0):
        self.n, self.start = int(n), int(start)
        super().__init__((), np.int64, seed)

    def sample(self):
        return np.int64(self.start + self.np_random.integers(self.n))

    def contains(self, x):
        return self.start <= int(x) < self.start + self.n
