"""Observation/action space objects.  Uses `gymnasium.spaces` when gymnasium is installed (so `isinstance` checks in
user code keep working) and a minimal compatible stand-in otherwise (this image has no gymnasium)."""
import numpy as np

try:  # pragma: no cover - depends on the environment
    from gymnasium.spaces import Box, Discrete, MultiDiscrete, Tuple  # noqa: F401

    HAVE_GYMNASIUM = True
except ImportError:
    HAVE_GYMNASIUM = False

    class _Space:
        shape = None
        dtype = None

        def __contains__(self, x):
            return self.contains(x)

    class Box(_Space):
        def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
            self.dtype = np.dtype(dtype)
            if shape is None:
                shape = np.broadcast(np.asarray(low), np.asarray(high)).shape
            self.shape = tuple(int(s) for s in shape)
            self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape).copy()
            self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape).copy()
            self._rng = np.random.default_rng(seed)

        def sample(self):
            return self._rng.uniform(self.low, self.high).astype(self.dtype)

        def seed(self, seed=None):
            self._rng = np.random.default_rng(seed)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

        def __eq__(self, other):
            return isinstance(other, Box) and self.shape == other.shape and np.array_equal(self.low, other.low) and np.array_equal(self.high, other.high)

        def __repr__(self):
            return f"Box({self.low}, {self.high}, {self.shape}, {self.dtype})"

    class Discrete(_Space):
        def __init__(self, n, seed=None, start=0):
            self.n, self.start, self.shape, self.dtype = int(n), int(start), (), np.dtype(np.int64)
            self._rng = np.random.default_rng(seed)

        def sample(self):
            return int(self._rng.integers(self.start, self.start + self.n))

        def seed(self, seed=None):
            self._rng = np.random.default_rng(seed)

        def contains(self, x):
            try:
                xi = int(x)
            except (TypeError, ValueError):
                return False
            return xi == x and self.start <= xi < self.start + self.n

        def __eq__(self, other):
            return isinstance(other, Discrete) and self.n == other.n and self.start == other.start

        def __repr__(self):
            return f"Discrete({self.n})"

    class MultiDiscrete(_Space):
        def __init__(self, nvec, dtype=np.int64, seed=None):
            self.nvec = np.asarray(nvec, dtype=dtype)
            self.shape, self.dtype = self.nvec.shape, np.dtype(dtype)
            self._rng = np.random.default_rng(seed)

        def sample(self):
            return (self._rng.random(self.nvec.shape) * self.nvec).astype(self.dtype)

        def seed(self, seed=None):
            self._rng = np.random.default_rng(seed)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= 0) and np.all(x < self.nvec))

        def __eq__(self, other):
            return isinstance(other, MultiDiscrete) and np.array_equal(self.nvec, other.nvec)

        def __repr__(self):
            return f"MultiDiscrete({self.nvec})"

    class Tuple(_Space):
        def __init__(self, spaces, seed=None):
            self.spaces = tuple(spaces)

        def sample(self):
            return tuple(s.sample() for s in self.spaces)

        def contains(self, x):
            return len(x) == len(self.spaces) and all(s.contains(v) for s, v in zip(self.spaces, x))

        def __getitem__(self, i):
            return self.spaces[i]

        def __len__(self):
            return len(self.spaces)

        def __eq__(self, other):
            return isinstance(other, Tuple) and self.spaces == other.spaces
