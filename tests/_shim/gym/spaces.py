"""TEST-ONLY: old-gym `spaces.Box` (shape + closed-interval `contains`, no dtype check)."""
import numpy as np


class Space:
    def __init__(self, shape=None, dtype=None):
        self.shape = None if shape is None else tuple(shape)
        self.dtype = None if dtype is None else np.dtype(dtype)
        self.np_random = np.random.RandomState()

    def seed(self, seed=None):
        self.np_random.seed(seed)
        return [seed]


class Box(Space):
    def __init__(self, low=None, high=None, shape=None, dtype=np.float32):
        dtype = np.dtype(dtype)
        if shape is None:
            low = np.asarray(low)
            high = np.asarray(high)
            assert low.shape == high.shape
            shape = low.shape
        else:
            shape = tuple(shape)
            low = np.full(shape, low) if np.isscalar(low) else np.asarray(low)
            high = np.full(shape, high) if np.isscalar(high) else np.asarray(high)
        self.low = low.astype(dtype)
        self.high = high.astype(dtype)
        super().__init__(shape, dtype)

    def sample(self):
        return self.np_random.uniform(low=self.low, high=self.high, size=self.shape).astype(self.dtype)

    def contains(self, x):
        if isinstance(x, list):
            x = np.array(x)
        return x.shape == self.shape and bool(np.all(x >= self.low)) and bool(np.all(x <= self.high))

    def __contains__(self, x):
        return self.contains(x)

    def __repr__(self):
        return "Box" + str(self.shape)

    def __eq__(self, other):
        return isinstance(other, Box) and np.allclose(self.low, other.low) and np.allclose(self.high, other.high)
