"""`Box` space with the semantics the reference was written against (old OpenAI gym).

The reference builds its spaces from `gym.spaces.Box` (msj_robot.py:9-16, roboy_env.py:31-36).
gym is an optional dependency here: this class provides the same attributes
(`low`, `high`, `shape`, `dtype`) and the same `contains` rule -- shape equality plus a closed
interval test, no dtype check (the reference's own tests pass float64 zeros into float32
spaces, test_roboy_env.py:64-65,175) -- so host code does not need gym to be installed.
"""
import numpy as np


class Box:
    def __init__(self, low, high, shape=None, dtype="float32"):
        dtype = np.dtype(dtype)
        if shape is None:
            low, high = np.asarray(low), np.asarray(high)
            if low.shape != high.shape:
                raise ValueError("low and high must have the same shape")
            shape = low.shape
        else:
            shape = tuple(shape)
            low = np.full(shape, low) if np.isscalar(low) else np.asarray(low)
            high = np.full(shape, high) if np.isscalar(high) else np.asarray(high)
        self.low = low.astype(dtype)
        self.high = high.astype(dtype)
        self.shape = tuple(shape)
        self.dtype = dtype
        self.np_random = np.random.RandomState()

    def seed(self, seed=None):
        self.np_random.seed(seed)
        return [seed]

    def sample(self):
        """Host-side convenience sampler (uniform, own RNG) -- not used by the device path."""
        return self.np_random.uniform(low=self.low, high=self.high, size=self.shape).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low)) and bool(np.all(x <= self.high))

    __contains__ = contains

    def __repr__(self):
        return "Box{}".format(self.shape)

    def __eq__(self, other):
        return isinstance(other, Box) and np.allclose(self.low, other.low) and np.allclose(self.high, other.high)
