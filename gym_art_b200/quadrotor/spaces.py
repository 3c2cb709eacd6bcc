"""Box space used by the facades: gym's when a gym is installed, else a minimal stand-in with the attributes the
reference touches (low, high, shape, dtype; quadrotor.py:540,890,935; quadrotor_control.py:15-19,72-84)."""
import numpy as np

try:                                    # pragma: no cover - no gym in the build image
    from gym.spaces import Box          # noqa: F401
except Exception:                       # noqa: BLE001
    class Box(object):
        def __init__(self, low, high, shape=None, dtype=np.float32):
            if shape is not None and np.isscalar(low):
                low, high = np.full(shape, low), np.full(shape, high)
            self.low = np.asarray(low, dtype=dtype)
            self.high = np.asarray(high, dtype=dtype)
            self.shape = self.low.shape
            self.dtype = np.dtype(dtype)
            self._rng = np.random.RandomState()

        def seed(self, seed=None):
            self._rng = np.random.RandomState(seed)
            return [seed]

        def sample(self):
            return self._rng.uniform(self.low, self.high).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low)) and bool(np.all(x <= self.high))

        def __repr__(self):
            return "Box%s" % (self.shape,)
