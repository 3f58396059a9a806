"""`Box` space: gymnasium's when it is installed, otherwise a minimal equivalent.

The reference declares its spaces with `gymnasium.spaces.Box` (newsvendor.py:82-88,
inventory_management.py:111-128, network_management.py:277-298).  gymnasium is not
part of this image, so the package carries a small stand-in with the same
attributes (`low`, `high`, `shape`, `dtype`, `sample`, `contains`).
"""
import numpy as np

try:  # pragma: no cover - not installed in the build image
    from gymnasium.spaces import Box  # type: ignore
    HAVE_GYMNASIUM = True
except Exception:  # noqa: BLE001
    HAVE_GYMNASIUM = False

    class Box:  # type: ignore
        def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
            self.dtype = np.dtype(dtype)
            if shape is None:
                shape = np.broadcast(np.asarray(low), np.asarray(high)).shape
            self.shape = tuple(int(s) for s in shape)
            self.low = np.broadcast_to(np.asarray(low), self.shape).astype(self.dtype).copy()
            self.high = np.broadcast_to(np.asarray(high), self.shape).astype(self.dtype).copy()
            self._np_random = np.random.default_rng(seed)

        def seed(self, seed=None):
            self._np_random = np.random.default_rng(seed)
            return seed

        def sample(self):
            if np.issubdtype(self.dtype, np.integer):
                return self._np_random.integers(self.low, self.high + 1, size=self.shape).astype(self.dtype)
            return self._np_random.uniform(self.low, self.high, size=self.shape).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

        def __eq__(self, other):
            return (isinstance(other, Box) and self.shape == other.shape and self.dtype == other.dtype
                    and np.array_equal(self.low, other.low) and np.array_equal(self.high, other.high))

        def __repr__(self):
            return f"Box({self.low.min() if self.low.size else 0}, {self.high.max() if self.high.size else 0}, {self.shape}, {self.dtype})"


class BatchedBox:
    """Batched space of a vector env: the single space broadcast along a leading axis of length n.

    `low` / `high` are zero-stride read-only views, so a 16M-instance env does not allocate gigabytes of bounds."""

    def __init__(self, space, n):
        self.single = space
        self.dtype = space.dtype
        self.shape = (int(n),) + tuple(space.shape)
        self.low = np.broadcast_to(space.low, self.shape)
        self.high = np.broadcast_to(space.high, self.shape)

    def sample(self):
        rng = getattr(self.single, "_np_random", None) or np.random.default_rng()
        if np.issubdtype(self.dtype, np.integer):
            return rng.integers(self.low, self.high + 1, size=self.shape).astype(self.dtype)
        return rng.uniform(self.low, self.high, size=self.shape).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    def __repr__(self):
        return f"BatchedBox({self.single!r}, n={self.shape[0]})"


def batch_box(space, n):
    return BatchedBox(space, n)
