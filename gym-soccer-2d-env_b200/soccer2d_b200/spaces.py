"""`spaces` used by the env classes: gym's if it is installed (the reference pins gym==0.26.2,
requirements.txt:3), gymnasium's as second choice, else a minimal stand-in with the attributes SB3 and the
reference's scripts read (shape, dtype, n, low, high, sample, contains)."""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - depends on the installation
    from gym import Env  # type: ignore
    from gym.spaces import Box, Discrete  # type: ignore
    BACKEND = "gym"
except Exception:  # noqa: BLE001
    try:  # pragma: no cover
        from gymnasium import Env  # type: ignore
        from gymnasium.spaces import Box, Discrete  # type: ignore
        BACKEND = "gymnasium"
    except Exception:  # noqa: BLE001
        BACKEND = "builtin"

        class Env:  # noqa: D101 - same surface as gym.Env as far as soccer_2d_env.py uses it
            metadata: dict = {}
            action_space = None
            observation_space = None

            def reset(self):
                raise NotImplementedError

            def step(self, action):
                raise NotImplementedError

            def render(self, mode="human"):
                return None

            def close(self):
                return None

        class Box:  # noqa: D101
            def __init__(self, low, high, shape=None, dtype=np.float32):
                self.dtype = np.dtype(dtype)
                if shape is None:
                    shape = np.broadcast(np.asarray(low), np.asarray(high)).shape
                self.shape = tuple(shape)
                self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape).copy()
                self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape).copy()
                self._rng = np.random.default_rng()

            def seed(self, seed=None):
                self._rng = np.random.default_rng(seed)
                return [seed]

            def sample(self):
                return self._rng.uniform(self.low, self.high).astype(self.dtype)

            def contains(self, x):
                x = np.asarray(x)
                return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

            def __repr__(self):
                return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"

        class Discrete:  # noqa: D101
            def __init__(self, n):
                self.n = int(n)
                self.shape = ()
                self.dtype = np.dtype(np.int64)
                self._rng = np.random.default_rng()

            def seed(self, seed=None):
                self._rng = np.random.default_rng(seed)
                return [seed]

            def sample(self):
                return int(self._rng.integers(self.n))

            def contains(self, x):
                try:
                    return 0 <= int(x) < self.n
                except (TypeError, ValueError):
                    return False

            def __repr__(self):
                return f"Discrete({self.n})"
