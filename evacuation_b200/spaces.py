"""Minimal Box / Dict space stand-ins (gymnasium is not a dependency of this package; the
reference uses gymnasium.spaces at env.py:69,86-96, wrappers.py:34-40,62-73)."""
from __future__ import annotations

import numpy as np


class Box:
    def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
        self.dtype = np.dtype(dtype)
        self.shape = tuple(shape) if shape is not None else tuple(np.shape(low))
        self.low = np.full(self.shape, low, dtype=self.dtype)
        self.high = np.full(self.shape, high, dtype=self.dtype)
        self._np_random = None
        if seed is not None:
            self.seed(seed)

    # Like gymnasium's spaces, a Box samples from ITS OWN generator -- never from the global `np.random` stream, which the
    # single-env face (rng="numpy") consumes in the reference's order (pedestrians.py:17-18, area.py:124): an
    # `action_space.sample()` between two steps (RandomAgent, random_agent.py:8-9) must not shift that stream.
    @property
    def np_random(self) -> np.random.Generator:
        if self._np_random is None:
            self.seed()
        return self._np_random

    def seed(self, seed=None):
        self._np_random = np.random.default_rng(seed)
        return [seed]

    def sample(self):
        return self.np_random.uniform(self.low, self.high).astype(self.dtype)

    def contains(self, x) -> bool:
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    def __repr__(self):
        return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"


class Dict(dict):
    def __init__(self, spaces=None, **kwargs):
        super().__init__()
        if spaces is not None:
            self.update(spaces)
        self.update(kwargs)

    def seed(self, seed=None):
        for i, v in enumerate(self.values()):
            v.seed(None if seed is None else seed + i)
        return [seed]

    def sample(self):
        return {k: v.sample() for k, v in self.items()}
