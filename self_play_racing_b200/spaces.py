"""Gymnasium surface used by the reference (`gym.Env`, `gym.Wrapper`,
`gym.spaces.Box/Dict`).  The real gymnasium package is used when importable;
this image does not ship it, so a minimal stand-in with the same attributes
(`shape`, `dtype`, `low`, `high`, `sample`, `seed`) is provided otherwise."""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - depends on the image
    import gymnasium as _gym
    Env, Wrapper, Box, Dict = _gym.Env, _gym.Wrapper, _gym.spaces.Box, _gym.spaces.Dict
    HAVE_GYMNASIUM = True
except ImportError:
    HAVE_GYMNASIUM = False

    class Env:
        observation_space = None
        action_space = None

        def reset(self, seed=None, options=None):
            return None

        def close(self):
            pass

    class Wrapper(Env):
        def __init__(self, env):
            self.env = env

        def __getattr__(self, name):
            if name.startswith('_') or name == 'env':
                raise AttributeError(name)
            return getattr(self.env, name)

        def reset(self, **kwargs):
            return self.env.reset(**kwargs)

        def step(self, action):
            return self.env.step(action)

        def close(self):
            return self.env.close()

    class Box:
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.dtype = np.dtype(dtype)
            self.shape = tuple(np.shape(low)) if shape is None else tuple(shape)
            self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape).copy()
            self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape).copy()
            self._rng = np.random.default_rng()

        def seed(self, seed=None):
            self._rng = np.random.default_rng(seed)
            return seed

        def sample(self):
            return self._rng.uniform(self.low, self.high).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

        def __repr__(self):
            return f'Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})'

    class Dict(dict):
        def __init__(self, spaces):
            super().__init__(spaces)
            self.spaces = dict(spaces)

        def seed(self, seed=None):
            for s in self.values():
                s.seed(seed)
            return seed

        def sample(self):
            return {k: s.sample() for k, s in self.items()}
