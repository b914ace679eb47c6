"""A minimal stand-in for ``gymnasium`` (absent from this image), installed into
``sys.modules`` ONLY by tools/make_golden.py and the reference-pinning tests so
that the unmodified reference under /root/reference can be imported.

It reproduces the parts of gymnasium 1.2.3 the reference touches:
``Env``/``Wrapper``, ``spaces.Box``/``spaces.Dict``, and -- written from the
documented gymnasium 1.x semantics -- ``vector.SyncVectorEnv``
with NEXT_STEP auto-reset and ``wrappers.RecordEpisodeStatistics``.
"""
import sys
import time
import types

import numpy as np


class Env:
    def reset(self, seed=None, options=None):
        return None

    def close(self):
        pass


class Wrapper(Env):
    def __init__(self, env):
        self.env = env

    def __getattr__(self, name):
        if name.startswith('_'):
            raise AttributeError(name)
        return getattr(self.env, name)

    def reset(self, **kw):
        return self.env.reset(**kw)

    def step(self, action):
        return self.env.step(action)

    def close(self):
        return self.env.close()


class Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        if shape is None:
            shape = np.shape(low)
        self.shape = tuple(shape)
        self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape).copy()
        self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape).copy()
        self._rng = np.random.default_rng()

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)

    def sample(self):
        return self._rng.uniform(self.low, self.high).astype(self.dtype)


class Dict(dict):
    def __init__(self, spaces):
        super().__init__(spaces)
        self.spaces = dict(spaces)

    def seed(self, seed=None):
        for s in self.values():
            s.seed(seed)


class RecordEpisodeStatistics(Wrapper):
    def __init__(self, env):
        super().__init__(env)
        self.episode_returns = 0.0
        self.episode_lengths = 0
        self._t0 = time.perf_counter()

    def reset(self, **kw):
        out = self.env.reset(**kw)
        self.episode_returns = 0.0
        self.episode_lengths = 0
        self._t0 = time.perf_counter()
        return out

    def step(self, action):
        obs, r, term, trunc, info = self.env.step(action)
        self.episode_returns += r
        self.episode_lengths += 1
        if term or trunc:
            info = dict(info)
            info['episode'] = {'r': self.episode_returns, 'l': self.episode_lengths,
                               't': round(time.perf_counter() - self._t0, 6)}
        return obs, r, term, trunc, info


class SyncVectorEnv:
    """Serial vector env, autoreset_mode=NEXT_STEP."""

    def __init__(self, env_fns):
        self.envs = [fn() for fn in env_fns]
        self.num_envs = len(self.envs)
        self.single_observation_space = self.envs[0].observation_space
        self.single_action_space = self.envs[0].action_space
        self._autoreset = np.zeros(self.num_envs, dtype=bool)

    def reset(self, seed=None, options=None):
        obs = []
        for env in self.envs:
            o, _ = env.reset()
            obs.append(o)
        self._autoreset[:] = False
        return np.stack(obs), {}

    def step(self, actions):
        obs, rew, term, trunc = [], [], [], []
        ep_r = np.zeros(self.num_envs)
        ep_l = np.zeros(self.num_envs, dtype=np.int64)
        ep_mask = np.zeros(self.num_envs, dtype=bool)
        for i, env in enumerate(self.envs):
            if self._autoreset[i]:
                o, info = env.reset()
                r, te, tr = 0.0, False, False
            else:
                o, r, te, tr, info = env.step(actions[i])
            if 'episode' in info:
                ep_r[i], ep_l[i], ep_mask[i] = info['episode']['r'], info['episode']['l'], True
            obs.append(o); rew.append(r); term.append(te); trunc.append(tr)
        term = np.array(term, dtype=bool)
        trunc = np.array(trunc, dtype=bool)
        self._autoreset = term | trunc
        infos = {}
        if ep_mask.any():
            infos['episode'] = {'r': ep_r, 'l': ep_l}
            infos['_episode'] = ep_mask
        return np.stack(obs), np.array(rew, dtype=np.float64), term, trunc, infos

    def close(self):
        pass


def install():
    """Register the stub as ``gymnasium`` unless the real package imports."""
    try:
        import gymnasium  # noqa: F401
        return False
    except ImportError:
        pass
    gym = types.ModuleType('gymnasium')
    gym.Env, gym.Wrapper = Env, Wrapper
    spaces = types.ModuleType('gymnasium.spaces')
    spaces.Box, spaces.Dict = Box, Dict
    wrappers = types.ModuleType('gymnasium.wrappers')
    wrappers.RecordEpisodeStatistics = RecordEpisodeStatistics
    vector = types.ModuleType('gymnasium.vector')
    vector.SyncVectorEnv = SyncVectorEnv
    gym.spaces, gym.wrappers, gym.vector = spaces, wrappers, vector
    for name, mod in (('gymnasium', gym), ('gymnasium.spaces', spaces),
                      ('gymnasium.wrappers', wrappers), ('gymnasium.vector', vector)):
        sys.modules[name] = mod
    return True
