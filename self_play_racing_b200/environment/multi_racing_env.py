"""`MultiRacingEnv`: the reference's N-car env (environment/multi_racing_env.py
:8-268) over the CUDA backend; dict observations/actions keyed "0", "1", ...
Start slots are drawn with the global ``np.random.shuffle`` exactly as the
reference does (multi_racing_env.py:127-133) and injected into the kernel."""
from __future__ import annotations

import numpy as np

from .. import spaces
from .racing_env import _CarView
from .track import Track, resolve_track


class MultiRacingEnv(spaces.Env):
    KIND = 'multi'

    def __init__(self, num_agents=2, num_sensors=11, track_pool=None, track_id=None, track_width=None):
        self.num_agents = num_agents
        self.num_sensors = num_sensors
        self.max_sensor_range = 50.0
        self.control_points, self.track_width, self.track_id = resolve_track(
            None, track_width, track_pool, track_id)
        obs_dim = num_sensors + 4 + (num_agents - 1) * 4
        self.action_space = spaces.Dict({
            f'{i}': spaces.Box(low=np.array([-1.0, 0.0]), high=np.array([1.0, 1.0]), shape=(2,), dtype=np.float32)
            for i in range(num_agents)})
        self.observation_space = spaces.Dict({
            f'{i}': spaces.Box(low=np.float32(-1.0), high=np.float32(1.0), shape=(obs_dim,), dtype=np.float32)
            for i in range(num_agents)})
        self._be = None
        self._track = None
        self.cars = [_CarView(self, i) for i in range(num_agents)]

    def _backend_ready(self):
        if self._be is None:
            from ..backend import RacingBackend
            self._be = RacingBackend(1, kind='multi', num_agents=self.num_agents, num_sensors=self.num_sensors,
                                     autoreset='disabled', query='exact')
            self._be.set_tracks_from_control_points([self.control_points], [self.track_width])
        return self._be

    @property
    def track(self):
        if self._track is None:
            self._track = Track(self._backend_ready().get_track(0))
        return self._track

    @property
    def steps(self):
        return int(self._backend_ready().get_state()['env_i32'][0, 0])

    def _infos(self, be):
        f = be.info_f64[0].cpu().numpy()
        i = be.info_i32[0].cpu().numpy()
        return [{'position': (float(f[a, 0]), float(f[a, 1])), 'speed': float(f[a, 2]),
                 'progress': float(f[a, 3]), 'crashed': bool(i[a, 0]), 'finished': bool(i[a, 1])}
                for a in range(self.num_agents)], i

    def reset(self, seed=None, options=None):
        import torch
        be = self._backend_ready()
        order = list(range(self.num_agents))
        np.random.shuffle(order)  # multi_racing_env.py:127-128, global stream
        slot = np.array([[order.index(i) for i in range(self.num_agents)]], dtype=np.int32)
        obs = be.reset(start_slot=torch.from_numpy(slot).to(be.device))[0].cpu().numpy()
        st = be.get_state()['car_f64'][0]
        observations = {f'{i}': obs[i] for i in range(self.num_agents)}
        infos = {f'{i}': {'position': (float(st[i, 0]), float(st[i, 1])), 'speed': 0.0, 'progress': 0.0,
                          'crashed': False, 'finished': False} for i in range(self.num_agents)}
        return observations, infos

    def step(self, actions):
        import torch
        be = self._backend_ready()
        a = np.stack([np.asarray(actions[f'{i}'], dtype=np.float32) for i in range(self.num_agents)])
        be.actions[0].copy_(torch.from_numpy(a))
        be.step()
        obs = be.obs[0].cpu().numpy()
        rew = be.reward64[0].cpu().numpy()
        terminated = bool(be.terminated[0].item())
        truncated = bool(be.truncated[0].item())
        info_list, ii = self._infos(be)
        observations, rewards, infos = {}, {}, {}
        for i in range(self.num_agents):
            observations[f'{i}'] = obs[i]
            rewards[f'{i}'] = float(rew[i])
            infos[f'{i}'] = info_list[i]
            infos[f'{i}']['reward'] = float(rew[i])
            if terminated or truncated:
                infos[f'{i}']['placement'] = int(ii[i, 2])
        dones = {f'{i}': terminated for i in range(self.num_agents)}
        dones['__all__'] = terminated or truncated
        return observations, rewards, dones, truncated, infos

    def close(self):
        if self._be is not None:
            self._be.close()
            self._be = None
