"""`RacingEnv`: the reference's single-car Gymnasium env (environment/
racing_env.py:8-166) as a thin host object over the CUDA backend.

Constructing one is cheap (it only records its track and sensor arguments), so
`BatchedRacingVecEnv` can take thousands of them from the reference's `env_fn`
factories and fuse them into ONE device-resident batch.  Used on its own
(evaluate.py, utils/metrics.py) it lazily creates a 1-environment backend.
"""
from __future__ import annotations

import numpy as np

from .. import spaces
from .track import Track, resolve_track


class _CarView:
    """env.car / env.cars[i]: x, y, angle, vx, vy, crashed, finished, progress,
    get_corners() -- what utils/visualization.py:108-114,225-244 reads."""
    LENGTH, WIDTH, MAX_SPEED, STEERING_SPEED = 4.0, 2.0, 30.0, 3.0

    def __init__(self, env, idx):
        self._env, self._idx = env, idx

    def _row(self):
        st = self._env._backend_ready().get_state()
        return st['car_f64'][0, self._idx], st['car_i32'][0, self._idx]

    x = property(lambda self: float(self._row()[0][0]))
    y = property(lambda self: float(self._row()[0][1]))
    angle = property(lambda self: float(self._row()[0][2]))
    vx = property(lambda self: float(self._row()[0][3]))
    vy = property(lambda self: float(self._row()[0][4]))
    angular_velocity = 0.0  # never updated by the reference (car.py:21,54)
    crashed = property(lambda self: bool(self._row()[1][2] & 1))
    finished = property(lambda self: bool(self._row()[1][2] & 2))

    @property
    def progress(self):
        return float(self._row()[1][0]) / len(self._env.track.waypoints)

    def get_corners(self):
        f, _ = self._row()
        c, s = np.cos(f[2]), np.sin(f[2])
        local = np.array([[2.0, 1.0], [2.0, -1.0], [-2.0, -1.0], [-2.0, 1.0]])
        return local @ np.array([[c, s], [-s, c]]) + f[:2]


class RacingEnv(spaces.Env):
    """Same constructor and step/reset contract as the reference class."""
    KIND = 'single'

    def __init__(self, num_sensors=7, track_pool=None, track_id=None, track_width=None, speed_weight=8.0):
        self.num_sensors = num_sensors
        self.num_agents = 1
        self.max_sensor_range = 50.0
        self.control_points, self.track_width, self.track_id = resolve_track(
            None, track_width, track_pool, track_id)
        self._speed_weight = float(speed_weight)
        self.action_space = spaces.Box(low=np.array([-1.0, 0.0]), high=np.array([1.0, 1.0]),
                                       shape=(2,), dtype=np.float32)
        self.observation_space = spaces.Box(low=np.float32(-1.0), high=np.float32(1.0),
                                            shape=(num_sensors + 4,), dtype=np.float32)
        self._be = None
        self._track = None
        self.car = _CarView(self, 0)

    # ---- backend plumbing -------------------------------------------------
    def _backend_ready(self):
        if self._be is None:
            from ..backend import RacingBackend
            self._be = RacingBackend(1, kind='single', num_sensors=self.num_sensors, autoreset='disabled',
                                     query='exact', speed_weight=self._speed_weight)
            self._be.set_tracks_from_control_points([self.control_points], [self.track_width])
        return self._be

    @property
    def track(self):
        if self._track is None:
            self._track = Track(self._backend_ready().get_track(0))
        return self._track

    @property
    def speed_weight(self):
        return self._speed_weight

    @speed_weight.setter
    def speed_weight(self, value):
        self._speed_weight = float(value)
        if self._be is not None:
            self._be.set_speed_weight(self._speed_weight)

    @property
    def steps(self):
        return int(self._backend_ready().get_state()['env_i32'][0, 0])

    # ---- gymnasium API ----------------------------------------------------
    def _info(self, be):
        f = be.info_f64[0, 0].cpu().numpy()
        i = be.info_i32[0, 0].cpu().numpy()
        return {'position': (float(f[0]), float(f[1])), 'speed': float(f[2]), 'progress': float(f[3]),
                'crashed': bool(i[0]), 'finished': bool(i[1])}, f

    def reset(self, seed=None, options=None):
        be = self._backend_ready()
        obs = be.reset()[0, 0].cpu().numpy()
        x0, y0, _ = self.track.get_start_pos()
        info = {'position': (float(x0), float(y0)), 'speed': 0.0, 'progress': 0.0, 'crashed': False,
                'finished': False}
        return obs, info

    def step(self, action):
        be = self._backend_ready()
        import torch
        be.actions[0, 0].copy_(torch.as_tensor(np.asarray(action, dtype=np.float32)))
        be.step()
        obs = be.obs[0, 0].cpu().numpy()
        info, f = self._info(be)
        reward = float(be.reward64[0, 0].item())
        info['reward'] = reward
        info['progress_delta'] = float(f[4])
        return obs, reward, bool(be.terminated[0].item()), bool(be.truncated[0].item()), info

    def close(self):
        if self._be is not None:
            self._be.close()
            self._be = None
