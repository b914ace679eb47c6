"""CPU oracle for the batched racing step path -- TEST INFRASTRUCTURE ONLY.

This file is a float64 numpy restatement, vectorised over environments, of the
reference's per-environment simulator.  It exists so that the CUDA path can be
checked against the reference's arithmetic on the GPU box, where
``/root/reference`` is not available.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it; nothing under ``self_play_racing_b200/`` does.

Parity status: **pinned by fixtures generated from the reference itself**.  The
reference ships no tests or golden vectors, so
``tools/make_golden.py`` imports the unmodified reference (with a gymnasium
stub) in the build container, records trajectories, and commits them under
``tests/golden/``; ``tests/test_oracle_golden.py`` replays them through this
oracle.  Third-party arithmetic that the reference calls and that is *called,
not restated,* here: ``scipy.interpolate.CubicSpline(bc_type='periodic')``
(reference ``environment/track.py:106-107``; pin scipy==1.16.3, image has
1.18.1) -- present in the image on both boxes.

Every function cites the reference lines it follows (paths relative to the
reference root).  Array conventions: E environments, A cars per environment,
R rays; all state is float64 ``[E, A]``.
"""
from __future__ import annotations

import numpy as np
from scipy.interpolate import CubicSpline

# environment/car.py:4-11
MAX_SPEED = 30.0
ACCELERATION = 10.0
STEERING_SPEED = 3.0
DRAG = 0.985
LATERAL_FRICTION = 0.85
GRIP = 0.9
CAR_LENGTH = 4.0
CAR_WIDTH = 2.0
DT = 0.05  # environment/car.py:45
MAX_SENSOR_RANGE = 50.0  # racing_env.py:15, multi_racing_env.py:16
MAX_EPISODE_STEPS = 3000  # racing_env.py:162, multi_racing_env.py:250

# environment/track.py:69-74 -- the fixed polygon used when no pool is given
DEFAULT_CONTROL_POINTS = np.array(
    [[0, 0], [50, 0], [70, 20], [60, 40], [70, 50],
     [50, 70], [20, 70], [10, 50], [10, 20], [0, 10]])

# car.py:31-36 corner order: FL(+,+) FR(+,-) RR(-,-) RL(-,+)
_CORNER_LX = np.array([CAR_LENGTH / 2, CAR_LENGTH / 2, -CAR_LENGTH / 2, -CAR_LENGTH / 2])
_CORNER_LY = np.array([CAR_WIDTH / 2, -CAR_WIDTH / 2, -CAR_WIDTH / 2, CAR_WIDTH / 2])


# --------------------------------------------------------------------------
# T0: procedural control points
# --------------------------------------------------------------------------
def gen_random_track(num_points=15, base_radius=50, radius_variation=15,
                     angle_jitter=0.2, smoothness=0.5, seed=None, rng=None):
    """environment/track.py:4-45.  ``rng`` defaults to the *global* legacy
    ``np.random`` module, which is what the reference draws from."""
    rng = np.random if rng is None else rng
    if seed is not None:
        rng.seed(seed)  # track.py:5-6 (re-seeds the global stream: SURVEY quirk 8)
    ang = np.linspace(0, 2 * np.pi, num_points, endpoint=False)  # :9
    if angle_jitter > 0:  # :10-18
        spacing = 2 * np.pi / num_points
        ang = ang + rng.uniform(-angle_jitter * spacing / 2, angle_jitter * spacing / 2, num_points)
        ang = np.sort(ang % (2 * np.pi))
    # :21-33 -- one uniform draw per point; the vector draw consumes the same stream
    var = rng.uniform(-radius_variation, radius_variation, num_points)
    radii = np.zeros(num_points)
    for i in range(num_points):
        raw = base_radius + var[i]
        if smoothness > 0 and i > 0:
            radii[i] = (1 - smoothness) * raw + (smoothness * radii[i - 1])
        else:
            radii[i] = raw
    if smoothness > 0:  # :36-37
        radii[0] = (radii[0] + radii[-1]) / 2
    return np.column_stack([radii * np.cos(ang), radii * np.sin(ang)])  # :40-43


def gen_tracks(num_tracks=10, seed=None, rng=None):
    """environment/track.py:47-56 (same draw order: n, base, variation, jitter, smooth)."""
    rng = np.random if rng is None else rng
    pool = []
    for _ in range(num_tracks):
        n = rng.randint(10, 15)
        base = rng.randint(50, 80)
        variation = rng.randint(10, base // 2 - 10)
        jitter = rng.uniform(0.2, 0.7)
        smooth = rng.uniform(0.2, 0.7)
        pool.append(gen_random_track(n, base, variation, jitter, smooth, seed, rng=rng))
    return pool


# --------------------------------------------------------------------------
# T1: spline -> waypoints -> normals -> boundaries -> segment table
# --------------------------------------------------------------------------
class TrackTables:
    """What ``Track.__init__`` leaves behind (environment/track.py:61-148)."""

    def __init__(self, control_points=None, track_width=None, factor=30):
        cp = DEFAULT_CONTROL_POINTS if control_points is None else np.asarray(control_points)
        self.control_points = cp
        self.track_width = 6.0 if track_width is None else track_width  # :77-80
        self.waypoints = self._waypoints(cp, factor)
        w = self.waypoints
        # :82-91 bounding-box diagonal
        self.max_track_distance = np.sqrt((w[:, 0].max() - w[:, 0].min()) ** 2 +
                                          (w[:, 1].max() - w[:, 1].min()) ** 2)
        # :117-124 forward-difference tangents (wrapping), unit length, rotated +90 deg
        tang = np.diff(w, axis=0, append=[w[0]])
        ln = np.linalg.norm(tang, axis=1, keepdims=True)
        ln = np.where(ln == 0, 1, ln)
        tang = tang / ln
        self.normals = np.column_stack((-tang[:, 1], tang[:, 0]))
        # :93-94
        self.left_boundary = w + self.normals * self.track_width
        self.right_boundary = w - self.normals * self.track_width
        # :126-148 closed polylines, left then right
        self.starts = np.vstack([self.left_boundary, self.right_boundary])
        self.ends = np.vstack([np.roll(self.left_boundary, -1, axis=0),
                               np.roll(self.right_boundary, -1, axis=0)])
        self.v2 = self.ends - self.starts

    @staticmethod
    def _waypoints(cp, factor):
        """track.py:100-115: chord-length knots, two periodic cubic splines."""
        closed = np.vstack((cp, cp[0]))
        t = np.concatenate(([0], np.cumsum(np.sqrt(np.sum(np.diff(closed, axis=0) ** 2, axis=1)))))
        sx = CubicSpline(t, closed[:, 0], bc_type='periodic')
        sy = CubicSpline(t, closed[:, 1], bc_type='periodic')
        tw = np.linspace(0, t[-1], len(cp) * factor, endpoint=False)
        return np.column_stack((sx(tw), sy(tw)))

    @property
    def num_waypoints(self):
        return len(self.waypoints)

    def start_pose(self):
        """track.py:154-157."""
        w = self.waypoints
        return w[0, 0], w[0, 1], np.arctan2(w[1, 1] - w[0, 1], w[1, 0] - w[0, 0])


def make_pool(control_point_pool, widths):
    """Build one TrackTables per pool entry.  ``widths`` may be a scalar or a
    per-track sequence (track.py:62-67)."""
    out = []
    for i, cp in enumerate(control_point_pool):
        w = widths[i] if isinstance(widths, (list, tuple, np.ndarray)) else widths
        out.append(TrackTables(cp, w))
    return out


# --------------------------------------------------------------------------
# W / C / R / RC / K / X : geometry queries, vectorised over a group of points
# --------------------------------------------------------------------------
def closest_waypoint_idx(trk, px, py):
    """track.py:150-152 for flat arrays of query points on ONE track.
    Returns (idx, best, second) where best/second are the smallest and second
    smallest squared distances (used by tests to flag epsilon-ties)."""
    w = trk.waypoints
    d = (w[None, :, 0] - px[:, None]) ** 2 + (w[None, :, 1] - py[:, None]) ** 2
    idx = d.argmin(axis=1)
    if d.shape[1] > 1:
        part = np.partition(d, 1, axis=1)
        return idx, part[:, 0], part[:, 1]
    return idx, d[:, 0], d[:, 0]


def corners_of(x, y, ang):
    """car.py:26-43 -> arrays [..., 4] of corner x and y."""
    c = np.cos(ang)[..., None]
    s = np.sin(ang)[..., None]
    cx = c * _CORNER_LX + (-s) * _CORNER_LY + x[..., None]
    cy = s * _CORNER_LX + c * _CORNER_LY + y[..., None]
    return cx, cy


def wall_test(trk, cx, cy):
    """track.py:163-171 for [G, 4] corners on ONE track.  Returns (crashed[G],
    margin[G]) where margin = min over corners of | |dist| - width |."""
    g = cx.shape[0]
    idx, _, _ = closest_waypoint_idx(trk, cx.reshape(-1), cy.reshape(-1))
    w = trk.waypoints[idx]
    n = trk.normals[idx]
    dist = np.abs((cx.reshape(-1) - w[:, 0]) * n[:, 0] + (cy.reshape(-1) - w[:, 1]) * n[:, 1])
    crashed = (dist > trk.track_width).reshape(g, 4).any(axis=1)
    margin = np.abs(dist - trk.track_width).reshape(g, 4).min(axis=1)
    return crashed, margin


def raycast_walls(trk, ox, oy, direction, max_dist=MAX_SENSOR_RANGE):
    """track.py:173-198.  ox, oy: [G]; direction: [G, R] world angles.
    Returns [G, R] float64 -- NOT clamped to max_dist."""
    dx = np.cos(direction)[..., None]  # [G,R,1]
    dy = np.sin(direction)[..., None]
    v3x, v3y = -dy, dx
    v1x = (ox[:, None] - trk.starts[None, :, 0])[:, None, :]  # [G,1,S]
    v1y = (oy[:, None] - trk.starts[None, :, 1])[:, None, :]
    v2x = trk.v2[None, None, :, 0]
    v2y = trk.v2[None, None, :, 1]
    dotp = v2x * v3x + v2y * v3y
    valid = np.abs(dotp) > 1e-10
    with np.errstate(divide='ignore', invalid='ignore'):
        t = np.where(valid, (v2x * v1y - v2y * v1x) / dotp, max_dist)
        s = np.where(valid, (v1x * v3x + v1y * v3y) / dotp, -1.0)
    hit = valid & (t >= 0) & (s >= 0) & (s <= 1)
    return _min_or_default(t, hit, max_dist)


def _min_or_default(t, hit, default):
    m = np.where(hit, t, np.inf).min(axis=-1)
    return np.where(np.isinf(m), default, m)


def raycast_car_edges(ox, oy, direction, ecx, ecy, skip, max_dist=MAX_SENSOR_RANGE):
    """multi_track.py:8-24,28-44: min over the 4 edges of each not-skipped car.
    ox, oy: [G]; direction [G,R]; ecx, ecy: [G, K, 4] corners of the K cars
    (including self); skip: [G, K] bool.  Returns [G, R] (max_dist if no hit)."""
    dx = np.cos(direction)[:, :, None, None]  # [G,R,1,1]
    dy = np.sin(direction)[:, :, None, None]
    v3x, v3y = -dy, dx
    sx = ecx[:, None, :, :]
    sy = ecy[:, None, :, :]
    ex = np.roll(ecx, -1, axis=2)[:, None, :, :]
    ey = np.roll(ecy, -1, axis=2)[:, None, :, :]
    v1x = ox[:, None, None, None] - sx
    v1y = oy[:, None, None, None] - sy
    v2x = ex - sx
    v2y = ey - sy
    dotp = v2x * v3x + v2y * v3y
    ok = ~(np.abs(dotp) < 1e-10)  # multi_track.py:35
    with np.errstate(divide='ignore', invalid='ignore'):
        t = (v2x * v1y - v2y * v1x) / dotp
        s = (v1x * v3x + v1y * v3y) / dotp
    hit = ok & (t >= 0) & (s >= 0) & (s <= 1) & (~skip[:, None, :, None])
    g, r = direction.shape
    # multi_track.py:8,24: the running minimum starts at max_dist, so car hits
    # beyond the sensor range never lower the reading
    return np.minimum(max_dist, _min_or_default(np.broadcast_to(t, hit.shape).reshape(g, r, -1),
                                                hit.reshape(g, r, -1), max_dist))


def rectangles_intersect(ax, ay, bx, by):
    """multi_car.py:16-43 separating-axis test on [G,4] corner arrays.
    Returns (hit[G], gap[G]) with gap = smallest |projection gap| seen, for
    epsilon-tie flagging."""
    hit = np.ones(ax.shape[0], dtype=bool)
    gap = np.full(ax.shape[0], np.inf)
    for (px, py) in ((ax, ay), (bx, by)):
        for i in range(2):  # multi_car.py:19-22: edges 0->1 and 1->2
            ex = px[:, (i + 1) % 4] - px[:, i]
            ey = py[:, (i + 1) % 4] - py[:, i]
            nx, ny = -ey, ex
            pa = ax * nx[:, None] + ay * ny[:, None]
            pb = bx * nx[:, None] + by * ny[:, None]
            g1 = pb.min(axis=1) - pa.max(axis=1)
            g2 = pa.min(axis=1) - pb.max(axis=1)
            sep = (g1 > 0) | (g2 > 0)  # :40 strict: touching counts as colliding
            hit &= ~sep
            gap = np.minimum(gap, np.minimum(np.abs(g1), np.abs(g2)))
    return hit, gap


# --------------------------------------------------------------------------
# E1 / E2 / V : the batched environment
# --------------------------------------------------------------------------
class OracleVecEnv:
    """E independent copies of ``RacingEnv`` (kind='single', A=1) or
    ``MultiRacingEnv`` (kind='multi'), stepped in lock-step, with gymnasium's
    NEXT_STEP auto-reset and RecordEpisodeStatistics counters (SURVEY 8b/8c;
    call sites agent/ppo.py:70,88,114-130).

    ``step`` returns the raw per-car outputs; ``selfplay_view`` reduces them to
    what ``SelfPlayWrapper.step`` returns (environment/wrappers.py:29-55).
    """

    def __init__(self, tracks, env_to_track, kind='single', num_agents=1, num_sensors=11,
                 speed_weight=8.0, autoreset='next_step', seed=0):
        assert kind in ('single', 'multi')
        self.kind = kind
        self.tracks = list(tracks)
        self.env_to_track = np.asarray(env_to_track, dtype=np.int64)
        self.E = len(self.env_to_track)
        self.A = 1 if kind == 'single' else int(num_agents)
        self.R = int(num_sensors)
        self.speed_weight = speed_weight
        self.autoreset = autoreset
        self.rng = np.random.RandomState(seed)
        # racing_env.py:45 (120 deg cone) / multi_racing_env.py:50 (180 deg cone)
        half = np.pi / 3 if kind == 'single' else np.pi / 2
        self.sensor_angles = np.linspace(-half, half, self.R)
        self.obs_dim = self.R + 4 + (self.A - 1) * 4 if kind == 'multi' else self.R + 4
        E, A = self.E, self.A
        z = lambda dt=np.float64: np.zeros((E, A), dtype=dt)
        self.x, self.y, self.angle, self.vx, self.vy = z(), z(), z(), z(), z()
        self.progress, self.last_progress = z(), z()
        self.progress_idx = z(np.int64)
        self.last_steering = z()
        self.crashed, self.finished, self.has_crashed = z(bool), z(bool), z(bool)
        self.checkpoints = np.zeros((E, A, 3), dtype=bool)
        self.finished_step = z(np.int64)  # 0 stands for None (multi_racing_env.py:24)
        self.steps = np.zeros(E, dtype=np.int64)
        self.needs_reset = np.zeros(E, dtype=bool)
        self.ep_return = np.zeros(E)
        self.ep_length = np.zeros(E, dtype=np.int64)
        self._groups = [np.nonzero(self.env_to_track == t)[0] for t in range(len(self.tracks))]
        self.tie = {}  # epsilon-tie diagnostics of the last step

    # ---- reset ----------------------------------------------------------
    def _draw_start_order(self, n):
        """multi_racing_env.py:127-133: shuffle agent ids, car i gets slot
        ``agent_order.index(i)``.  Own RandomState (the reference uses the
        global stream; parity runs inject the permutation instead)."""
        out = np.zeros((n, self.A), dtype=np.int64)
        for k in range(n):
            order = list(range(self.A))
            self.rng.shuffle(order)
            for i in range(self.A):
                out[k, i] = order.index(i)
        return out

    def _reset_envs(self, envs, start_order):
        """racing_env.py:86-102 / multi_racing_env.py:118-153 + car.py:17-24."""
        if len(envs) == 0:
            return
        for e in envs:
            trk = self.tracks[self.env_to_track[e]]
            x0, y0, a0 = trk.start_pose()
            self.x[e, :], self.y[e, :], self.angle[e, :] = x0, y0, a0
            if self.kind == 'multi':
                spacing = CAR_WIDTH + 1.5
                center = (self.A - 1) / 2.0
                nrm = trk.normals[0]
                for i in range(self.A):
                    off = (int(start_order[e, i]) - center) * spacing
                    self.x[e, i] = trk.waypoints[0][0] + nrm[0] * off
                    self.y[e, i] = trk.waypoints[0][1] + nrm[1] * off
        for arr in (self.vx, self.vy, self.progress, self.last_progress, self.last_steering):
            arr[envs] = 0.0
        self.progress_idx[envs] = 0
        for arr in (self.crashed, self.finished, self.has_crashed):
            arr[envs] = False
        self.checkpoints[envs] = False
        self.finished_step[envs] = 0
        self.steps[envs] = 0
        self.ep_return[envs] = 0.0
        self.ep_length[envs] = 0
        self.needs_reset[envs] = False

    def reset(self, start_order=None):
        if start_order is None:
            start_order = self._draw_start_order(self.E)
        self._reset_envs(np.arange(self.E), np.asarray(start_order))
        return self._observe(np.arange(self.E)), self._infos()

    # ---- D: vehicle dynamics -------------------------------------------
    def _car_update(self, env_mask, steering, throttle):
        """car.py:45-80 for every car of the environments in env_mask."""
        act = env_mask[:, None] & ~self.crashed  # :51-52 crashed cars stay frozen
        ang = (self.angle + ((steering * STEERING_SPEED) * DT)) % (2 * np.pi)  # :54-56
        c, s = np.cos(ang), np.sin(ang)
        vf = self.vx * c + self.vy * s  # :59
        vl = self.vx * (-s) + self.vy * c  # :60
        vf = (vf + ((throttle * ACCELERATION) * DT)) * DRAG  # :61-62
        vl = vl * LATERAL_FRICTION * GRIP  # :63
        nvx = vf * c - vl * s  # :66-67
        nvy = vf * s + vl * c
        speed = np.sqrt((nvx ** 2) + (nvy ** 2))  # :70
        over = speed > MAX_SPEED
        with np.errstate(divide='ignore', invalid='ignore'):
            scale = MAX_SPEED / speed
        nvx = np.where(over, nvx * scale, nvx)  # :71-74
        nvy = np.where(over, nvy * scale, nvy)
        self.angle = np.where(act, ang, self.angle)
        self.vx = np.where(act, nvx, self.vx)
        self.vy = np.where(act, nvy, self.vy)
        self.x = np.where(act, self.x + (self.vx * DT), self.x)  # :77-78
        self.y = np.where(act, self.y + (self.vy * DT), self.y)
        # :79-80 progress and wall test, grouped by track
        argmin_gap = np.full((self.E, self.A), np.inf)
        wall_margin = np.full((self.E, self.A), np.inf)
        for t, envs in enumerate(self._groups):
            if len(envs) == 0:
                continue
            trk = self.tracks[t]
            sel = act[envs]  # [g, A]
            if not sel.any():
                continue
            ge, ga = np.nonzero(sel)
            ee = envs[ge]
            idx, best, second = closest_waypoint_idx(trk, self.x[ee, ga], self.y[ee, ga])
            self.progress_idx[ee, ga] = idx
            self.progress[ee, ga] = idx / trk.num_waypoints  # track.py:159-161
            argmin_gap[ee, ga] = second - best
            cx, cy = corners_of(self.x[ee, ga], self.y[ee, ga], self.angle[ee, ga])
            crashed, margin = wall_test(trk, cx, cy)
            self.crashed[ee, ga] = crashed
            wall_margin[ee, ga] = margin
        self.tie['argmin_gap'] = argmin_gap
        self.tie['wall_margin'] = wall_margin

    # ---- progress delta shared by both reward functions ------------------
    def _progress_delta(self):
        """racing_env.py:112-116 / multi_racing_env.py:159-163."""
        p, lp = self.progress, self.last_progress
        d = p - lp
        fwd = (lp > 0.9) & (p < 0.1)
        bwd = ~fwd & (lp < 0.1) & (p > 0.9)
        d = np.where(fwd, (1.0 - lp) + p, d)
        return np.where(bwd, -((1.0 - p) + lp), d)

    def _checkpoint_logic(self, live, bonus):
        """racing_env.py:123-135 / multi_racing_env.py:175-183; returns reward add."""
        p = self.progress
        cp = self.checkpoints
        add = np.zeros_like(p)
        h0 = live & ~cp[..., 0] & (0.25 <= p) & (p < 0.35)
        cp[..., 0] |= h0
        add += np.where(h0, bonus, 0)
        h1 = live & cp[..., 0] & ~cp[..., 1] & (0.50 <= p) & (p < 0.60)
        cp[..., 1] |= h1
        add += np.where(h1, bonus, 0)
        h2 = live & cp[..., 1] & ~cp[..., 2] & (0.75 <= p) & (p < 0.85)
        cp[..., 2] |= h2
        add += np.where(h2, bonus, 0)
        return add

    # ---- step -----------------------------------------------------------
    def step(self, actions, start_order=None):
        """actions: float32 [E, A, 2].  Returns obs float32 [E,A,D], reward
        float64 [E,A], terminated [E], truncated [E], infos."""
        actions = np.asarray(actions, dtype=np.float32).reshape(self.E, self.A, 2)
        resetting = self.needs_reset.copy() if self.autoreset == 'next_step' else np.zeros(self.E, bool)
        live_env = ~resetting
        self.tie = {}
        # action decoding: racing_env.py:106-107 / multi_racing_env.py:216-217
        steering = np.clip(actions[..., 0], -1.0, 1.0).astype(np.float64)
        if self.kind == 'single':
            throttle = np.clip(actions[..., 1], 0.0, 1.0).astype(np.float64)
        else:  # evaluated in float32 (NumPy-2 weak promotion)
            throttle = np.clip((actions[..., 1] + np.float32(1.0)) / np.float32(2.0), 0.0, 1.0).astype(np.float64)
        self.last_steering = np.where(live_env[:, None], steering, self.last_steering)
        self._car_update(live_env, steering, throttle)

        reward = np.zeros((self.E, self.A))
        live = np.broadcast_to(live_env[:, None], (self.E, self.A))
        if self.kind == 'multi':
            touching = self._car_collisions(live_env)  # multi_racing_env.py:222-231
        self.steps = np.where(live_env, self.steps + 1, self.steps)
        delta = self._progress_delta()
        speed = np.sqrt(self.vx ** 2 + self.vy ** 2)
        speed_ratio = np.clip(speed / MAX_SPEED, 0.0, 1.0)
        steps_f = self.steps[:, None].astype(np.float64)
        if self.kind == 'single':
            reward = delta * 200  # racing_env.py:121
            reward = reward + self._checkpoint_logic(live, 20)
            reward = reward + np.where(~self.crashed & (delta > 0), speed_ratio * self.speed_weight, 0.0)  # :137-140
            reward = reward - np.where(self.crashed, 60, 0)  # :142-143
            fin = live & self.checkpoints.all(axis=2) & (self.last_progress > 0.9) & \
                (self.progress < 0.1) & (delta > 0)  # :145-146
            self.finished |= fin
            reward = reward + np.where(fin, 100, 0)
            reward = reward + np.where(fin, np.maximum(0, 200 - (steps_f / 10)), 0)  # :149-150
        else:
            reward = 0.0 + delta * 200  # multi_racing_env.py:165-167
            reward = reward + np.where(~self.crashed & (delta > 0), speed_ratio * 18, 0.0)  # :169-172
            reward = reward + self._checkpoint_logic(live, 25)
            fin = live & self.checkpoints.all(axis=2) & (self.last_progress > 0.9) & \
                (self.progress < 0.1) & (delta > 0)  # :185-186
            self.finished |= fin
            self.finished_step = np.where(fin, self.steps[:, None], self.finished_step)
            reward = reward + np.where(fin, 100 + np.maximum(0, 300 - (steps_f / 15)), 0)  # :189-190
            first_crash = live & self.crashed & ~self.has_crashed  # :192-194
            reward = reward - np.where(first_crash, 160, 0)
            self.has_crashed |= first_crash
            reward = reward + touching  # :240
        reward = np.where(live, reward, 0.0)

        if self.kind == 'single':
            terminated = (self.crashed | self.finished)[:, 0]  # racing_env.py:161
        else:
            terminated = self.finished.any(axis=1) | self.crashed.all(axis=1)  # multi:247-249
        truncated = self.steps >= MAX_EPISODE_STEPS
        terminated = terminated & live_env
        truncated = truncated & live_env
        placement = np.zeros((self.E, self.A), dtype=np.int64)
        if self.kind == 'multi':
            ended = terminated | truncated
            placement = self._place()  # multi_racing_env.py:198-211
            placement = np.where(ended[:, None], placement, 0)
            reward = reward + np.where(placement == 1, 250, 0)  # :256-257

        # RecordEpisodeStatistics: accumulate agent 0's reward, emit on episode end
        ended = terminated | truncated
        self.ep_return = np.where(live_env, self.ep_return + reward[:, 0], self.ep_return)
        self.ep_length = np.where(live_env, self.ep_length + 1, self.ep_length)
        episode_r = np.where(ended, self.ep_return, 0.0)
        episode_l = np.where(ended, self.ep_length, 0)
        self.last_progress = np.where(live, self.progress, self.last_progress)  # racing_env.py:165

        # auto-reset: envs flagged by the previous step (gymnasium NEXT_STEP), or
        # the envs that just ended (SAME_STEP: the terminal observation is dropped)
        final_infos = self._infos() if self.autoreset == 'same_step' else None
        if self.autoreset == 'same_step':
            resetting = ended
        if resetting.any():
            envs = np.nonzero(resetting)[0]
            so = np.zeros((self.E, self.A), dtype=np.int64)
            if self.kind == 'multi':
                if start_order is None:
                    so[envs] = self._draw_start_order(len(envs))
                else:
                    so = np.asarray(start_order)
            self._reset_envs(envs, so)
        if self.autoreset == 'next_step':
            self.needs_reset = ended

        obs = self._observe(np.arange(self.E))
        infos = final_infos if final_infos is not None else self._infos()
        infos['reward'] = reward.copy()
        infos['progress_delta'] = np.where(live, delta, 0.0)
        infos['placement'] = placement
        infos['_episode'] = ended.copy()
        infos['episode_r'] = episode_r
        infos['episode_l'] = episode_l
        return obs, reward, terminated, truncated, infos

    def _car_collisions(self, live_env):
        """multi_racing_env.py:222-231 -- every pair i<j, no crashed/finished filter."""
        touching = np.zeros((self.E, self.A))
        cx, cy = corners_of(self.x, self.y, self.angle)  # [E,A,4]
        gaps = np.full(self.E, np.inf)
        for i in range(self.A):
            for j in range(i + 1, self.A):
                hit, gap = rectangles_intersect(cx[:, i], cy[:, i], cx[:, j], cy[:, j])
                hit &= live_env
                gaps = np.minimum(gaps, gap)
                for k in (i, j):
                    self.vx[:, k] = np.where(hit, self.vx[:, k] * 0.92, self.vx[:, k])
                    self.vy[:, k] = np.where(hit, self.vy[:, k] * 0.92, self.vy[:, k])
                    touching[:, k] += np.where(hit, -5.0, 0.0)
        self.tie['sat_gap'] = gaps
        return touching

    def _place(self):
        """multi_racing_env.py:198-211: descending (score, idx) => exact ties
        go to the higher car index."""
        fs = np.where(self.finished_step == 0, 10000, self.finished_step)
        score = (self.finished * 10000 + self.progress * 100 + (~self.crashed) * 10 + (1.0 / fs))
        idx = np.arange(self.A)[None, :]
        better = (score[:, None, :] > score[:, :, None]) | \
                 ((score[:, None, :] == score[:, :, None]) & (idx[:, None, :] > idx[:, :, None]))
        return 1 + better.sum(axis=2)  # [E, A]: number of cars ranked ahead + 1

    # ---- observations ---------------------------------------------------
    def _observe(self, envs):
        """racing_env.py:44-75 / multi_racing_env.py:48-105."""
        E, A, R = self.E, self.A, self.R
        rays = np.zeros((E, A, R), dtype=np.float32)
        cx, cy = corners_of(self.x, self.y, self.angle)
        for t, ge in enumerate(self._groups):
            if len(ge) == 0:
                continue
            trk = self.tracks[t]
            for a in range(A):
                ox, oy = self.x[ge, a], self.y[ge, a]
                direction = self.angle[ge, a][:, None] + self.sensor_angles[None, :]
                d = raycast_walls(trk, ox, oy, direction)
                if self.kind == 'multi':
                    # multi_track.py:13: a car is skipped iff its centre is within 0.5 of the origin
                    dist = np.sqrt((self.x[ge] - ox[:, None]) ** 2 + (self.y[ge] - oy[:, None]) ** 2)
                    car_d = raycast_car_edges(ox, oy, direction, cx[ge], cy[ge], dist < 0.5)
                    d = np.minimum(d, car_d)  # multi_track.py:26
                rays[ge, a, :] = d.astype(np.float32)
        rays = rays / np.float32(MAX_SENSOR_RANGE)  # racing_env.py:53 (float32 divide)
        c, s = np.cos(self.angle), np.sin(self.angle)
        vf = np.clip((self.vx * c + self.vy * s) / MAX_SPEED, -1.0, 1.0)
        vl = np.clip((-self.vx * s + self.vy * c) / MAX_SPEED, -1.0, 1.0)
        parts = [rays.astype(np.float64), vf[..., None], vl[..., None],
                 np.zeros((E, A, 1)),  # angular_velocity is always 0
                 self.last_steering[..., None]]
        if self.kind == 'multi' and A > 1:
            mtd = np.array([self.tracks[t].max_track_distance for t in self.env_to_track])[:, None]
            opp = np.zeros((E, A, 4 * (A - 1)))
            for a in range(A):
                feats = []
                for o in range(A):
                    if o == a:
                        continue
                    rx, ry = self.x[:, o] - self.x[:, a], self.y[:, o] - self.y[:, a]
                    lx = np.clip((rx * c[:, a] + ry * s[:, a]) / mtd[:, 0], -1.0, 1.0)
                    ly = np.clip((-rx * s[:, a] + ry * c[:, a]) / mtd[:, 0], -1.0, 1.0)
                    rvx, rvy = self.vx[:, o] - self.vx[:, a], self.vy[:, o] - self.vy[:, a]
                    lvx = np.clip((rvx * c[:, a] + rvy * s[:, a]) / MAX_SPEED, -1.0, 1.0)
                    lvy = np.clip((-rvx * s[:, a] + rvy * c[:, a]) / MAX_SPEED, -1.0, 1.0)
                    feats += [lx, ly, lvx, lvy]
                opp[:, a, :] = np.stack(feats, axis=1)
            parts.append(opp)
        return np.concatenate(parts, axis=2).astype(np.float32)

    def _infos(self):
        """racing_env.py:77-84,156-159 / multi_racing_env.py:107-116,244-245."""
        return {
            'position': np.stack([self.x, self.y], axis=2),
            'speed': np.sqrt(self.vx ** 2 + self.vy ** 2),
            'progress': np.where(self.finished, 1.0, self.progress),
            'crashed': self.crashed.copy(),
            'finished': self.finished.copy(),
        }

    # ---- SP: the single-agent view of a 2-car env -----------------------
    @staticmethod
    def selfplay_view(obs, reward, terminated, truncated, agent_idx=0):
        """environment/wrappers.py:46-55: (obs_i, r_i, done=__all__, truncated)."""
        return obs[:, agent_idx], reward[:, agent_idx], terminated | truncated, truncated


# --------------------------------------------------------------------------
# G: generalised advantage estimation
# --------------------------------------------------------------------------
def gae(rewards, dones, values, next_value, next_done, gamma, lam):
    """agent/ppo.py:134-154 in float32 (the reference runs it on float32
    torch tensors; python-float coefficients are applied as float32 scalars)."""
    rewards = np.asarray(rewards, np.float32)
    dones = np.asarray(dones, np.float32)
    values = np.asarray(values, np.float32)
    T = rewards.shape[0]
    adv = np.zeros_like(rewards)
    running = np.zeros_like(rewards[0])
    g = np.float32(gamma)
    gl = np.float32(gamma * lam)
    for t in reversed(range(T)):
        if t == T - 1:
            nnt = np.float32(1.0) - np.asarray(next_done, np.float32)
            nv = np.asarray(next_value, np.float32)
        else:
            nnt = np.float32(1.0) - dones[t + 1]
            nv = values[t + 1]
        delta = rewards[t] + (g * nnt * nv) - values[t]
        running = delta + gl * nnt * running
        adv[t] = running
    return adv, adv + values


# --------------------------------------------------------------------------
# P: Agent forward pass (agent/ppo.py:11-62), float64
# --------------------------------------------------------------------------
def agent_forward(state_dict, obs, action=None):
    """Agent.get_action_and_value's deterministic part for a state_dict of numpy arrays (torch's key names):
    mu = actor_mu(obs) -- Linear-Tanh-Linear-Tanh-Linear-Tanh (agent/ppo.py:18-26), value = critic(obs) --
    Linear-Tanh-Linear-Tanh-Linear (ppo.py:31-37); with `action` also Normal(mu, exp(log_std)).log_prob(action)
    summed over the two action dimensions (ppo.py:45-56).  Returns (mu [n, 2], value [n], logp [n] or None)."""
    sd = {k: np.asarray(v, dtype=np.float64) for k, v in state_dict.items()}
    x = np.asarray(obs, dtype=np.float64)

    def mlp(prefix, last_tanh):
        h = np.tanh(x @ sd[prefix + '.0.weight'].T + sd[prefix + '.0.bias'])
        h = np.tanh(h @ sd[prefix + '.2.weight'].T + sd[prefix + '.2.bias'])
        o = h @ sd[prefix + '.4.weight'].T + sd[prefix + '.4.bias']
        return np.tanh(o) if last_tanh else o

    mu = mlp('actor_mu', True)
    value = mlp('critic', False)[:, 0]
    logp = None
    if action is not None:
        ls = sd['log_std']
        a = np.asarray(action, dtype=np.float64)
        logp = (-((a - mu) ** 2) / (2.0 * np.exp(ls) ** 2) - ls - 0.5 * np.log(2.0 * np.pi)).sum(1)
    return mu, value, logp
