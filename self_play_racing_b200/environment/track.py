"""Host side of the track path: procedural control points drawn from the
global legacy numpy stream (so pools equal the reference's draw for draw,
including the re-seeding collapse of gen_tracks -- SURVEY quirk 8), and `Track`,
a read-only view of a track whose tables were built ON THE DEVICE.

Reference: environment/track.py:4-56 (generation), :58-171 (Track attributes
that utils/visualization.py and the envs read).
"""
from __future__ import annotations

import numpy as np

# environment/track.py:69-74 -- polygon used when neither pool nor points are given
DEFAULT_CONTROL_POINTS = np.array([[0, 0], [50, 0], [70, 20], [60, 40], [70, 50],
                                   [50, 70], [20, 70], [10, 50], [10, 20], [0, 10]], dtype=np.float64)


def gen_random_track(num_points=15, base_radius=50, radius_variation=15, angle_jitter=0.2,
                     smoothness=0.5, seed=None):
    """Jittered polar control points with a smoothed radius walk
    (environment/track.py:4-45).  Draws from the global ``np.random`` stream in
    the reference's order: one vector of angle offsets, then one radius
    variation per point."""
    if seed is not None:
        np.random.seed(seed)
    theta = np.linspace(0, 2 * np.pi, num_points, endpoint=False)
    if angle_jitter > 0:
        half_span = angle_jitter * (2 * np.pi / num_points) / 2
        theta = np.sort((theta + np.random.uniform(-half_span, half_span, num_points)) % (2 * np.pi))
    raw = base_radius + np.random.uniform(-radius_variation, radius_variation, num_points)
    if smoothness > 0:
        radius = raw.copy()
        keep = 1 - smoothness
        for k in range(1, num_points):  # first-order recursive smoothing, sequential by construction
            radius[k] = keep * raw[k] + (smoothness * radius[k - 1])
        radius[0] = (radius[0] + radius[-1]) / 2
    else:
        radius = raw
    return np.column_stack([radius * np.cos(theta), radius * np.sin(theta)])


def gen_tracks(num_tracks=10, seed=None):
    """A pool of control-point arrays (environment/track.py:47-56)."""
    pool = []
    for _ in range(num_tracks):
        n_pts = np.random.randint(10, 15)
        base = np.random.randint(50, 80)
        variation = np.random.randint(10, base // 2 - 10)
        jitter = np.random.uniform(0.2, 0.7)
        smooth = np.random.uniform(0.2, 0.7)
        pool.append(gen_random_track(n_pts, base, variation, jitter, smooth, seed))
    return pool


def resolve_track(control_points=None, track_width=None, track_pool=None, track_id=None):
    """Argument handling of Track.__init__ (environment/track.py:61-80):
    returns (control_points float64 [n,2], width float, track_id)."""
    if track_pool is not None:
        if track_id is None:
            track_id = np.random.randint(0, len(track_pool))
        control_points = track_pool[track_id]
        if track_width is not None and isinstance(track_width, list):
            track_width = track_width[track_id]
    if control_points is None:
        control_points = DEFAULT_CONTROL_POINTS
    width = 6.0 if track_width is None else float(track_width)
    return np.asarray(control_points, dtype=np.float64), width, track_id


class Track:
    """Read-only host view of one device-built track: ``waypoints``,
    ``normals``, ``left_boundary``, ``right_boundary``, ``track_width``,
    ``max_track_distance``, ``control_points`` (what utils/visualization.py:12-59
    reads) and ``get_start_pos``.  Queries (closest waypoint, wall test,
    raycast) exist only inside the step kernel -- there is no host path."""

    def __init__(self, tables):
        self.control_points = tables['control_points']
        self.track_width = tables['track_width']
        self.waypoints = tables['waypoints']
        self.normals = tables['normals']
        self.left_boundary = tables['left_boundary']
        self.right_boundary = tables['right_boundary']
        self.max_track_distance = tables['max_track_distance']
        self._start = tables['start_pos']
        w = self.waypoints
        self.track_bounds = {'min_x': w[:, 0].min(), 'max_x': w[:, 0].max(),
                             'min_y': w[:, 1].min(), 'max_y': w[:, 1].max()}

    def get_start_pos(self):
        return self._start
