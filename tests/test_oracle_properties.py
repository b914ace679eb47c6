"""Property tests of the oracle's geometric primitives (SURVEY.md section 4,
tier 3): they guard the checker itself, and state the invariants the GPU
invariants test relies on."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import racing_oracle as O

TRACK = O.TrackTables()  # the reference's fixed default polygon


@settings(max_examples=60, deadline=None)
@given(st.floats(0, 1), st.floats(-1, 1), st.floats(0, 2 * np.pi))
def test_ray_distance_is_non_negative_and_hits_a_wall_from_inside(frac, lateral, angle):
    k = int(frac * (TRACK.num_waypoints - 1))
    p = TRACK.waypoints[k] + lateral * 0.9 * TRACK.track_width * TRACK.normals[k]
    d = O.raycast_walls(TRACK, p[:1], p[1:], np.array([[angle]]))[0, 0]
    assert d >= 0.0
    assert d < TRACK.max_track_distance  # a point inside the closed corridor always sees a wall


@settings(max_examples=60, deadline=None)
@given(st.floats(-20, 20), st.floats(-20, 20), st.floats(0, 2 * np.pi), st.floats(-20, 20), st.floats(-20, 20),
       st.floats(0, 2 * np.pi))
def test_sat_is_symmetric_and_detects_overlap_of_centres(x0, y0, a0, x1, y1, a1):
    ax, ay = O.corners_of(np.array([x0]), np.array([y0]), np.array([a0]))
    bx, by = O.corners_of(np.array([x1]), np.array([y1]), np.array([a1]))
    h1, _ = O.rectangles_intersect(ax, ay, bx, by)
    h2, _ = O.rectangles_intersect(bx, by, ax, ay)
    assert h1[0] == h2[0]
    dist = np.hypot(x1 - x0, y1 - y0)
    if dist < 2.0:          # closer than the car width: must overlap
        assert h1[0]
    if dist > 2 * np.sqrt(5) + 1e-9:  # farther than two half-diagonals: can not overlap
        assert not h1[0]


@settings(max_examples=40, deadline=None)
@given(st.floats(0, 1), st.floats(-1.5, 1.5))
def test_closest_waypoint_is_a_global_minimum(frac, lateral):
    k = int(frac * (TRACK.num_waypoints - 1))
    p = TRACK.waypoints[k] + lateral * TRACK.track_width * TRACK.normals[k]
    idx, best, second = O.closest_waypoint_idx(TRACK, p[:1], p[1:])
    d = ((TRACK.waypoints - p) ** 2).sum(1)
    assert d[idx[0]] == d.min() == best[0] and second[0] >= best[0]


def test_gae_closed_form_without_dones():
    rs = np.random.RandomState(0)
    T, E = 32, 5
    r, v = rs.normal(size=(T, E)).astype(np.float32), rs.normal(size=(T, E)).astype(np.float32)
    nv = rs.normal(size=E).astype(np.float32)
    adv, ret = O.gae(r, np.zeros((T, E), np.float32), v, nv, np.zeros(E, bool), 0.99, 0.95)
    vn = np.concatenate([v[1:], nv[None]], 0)
    delta = r + 0.99 * vn - v
    w = (0.99 * 0.95) ** np.arange(T)
    closed = np.array([(delta[t:] * w[:T - t, None]).sum(0) for t in range(T)])
    np.testing.assert_allclose(adv, closed, rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(ret, adv + v, rtol=0, atol=1e-6)
