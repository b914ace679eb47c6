"""Pins oracle/racing_oracle.py to fixtures recorded from the unmodified
reference (tools/make_golden.py).  Bar: discrete outputs exact; float64 state
and rewards <= 1e-9 absolute (observed: bit-equal); observations equal as
float32 up to 1 ulp-level slack of 1e-6."""
import numpy as np
import pytest

from oracle import racing_oracle as O


def _replay_single(g):
    trk = O.TrackTables(g['control_points'], float(g['width']))
    np.testing.assert_array_equal(trk.waypoints, g['waypoints'])
    env = O.OracleVecEnv([trk], [0], kind='single', num_sensors=11)
    obs0, _ = env.reset()
    np.testing.assert_array_equal(obs0[0, 0], g['obs0'])
    n = len(g['actions'])
    for k in range(n):
        obs, r, te, tr, info = env.step(g['actions'][k][None, None, :])
        assert te[0] == g['terminated'][k] and tr[0] == g['truncated'][k], k
        np.testing.assert_allclose(obs[0, 0], g['obs'][k], rtol=0, atol=1e-6, err_msg=f'step {k}')
        assert abs(r[0, 0] - g['reward'][k]) <= 1e-9, k
        st = np.array([env.x[0, 0], env.y[0, 0], env.angle[0, 0], env.vx[0, 0], env.vy[0, 0]])
        np.testing.assert_allclose(st, g['state'][k], rtol=0, atol=1e-9, err_msg=f'step {k}')
        assert env.progress_idx[0, 0] == g['progress_idx'][k], k
    return n


def test_single_default_10k(golden):
    """BASELINE config 1: RacingEnv(num_sensors=11), default track, 10k steps."""
    g = golden('single_default_10k')
    assert _replay_single(g) == 10000
    assert g['terminated'].sum() > 100  # many episodes, mostly crashes


@pytest.mark.parametrize('i', range(4))
def test_single_procedural(golden, i):
    _replay_single(golden(f'single_proc{i}'))


@pytest.mark.parametrize('name,A', [('multi2_default', 2), ('multi2_proc1', 2),
                                    ('multi2_proc2', 2), ('multi3_proc2', 3)])
def test_multi(golden, name, A):
    g = golden(name)
    trk = O.TrackTables(g['control_points'], float(g['width']))
    env = O.OracleVecEnv([trk], [0], kind='multi', num_agents=A, num_sensors=11)
    obs0, _ = env.reset(start_order=g['start_order0'][None])
    np.testing.assert_array_equal(obs0[0], g['obs0'])
    ended = 0
    for k in range(len(g['actions'])):
        obs, r, te, tr, info = env.step(g['actions'][k][None], start_order=g['start_order'][k][None])
        assert te[0] == g['terminated'][k] and tr[0] == g['truncated'][k], k
        np.testing.assert_allclose(obs[0], g['obs'][k], rtol=0, atol=1e-6, err_msg=f'step {k}')
        np.testing.assert_allclose(r[0], g['reward'][k], rtol=0, atol=1e-9, err_msg=f'step {k}')
        st = np.stack([env.x[0], env.y[0], env.angle[0], env.vx[0], env.vy[0]], axis=1)
        np.testing.assert_allclose(st, g['state'][k], rtol=0, atol=1e-9, err_msg=f'step {k}')
        np.testing.assert_array_equal(np.stack([env.crashed[0], env.finished[0]], axis=1), g['flags'][k])
        if te[0] or tr[0]:
            ended += 1
            np.testing.assert_array_equal(info['placement'][0], g['placement'][k])
    assert ended >= 5


def test_vector_autoreset_and_episode_stats(golden):
    """SyncVectorEnv(NEXT_STEP) + RecordEpisodeStatistics over 4 envs on 4 tracks."""
    g = golden('vector_single4')
    sizes = g['pool_sizes']
    cps = np.split(g['pool'], np.cumsum(sizes)[:-1])
    tracks = O.make_pool(cps, list(g['widths']))
    env = O.OracleVecEnv(tracks, [0, 1, 2, 3], kind='single', num_sensors=11)
    obs0, _ = env.reset()
    np.testing.assert_array_equal(obs0[:, 0], g['obs0'])
    for k in range(len(g['actions'])):
        obs, r, te, tr, info = env.step(g['actions'][k][:, None, :])
        np.testing.assert_array_equal(te, g['terminated'][k])
        np.testing.assert_array_equal(tr, g['truncated'][k])
        np.testing.assert_allclose(obs[:, 0], g['obs'][k], rtol=0, atol=1e-6)
        np.testing.assert_allclose(r[:, 0], g['reward'][k], rtol=0, atol=1e-9)
        np.testing.assert_array_equal(info['_episode'], g['ep_mask'][k])
        np.testing.assert_allclose(info['episode_r'], g['ep_r'][k], rtol=0, atol=1e-9)
        np.testing.assert_array_equal(info['episode_l'], g['ep_l'][k])
    assert g['ep_mask'].sum() > 20


def test_track_tables_and_queries(golden):
    g = golden('tracks')
    for i in range(int(g['n'])):
        trk = O.TrackTables(g[f'cp{i}'], float(g[f'width{i}']))
        np.testing.assert_array_equal(trk.waypoints, g[f'wp{i}'])
        np.testing.assert_array_equal(trk.normals, g[f'nrm{i}'])
        np.testing.assert_array_equal(trk.starts, g[f'starts{i}'])
        np.testing.assert_array_equal(trk.v2, g[f'v2{i}'])
        assert trk.max_track_distance == g[f'mtd{i}']
        np.testing.assert_array_equal(np.array(trk.start_pose()), g[f'start{i}'])
        pts, ang = g[f'q_pts{i}'], g[f'q_ang{i}']
        idx, _, _ = O.closest_waypoint_idx(trk, pts[:, 0], pts[:, 1])
        np.testing.assert_array_equal(idx, g[f'q_idx{i}'])
        ray = O.raycast_walls(trk, pts[:, 0], pts[:, 1], ang[:, None])[:, 0]
        np.testing.assert_allclose(ray, g[f'q_ray{i}'], rtol=1e-13, atol=0)


def test_procedural_generator_draw_for_draw(golden):
    """gen_tracks(16, seed=1) incl. the re-seeding collapse."""
    g = golden('tracks')
    np.random.seed(1)
    pool = O.gen_tracks(num_tracks=16, seed=1)
    widths = [np.random.randint(6, 10) for _ in range(16)]
    np.testing.assert_array_equal(np.array([len(p) for p in pool]), g['train_pool_sizes'])
    np.testing.assert_array_equal(np.concatenate(pool), g['train_pool'])
    np.testing.assert_array_equal(np.array(widths), g['train_widths'])
    assert len({p.tobytes() for p in pool}) == 4
    np.testing.assert_array_equal(O.gen_random_track(13, 62, 17, 0.45, 0.35, seed=9), g['rand_track'])


def test_gae(golden):
    g = golden('gae')
    for lam, tag in ((0.97, 'sp'), (0.95, 'single')):
        adv, ret = O.gae(g['rewards'], g['dones'], g['values'], g['next_value'], g['next_done'], 0.99, lam)
        np.testing.assert_allclose(adv, g[f'adv_{tag}'], rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(ret, g[f'ret_{tag}'], rtol=1e-6, atol=1e-6)
