"""Pins oracle/racing_oracle.py to fixtures recorded from the unmodified
reference (tools/make_golden.py).  Bar: discrete outputs exact; float64 state
and rewards <= 1e-9 absolute (observed: bit-equal); observations equal as
float32 up to 1 ulp-level slack of 1e-6."""
import numpy as np
import pytest

from oracle import racing_oracle as O
from tests import _finish_line as FL


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


# ---------------------------------------------------------------- the finish line
# Scripted-driver recordings of the reference (tools/make_golden.py record_laps): checkpoints, finish +
# time bonus, both wraps, the 3000-step truncation, finished_step-driven placement, +250 on
# termination and truncation (racing_env.py:112-162, multi_racing_env.py:155-211,247-259).
LAP_EVENTS_SINGLE = ('finish', 'cp', 'trunc', 'crash', 'bwd', 'fwd_nofinish')
LAP_EVENTS_MULTI = ('finish', 'trunc', 'first_crash', 'bwd', 'fwd_nofinish', 'win_bonus')


@pytest.mark.parametrize('name,min_finish', [('single_laps_default', 10), ('single_laps_proc1', 3)])
def test_single_laps(golden, name, min_finish):
    g = golden(name)
    ev = dict(zip(LAP_EVENTS_SINGLE, g['events']))
    assert ev['finish'] >= min_finish and ev['cp'] >= 3 * min_finish and ev['trunc'] >= 1 and ev['crash'] >= 1
    assert ev['bwd'] >= 1 and ev['fwd_nofinish'] >= 1
    assert g['reward'].max() > 200 and (g['truncated'] & ~g['terminated']).any()
    trk = O.TrackTables(g['control_points'], float(g['width']))
    env = O.OracleVecEnv([trk], [0], kind='single', num_sensors=11)
    obs0, _ = env.reset()
    np.testing.assert_array_equal(obs0[0, 0], g['obs0'])
    for k in range(len(g['actions'])):
        obs, r, te, tr, info = env.step(g['actions'][k][None, None, :])
        assert te[0] == g['terminated'][k] and tr[0] == g['truncated'][k], k
        np.testing.assert_allclose(obs[0, 0], g['obs'][k], rtol=0, atol=1e-6, err_msg=f'step {k}')
        assert abs(r[0, 0] - g['reward'][k]) <= 1e-9, k
        st = np.array([env.x[0, 0], env.y[0, 0], env.angle[0, 0], env.vx[0, 0], env.vy[0, 0]])
        np.testing.assert_allclose(st, g['state'][k], rtol=0, atol=1e-9, err_msg=f'step {k}')
        assert env.progress_idx[0, 0] == g['progress_idx'][k], k
        assert env.finished[0, 0] == g['finished'][k] and env.crashed[0, 0] == g['crashed'][k], k
        np.testing.assert_array_equal(env.checkpoints[0, 0], g['checkpoints'][k], err_msg=f'step {k}')


@pytest.mark.parametrize('name,A,min_finish', [('multi2_laps_default', 2, 10), ('multi2_laps_proc2', 2, 3),
                                               ('multi3_laps_proc1', 3, 3)])
def test_multi_laps(golden, name, A, min_finish):
    g = golden(name)
    ev = dict(zip(LAP_EVENTS_MULTI, g['events']))
    assert ev['finish'] >= min_finish and ev['bwd'] >= 1 and ev['fwd_nofinish'] >= 1
    if A == 2:
        assert ev['trunc'] >= 1 and (g['reward'][g['truncated']] >= 250).any()   # +250 on truncation
    if name == 'multi2_laps_default':
        assert ev['first_crash'] >= 1
    assert (g['finished_step'] > 0).any() and (g['reward'] == g['reward']).all()
    trk = O.TrackTables(g['control_points'], float(g['width']))
    env = O.OracleVecEnv([trk], [0], kind='multi', num_agents=A, num_sensors=11)
    obs0, _ = env.reset(start_order=g['start_order0'][None])
    np.testing.assert_array_equal(obs0[0], g['obs0'])
    for k in range(len(g['actions'])):
        obs, r, te, tr, info = env.step(g['actions'][k][None], start_order=g['start_order'][k][None])
        assert te[0] == g['terminated'][k] and tr[0] == g['truncated'][k], k
        np.testing.assert_allclose(obs[0], g['obs'][k], rtol=0, atol=1e-6, err_msg=f'step {k}')
        np.testing.assert_allclose(r[0], g['reward'][k], rtol=0, atol=1e-9, err_msg=f'step {k}')
        st = np.stack([env.x[0], env.y[0], env.angle[0], env.vx[0], env.vy[0]], axis=1)
        np.testing.assert_allclose(st, g['state'][k], rtol=0, atol=1e-9, err_msg=f'step {k}')
        np.testing.assert_array_equal(np.stack([env.crashed[0], env.finished[0]], axis=1), g['flags'][k])
        np.testing.assert_array_equal(env.finished_step[0], g['finished_step'][k], err_msg=f'step {k}')
        np.testing.assert_array_equal(env.checkpoints[0], g['checkpoints'][k], err_msg=f'step {k}')
        np.testing.assert_array_equal(env.progress_idx[0], g['progress_idx'][k], err_msg=f'step {k}')
        if te[0] or tr[0]:
            np.testing.assert_array_equal(info['placement'][0], g['placement'][k], err_msg=f'step {k}')


def test_injected_single_branches(golden):
    """One hand-built state per branch, scenario s = environment s (names in the fixture)."""
    g = golden('injected_single')
    S = len(g['names'])
    trk = O.TrackTables()
    np.testing.assert_array_equal(trk.waypoints, g['waypoints'])
    env = O.OracleVecEnv([trk], [0] * S, kind='single', num_sensors=11)
    env.reset()
    FL.inject_oracle_single(env, g)
    for t in range(g['actions'].shape[0]):
        obs, r, te, tr, info = env.step(g['actions'][t][:, None, :])
        np.testing.assert_array_equal(te, g['terminated'][t])
        np.testing.assert_array_equal(tr, g['truncated'][t])
        np.testing.assert_allclose(obs[:, 0], g['obs'][t], rtol=0, atol=1e-6)
        np.testing.assert_allclose(r[:, 0], g['reward'][t], rtol=0, atol=1e-9)
        st = np.stack([env.x[:, 0], env.y[:, 0], env.angle[:, 0], env.vx[:, 0], env.vy[:, 0]], axis=1)
        np.testing.assert_allclose(st, g['state'][t], rtol=0, atol=1e-9)
        np.testing.assert_array_equal(env.progress_idx[:, 0], g['progress_idx'][t])
        np.testing.assert_array_equal(env.finished[:, 0], g['finished'][t])
        np.testing.assert_array_equal(env.crashed[:, 0], g['crashed'][t])
        np.testing.assert_array_equal(env.checkpoints[:, 0], g['checkpoints'][t])
        if t == 0:
            np.testing.assert_allclose(info['progress'][:, 0], g['info_progress'][t], rtol=0, atol=1e-12)
            np.testing.assert_allclose(info['progress_delta'][:, 0], g['info_delta'][t], rtol=0, atol=1e-12)
    names = list(g['names'])
    r0 = dict(zip(names, g['reward'][0]))
    assert r0['finish_bonus_floor'] < 120 < r0['finish_time_bonus']           # max(0, 200 - steps/10) hit its floor
    assert g['terminated'][0][names.index('finish_and_truncate')] and g['truncated'][0][names.index('finish_and_truncate')]
    assert r0['bwd_wrap'] < 0 and abs(r0['cp50_hit'] - r0['cp50_needs_cp25'] - 20) < 1e-9


def test_injected_multi_branches(golden):
    g = golden('injected_multi2')
    S = len(g['names'])
    trk = O.TrackTables()
    env = O.OracleVecEnv([trk], [0] * S, kind='multi', num_agents=2, num_sensors=11)
    env.reset(start_order=np.tile([0, 1], (S, 1)))
    FL.inject_oracle_multi(env, g)
    for t in range(g['actions'].shape[0]):
        obs, r, te, tr, info = env.step(g['actions'][t], start_order=g['start_order'][t])
        np.testing.assert_array_equal(te, g['terminated'][t])
        np.testing.assert_array_equal(tr, g['truncated'][t])
        np.testing.assert_allclose(obs, g['obs'][t], rtol=0, atol=1e-6)
        np.testing.assert_allclose(r, g['reward'][t], rtol=0, atol=1e-9)
        st = np.stack([env.x, env.y, env.angle, env.vx, env.vy], axis=2)
        np.testing.assert_allclose(st, g['state'][t], rtol=0, atol=1e-9)
        np.testing.assert_array_equal(np.stack([env.crashed, env.finished], axis=2), g['flags'][t])
        np.testing.assert_array_equal(env.finished_step, g['finished_step'][t])
        np.testing.assert_array_equal(env.checkpoints, g['checkpoints'][t])
        np.testing.assert_array_equal(info['placement'], g['placement'][t])
    names = list(g['names'])
    tie = names.index('tie_on_truncation')
    assert list(g['placement'][0][tie]) == [2, 1] and list(g['reward'][0][tie]) == [0.0, 250.0]   # exact tie -> higher index
    both = names.index('both_finish_exact_tie')
    assert g['flags'][0][both][:, 1].all() and list(g['placement'][0][both]) == [2, 1]


def test_rollout_buffers_of_the_reference_loop(golden):
    """SelfPlayPPO.collect_rollout of the reference (tools/make_golden.py record_rollout_buffers) replayed through the
    oracle: buffer slot semantics, NEXT_STEP resets inside SelfPlayWrapper envs, RecordEpisodeStatistics."""
    g = golden('rollout_selfplay16')
    cps = np.split(g['pool'], np.cumsum(g['pool_sizes'])[:-1])
    tracks = O.make_pool(cps, list(g['widths']))
    T, E = g['actions0'].shape[:2]
    for it in range(2):
        env = O.OracleVecEnv(tracks, np.arange(E) % len(tracks), kind='multi', num_agents=2, num_sensors=11)
        env.reset(start_order=g[f'slots_init{it}'])
        ep_r, ep_l = [], []
        for t in range(T):
            a = np.stack([g[f'actions{it}'][t], g[f'opp_actions{it}'][t]], axis=1)
            obs, r, te, tr, info = env.step(a, start_order=g[f'slots{it}'][t])
            o, rew, done, _ = O.OracleVecEnv.selfplay_view(obs, r, te, tr)
            nxt = g[f'obs{it}'][t + 1] if t + 1 < T else g[f'next_obs{it}']
            nd = g[f'dones{it}'][t + 1] if t + 1 < T else g[f'next_done{it}']
            np.testing.assert_allclose(o, nxt, rtol=0, atol=1e-6, err_msg=f'rollout {it} step {t}')
            np.testing.assert_array_equal(done, nd.astype(bool))
            np.testing.assert_allclose(rew, g[f'rewards{it}'][t], rtol=1e-6, atol=1e-5)
            ep_r += list(info['episode_r'][info['_episode']]); ep_l += list(info['episode_l'][info['_episode']])
        np.testing.assert_allclose(ep_r, g[f'ep_r{it}'], rtol=0, atol=1e-9)
        np.testing.assert_array_equal(ep_l, g[f'ep_l{it}'])


def test_agent_forward_matches_reference_agent(golden):
    """oracle.agent_forward (float64 restatement of agent/ppo.py:11-62) against the outputs the reference's torch Agent
    recorded for the same parameters and observations: mean, value and the log-probability of the recorded action."""
    g = golden('agent')
    sd = {k[3:]: v for k, v in g.items() if k.startswith('sd.')}
    mu, value, logp = O.agent_forward(sd, g['obs'], action=g['act'])
    np.testing.assert_allclose(mu, g['mu'], rtol=0, atol=5e-7)          # float32 torch vs float64 numpy
    np.testing.assert_allclose(value, g['value'][:, 0], rtol=0, atol=2e-6)
    np.testing.assert_allclose(logp, g['logp'], rtol=0, atol=2e-5)
