"""GPU parity tests: the CUDA path (through the C ABI) against the oracle and
against the golden trajectories recorded from the reference.

Bars (SURVEY.md 8c / BASELINE.json north_star):
  * discrete outputs (terminated, truncated, crashed, finished, placement,
    progress index, auto-reset) -- exact;
  * float64 state and rewards -- <= 1e-9 absolute (the kernel follows the
    reference's float64 operation order; only sin/cos/atan2 may differ by an ulp);
  * observations (float32) -- <= 1e-6 absolute, far inside the stated 1e-4
    relative tolerance;
  * GAE -- <= 1e-5.
"""
import numpy as np
import pytest

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu

from oracle import racing_oracle as O  # noqa: E402
from tests import _finish_line as FL  # noqa: E402


@pytest.fixture(scope='module')
def B():
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    from self_play_racing_b200 import backend
    return backend


QUERY_MODES = ['exact', 'culled', 'grid']
OBS_ATOL = 1e-6
STATE_ATOL = 1e-9


def _state(be):
    s = be.get_state()
    return s['car_f64'][..., :5], s['car_i32']


# ------------------------------------------------------------------ tracks
def test_track_build_matches_scipy(B, golden):
    """Device periodic spline + tables vs scipy/numpy (track.py:61-148)."""
    g = golden('tracks')
    n = int(g['n'])
    cps = [g[f'cp{i}'] for i in range(n)]
    widths = [float(g[f'width{i}']) for i in range(n)]
    be = B.RacingBackend(n, kind='single')
    be.set_tracks_from_control_points(cps, widths)
    for i in range(n):
        t = be.get_track(i)
        np.testing.assert_allclose(t['waypoints'], g[f'wp{i}'], rtol=0, atol=1e-11)
        np.testing.assert_allclose(t['normals'], g[f'nrm{i}'], rtol=0, atol=1e-9)
        np.testing.assert_allclose(t['left_boundary'], g[f'starts{i}'][:len(t['waypoints'])], rtol=0, atol=1e-9)
        assert abs(t['max_track_distance'] - float(g[f'mtd{i}'])) < 1e-10
        np.testing.assert_allclose(np.array(t['start_pos']), g[f'start{i}'], rtol=0, atol=1e-12)
    # from the reference's own waypoints every table is bit-identical
    be2 = B.RacingBackend(n, kind='single')
    be2.set_tracks_from_waypoints([g[f'wp{i}'] for i in range(n)], widths)
    for i in range(n):
        t = be2.get_track(i)
        nw = len(t['waypoints'])
        np.testing.assert_array_equal(t['waypoints'], g[f'wp{i}'])
        np.testing.assert_array_equal(t['normals'], g[f'nrm{i}'])
        np.testing.assert_array_equal(t['left_boundary'], g[f'starts{i}'][:nw])
        np.testing.assert_array_equal(t['right_boundary'], g[f'starts{i}'][nw:])
        assert t['max_track_distance'] == float(g[f'mtd{i}'])
    be.close()
    be2.close()


def test_generated_tracks_are_valid(B):
    be = B.RacingBackend(64, kind='single')
    be.generate_tracks(seed=3, n_tracks=16)
    assert be.num_tracks == 16
    seen = set()
    for i in range(16):
        t = be.get_track(i)
        n_ctrl = len(t['control_points'])
        assert 10 <= n_ctrl <= 14 and len(t['waypoints']) == 30 * n_ctrl
        r = np.hypot(t['control_points'][:, 0], t['control_points'][:, 1])
        assert r.min() > 15 and r.max() < 125
        np.testing.assert_allclose(np.hypot(t['normals'][:, 0], t['normals'][:, 1]), 1.0, atol=1e-12)
        # must equal the oracle's tables for the same control points
        o = O.TrackTables(t['control_points'], t['track_width'])
        np.testing.assert_allclose(t['waypoints'], o.waypoints, rtol=0, atol=1e-10)
        seen.add(t['control_points'].tobytes())
    assert len(seen) == 16
    be.close()


# ------------------------------------------------- golden trajectories (reference)
@pytest.mark.parametrize('query', QUERY_MODES)
@pytest.mark.parametrize('name', ['single_default_10k', 'single_proc0', 'single_proc1', 'single_proc2', 'single_proc3',
                                  'single_laps_default', 'single_laps_proc1'])
def test_single_env_golden(B, golden, name, query):
    """BASELINE config 1: lock-step with the reference over a free-running
    trajectory (10k steps on the default track), identical actions.  The *_laps_*
    recordings are driven by a pure-pursuit script and run the whole reward state
    machine: checkpoints, finish + time bonus, both wraps, 3000-step truncation
    (racing_env.py:112-162)."""
    g = golden(name)
    be = B.RacingBackend(1, kind='single', num_sensors=11, query=query)
    be.set_tracks_from_waypoints([g['waypoints']], [float(g['width'])])
    obs0 = be.reset().cpu().numpy()
    np.testing.assert_allclose(obs0[0, 0], g['obs0'], rtol=0, atol=OBS_ATOL)
    n = len(g['actions'])
    acts = torch.from_numpy(g['actions']).cuda()
    OBS = torch.zeros(n, 15, device='cuda')
    REW = torch.zeros(n, dtype=torch.float64, device='cuda')
    TERM = torch.zeros(n, dtype=torch.uint8, device='cuda')
    TRUNC = torch.zeros(n, dtype=torch.uint8, device='cuda')
    ST = torch.zeros(n, 4, dtype=torch.float64, device='cuda')
    PIDX = torch.zeros(n, dtype=torch.int32, device='cuda')
    INFO = torch.zeros(n, 4, dtype=torch.int32, device='cuda')
    for k in range(n):
        be.actions[0, 0].copy_(acts[k])
        be.step()
        OBS[k] = be.obs[0, 0]; REW[k] = be.reward64[0, 0]
        TERM[k] = be.terminated[0]; TRUNC[k] = be.truncated[0]
        ST[k, :2] = be.info_f64[0, 0, :2]; PIDX[k] = be.info_i32[0, 0, 3]; INFO[k] = be.info_i32[0, 0]
    np.testing.assert_array_equal(TERM.cpu().numpy().astype(bool), g['terminated'])
    np.testing.assert_array_equal(TRUNC.cpu().numpy().astype(bool), g['truncated'])
    np.testing.assert_allclose(OBS.cpu().numpy(), g['obs'], rtol=0, atol=OBS_ATOL)
    np.testing.assert_allclose(REW.cpu().numpy(), g['reward'], rtol=0, atol=STATE_ATOL)
    # info position is the step's own (pre-reset) position; golden state is post-reset on reset steps
    prev_done = np.concatenate([[False], (g['terminated'] | g['truncated'])[:-1]])
    np.testing.assert_allclose(ST.cpu().numpy()[~prev_done, :2], g['state'][~prev_done, :2], rtol=0, atol=STATE_ATOL)
    np.testing.assert_array_equal(PIDX.cpu().numpy()[~prev_done], g['progress_idx'][~prev_done])
    if 'finished' in g:   # scripted-driver recordings: the finish line was crossed, a crawl was truncated
        info = INFO.cpu().numpy()
        np.testing.assert_array_equal(info[~prev_done, 0].astype(bool), g['crashed'][~prev_done])
        np.testing.assert_array_equal(info[~prev_done, 1].astype(bool), g['finished'][~prev_done])
        assert g['finished'].sum() >= 3 and (g['truncated'] & ~g['terminated']).sum() >= 1
    be.close()


@pytest.mark.parametrize('query', QUERY_MODES)
@pytest.mark.parametrize('name,A', [('multi2_default', 2), ('multi2_proc1', 2), ('multi2_proc2', 2), ('multi3_proc2', 3),
                                    ('multi2_laps_default', 2), ('multi2_laps_proc2', 2), ('multi3_laps_proc1', 3)])
def test_multi_env_golden(B, golden, name, A, query):
    """Free-running against the reference's recordings.  The *_laps_* files hold laps driven by a
    pure-pursuit script: checkpoints, finish + time bonus, finished_step-driven placement, +250 on
    termination and truncation, wraps, cars bumping (multi_racing_env.py:155-211,222-259)."""
    g = golden(name)
    trk = O.TrackTables(g['control_points'], float(g['width']))
    be = B.RacingBackend(1, kind='multi', num_agents=A, num_sensors=11, query=query)
    be.set_tracks_from_waypoints([trk.waypoints], [float(g['width'])])
    slot0 = torch.from_numpy(g['start_order0'].astype(np.int32)[None]).cuda()
    obs0 = be.reset(start_slot=slot0).cpu().numpy()
    np.testing.assert_allclose(obs0[0], g['obs0'], rtol=0, atol=OBS_ATOL)
    n = len(g['actions'])
    D = g['obs'].shape[-1]
    acts = torch.from_numpy(g['actions']).cuda()
    slots = torch.from_numpy(g['start_order'].astype(np.int32)).cuda()
    OBS = torch.zeros(n, A, D, device='cuda')
    REW = torch.zeros(n, A, dtype=torch.float64, device='cuda')
    TERM = torch.zeros(n, dtype=torch.uint8, device='cuda')
    TRUNC = torch.zeros(n, dtype=torch.uint8, device='cuda')
    INFO = torch.zeros(n, A, 4, dtype=torch.int32, device='cuda')
    for k in range(n):
        be.actions[0].copy_(acts[k])
        be.step(start_slot=slots[k:k + 1])
        OBS[k] = be.obs[0]; REW[k] = be.reward64[0]
        TERM[k] = be.terminated[0]; TRUNC[k] = be.truncated[0]; INFO[k] = be.info_i32[0]
    np.testing.assert_array_equal(TERM.cpu().numpy().astype(bool), g['terminated'])
    np.testing.assert_array_equal(TRUNC.cpu().numpy().astype(bool), g['truncated'])
    np.testing.assert_allclose(OBS.cpu().numpy(), g['obs'], rtol=0, atol=OBS_ATOL)
    np.testing.assert_allclose(REW.cpu().numpy(), g['reward'], rtol=0, atol=STATE_ATOL)
    info = INFO.cpu().numpy()
    ended = g['terminated'] | g['truncated']
    np.testing.assert_array_equal(info[ended, :, 2], g['placement'][ended])
    prev_done = np.concatenate([[False], ended[:-1]])
    np.testing.assert_array_equal(info[~prev_done, :, 0].astype(bool), g['flags'][~prev_done, :, 0])
    np.testing.assert_array_equal(info[~prev_done, :, 1].astype(bool), g['flags'][~prev_done, :, 1])
    if 'progress_idx' in g:
        np.testing.assert_array_equal(info[~prev_done, :, 3], g['progress_idx'][~prev_done])
        assert g['flags'][..., 1].any() and (g['placement'] == 1).sum() == ended.sum()
    be.close()


def test_vector_autoreset_golden(B, golden):
    """NEXT_STEP auto-reset + RecordEpisodeStatistics over 4 envs / 4 tracks."""
    g = golden('vector_single4')
    cps = np.split(g['pool'], np.cumsum(g['pool_sizes'])[:-1])
    tracks = O.make_pool(cps, list(g['widths']))
    be = B.RacingBackend(4, kind='single', num_sensors=11)
    be.set_tracks_from_waypoints([t.waypoints for t in tracks], list(g['widths']), env_to_track=[0, 1, 2, 3])
    obs0 = be.reset().cpu().numpy()
    np.testing.assert_allclose(obs0[:, 0], g['obs0'], rtol=0, atol=OBS_ATOL)
    for k in range(len(g['actions'])):
        be.actions[:, 0].copy_(torch.from_numpy(g['actions'][k]))
        be.step()
        np.testing.assert_array_equal(be.terminated.cpu().numpy().astype(bool), g['terminated'][k])
        np.testing.assert_array_equal(be.truncated.cpu().numpy().astype(bool), g['truncated'][k])
        np.testing.assert_allclose(be.obs[:, 0].cpu().numpy(), g['obs'][k], rtol=0, atol=OBS_ATOL)
        np.testing.assert_allclose(be.reward64[:, 0].cpu().numpy(), g['reward'][k], rtol=0, atol=STATE_ATOL)
        m = be.ep_mask.cpu().numpy().astype(bool)
        np.testing.assert_array_equal(m, g['ep_mask'][k])
        np.testing.assert_allclose(be.ep_return.cpu().numpy()[m], g['ep_r'][k][m], rtol=0, atol=STATE_ATOL)
        np.testing.assert_array_equal(be.ep_length.cpu().numpy()[m], g['ep_l'][k][m])
    be.close()


# ------------------------------------------------------ batched, live oracle
def _pool(n_tracks, seed):
    rs = np.random.RandomState(seed)
    cps = [O.gen_random_track(rs.randint(10, 15), rs.randint(50, 80), rs.randint(10, 20),
                              rs.uniform(0.2, 0.7), rs.uniform(0.2, 0.7), rng=rs) for _ in range(n_tracks)]
    widths = [float(rs.randint(6, 10)) for _ in range(n_tracks)]
    return cps, widths


@pytest.mark.parametrize('query', QUERY_MODES + ['culled-staged'])
@pytest.mark.parametrize('kind,A,E,steps', [('single', 1, 192, 260), ('multi', 2, 160, 260), ('multi', 4, 48, 120)])
def test_batched_lockstep_vs_oracle(B, kind, A, E, steps, query, monkeypatch):
    """E envs over 6 procedural tracks, free-running against the oracle on the
    same seeded actions and injected start slots; ragged N (300..420)."""
    if query == 'culled-staged':  # the opt-in launch that stages each CTA's track tables with bulk async copies
        monkeypatch.setenv('RK_B200_STAGED', '1')
        query = 'culled'
    cps, widths = _pool(6, seed=21 + A)
    if A == 4:
        widths = [w + 3 for w in widths]  # 4 cars abreast need width >= 8
    tracks = O.make_pool(cps, widths)
    e2t = np.arange(E) % 6
    orc = O.OracleVecEnv(tracks, e2t, kind=kind, num_agents=A, num_sensors=11, seed=5)
    be = B.RacingBackend(E, kind=kind, num_agents=A, num_sensors=11, query=query)
    be.set_tracks_from_waypoints([t.waypoints for t in tracks], widths, env_to_track=e2t)
    rs = np.random.RandomState(9)
    so = orc._draw_start_order(E)
    oobs, _ = orc.reset(start_order=so)
    gobs = be.reset(start_slot=torch.from_numpy(so.astype(np.int32)).cuda()).cpu().numpy()
    np.testing.assert_allclose(gobs, oobs, rtol=0, atol=OBS_ATOL)
    n_done = 0
    for k in range(steps):
        a = rs.uniform(-1, 1, size=(E, A, 2)).astype(np.float32)
        a[..., 1] = np.abs(a[..., 1]) if kind == 'multi' else a[..., 1] * 0.5 + 0.5
        so = orc._draw_start_order(E)
        oobs, orew, ote, otr, oinf = orc.step(a, start_order=so)
        be.actions.copy_(torch.from_numpy(a))
        be.step(start_slot=torch.from_numpy(so.astype(np.int32)).cuda())
        np.testing.assert_array_equal(be.terminated.cpu().numpy().astype(bool), ote, err_msg=f'step {k}')
        np.testing.assert_array_equal(be.truncated.cpu().numpy().astype(bool), otr, err_msg=f'step {k}')
        np.testing.assert_allclose(be.obs.cpu().numpy(), oobs, rtol=0, atol=OBS_ATOL, err_msg=f'step {k}')
        np.testing.assert_allclose(be.reward64.cpu().numpy(), orew, rtol=0, atol=STATE_ATOL, err_msg=f'step {k}')
        n_done += int((ote | otr).sum())
    st, sti = _state(be)
    ost = np.stack([orc.x, orc.y, orc.angle, orc.vx, orc.vy], axis=2)
    np.testing.assert_allclose(st, ost, rtol=0, atol=STATE_ATOL)
    np.testing.assert_array_equal(sti[..., 0], orc.progress_idx)
    assert n_done > E // 4
    be.close()


def test_same_step_autoreset_and_philox_slots(B):
    """SAME_STEP mode returns the reset observation on the terminal step; the
    built-in Philox shuffle puts the two cars on distinct slots +-1.75 from the line."""
    cps, widths = _pool(2, seed=3)
    tracks = O.make_pool(cps, widths)
    E = 64
    be = B.RacingBackend(E, kind='multi', num_agents=2, autoreset='same_step', seed=11)
    be.set_tracks_from_waypoints([t.waypoints for t in tracks], widths)
    be.reset()
    s = be.get_state()['car_f64']
    for e in range(E):
        t = tracks[e % 2]
        off = (s[e, :, 0] - t.waypoints[0, 0]) * t.normals[0, 0] + (s[e, :, 1] - t.waypoints[0, 1]) * t.normals[0, 1]
        assert sorted(np.round(off, 9)) == [-1.75, 1.75]
    first = s[:, 0, 0].copy()
    be.actions[..., 0] = 0.0   # straight ahead at full throttle: everybody leaves the track at the first bend
    be.actions[..., 1] = 1.0
    saw = 0
    for _ in range(400):
        be.step()
        done = be.done.cpu().numpy().astype(bool)
        if done.any():
            st = be.get_state()
            assert (st['env_i32'][done, 0] == 0).all()  # steps reset in the same call
            assert (st['car_f64'][done][..., 3:5] == 0).all()
            saw += int(done.sum())
    assert saw > E
    be.close()
    del first


def test_state_roundtrip_and_observe(B):
    cps, widths = _pool(3, seed=8)
    tracks = O.make_pool(cps, widths)
    be = B.RacingBackend(9, kind='multi', num_agents=2)
    be.set_tracks_from_waypoints([t.waypoints for t in tracks], widths)
    be.reset()
    be.actions.uniform_(-1, 1)
    for _ in range(20):
        be.step()
    s = be.get_state()
    obs = be.obs.clone()
    be2 = B.RacingBackend(9, kind='multi', num_agents=2)
    be2.set_tracks_from_waypoints([t.waypoints for t in tracks], widths)
    be2.reset()
    be2.set_state(**s)
    s2 = be2.get_state()
    for k in s:
        np.testing.assert_array_equal(s[k], s2[k])
    assert torch.equal(be2.observe(), obs)
    be.close(); be2.close()


# ---------------------------------------------------------------- GAE, policy
def test_gae_golden_and_oracle(B, golden):
    g = golden('gae')
    dev = 'cuda'
    t = lambda k: torch.from_numpy(g[k]).to(dev)
    for lam, tag in ((0.97, 'sp'), (0.95, 'single')):
        adv, ret = B.gae(t('rewards'), t('values'), t('dones'), t('next_value'), t('next_done'), 0.99, lam)
        np.testing.assert_allclose(adv.cpu().numpy(), g[f'adv_{tag}'], rtol=0, atol=1e-5)
        np.testing.assert_allclose(ret.cpu().numpy(), g[f'ret_{tag}'], rtol=0, atol=1e-5)
    rs = np.random.RandomState(1)
    T, E = 128, 1000
    r = rs.normal(0, 5, (T, E)).astype(np.float32); v = rs.normal(0, 5, (T, E)).astype(np.float32)
    d = (rs.uniform(size=(T, E)) < 0.05).astype(np.float32)
    nv = rs.normal(0, 5, E).astype(np.float32); nd = (rs.uniform(size=E) < 0.5)
    oadv, oret = O.gae(r, d, v, nv, nd, 0.99, 0.97)
    adv, ret = B.gae(*(torch.from_numpy(x).to(dev) for x in (r, v, d, nv)), torch.from_numpy(nd).to(dev), 0.99, 0.97)
    np.testing.assert_allclose(adv.cpu().numpy(), oadv, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(ret.cpu().numpy(), oret, rtol=1e-5, atol=1e-5)
    # dones == 0 everywhere: closed form sum_k (gamma*lambda)^k delta_{t+k}
    d0 = np.zeros_like(d)
    adv0, _ = B.gae(*(torch.from_numpy(x).to(dev) for x in (r, v, d0, nv)), torch.zeros(E, device=dev), 0.99, 0.97)
    vn = np.concatenate([v[1:], nv[None]], 0)
    delta = r + 0.99 * vn - v
    w = (0.99 * 0.97) ** np.arange(T)
    closed = np.array([(delta[t0:] * w[:T - t0, None]).sum(0) for t0 in range(T)])
    np.testing.assert_allclose(adv0.cpu().numpy(), closed, rtol=1e-3, atol=1e-3)


def test_policy_act_matches_torch_agent(B, golden):
    """Fused MLP inference vs the reference Agent's recorded outputs (agent/ppo.py:43-56)."""
    g = golden('agent')
    sd = {k[3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith('sd.')}
    params = B.flatten_agent(sd).cuda()
    assert params.numel() == 11080  # 11,075 parameters + log_std (2), padded to a multiple of 4
    obs = torch.from_numpy(g['obs']).cuda()
    n = obs.shape[0]
    act = torch.zeros(n, 2, device='cuda'); lp = torch.zeros(n, device='cuda')
    val = torch.zeros(n, device='cuda'); mu = torch.zeros(n, 2, device='cuda')
    B.policy_act(params, obs, act, seed=1, counter=0, logprob=lp, value=val, mean=mu)
    np.testing.assert_allclose(mu.cpu().numpy(), g['mu'], rtol=0, atol=2e-6)
    np.testing.assert_allclose(val.cpu().numpy(), g['value'][:, 0], rtol=0, atol=2e-5)
    assert act.abs().max() <= 1.0
    # log-prob of the returned (clamped) action under N(mu, exp(-0.3))
    std = np.exp(np.float32(-0.3))
    a = act.cpu().numpy()
    ref_lp = (-((a - g['mu']) ** 2) / (2 * std * std) - np.float32(-0.3) - 0.5 * np.log(2 * np.pi)).sum(1)
    np.testing.assert_allclose(lp.cpu().numpy(), ref_lp, rtol=0, atol=1e-4)
    # noise statistics over a large batch: z = (a - mu)/std ~ N(0,1) where not clamped
    # (log_std lowered to -2 so that the [-1, 1] clamp does not truncate the sample)
    big = obs.repeat(4096, 1)
    nb = big.shape[0]
    act = torch.zeros(nb, 2, device='cuda'); mu = torch.zeros(nb, 2, device='cuda')
    quiet = params.clone()
    quiet[5570:5572] = -2.0
    B.policy_act(quiet, big, act, seed=2, counter=7, mean=mu)
    z = (act - mu) / float(np.exp(-2.0))
    assert float((act.abs() >= 1).float().mean()) < 1e-3
    assert abs(float(z.mean())) < 0.02 and abs(float(z.std()) - 1) < 0.02
    assert abs(float((z[:, 0] * z[:, 1]).mean())) < 0.02   # the two action dims are independent
    act2 = torch.zeros_like(act)
    B.policy_act(quiet, big, act2, seed=2, counter=7)
    assert torch.equal(act, act2)                      # counter-based: reproducible
    B.policy_act(quiet, big, act2, seed=2, counter=8)
    assert not torch.equal(act, act2)
    # strided views: write car 1's action slice from car 1's obs slice of [E,A,*] tensors
    E = 128
    obs3 = torch.randn(E, 2, 19, device='cuda').clamp(-1, 1); act3 = torch.zeros(E, 2, 2, device='cuda')
    B.policy_act(params, obs3[:, 1], act3[:, 1], seed=3, counter=0)
    ref = torch.zeros(E, 2, device='cuda')
    B.policy_act(params, obs3[:, 1].contiguous(), ref, seed=3, counter=0)
    assert torch.equal(act3[:, 1], ref) and (act3[:, 0] == 0).all()
    # pool-empty opponent: uniform Box([-1,0],[1,1])
    B.policy_act(None, None, act3[:, 1], seed=4, counter=0)
    r = act3[:, 1].cpu().numpy()
    assert (r[:, 0] >= -1).all() and (r[:, 0] <= 1).all() and (r[:, 1] >= 0).all() and (r[:, 1] <= 1).all()
    assert r[:, 0].std() > 0.4


@pytest.mark.parametrize('query', QUERY_MODES)
def test_sweep_shape_64_rays_4_cars_vs_oracle(B, query):
    """BASELINE config 5 shape (4 cars, 64 rays, obs 80-dim) at a size the oracle
    finishes in seconds; E = 21 is deliberately not a multiple of the 8 envs a warp holds."""
    cps, widths = _pool(3, seed=31)
    widths = [w + 3 for w in widths]
    tracks = O.make_pool(cps, widths)
    E, A, R = 21, 4, 64
    e2t = np.arange(E) % 3
    orc = O.OracleVecEnv(tracks, e2t, kind='multi', num_agents=A, num_sensors=R, seed=2)
    be = B.RacingBackend(E, kind='multi', num_agents=A, num_sensors=R, query=query)
    assert be.D == 80
    be.set_tracks_from_waypoints([t.waypoints for t in tracks], widths, env_to_track=e2t)
    rs = np.random.RandomState(4)
    so = orc._draw_start_order(E)
    oobs, _ = orc.reset(start_order=so)
    gobs = be.reset(start_slot=torch.from_numpy(so.astype(np.int32)).cuda()).cpu().numpy()
    np.testing.assert_allclose(gobs, oobs, rtol=0, atol=OBS_ATOL)
    for k in range(60):
        a = rs.uniform(-1, 1, size=(E, A, 2)).astype(np.float32)
        a[..., 1] = np.abs(a[..., 1])
        so = orc._draw_start_order(E)
        oobs, orew, ote, otr, _ = orc.step(a, start_order=so)
        be.actions.copy_(torch.from_numpy(a))
        be.step(start_slot=torch.from_numpy(so.astype(np.int32)).cuda())
        np.testing.assert_array_equal(be.terminated.cpu().numpy().astype(bool), ote, err_msg=f'step {k}')
        np.testing.assert_allclose(be.obs.cpu().numpy(), oobs, rtol=0, atol=OBS_ATOL, err_msg=f'step {k}')
        np.testing.assert_allclose(be.reward64.cpu().numpy(), orew, rtol=0, atol=STATE_ATOL, err_msg=f'step {k}')
    be.close()


def test_step_invariants_at_scale(B):
    """Size-independent properties at the full benchmark size (65,536 two-car
    envs, device-generated pool): all query modes agree bit for bit on every
    discrete output and to 1e-6 on observations over a free-running rollout;
    readings stay in [0, 1]; episode bookkeeping is consistent."""
    E = 65536
    outs = []
    for query in QUERY_MODES:
        be = B.RacingBackend(E, kind='multi', num_agents=2, num_sensors=11, query=query, seed=5)
        be.generate_tracks(seed=9, n_tracks=16)
        be.reset()
        g = torch.Generator(device='cuda').manual_seed(1)
        term_count = torch.zeros((), dtype=torch.int64, device='cuda')
        for k in range(40):
            be.actions.copy_(torch.rand(be.actions.shape, device='cuda', generator=g) * 2 - 1)
            be.actions[..., 1].abs_()
            be.step()
            term_count += be.done.sum()
        assert float(be.obs[..., :11].min()) >= 0.0 and float(be.obs[..., :11].max()) <= 1.0
        assert float(be.obs.abs().max()) <= 1.0
        assert int(be.ep_stats[2]) == int(term_count) > 1000
        st = be.get_state()
        outs.append((be.obs.clone(), be.reward64.clone(), st['car_i32'].copy(), st['car_f64'].copy(), int(term_count)))
        be.close()
    o0, r0, i0, f0, n0 = outs[0]
    for o1, r1, i1, f1, n1 in outs[1:]:
        assert n0 == n1
        np.testing.assert_array_equal(i0, i1)
        np.testing.assert_array_equal(f0, f1)           # float64 state is bit-identical between the query modes
        assert torch.equal(r0, r1)
        assert float((o0 - o1).abs().max()) <= 1e-6


def test_error_behaviour_through_the_abi(B):
    """Errors never throw across the C ABI: non-zero status + rk_last_error text,
    surfaced as RuntimeError by the Python layer."""
    be = B.RacingBackend(8, kind='multi', num_agents=2)
    with pytest.raises(RuntimeError, match='no tracks set'):
        be.step()
    with pytest.raises(RuntimeError, match='no tracks set'):
        be.reset()
    with pytest.raises(RuntimeError, match='control points'):
        be.set_tracks_from_control_points([np.zeros((2, 2))], [6.0])          # fewer than 3 control points
    with pytest.raises(RuntimeError, match='out of range'):
        cps, widths = _pool(2, seed=1)
        be.set_tracks_from_control_points(cps, widths, env_to_track=np.full(8, 5))
    cps, widths = _pool(2, seed=1)
    be.set_tracks_from_control_points(cps, widths)
    with pytest.raises(RuntimeError, match='out of range'):
        be.get_track(7)
    be.reset()
    be.step()                                                                 # and it still works afterwards
    be.close()
    with pytest.raises(RuntimeError, match='invalid config'):
        B.RacingBackend(8, kind='single', num_sensors=65)                     # RK_MAX_SENSORS = 64
    with pytest.raises(RuntimeError, match='invalid config'):
        B.RacingBackend(8, kind='multi', num_agents=9)                        # RK_MAX_AGENTS = 8
    # update-path entry points
    import ctypes as C
    from self_play_racing_b200 import _lib
    lib = _lib.load()
    io = _lib.RkPpoGradIO()
    io.struct_size = 8
    assert lib.rk_ppo_minibatch_grad(C.byref(io), None) != 0 and b'struct_size' in lib.rk_last_error(None)
    io.struct_size = C.sizeof(_lib.RkPpoGradIO)
    io.obs_dim, io.n = 21, 128                                                # RK_PPO_MAX_OBS_DIM = 20
    assert lib.rk_ppo_minibatch_grad(C.byref(io), None) != 0 and b'obs_dim' in lib.rk_last_error(None)
    # the native rollout: struct sizes, missing buffers, layout and self-play preconditions are refused with a message
    be2 = B.RacingBackend(8, kind='multi', num_agents=2, agent_major=False)
    be2.set_tracks_from_control_points(cps, widths)
    ro = _lib.RkRolloutIO()
    ro.struct_size = 4
    assert lib.rk_rollout(be2.h, C.byref(be2._io), C.byref(ro), None) != 0 and b'struct_size' in lib.rk_last_error(be2.h)
    ro.struct_size, ro.T = C.sizeof(_lib.RkRolloutIO), 4
    assert lib.rk_rollout(be2.h, C.byref(be2._io), C.byref(ro), None) != 0 and b'needs T > 0' in lib.rk_last_error(be2.h)
    be2.reset(); be2.step()                                                   # the handle keeps working
    assert lib.rk_set_seed(None, 1) != 0
    be2.close()
    ad = _lib.RkAdamIO()
    ad.struct_size = C.sizeof(_lib.RkAdamIO)
    assert lib.rk_ppo_adam_step(C.byref(ad), None) != 0 and b'invalid arguments' in lib.rk_last_error(None)
    out = torch.zeros(300, 2, device='cuda')
    with pytest.raises(RuntimeError, match='multiple of 256'):
        B.policy_act_pool(torch.zeros(2, 11080, device='cuda'), torch.zeros(3, dtype=torch.int32, device='cuda'), 100,
                          torch.zeros(300, 19, device='cuda'), out, seed=1, counter=1)
    assert B.random_permutation(0, 1, 1, device='cuda').numel() == 0          # empty input is not an error
    with pytest.raises(RuntimeError, match='invalid arguments'):
        lib_rc = lib.rk_random_permutation(1, 1, 5, None, None)
        _lib.check(lib_rc, None, 'rk_random_permutation')


def test_extreme_track_shapes(B):
    """Smallest (3 control points) and a long (60 control points, 1800 waypoints,
    3600 segments) track, one env each, against the oracle in both query modes."""
    th = np.linspace(0, 2 * np.pi, 60, endpoint=False)
    big = np.column_stack([(90 + 12 * np.sin(5 * th)) * np.cos(th), (70 + 9 * np.cos(3 * th)) * np.sin(th)])
    tri = np.array([[0.0, 0.0], [80.0, 0.0], [40.0, 70.0]])
    cps, widths = [tri, big], [7.0, 8.0]
    tracks = O.make_pool(cps, widths)
    rs = np.random.RandomState(0)
    for query in QUERY_MODES:
        orc = O.OracleVecEnv(tracks, [0, 1, 0, 1], kind='single', num_sensors=11)
        be = B.RacingBackend(4, kind='single', num_sensors=11, query=query)
        be.set_tracks_from_control_points(cps, widths, env_to_track=[0, 1, 0, 1])
        for i in range(2):
            np.testing.assert_allclose(be.get_track(i)['waypoints'], tracks[i].waypoints, rtol=0, atol=1e-10)
        oobs, _ = orc.reset()
        np.testing.assert_allclose(be.reset().cpu().numpy(), oobs, atol=OBS_ATOL)
        for k in range(150):
            a = rs.uniform([-0.4, 0.3], [0.4, 1.0], size=(4, 1, 2)).astype(np.float32)
            oobs, orew, ote, otr, _ = orc.step(a)
            be.actions.copy_(torch.from_numpy(a))
            be.step()
            np.testing.assert_array_equal(be.terminated.cpu().numpy().astype(bool), ote, err_msg=f'step {k}')
            np.testing.assert_allclose(be.obs.cpu().numpy(), oobs, rtol=0, atol=OBS_ATOL, err_msg=f'step {k}')
            np.testing.assert_allclose(be.reward64.cpu().numpy(), orew, rtol=0, atol=STATE_ATOL, err_msg=f'step {k}')
        be.close()


# ------------------------------------------------ injected states: one reward/termination branch each
@pytest.mark.parametrize('query', QUERY_MODES)
def test_injected_single_branches(B, golden, query):
    """Hand-built states (rk_set_state) recorded on the unmodified reference: finish + time bonus and
    its floor, finish on the truncation step, both wraps with and without checkpoints, checkpoint order
    and upper edges, crash on the truncation step, speed clamp (racing_env.py:112-162)."""
    g = golden('injected_single')
    S = len(g['names'])
    be = B.RacingBackend(S, kind='single', num_sensors=11, query=query)
    be.set_tracks_from_waypoints([g['waypoints']], [float(g['width'])])
    be.reset()
    car_f, car_i, env_i = FL.backend_state_single(g)
    be.set_state(car_f64=car_f, car_i32=car_i, env_i32=env_i, env_f64=np.zeros(S))
    for t in range(g['actions'].shape[0]):
        be.actions[:, 0].copy_(torch.from_numpy(g['actions'][t]))
        be.step()
        names = [f'{n} (step {t})' for n in g['names']]
        np.testing.assert_array_equal(be.terminated.cpu().numpy().astype(bool), g['terminated'][t], err_msg=str(names))
        np.testing.assert_array_equal(be.truncated.cpu().numpy().astype(bool), g['truncated'][t])
        np.testing.assert_allclose(be.obs[:, 0].cpu().numpy(), g['obs'][t], rtol=0, atol=OBS_ATOL)
        np.testing.assert_allclose(be.reward64[:, 0].cpu().numpy(), g['reward'][t], rtol=0, atol=STATE_ATOL)
        st, sti = _state(be)
        ended_before = (g['terminated'][t - 1] | g['truncated'][t - 1]) if t else np.zeros(S, bool)
        np.testing.assert_allclose(st[:, 0], g['state'][t], rtol=0, atol=STATE_ATOL)
        np.testing.assert_array_equal(sti[:, 0, 0], g['progress_idx'][t])
        fl = sti[:, 0, 2]
        np.testing.assert_array_equal((fl & FL.F_FINISHED) != 0, g['finished'][t])
        np.testing.assert_array_equal((fl & FL.F_CRASHED) != 0, g['crashed'][t])
        np.testing.assert_array_equal(np.stack([(fl & m) != 0 for m in (FL.F_CP25, FL.F_CP50, FL.F_CP75)], 1), g['checkpoints'][t])
        if t == 0:
            np.testing.assert_allclose(be.info_f64[:, 0, 3].cpu().numpy(), g['info_progress'][t], rtol=0, atol=1e-12)
            np.testing.assert_allclose(be.info_f64[:, 0, 4].cpu().numpy(), g['info_delta'][t], rtol=0, atol=1e-12)
        del ended_before
    be.close()


@pytest.mark.parametrize('query', QUERY_MODES)
def test_injected_multi_branches(B, golden, query):
    """Placement with exact ties (higher index wins), +250 on truncation and on termination, finished_step,
    finish against a crashed car, all-crashed termination, -160 only once, checkpoints collected by a
    crashed car, touching penalty (multi_racing_env.py:155-211,222-259)."""
    g = golden('injected_multi2')
    S = len(g['names'])
    be = B.RacingBackend(S, kind='multi', num_agents=2, num_sensors=11, query=query)
    be.set_tracks_from_waypoints([g['waypoints']], [float(g['width'])])
    be.reset(start_slot=torch.from_numpy(np.tile([0, 1], (S, 1)).astype(np.int32)).cuda())
    car_f, car_i, env_i = FL.backend_state_multi(g)
    be.set_state(car_f64=car_f, car_i32=car_i, env_i32=env_i, env_f64=np.zeros(S))
    for t in range(g['actions'].shape[0]):
        be.actions.copy_(torch.from_numpy(g['actions'][t]))
        be.step(start_slot=torch.from_numpy(g['start_order'][t].astype(np.int32)).cuda())
        np.testing.assert_array_equal(be.terminated.cpu().numpy().astype(bool), g['terminated'][t], err_msg=str(list(g['names'])))
        np.testing.assert_array_equal(be.truncated.cpu().numpy().astype(bool), g['truncated'][t])
        np.testing.assert_allclose(be.obs.cpu().numpy(), g['obs'][t], rtol=0, atol=OBS_ATOL)
        np.testing.assert_allclose(be.reward64.cpu().numpy(), g['reward'][t], rtol=0, atol=STATE_ATOL)
        np.testing.assert_array_equal(be.info_i32[..., 2].cpu().numpy(), g['placement'][t])
        st, sti = _state(be)
        np.testing.assert_allclose(st, g['state'][t], rtol=0, atol=STATE_ATOL)
        fl = sti[..., 2]
        np.testing.assert_array_equal(np.stack([(fl & FL.F_CRASHED) != 0, (fl & FL.F_FINISHED) != 0], 2), g['flags'][t])
        np.testing.assert_array_equal(np.stack([(fl & m) != 0 for m in (FL.F_CP25, FL.F_CP50, FL.F_CP75)], 2), g['checkpoints'][t])
        np.testing.assert_array_equal(sti[..., 3], g['finished_step'][t])
    be.close()
