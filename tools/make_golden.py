"""Generate tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, imported read-only with tools/gym_stub.py standing in for the
absent gymnasium package).  Run in the build container only:

    PYTHONDONTWRITEBYTECODE=1 python tools/make_golden.py

The reference ships no golden vectors; these files are
the pin for oracle/racing_oracle.py and, through it, for the CUDA path.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import gym_stub  # noqa: E402

gym_stub.install()
sys.path.insert(0, '/root/reference')
sys.dont_write_bytecode = True

import torch  # noqa: E402
from environment.racing_env import RacingEnv  # noqa: E402
from environment.multi_racing_env import MultiRacingEnv  # noqa: E402
from environment.track import Track, gen_random_track, gen_tracks  # noqa: E402
from environment.wrappers import SelfPlayWrapper  # noqa: E402
from agent.ppo import Agent, PPO  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')
os.makedirs(OUT, exist_ok=True)


def save(name, **arrays):
    path = os.path.join(OUT, name + '.npz')
    np.savez_compressed(path, **arrays)
    print(f'{name}: {os.path.getsize(path) / 1024:.0f} KiB')


def procedural_pool():
    """Distinct procedural tracks (gen_tracks collapses)."""
    params = [(10, 55, 12, 0.3, 0.5, 11), (12, 60, 15, 0.4, 0.5, 3), (14, 70, 20, 0.6, 0.3, 5),
              (11, 65, 14, 0.25, 0.65, 8)]
    return [gen_random_track(*p[:5], seed=p[5]) for p in params], [6, 8, 9, 7]


def car_state(cars):
    return np.array([[c.x, c.y, c.angle, c.vx, c.vy] for c in cars], dtype=np.float64)


# ---------------------------------------------------------------- single env
def record_single(name, steps, pool=None, track_id=None, width=None, seed=0):
    rng = np.random.default_rng(seed)
    env = RacingEnv(num_sensors=11, track_pool=pool, track_id=track_id, track_width=width)
    obs0, _ = env.reset()
    A = np.zeros((steps, 2), np.float32)
    OBS = np.zeros((steps, 15), np.float32)
    REW = np.zeros(steps)
    TERM = np.zeros(steps, bool)
    TRUNC = np.zeros(steps, bool)
    ST = np.zeros((steps, 5))
    PIDX = np.zeros(steps, np.int64)
    need_reset = False
    for k in range(steps):
        a = rng.uniform([-1, 0], [1, 1]).astype(np.float32)
        if k % 7 == 3:  # exercise the clip paths too
            a = (a * 1.5).astype(np.float32)
        A[k] = a
        if need_reset:  # NEXT_STEP auto-reset, as SyncVectorEnv does it
            obs, _ = env.reset()
            r, te, tr = 0.0, False, False
        else:
            obs, r, te, tr, info = env.step(a)
        need_reset = te or tr
        OBS[k], REW[k], TERM[k], TRUNC[k] = obs, r, te, tr
        ST[k] = car_state([env.car])[0]
        PIDX[k] = round(env.car.progress * len(env.track.waypoints))
    save(name, control_points=np.asarray(env.track.control_points, np.float64),
         width=np.float64(env.track.track_width), waypoints=env.track.waypoints,
         obs0=obs0, actions=A, obs=OBS, reward=REW, terminated=TERM, truncated=TRUNC,
         state=ST, progress_idx=PIDX)


# ----------------------------------------------------------------- multi env
def ref_reset(env):
    """Reset and recover the start slots the global np.random shuffle produced."""
    st = np.random.get_state()
    out = env.reset()
    rs = np.random.RandomState()
    rs.set_state(st)
    order = list(range(env.num_agents))
    rs.shuffle(order)
    return out, np.array([order.index(i) for i in range(env.num_agents)], np.int64)


def record_multi(name, steps, num_agents, pool=None, track_id=None, width=None, seed=0):
    np.random.seed(100 + seed)
    rng = np.random.default_rng(seed)
    n = num_agents
    env = MultiRacingEnv(num_agents=n, num_sensors=11, track_pool=pool, track_id=track_id, track_width=width)
    (obs, _), so0 = ref_reset(env)
    D = 11 + 4 + 4 * (n - 1)
    obs0 = np.stack([obs[str(i)] for i in range(n)])
    A = np.zeros((steps, n, 2), np.float32)
    OBS = np.zeros((steps, n, D), np.float32)
    REW = np.zeros((steps, n))
    TERM = np.zeros(steps, bool)
    TRUNC = np.zeros(steps, bool)
    ST = np.zeros((steps, n, 5))
    SO = np.zeros((steps, n), np.int64)
    PL = np.zeros((steps, n), np.int64)
    FLAGS = np.zeros((steps, n, 2), bool)
    need_reset = False
    for k in range(steps):
        a = rng.uniform(-1, 1, size=(n, 2)).astype(np.float32)
        if k % 5 != 0:
            a[:, 1] = np.abs(a[:, 1])  # mostly accelerate, so that cars travel
        if k % 11 == 4:
            a = (a * 1.4).astype(np.float32)
        A[k] = a
        if need_reset:
            (obs, _), SO[k] = ref_reset(env)
            rew = {str(i): 0.0 for i in range(n)}
            te = tr = False
            info = None
        else:
            obs, rew, dones, tr, info = env.step({str(i): a[i] for i in range(n)})
            te = dones['0']
        need_reset = te or tr
        OBS[k] = np.stack([obs[str(i)] for i in range(n)])
        REW[k] = [rew[str(i)] for i in range(n)]
        TERM[k], TRUNC[k] = te, tr
        ST[k] = car_state(env.cars)
        FLAGS[k] = [[c.crashed, c.finished] for c in env.cars]
        if need_reset:
            PL[k] = [info[str(i)]['placement'] for i in range(n)]
    save(name, control_points=np.asarray(env.track.control_points, np.float64),
         width=np.float64(env.track.track_width), obs0=obs0, start_order0=so0,
         actions=A, obs=OBS, reward=REW, terminated=TERM, truncated=TRUNC, state=ST,
         start_order=SO, placement=PL, flags=FLAGS)


# ------------------------------------------------- vector env + episode stats
def record_vector(name, steps, num_envs=4):
    import gymnasium as gym
    pool, widths = procedural_pool()

    def thunk(i):
        def f():
            env = RacingEnv(num_sensors=11, track_pool=pool, track_id=i % len(pool), track_width=widths[i % len(pool)])
            env = gym.wrappers.RecordEpisodeStatistics(env)
            env.reset(seed=i)
            return env
        return f
    envs = gym.vector.SyncVectorEnv([thunk(i) for i in range(num_envs)])
    rng = np.random.default_rng(5)
    obs0, _ = envs.reset()
    A = np.zeros((steps, num_envs, 2), np.float32)
    OBS = np.zeros((steps, num_envs, 15), np.float32)
    REW = np.zeros((steps, num_envs))
    TERM = np.zeros((steps, num_envs), bool)
    TRUNC = np.zeros((steps, num_envs), bool)
    EPM = np.zeros((steps, num_envs), bool)
    EPR = np.zeros((steps, num_envs))
    EPL = np.zeros((steps, num_envs), np.int64)
    for k in range(steps):
        a = rng.uniform([-1, 0], [1, 1], size=(num_envs, 2)).astype(np.float32)
        A[k] = a
        OBS[k], REW[k], TERM[k], TRUNC[k], infos = envs.step(a)
        if 'episode' in infos:
            EPM[k] = infos['_episode']
            EPR[k] = infos['episode']['r']
            EPL[k] = infos['episode']['l']
    save(name, pool_sizes=np.array([len(p) for p in pool]), pool=np.concatenate(pool),
         widths=np.array(widths, np.float64), obs0=obs0, actions=A, obs=OBS, reward=REW,
         terminated=TERM, truncated=TRUNC, ep_mask=EPM, ep_r=EPR, ep_l=EPL)


# ---------------------------------------------------------- track-level pins
def record_tracks():
    pool, widths = procedural_pool()
    arrs = {}
    for i, (cp, w) in enumerate(zip(pool + [None], widths + [None])):
        t = Track(control_points=cp, track_width=w)
        arrs[f'cp{i}'] = np.asarray(t.control_points, np.float64)
        arrs[f'width{i}'] = np.float64(t.track_width)
        arrs[f'wp{i}'] = t.waypoints
        arrs[f'nrm{i}'] = t.normals
        arrs[f'starts{i}'] = t.segment_cache['starts']
        arrs[f'v2{i}'] = t.segment_cache['v2']
        arrs[f'mtd{i}'] = np.float64(t.max_track_distance)
        arrs[f'start{i}'] = np.array(t.get_start_pos(), np.float64)
        # query pins: closest waypoint, wall test, ray distances on random probes
        rs = np.random.RandomState(40 + i)
        k = rs.randint(0, len(t.waypoints), 64)
        pts = t.waypoints[k] + rs.uniform(-1.2, 1.2, (64, 1)) * t.track_width * t.normals[k]
        ang = rs.uniform(0, 2 * np.pi, 64)
        arrs[f'q_pts{i}'] = pts
        arrs[f'q_ang{i}'] = ang
        arrs[f'q_idx{i}'] = np.array([t.closest_waypoint_idx(p[0], p[1]) for p in pts])
        arrs[f'q_ray{i}'] = np.array([t.raycast(p, a, 50.0) for p, a in zip(pts, ang)])
    arrs['n'] = np.int64(len(pool) + 1)
    # the degenerate pool of train.py:29 and its widths (train.py:30)
    np.random.seed(1)
    tp = gen_tracks(num_tracks=16, seed=1)
    tw = [np.random.randint(6, 10) for _ in range(16)]
    arrs['train_pool_sizes'] = np.array([len(p) for p in tp])
    arrs['train_pool'] = np.concatenate(tp)
    arrs['train_widths'] = np.array(tw)
    arrs['rand_track'] = gen_random_track(13, 62, 17, 0.45, 0.35, seed=9)
    save('tracks', **arrs)


# ------------------------------------------------ scripted drivers: the finish line
# Random actions crash within 50-90 steps, so the trajectories above never reach a
# checkpoint.  These recordings drive the UNMODIFIED reference with a pure-pursuit
# script on track.waypoints so that every branch of the reward state machine past
# "crash" runs: the three checkpoints, finish + time bonus, forward wrap without a
# finish, backward wrap, the 3000-step truncation, finished_step-driven placement,
# +250 on termination and on truncation (racing_env.py:112-162,
# multi_racing_env.py:155-211,247-259).  Only the actions are stored with the
# outputs: replays are action-driven, the script itself is not needed by any test.
def _wrap(a):
    return (a + np.pi) % (2 * np.pi) - np.pi


def pursuit_action(track, car, plan, k, multi):
    """One float32 action of the scripted driver.  plan: dict(vmax, look, off, gain,
    schedule=[(from_step, direction, vmax)], crash_at)."""
    direction, vmax = 1, plan['vmax']
    for k0, d, v in plan.get('schedule', ()):
        if k >= k0:
            direction, vmax = d, v
    n = len(track.waypoints)
    idx = track.closest_waypoint_idx(car.x, car.y)
    tgt_i = (idx + direction * plan['look']) % n
    tgt = track.waypoints[tgt_i] + track.normals[tgt_i] * plan['off'] * direction
    dth = _wrap(np.arctan2(tgt[1] - car.y, tgt[0] - car.x) - car.angle)
    steer = float(np.clip(plan['gain'] * dth, -1.0, 1.0))
    speed = float(np.hypot(car.vx, car.vy))
    # slow down while the heading error is large (U-turns of the backward-wrap plans)
    vlim = vmax if abs(dth) < 0.6 else min(vmax, 2.5)
    thr = 1.0 if speed < vlim else 0.0
    if plan.get('crash_at') is not None and k >= plan['crash_at']:
        steer, thr = 1.0, 1.0
    if multi:
        thr = 2.0 * thr - 1.0           # multi_racing_env.py:217 maps [-1, 1] -> [0, 1]
    return np.array([steer, thr], np.float32)


def lap_plans(rng, kinds):
    """One script per episode.  kinds: 'lap' (clean lap at a random pace), 'wrap' (backward over the
    start line, forward again -- both wraps without a finish -- then a crash), 'crawl' (too slow to
    finish: the 3000-step truncation), 'crash' (crash part of the way round), 'bump' (multi: identical
    scripts on the centre line, so the cars collide)."""
    plans = []
    for kind in kinds:
        p = dict(kind=kind, vmax=float(rng.uniform(10.0, 16.0)), look=int(rng.integers(8, 14)), off=0.0, gain=2.0,
                 schedule=[], crash_at=None)
        if kind == 'wrap':
            p['schedule'] = [(0, 1, 3.0), (12, -1, 4.0), (150, 1, 6.0)]
            p['crash_at'] = 330
        elif kind == 'crawl':
            p['vmax'] = float(rng.uniform(0.9, 1.3))
        elif kind == 'crash':
            p['crash_at'] = int(rng.integers(150, 400))
        plans.append(p)
    return plans


def record_single_laps(name, kinds, pool=None, track_id=None, width=None, seed=0, max_steps=16000):
    rng = np.random.default_rng(seed)
    env = RacingEnv(num_sensors=11, track_pool=pool, track_id=track_id, track_width=width)
    obs0, _ = env.reset()
    n_episodes = len(kinds)
    plans = lap_plans(rng, kinds)
    rec = {k: [] for k in ('actions', 'obs', 'reward', 'terminated', 'truncated', 'state', 'progress_idx',
                           'finished', 'crashed', 'checkpoints')}
    ep, k_ep, need_reset = 0, 0, False
    ev = dict(finish=0, cp=0, trunc=0, crash=0, bwd=0, fwd_nofinish=0)
    while ep < n_episodes and len(rec['actions']) < max_steps:
        a = pursuit_action(env.track, env.car, plans[ep], k_ep, multi=False)
        if need_reset:
            obs, _ = env.reset()
            r, te, tr = 0.0, False, False
            ep, k_ep = ep + 1, 0
        else:
            lp = env.last_progress
            obs, r, te, tr, info = env.step(a)
            k_ep += 1
            ev['finish'] += env.car.finished
            ev['trunc'] += tr
            ev['crash'] += env.car.crashed
            ev['bwd'] += lp < 0.1 and env.car.progress > 0.9
            ev['fwd_nofinish'] += lp > 0.9 and env.car.progress < 0.1 and not env.car.finished
        need_reset = te or tr
        rec['actions'].append(a); rec['obs'].append(obs); rec['reward'].append(r)
        rec['terminated'].append(te); rec['truncated'].append(tr)
        rec['state'].append(car_state([env.car])[0])
        rec['progress_idx'].append(round(env.car.progress * len(env.track.waypoints)))
        rec['finished'].append(env.car.finished); rec['crashed'].append(env.car.crashed)
        rec['checkpoints'].append([env.checkpoints[0.25], env.checkpoints[0.50], env.checkpoints[0.75]])
    ev['cp'] = int(np.diff(np.array(rec['checkpoints'], np.int8), axis=0).clip(0).sum())
    print(name, 'events:', ev, 'steps:', len(rec['actions']))
    save(name, control_points=np.asarray(env.track.control_points, np.float64),
         width=np.float64(env.track.track_width), waypoints=env.track.waypoints, obs0=obs0,
         actions=np.array(rec['actions'], np.float32), obs=np.array(rec['obs'], np.float32),
         reward=np.array(rec['reward']), terminated=np.array(rec['terminated'], bool),
         truncated=np.array(rec['truncated'], bool), state=np.array(rec['state']),
         progress_idx=np.array(rec['progress_idx'], np.int64), finished=np.array(rec['finished'], bool),
         crashed=np.array(rec['crashed'], bool), checkpoints=np.array(rec['checkpoints'], bool),
         events=np.array([ev[k] for k in ('finish', 'cp', 'trunc', 'crash', 'bwd', 'fwd_nofinish')]))


def record_multi_laps(name, kinds, num_agents=2, pool=None, track_id=None, width=None, seed=0, max_steps=20000):
    np.random.seed(300 + seed)
    rng = np.random.default_rng(seed)
    n = num_agents
    env = MultiRacingEnv(num_agents=n, num_sensors=11, track_pool=pool, track_id=track_id, track_width=width)
    (obs, _), so0 = ref_reset(env)
    obs0 = np.stack([obs[str(i)] for i in range(n)])
    # one script per car and episode; lateral offsets keep the cars apart most of the time
    n_episodes = len(kinds)
    plans = [lap_plans(rng, kinds) for _ in range(n)]
    offs = np.linspace(-1.6, 1.6, n)
    for ep, kind in enumerate(kinds):
        for i in range(n):
            pl = plans[i][ep]
            pl['off'] = float(offs[i])
            if kind == 'crawl':            # all crawl (truncation with a leader: +250 on truncation)
                pl['vmax'] = 0.7 + 0.25 * i
            elif kind == 'wrap' and i > 0:  # only car 0 does the wrap manoeuvre, the others race on slowly
                pl['schedule'], pl['crash_at'], pl['vmax'] = [], None, 5.0
            elif kind == 'crash':           # one car crashes early, the others finish
                pl['crash_at'] = int(rng.integers(60, 200)) if i == ep % n else None
            elif kind == 'bump':            # same script for every car: they bump into each other (SAT, -5, x0.92)
                pl['off'], pl['vmax'], pl['look'] = 0.0, plans[0][ep]['vmax'], plans[0][ep]['look']
    rec = {k: [] for k in ('actions', 'obs', 'reward', 'terminated', 'truncated', 'state', 'start_order',
                           'placement', 'flags', 'finished_step', 'checkpoints', 'progress_idx')}
    ep, k_ep, need_reset = 0, 0, False
    ev = dict(finish=0, trunc=0, first_crash=0, bwd=0, fwd_nofinish=0, win_bonus=0)
    while ep < n_episodes and len(rec['actions']) < max_steps:
        a = np.stack([pursuit_action(env.track, env.cars[i], plans[i][ep], k_ep, multi=True) for i in range(n)])
        so = np.zeros(n, np.int64)
        pl = np.zeros(n, np.int64)
        if need_reset:
            (obs, _), so = ref_reset(env)
            rew = {str(i): 0.0 for i in range(n)}
            te = tr = False
            ep, k_ep = ep + 1, 0
        else:
            lps = [env.agents_data[i]['last_progress'] for i in range(n)]
            had = [env.agents_data[i]['has_crashed'] for i in range(n)]
            obs, rew, dones, tr, info = env.step({str(i): a[i] for i in range(n)})
            te = dones['0']
            k_ep += 1
            for i in range(n):
                c = env.cars[i]
                ev['first_crash'] += c.crashed and not had[i]
                ev['bwd'] += lps[i] < 0.1 and c.progress > 0.9
                ev['fwd_nofinish'] += lps[i] > 0.9 and c.progress < 0.1 and not c.finished
            ev['finish'] += sum(c.finished for c in env.cars)
            ev['trunc'] += tr
            if te or tr:
                pl = np.array([info[str(i)]['placement'] for i in range(n)])
                ev['win_bonus'] += 1
        need_reset = te or tr
        rec['actions'].append(a); rec['obs'].append(np.stack([obs[str(i)] for i in range(n)]))
        rec['reward'].append([rew[str(i)] for i in range(n)])
        rec['terminated'].append(te); rec['truncated'].append(tr)
        rec['state'].append(car_state(env.cars)); rec['start_order'].append(so); rec['placement'].append(pl)
        rec['flags'].append([[c.crashed, c.finished] for c in env.cars])
        rec['finished_step'].append([env.agents_data[i]['finished_step'] or 0 for i in range(n)])
        rec['checkpoints'].append([[env.agents_data[i]['checkpoints'][q] for q in (0.25, 0.50, 0.75)] for i in range(n)])
        rec['progress_idx'].append([round(c.progress * len(env.track.waypoints)) for c in env.cars])
    print(name, 'events:', ev, 'steps:', len(rec['actions']))
    save(name, control_points=np.asarray(env.track.control_points, np.float64),
         width=np.float64(env.track.track_width), obs0=obs0, start_order0=so0,
         actions=np.array(rec['actions'], np.float32), obs=np.array(rec['obs'], np.float32),
         reward=np.array(rec['reward']), terminated=np.array(rec['terminated'], bool),
         truncated=np.array(rec['truncated'], bool), state=np.array(rec['state']),
         start_order=np.array(rec['start_order'], np.int64), placement=np.array(rec['placement'], np.int64),
         flags=np.array(rec['flags'], bool), finished_step=np.array(rec['finished_step'], np.int64),
         checkpoints=np.array(rec['checkpoints'], bool), progress_idx=np.array(rec['progress_idx'], np.int64),
         events=np.array([ev[k] for k in ('finish', 'trunc', 'first_crash', 'bwd', 'fwd_nofinish', 'win_bonus')]))


def record_laps():
    pool, widths = procedural_pool()
    L = 'lap'
    record_single_laps('single_laps_default', [L, L, 'wrap', L, 'crawl', 'crash', L, L, L, L, L, L, L], seed=21)
    record_single_laps('single_laps_proc1', [L, 'wrap', 'crawl', L, 'crash', L], pool=pool, track_id=1,
                       width=widths[1], seed=22)
    record_multi_laps('multi2_laps_default', [L, L, 'wrap', L, 'crawl', 'crash', 'bump', L, L, L, L, L, 'crash'], 2, seed=23)
    record_multi_laps('multi2_laps_proc2', [L, 'crawl', 'crash', 'bump', 'wrap'], 2, pool=pool, track_id=2,
                      width=widths[2], seed=24)
    record_multi_laps('multi3_laps_proc1', [L, 'wrap', 'crash', 'bump', L], 3, pool=pool, track_id=1,
                      width=widths[1], seed=25)


# ------------------------------------------------------ injected states, one branch each
# Hand-built states pushed into the UNMODIFIED reference envs, then stepped twice (the step
# under test and the NEXT_STEP reset that follows an ended episode).  Scenario s becomes
# environment s of a batch in the oracle / CUDA tests (rk_set_state).  Each state is given as
# waypoint index + lateral offset + heading relative to the track tangent + forward speed.
def _pose(track, k, lateral, rel_heading, speed):
    n = len(track.waypoints)
    k %= n
    tang = track.waypoints[(k + 1) % n] - track.waypoints[k]
    ang = (np.arctan2(tang[1], tang[0]) + rel_heading) % (2 * np.pi)
    x, y = track.waypoints[k] + track.normals[k] * lateral
    return float(x), float(y), float(ang), float(speed * np.cos(ang)), float(speed * np.sin(ang))


def _single_scenarios(n):
    full, none_ = (True, True, True), (False, False, False)
    #     name                     wp idx            lat  heading speed  last idx         checkpoints            steps action
    return [
        ('finish_time_bonus',      2,                0.5, 0.0,    12.0,  n - 3,           full,                  700,  (0.1, 1.0)),
        ('finish_bonus_floor',     1,               -1.0, 0.0,    8.0,   n - 2,           full,                  2500, (0.0, 0.5)),   # max(0, 200 - 250) = 0
        ('finish_and_truncate',    2,                0.0, 0.0,    10.0,  n - 2,           full,                  2999, (0.0, 1.0)),
        ('fwd_wrap_no_cp',         1,                0.0, 0.0,    9.0,   n - 2,           none_,                 40,   (0.0, 1.0)),
        ('fwd_wrap_two_cp',        1,                0.0, 0.0,    9.0,   n - 2,           (True, True, False),   900,  (0.0, 1.0)),
        ('bwd_wrap',               n - 2,            0.0, np.pi,  9.0,   2,               none_,                 55,   (0.0, 1.0)),
        ('bwd_wrap_all_cp',        n - 2,            0.0, np.pi,  9.0,   2,               full,                  1500, (0.0, 1.0)),
        ('cp25_hit',               int(0.26 * n),    0.0, 0.0,    11.0,  int(0.24 * n),   none_,                 300,  (0.0, 1.0)),
        ('cp50_needs_cp25',        int(0.51 * n),    0.0, 0.0,    11.0,  int(0.49 * n),   none_,                 500,  (0.0, 1.0)),
        ('cp50_hit',               int(0.51 * n),    0.0, 0.0,    11.0,  int(0.49 * n),   (True, False, False),  500,  (0.0, 1.0)),
        ('cp75_needs_cp50',        int(0.76 * n),    0.0, 0.0,    11.0,  int(0.74 * n),   (True, False, False),  700,  (0.0, 1.0)),
        ('cp75_hit',               int(0.76 * n),    0.0, 0.0,    11.0,  int(0.74 * n),   (True, True, False),   700,  (0.0, 1.0)),
        ('cp25_upper_edge',        int(0.35 * n) + 1, 0.0, 0.0,   11.0,  int(0.34 * n),   none_,                 300,  (0.0, 1.0)),   # progress >= 0.35: no bonus
        ('truncate_plain',         int(0.4 * n),     0.0, 0.0,    5.0,   int(0.4 * n),    (True, False, False),  2999, (0.0, 0.0)),
        ('crash_and_truncate',     int(0.4 * n),     5.2, 1.2,    14.0,  int(0.4 * n),    (True, False, False),  2999, (1.0, 1.0)),
        ('speed_clamp',            int(0.6 * n),     0.0, 0.0,    30.0,  int(0.6 * n) - 1, (True, True, False),  200,  (0.0, 1.0)),
        ('standing_still',         int(0.1 * n),     0.0, 0.0,    0.0,   int(0.1 * n),    none_,                 10,   (0.0, 0.0)),   # delta = 0: no speed bonus
        ('reversing',              int(0.1 * n),     0.0, np.pi,  6.0,   int(0.1 * n) + 1, none_,                10,   (0.0, 1.0)),   # delta < 0
    ]


def record_injected_single(name='injected_single'):
    env0 = RacingEnv(num_sensors=11)
    n = len(env0.track.waypoints)
    sc = _single_scenarios(n)
    S, K = len(sc), 2
    init_f = np.zeros((S, 5)); init_i = np.zeros((S, 6), np.int64)  # pidx, lpidx, cp25, cp50, cp75, steps
    A = np.zeros((K, S, 2), np.float32); OBS = np.zeros((K, S, 15), np.float32); REW = np.zeros((K, S))
    TERM = np.zeros((K, S), bool); TRUNC = np.zeros((K, S), bool); ST = np.zeros((K, S, 5))
    PIDX = np.zeros((K, S), np.int64); FIN = np.zeros((K, S), bool); CR = np.zeros((K, S), bool)
    CP = np.zeros((K, S, 3), bool); PROG = np.zeros((K, S)); DELTA = np.zeros((K, S))
    for s_, (nm, k, lat, hd, v, lk, cps, steps, act) in enumerate(sc):
        env = RacingEnv(num_sensors=11)
        env.reset()
        c = env.car
        c.x, c.y, c.angle, c.vx, c.vy = _pose(env.track, k, lat, hd, v)
        pidx = env.track.closest_waypoint_idx(c.x, c.y)
        c.progress = pidx / n
        env.last_progress = (lk % n) / n
        env.checkpoints = {0.25: cps[0], 0.50: cps[1], 0.75: cps[2]}
        env.steps = steps
        init_f[s_] = car_state([c])[0]
        init_i[s_] = [pidx, lk % n, *cps, steps]
        need_reset = False
        for t in range(K):
            A[t, s_] = act
            if need_reset:
                obs, _ = env.reset()
                r, te, tr, info = 0.0, False, False, {'progress': 0.0, 'progress_delta': 0.0}
            else:
                obs, r, te, tr, info = env.step(np.array(act, np.float32))
            need_reset = te or tr
            OBS[t, s_], REW[t, s_], TERM[t, s_], TRUNC[t, s_] = obs, r, te, tr
            ST[t, s_] = car_state([c])[0]
            PIDX[t, s_] = round(c.progress * n)
            FIN[t, s_], CR[t, s_] = c.finished, c.crashed
            CP[t, s_] = [env.checkpoints[q] for q in (0.25, 0.50, 0.75)]
            PROG[t, s_], DELTA[t, s_] = info['progress'], info['progress_delta']
        print(f'  single {nm:22s} r={REW[0, s_]:9.3f} term={TERM[0, s_]} trunc={TRUNC[0, s_]} fin={FIN[0, s_]} cr={CR[0, s_]} cp={CP[0, s_].astype(int)}')
    save(name, names=np.array([x[0] for x in sc]), waypoints=env0.track.waypoints, width=np.float64(env0.track.track_width),
         init_f=init_f, init_i=init_i, actions=A, obs=OBS, reward=REW, terminated=TERM, truncated=TRUNC, state=ST,
         progress_idx=PIDX, finished=FIN, crashed=CR, checkpoints=CP, info_progress=PROG, info_delta=DELTA)


def _multi_scenarios(n):
    full, none_ = (True, True, True), (False, False, False)
    fwd, stop = (0.0, 1.0), (0.0, -1.0)
    q = lambda f: int(f * n)
    # per car: (wp idx, lateral, heading, speed, last idx, checkpoints, has_crashed, crashed, finished_step, action)
    return [
        ('tie_on_truncation',       2999, [(q(.4), -1.75, 0, 0.0, q(.4), none_, False, False, 0, stop),
                                           (q(.4), 1.75, 0, 0.0, q(.4), none_, False, False, 0, stop)]),      # equal scores: the higher index wins
        ('leader_on_truncation',    2999, [(q(.45), -1.75, 0, 3.0, q(.45), (True, False, False), False, False, 0, fwd),
                                           (q(.40), 1.75, 0, 3.0, q(.40), (True, False, False), False, False, 0, fwd)]),
        ('crashed_leader_truncation', 2999, [(q(.6), 0.0, 0, 0.0, q(.6), (True, True, False), True, True, 0, fwd),
                                             (q(.3), 0.0, 0, 4.0, q(.3), (True, False, False), False, False, 0, fwd)]),
        ('finish_car0',             800,  [(2, -1.75, 0, 12.0, n - 3, full, False, False, 0, fwd),
                                           (q(.9), 1.75, 0, 12.0, q(.9) - 1, full, False, False, 0, fwd)]),
        ('finish_car1_bonus_floor', 2900, [(q(.5), -1.75, 0, 5.0, q(.5), (True, False, False), False, False, 0, fwd),
                                           (1, 1.75, 0, 9.0, n - 2, full, False, False, 0, fwd)]),
        ('both_finish_same_step',   1200, [(1, -1.75, 0, 9.0, n - 2, full, False, False, 0, fwd),
                                           (2, 1.75, 0, 11.0, n - 2, full, False, False, 0, fwd)]),             # both finished, same finished_step: progress breaks the tie
        ('both_finish_exact_tie',   1200, [(1, -1.75, 0, 9.0, n - 2, full, False, False, 0, fwd),
                                           (1, 1.75, 0, 9.0, n - 2, full, False, False, 0, fwd)]),
        ('finish_and_truncate',     2999, [(2, -1.75, 0, 12.0, n - 3, full, False, False, 0, fwd),
                                           (q(.7), 1.75, 0, 12.0, q(.7) - 1, (True, True, False), False, False, 0, fwd)]),
        ('finish_vs_crashed',       600,  [(q(.8), 0.0, 0, 0.0, q(.8), full, True, True, 0, fwd),
                                           (2, 1.75, 0, 12.0, n - 3, full, False, False, 0, fwd)]),
        ('last_car_crashes',        400,  [(q(.5), 0.0, 0, 0.0, q(.5), (True, False, False), True, True, 0, fwd),
                                           (q(.3), 5.4, 1.2, 14.0, q(.3), (True, False, False), False, False, 0, (1.0, 1.0))]),   # all crashed: -160 once, winner by progress
        ('both_crash_same_progress', 90,  [(q(.2), 5.3, 1.2, 14.0, q(.2), none_, False, False, 0, (1.0, 1.0)),
                                           (q(.2), -5.3, -1.2, 14.0, q(.2), none_, False, False, 0, (-1.0, 1.0))]),
        ('fwd_wrap_no_cp',          60,   [(1, -1.75, 0, 9.0, n - 2, none_, False, False, 0, fwd),
                                           (q(.1), 1.75, 0, 9.0, q(.1) - 1, none_, False, False, 0, fwd)]),
        ('bwd_wrap',                60,   [(n - 2, -1.75, np.pi, 9.0, 2, none_, False, False, 0, fwd),
                                           (q(.1), 1.75, 0, 9.0, q(.1) - 1, none_, False, False, 0, fwd)]),
        ('checkpoints_25_50',       500,  [(q(.26), -1.75, 0, 11.0, q(.24), none_, False, False, 0, fwd),
                                           (q(.51), 1.75, 0, 11.0, q(.49), (True, False, False), False, False, 0, fwd)]),
        ('checkpoints_75_order',    700,  [(q(.76), -1.75, 0, 11.0, q(.74), (True, True, False), False, False, 0, fwd),
                                           (q(.76), 1.75, 0, 11.0, q(.74), (True, False, False), False, False, 0, fwd)]),
        ('touching',                300,  [(q(.3), 0.0, 0, 10.0, q(.3) - 1, (True, False, False), False, False, 0, fwd),
                                           (q(.3), 1.2, 0.2, 10.0, q(.3) - 1, (True, False, False), False, False, 0, fwd)]),
        ('crashed_car_keeps_scoring', 300, [(q(.3), 0.0, 0, 6.0, q(.3), (True, False, False), True, True, 0, fwd),
                                            (q(.35), 1.75, 0, 10.0, q(.35) - 1, (True, False, False), False, False, 0, fwd)]),
    ]


def record_injected_multi(name='injected_multi2'):
    env0 = MultiRacingEnv(num_agents=2, num_sensors=11)
    n = len(env0.track.waypoints)
    sc = _multi_scenarios(n)
    S, K, NA, D = len(sc), 2, 2, 19
    init_f = np.zeros((S, NA, 5)); init_i = np.zeros((S, NA, 8), np.int64)  # pidx, lpidx, cp x3, has_crashed, crashed, finished_step
    init_steps = np.zeros(S, np.int64)
    A = np.zeros((K, S, NA, 2), np.float32); OBS = np.zeros((K, S, NA, D), np.float32); REW = np.zeros((K, S, NA))
    TERM = np.zeros((K, S), bool); TRUNC = np.zeros((K, S), bool); ST = np.zeros((K, S, NA, 5))
    SO = np.zeros((K, S, NA), np.int64); PL = np.zeros((K, S, NA), np.int64); FLAGS = np.zeros((K, S, NA, 2), bool)
    FS = np.zeros((K, S, NA), np.int64); CP = np.zeros((K, S, NA, 3), bool)
    np.random.seed(77)
    for s_, (nm, steps, cars) in enumerate(sc):
        env = MultiRacingEnv(num_agents=2, num_sensors=11)
        env.reset()
        for i, (k, lat, hd, v, lk, cps, has_cr, cr, fs, act) in enumerate(cars):
            c = env.cars[i]
            c.x, c.y, c.angle, c.vx, c.vy = _pose(env.track, k, lat, hd, v)
            pidx = env.track.closest_waypoint_idx(c.x, c.y)
            c.progress, c.crashed = pidx / n, cr
            env.agents_data[i].update(last_progress=(lk % n) / n, checkpoints={0.25: cps[0], 0.50: cps[1], 0.75: cps[2]},
                                      has_crashed=has_cr, finished_step=fs or None)
            init_f[s_, i] = car_state([c])[0]
            init_i[s_, i] = [pidx, lk % n, *cps, has_cr, cr, fs]
            A[:, s_, i] = act
        env.steps = steps
        init_steps[s_] = steps
        need_reset = False
        for t in range(K):
            if need_reset:
                (obs, _), SO[t, s_] = ref_reset(env)
                rew, te, tr, info = {str(i): 0.0 for i in range(NA)}, False, False, None
            else:
                obs, rew, dones, tr, info = env.step({str(i): A[t, s_, i] for i in range(NA)})
                te = dones['0']
            need_reset = te or tr
            OBS[t, s_] = np.stack([obs[str(i)] for i in range(NA)])
            REW[t, s_] = [rew[str(i)] for i in range(NA)]
            TERM[t, s_], TRUNC[t, s_] = te, tr
            ST[t, s_] = car_state(env.cars)
            FLAGS[t, s_] = [[c.crashed, c.finished] for c in env.cars]
            FS[t, s_] = [env.agents_data[i]['finished_step'] or 0 for i in range(NA)]
            CP[t, s_] = [[env.agents_data[i]['checkpoints'][q] for q in (0.25, 0.50, 0.75)] for i in range(NA)]
            if need_reset:
                PL[t, s_] = [info[str(i)]['placement'] for i in range(NA)]
        print(f'  multi {nm:26s} r={np.round(REW[0, s_], 3)} term={TERM[0, s_]} trunc={TRUNC[0, s_]} place={PL[0, s_]} '
              f'flags={FLAGS[0, s_].astype(int).tolist()}')
    save(name, names=np.array([x[0] for x in sc]), waypoints=env0.track.waypoints, width=np.float64(env0.track.track_width),
         init_f=init_f, init_i=init_i, init_steps=init_steps, actions=A, obs=OBS, reward=REW, terminated=TERM,
         truncated=TRUNC, state=ST, start_order=SO, placement=PL, flags=FLAGS, finished_step=FS, checkpoints=CP)


# ------------------------------------------- rollout buffers through the reference's own PPO loop
# The UNMODIFIED SelfPlayPPO.collect_rollout (agent/ppo.py:97-132) over SyncVectorEnv(RecordEpisodeStatistics(
# SelfPlayWrapper(MultiRacingEnv))) (agent/self_play_ppo.py:19-29), two consecutive rollouts with update_opponent()
# in between (the envs are rebuilt but next_obs is carried over: SURVEY quirk 10).  Observer subclasses record what
# the device replay has to inject: the learner's and the opponent's sampled actions and the start slots of every
# reset; they do not change any arithmetic.
def record_rollout_buffers(name='rollout_selfplay16', T=96, E=16):
    import gymnasium as gym
    from agent.self_play_ppo import SelfPlayPPO
    from configs.self_play_config import hyperparams_config
    pool, widths = procedural_pool()
    log = {'slots': [], 'opp': []}

    class RecMulti(MultiRacingEnv):            # records the slot permutation of every reset
        def reset(self, seed=None, options=None):
            st = np.random.get_state()
            out = super().reset(seed=seed, options=options)
            rs = np.random.RandomState(); rs.set_state(st)
            order = list(range(self.num_agents)); rs.shuffle(order)
            self.last_slots = np.array([order.index(i) for i in range(self.num_agents)], np.int64)
            self.n_resets = getattr(self, 'n_resets', 0) + 1
            return out

    class RecOpponent:                          # a frozen snapshot whose sampled actions are written down
        def __init__(self, agent):
            self.agent, self.last = agent, None
        def get_action_and_value(self, obs):
            out = self.agent.get_action_and_value(obs)
            self.last = out[0].squeeze(0).numpy().copy()
            return out

    cfg = hyperparams_config()
    cfg.update(num_envs=E, num_steps=T, cuda=False)
    cfg['batch_size'] = T * E
    cfg['minibatch_size'] = cfg['batch_size'] // cfg['num_minibatches']
    env_fn = lambda i: RecMulti(num_agents=2, num_sensors=11, track_pool=pool, track_id=i % len(pool), track_width=widths)
    np.random.seed(5)
    trainer = SelfPlayPPO(env_fn, cfg, device='cpu')
    for w in trainer.envs.envs:                 # the CPU arm: opponent tensors stay on the host
        w.env.device = torch.device('cpu')
    trainer.agent.log_std.data.fill_(-0.3)
    torch.manual_seed(11)
    c = cfg
    z = lambda *sh: torch.zeros(*sh)
    obs, actions = z(T, E, 19), z(T, E, 2)
    logprobs, dones, rewards, values = (z(T, E) for _ in range(4))
    init_obs, _ = trainer.envs.reset()
    next_obs, next_done = torch.from_numpy(init_obs), torch.zeros(E, dtype=torch.bool)
    out = {}
    snap = trainer.snapshot_agent()
    with torch.no_grad():
        for p_ in snap.parameters():
            p_.add_(0.3 * torch.randn_like(p_))    # an opponent that differs from the learner
    for it in range(2):
        trainer.opponent_pool = [] if it == 0 else [RecOpponent(snap)]   # rollout 0: pool-empty random opponent
        trainer.update_opponent()                  # rebuilds the envs, does NOT reset next_obs (quirk 10)
        envs = trainer.envs.envs                   # RecordEpisodeStatistics(SelfPlayWrapper(RecMulti))
        for w in envs:
            w.env.device = torch.device('cpu')
        slots0 = np.stack([w.env.env.last_slots for w in envs])      # the rebuild's own reset
        # observer of the vector env's step: which envs reset, with which slots; which opponent actions were used
        opp_act = np.zeros((T, E, 2), np.float32)
        slot_log = np.zeros((T, E, 2), np.int64)
        reset_log = np.zeros((T, E), bool)
        step_no = [0]
        orig_samplers = []
        for i, w in enumerate(envs):
            sp = w.env.opponent_action_space
            orig = sp.sample
            def sample(orig=orig, i=i):
                a = orig()
                opp_act[step_no[0], i] = a
                return a
            sp.sample = sample
        vec_step = trainer.envs.step
        def step(acts):
            t = step_no[0]
            before = [w.env.env.n_resets for w in envs]
            res = vec_step(acts)
            for i, w in enumerate(envs):
                if w.env.env.n_resets != before[i]:
                    reset_log[t, i] = True
                    slot_log[t, i] = w.env.env.last_slots
                elif trainer.curr_opponent is not None:
                    opp_act[t, i] = trainer.curr_opponent.last if False else opp_act[t, i]
            step_no[0] += 1
            return res
        trainer.envs.step = step
        if trainer.curr_opponent is not None:       # per-env opponent actions: wrap the wrapper's policy call
            rec = trainer.curr_opponent
            for i, w in enumerate(envs):
                class PerEnv:
                    def __init__(self, i): self.i = i
                    def get_action_and_value(self, o):
                        res = rec.agent.get_action_and_value(o)
                        opp_act[step_no[0], self.i] = res[0].squeeze(0).numpy()
                        return res
                w.env.set_opponent(PerEnv(i))
        res = trainer.collect_rollout(obs, actions, logprobs, dones, rewards, values, next_obs, next_done)
        obs, actions, logprobs, dones, rewards, values, next_obs, next_done, ep_info = res
        out.update({f'obs{it}': obs.numpy().copy(), f'actions{it}': actions.numpy().copy(),
                    f'logprobs{it}': logprobs.numpy().copy(), f'dones{it}': dones.numpy().copy(),
                    f'rewards{it}': rewards.numpy().copy(), f'values{it}': values.numpy().copy(),
                    f'next_obs{it}': next_obs.numpy().copy(), f'next_done{it}': next_done.numpy().copy(),
                    f'opp_actions{it}': opp_act.copy(), f'reset{it}': reset_log.copy(), f'slots{it}': slot_log.copy(),
                    f'slots_init{it}': slots0,
                    f'ep_r{it}': np.array([e['reward'] for e in ep_info]), f'ep_l{it}': np.array([e['length'] for e in ep_info])})
        print(f'  rollout {it}: {len(ep_info)} episodes, dones {int(dones.sum())}, resets {int(reset_log.sum())}')
    sd = {k: v.numpy() for k, v in trainer.agent.state_dict().items()}
    save(name, pool_sizes=np.array([len(p) for p in pool]), pool=np.concatenate(pool), widths=np.array(widths, np.float64),
         **out, **{'sd.' + k: v for k, v in sd.items()})


# ------------------------------------------------ evaluate.py's protocol on the reference envs
def wired_agent(obs_space, act_space, multi):
    """An Agent (reference class, reference state_dict keys) whose weights are set by hand to a ray-balancing
    controller: steer towards the side whose rays see farther, constant throttle.  Deterministic with log_std = -100
    (Normal(mu, 4e-44).sample() == mu), so that the reference loop and the device batch act identically."""
    ag = Agent(obs_space, act_space)
    R = 11
    with torch.no_grad():
        for p_ in ag.actor_mu.parameters():
            p_.zero_()
        w = torch.linspace(-1.0, 1.0, R)            # ray i points to relative angle ~ w_i: positive = left
        ag.actor_mu[0].weight[0, :R] = 0.05 * w      # h0 ~ 0.05 * sum_i w_i ray_i  (tanh in its linear range)
        ag.actor_mu[2].weight[0, 0] = 1.0
        ag.actor_mu[4].weight[0, 0] = 60.0           # steer = tanh(60 h)
        ag.actor_mu[4].bias[1] = float(np.arctanh(0.35 if not multi else 2 * 0.35 - 1))   # throttle 0.35 after either mapping
        ag.log_std.fill_(-100.0)
    return ag


def record_eval(name='eval_protocol'):
    sys.path.insert(0, '/root/reference')
    import types as _t
    for mod in ('matplotlib', 'matplotlib.pyplot'):       # utils/metrics.py imports pyplot at module level; never called here
        sys.modules.setdefault(mod, _t.ModuleType(mod))
    from utils.metrics import eval_single_agent, eval_multi_agent
    pool, widths = procedural_pool()
    pool, runs_w = pool[:3], [7, 9]
    out = {}
    dev = torch.device('cpu')
    env = RacingEnv(num_sensors=11)
    ag1 = wired_agent(env.observation_space, env.action_space, multi=False)
    rows = []
    for t in range(len(pool)):
        for r, w in enumerate(runs_w):      # evaluate.py:21-31: width indexed by RUN
            m = eval_single_agent(RacingEnv(num_sensors=11, track_pool=pool, track_id=t, track_width=w), ag1, dev)
            rows.append([m['total_reward'], m['steps'], m['progress'], m['finished'], m['crashed'], m['speed'], m['total_distance']])
            print('  single', t, r, {k: (round(v, 3) if isinstance(v, float) else v) for k, v in m.items()})
    out['single'] = np.array(rows, np.float64)
    menv = MultiRacingEnv(num_agents=2, num_sensors=11)
    ag2 = wired_agent(menv.observation_space['0'], menv.action_space['0'], multi=True)
    rows, slots = [], []
    np.random.seed(9)
    for t in range(2):
        for r, w in enumerate(runs_w):
            e = MultiRacingEnv(num_agents=2, num_sensors=11, track_pool=pool, track_id=t, track_width=w)
            st = np.random.get_state()
            m = eval_multi_agent(e, ag2, dev)
            rs = np.random.RandomState(); rs.set_state(st)
            order = [0, 1]; rs.shuffle(order)
            slots.append([order.index(0), order.index(1)])
            rows.append([m['total_reward'], m['steps'], m['progress'], m['finished'], m['crashed'], m['speed'], m['total_distance'],
                         m['placement'] or 0])
            print('  multi', t, r, {k: (round(v, 3) if isinstance(v, float) else v) for k, v in m.items()})
    out['multi'] = np.array(rows, np.float64)
    out['multi_slots'] = np.array(slots, np.int64)
    save(name, pool_sizes=np.array([len(p) for p in pool]), pool=np.concatenate(pool), run_widths=np.array(runs_w, np.float64),
         **out, **{'sd1.' + k: v.numpy() for k, v in ag1.state_dict().items()},
         **{'sd2.' + k: v.numpy() for k, v in ag2.state_dict().items()})


# ------------------------------------------------------- GAE and Agent pins
def record_gae_agent():
    rs = np.random.RandomState(0)
    T, E = 48, 6
    rewards = torch.tensor(rs.normal(0, 3, (T, E)), dtype=torch.float32)
    values = torch.tensor(rs.normal(0, 2, (T, E)), dtype=torch.float32)
    dones = torch.tensor(rs.uniform(size=(T, E)) < 0.08, dtype=torch.float32)
    next_value = torch.tensor(rs.normal(0, 2, E), dtype=torch.float32)
    next_done = torch.tensor(rs.uniform(size=E) < 0.3)
    out = {}
    for lam, tag in ((0.97, 'sp'), (0.95, 'single')):
        me = types.SimpleNamespace(config={'num_steps': T, 'gamma': 0.99, 'gae_lambda': lam}, device=torch.device('cpu'))
        adv, ret = PPO.compute_advantages(me, rewards, dones, values, next_value, next_done)
        out[f'adv_{tag}'] = adv.numpy()
        out[f'ret_{tag}'] = ret.numpy()
    save('gae', rewards=rewards.numpy(), values=values.numpy(), dones=dones.numpy(),
         next_value=next_value.numpy(), next_done=next_done.numpy(), **out)

    torch.manual_seed(1)
    env = MultiRacingEnv(num_agents=2, num_sensors=11)
    agent = Agent(env.observation_space['0'], env.action_space['0'])
    agent.log_std.data.fill_(-0.3)
    obs = torch.tensor(rs.uniform(-1, 1, (32, 19)), dtype=torch.float32)
    act = torch.tensor(rs.uniform(-1, 1, (32, 2)), dtype=torch.float32)
    with torch.no_grad():
        mu = agent.actor_mu(obs)
        _, logp, ent, val = agent.get_action_and_value(obs, act)
    sd = {k: v.numpy() for k, v in agent.state_dict().items()}
    save('agent', obs=obs.numpy(), act=act.numpy(), mu=mu.numpy(), logp=logp.numpy(),
         entropy=ent.numpy(), value=val.numpy(), **{'sd.' + k: v for k, v in sd.items()})


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'laps':      # only the scripted-driver recordings
        record_laps()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'rollout':
        record_rollout_buffers()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'eval':
        record_eval()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'injected':  # only the injected-state recordings
        record_injected_single()
        record_injected_multi()
        sys.exit(0)
    pool, widths = procedural_pool()
    record_single('single_default_10k', 10000)                       # BASELINE config 1
    for i in range(len(pool)):
        record_single(f'single_proc{i}', 1500, pool=pool, track_id=i, width=widths[i], seed=10 + i)
    record_multi('multi2_default', 2500, 2, seed=1)
    record_multi('multi2_proc1', 2500, 2, pool=pool, track_id=1, width=widths[1], seed=2)
    record_multi('multi2_proc2', 1500, 2, pool=pool, track_id=2, width=widths[2], seed=3)
    record_multi('multi3_proc2', 1000, 3, pool=pool, track_id=2, width=widths[2], seed=4)
    record_vector('vector_single4', 1200)
    record_tracks()
    record_gae_agent()
    record_laps()
    record_injected_single()
    record_injected_multi()
    record_rollout_buffers()
    record_eval()
