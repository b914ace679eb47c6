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
