"""CPU-only tests: the C-ABI library loads and exports every symbol the header
declares (no compute without a GPU), the host-side mirrors of the reference
(track generator, Agent, configs, spaces), and the PPO update math -- against the
reference's own ppo_update when /root/reference is present in the container."""
import ctypes
import os
import re
import sys
import types

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'racing_b200.h')


def _declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r'^RK_API [^;(]*?\b(rk_[a-z0-9_]+)\(', src, flags=re.M)))


def test_library_exports_every_declared_symbol():
    from self_play_racing_b200 import _lib
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f'{n} declared in racing_b200.h but not exported'
        assert n in _lib.SIGNATURES, f'{n} has no ctypes signature in _lib.py'
    assert set(_lib.SIGNATURES) == set(names)
    assert lib.rk_abi_version() == 1
    assert ctypes.sizeof(_lib.RkConfig) == 56 and ctypes.sizeof(_lib.RkStepIO) == 8 + 15 * 8 + 8


def test_no_cpu_fallback_without_a_device():
    """rk_create must refuse to run without a CUDA device (there is no CPU path)."""
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    from self_play_racing_b200 import _lib
    lib = _lib.load()
    cfg = _lib.RkConfig(struct_size=ctypes.sizeof(_lib.RkConfig), device=0, num_envs=4, num_agents=1,
                        num_sensors=11, env_kind=0, autoreset_mode=0, query_mode=1, max_episode_steps=0,
                        reserved0=0, speed_weight=8.0, seed=0)
    h = ctypes.c_void_p()
    assert lib.rk_create(ctypes.byref(cfg), ctypes.byref(h)) != 0
    assert b'no usable CUDA device' in lib.rk_last_error(None)
    bad = _lib.RkConfig(struct_size=12)
    assert lib.rk_create(ctypes.byref(bad), ctypes.byref(h)) != 0 and b'ABI mismatch' in lib.rk_last_error(None)
    from self_play_racing_b200.backend import RacingBackend
    with pytest.raises(RuntimeError, match='no CPU path'):
        RacingBackend(4)


def test_product_never_imports_the_oracle():
    for dp, _, files in os.walk(os.path.join(ROOT, 'self_play_racing_b200')):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh')):
                text = open(os.path.join(dp, f)).read()
                assert 'import oracle' not in text and 'from oracle' not in text, f


def test_track_generator_matches_reference_draw_for_draw(golden):
    """gen_tracks(16, seed=1) incl. the re-seeding collapse and train.py:30 widths."""
    from self_play_racing_b200.environment.track import gen_random_track, gen_tracks, resolve_track
    g = golden('tracks')
    np.random.seed(1)
    pool = gen_tracks(num_tracks=16, seed=1)
    widths = [np.random.randint(6, 10) for _ in range(16)]
    np.testing.assert_array_equal(np.concatenate(pool), g['train_pool'])
    np.testing.assert_array_equal(np.array(widths), g['train_widths'])
    assert len({p.tobytes() for p in pool}) == 4
    np.testing.assert_array_equal(gen_random_track(13, 62, 17, 0.45, 0.35, seed=9), g['rand_track'])
    cp, w, tid = resolve_track(None, widths, pool, 3)      # width list indexed by track id (track.py:66-67)
    assert w == float(widths[3]) and tid == 3 and cp is not None
    cp, w, _ = resolve_track()                              # fixed default polygon, width 6 (track.py:69-80)
    assert cp.shape == (10, 2) and w == 6.0


def test_spaces_and_env_construction_are_cheap_and_lazy():
    from self_play_racing_b200.environment import MultiRacingEnv, RacingEnv, SelfPlayWrapper
    e = RacingEnv(num_sensors=11)
    assert e.observation_space.shape == (15,) and e.action_space.shape == (2,) and e._be is None
    m = MultiRacingEnv(num_agents=2, num_sensors=11)
    assert m.observation_space['0'].shape == (19,) and set(m.action_space) == {'0', '1'}
    w = SelfPlayWrapper(m, 0)
    assert w.observation_space.shape == (19,) and w.opponent_idx == 1
    a = w.opponent_action_space.sample()
    assert a.dtype == np.float32 and -1 <= a[0] <= 1 and 0 <= a[1] <= 1


def test_agent_matches_reference_init_and_forward(golden):
    """Same seed => same parameters as the reference Agent; same forward outputs."""
    from self_play_racing_b200.agent.ppo import Agent
    from self_play_racing_b200.backend import PARAM_ORDER, flatten_agent
    from self_play_racing_b200.environment import MultiRacingEnv
    g = golden('agent')
    env = MultiRacingEnv(num_agents=2, num_sensors=11)
    torch.manual_seed(1)
    agent = Agent(env.observation_space['0'], env.action_space['0'])
    sd = agent.state_dict()
    assert list(sd.keys())[0] == 'log_std' and len(sd) == 13
    for k, v in sd.items():
        if k != 'log_std':
            np.testing.assert_array_equal(v.numpy(), g['sd.' + k])
    agent.log_std.data.fill_(-0.3)
    obs, act = torch.from_numpy(g['obs']), torch.from_numpy(g['act'])
    with torch.no_grad():
        _, logp, ent, val = agent.get_action_and_value(obs, act)
        a, _, _, _ = agent.get_action_and_value(obs)
    np.testing.assert_allclose(logp.numpy(), g['logp'], atol=1e-6)
    np.testing.assert_allclose(ent.numpy(), g['entropy'], atol=1e-6)
    np.testing.assert_allclose(val.numpy(), g['value'], atol=1e-6)
    assert a.abs().max() <= 1
    flat = flatten_agent(sd)
    assert flat.numel() == 11080 and set(PARAM_ORDER) == set(sd)
    np.testing.assert_array_equal(flat[:19 * 64].view(19, 64).numpy(), sd['actor_mu.0.weight'].t().numpy())


def test_configs_match_reference_values():
    from self_play_racing_b200 import configs
    b, s = configs.base_config(), configs.self_play_config()
    assert b['batch_size'] == 32768 and b['minibatch_size'] == 2048 and b['gae_lambda'] == 0.95 and b['ent_coef'] == 0.01
    assert s['gae_lambda'] == 0.97 and s['ent_coef'] == 0.02 and s['snapshot_freq'] == 15 and s['pool_size'] == 5
    assert s['total_timesteps'] == 3_000_000 and b['total_timesteps'] == 5_000_000
    big = configs.self_play_config(num_envs=65536, num_steps=64)
    assert big['batch_size'] == 65536 * 64 and big['minibatch_size'] == big['batch_size'] // 16


def test_compat_aliases_resolve_reference_imports():
    """The import lines of the reference's train.py / evaluate.py resolve to the mirrors."""
    import subprocess
    code = ('from self_play_racing_b200.compat import install_as_reference_modules as f; f();'
            'from environment.racing_env import RacingEnv; from environment.multi_racing_env import MultiRacingEnv;'
            'from environment.track import gen_tracks; from environment.wrappers import SelfPlayWrapper;'
            'from agent.ppo import PPO, Agent; from agent.self_play_ppo import SelfPlayPPO;'
            'from configs.base_config import hyperparams_config as b; from configs.self_play_config import hyperparams_config as s;'
            'import self_play_racing_b200.environment.racing_env as m; assert RacingEnv is m.RacingEnv; print(s()["pool_size"])')
    out = subprocess.run([sys.executable, '-c', code], cwd=ROOT, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert out.stdout.strip() == '5'


# ------------------------------------------------------------------ PPO update
def _fake_ppo(config, seed=3):
    """A PPO object with an Agent but no environment (update math is pure torch)."""
    from self_play_racing_b200.agent.ppo import PPO, Agent
    from self_play_racing_b200 import spaces
    ppo = PPO.__new__(PPO)
    ppo.config, ppo.device, ppo.world, ppo.rank, ppo._perm_gen = config, torch.device('cpu'), 1, 0, None
    torch.manual_seed(seed)
    obs_space = spaces.Box(low=np.float32(-1), high=np.float32(1), shape=(19,), dtype=np.float32)
    act_space = spaces.Box(low=np.array([-1.0, 0.0]), high=np.array([1.0, 1.0]), shape=(2,), dtype=np.float32)
    ppo.agent = Agent(obs_space, act_space)
    ppo.agent.log_std.data.fill_(-0.5)
    ppo.optimizer = torch.optim.Adam(ppo.agent.parameters(), lr=config['learning_rate'], eps=1e-5)
    return ppo


def _batch(n, seed=0):
    g = torch.Generator().manual_seed(seed)
    obs = torch.rand(n, 19, generator=g) * 2 - 1
    actions = torch.rand(n, 2, generator=g) * 2 - 1
    adv = torch.randn(n, generator=g)
    values = torch.randn(n, generator=g)
    returns = values + adv
    return obs, actions, adv, values, returns


def test_ppo_update_matches_reference_implementation():
    """Same data, same minibatch order => same parameters as reference PPO.ppo_update."""
    if not os.path.isdir('/root/reference/agent'):
        pytest.skip('reference checkout not present on this box')
    sys.path.insert(0, os.path.join(ROOT, 'tools'))
    import gym_stub
    gym_stub.install()
    sys.dont_write_bytecode = True
    sys.path.insert(0, '/root/reference')
    try:
        for m in [m for m in sys.modules if m == 'agent' or m.startswith('agent.')]:
            del sys.modules[m]
        from agent.ppo import PPO as RefPPO, Agent as RefAgent
    finally:
        sys.path.remove('/root/reference')
    from self_play_racing_b200 import configs
    cfg = configs.self_play_config(num_envs=8, num_steps=64, update_epochs=3, num_minibatches=4, kl_target=1e9)
    n = cfg['batch_size']
    obs, actions, adv, values, returns = _batch(n)
    mine = _fake_ppo(cfg)
    with torch.no_grad():
        _, logp, _, _ = mine.agent.get_action_and_value(obs, actions)
    logp = logp + 0.05 * torch.randn(n, generator=torch.Generator().manual_seed(9))
    ref = types.SimpleNamespace(config=cfg)
    ref.agent = RefAgent(types.SimpleNamespace(shape=(19,)), types.SimpleNamespace(shape=(2,)))
    ref.agent.load_state_dict(mine.agent.state_dict())
    ref.optimizer = torch.optim.Adam(ref.agent.parameters(), lr=cfg['learning_rate'], eps=1e-5)
    # the reference shuffles with np.random.shuffle per epoch; replay the same permutations
    np.random.seed(11)
    perms = []
    inds = np.arange(n)
    for _ in range(cfg['update_epochs']):
        np.random.shuffle(inds)
        perms.append(torch.from_numpy(inds.copy()))
    np.random.seed(11)
    T, E = cfg['num_steps'], cfg['num_envs']  # the reference flattens [T, E, ...] buffers itself
    RefPPO.ppo_update(ref, adv.view(T, E), returns.view(T, E), values.view(T, E), logp.view(T, E),
                      actions.view(T, E, 2), obs.view(T, E, 19))
    steps = mine.ppo_update(adv, returns, values, logp, actions, obs, permutation=lambda ep: perms[ep])
    assert steps == cfg['update_epochs'] * cfg['num_minibatches']
    for (k, a), b in zip(mine.agent.state_dict().items(), ref.agent.state_dict().values()):
        torch.testing.assert_close(a, b, rtol=1e-4, atol=2e-6, msg=k)


def test_ppo_update_kl_early_stop():
    from self_play_racing_b200 import configs
    cfg = configs.self_play_config(num_envs=4, num_steps=32, kl_target=0.015)
    n = cfg['batch_size']
    obs, actions, adv, values, returns = _batch(n, seed=2)
    ppo = _fake_ppo(cfg)
    with torch.no_grad():
        _, logp, _, _ = ppo.agent.get_action_and_value(obs, actions)
    before = [p.detach().clone() for p in ppo.agent.parameters()]
    assert ppo.ppo_update(adv, returns, values, logp + 1.0, actions, obs) == 0   # mean(old - new) = 1 > target
    for a, b in zip(before, ppo.agent.parameters()):
        assert torch.equal(a, b)
    assert ppo.ppo_update(adv, returns, values, logp, actions, obs) > 0


def test_chunked_linear_backward_equals_plain_linear():
    from self_play_racing_b200.agent.ppo import _Linear, _SplitKLinearFn
    torch.manual_seed(0)
    lin = _Linear(19, 64).double()
    ref = torch.nn.Linear(19, 64).double()
    ref.load_state_dict(lin.state_dict())
    n = 4 * _SplitKLinearFn.CHUNK + 37            # ragged tail exercises the remainder path
    x = torch.randn(n, 19, dtype=torch.float64, requires_grad=True)
    x2 = x.detach().clone().requires_grad_(True)
    g = torch.randn(n, 64, dtype=torch.float64)
    lin(x).backward(g)
    ref(x2).backward(g)
    torch.testing.assert_close(x.grad, x2.grad)
    torch.testing.assert_close(lin.weight.grad, ref.weight.grad, rtol=1e-12, atol=1e-10)
    torch.testing.assert_close(lin.bias.grad, ref.bias.grad, rtol=1e-12, atol=1e-10)


# ------------------------------------------------------------------ data parallel, gloo world size 2
def _dp_worker(rank, world, port, tmp, cfg, n):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(1)
    ppo = _fake_ppo(cfg)
    ppo.world, ppo.rank = world, rank
    obs, actions, adv, values, returns = _batch(n)
    with torch.no_grad():
        _, logp, _, _ = ppo.agent.get_action_and_value(obs, actions)
    logp = logp + 0.05 * torch.randn(n, generator=torch.Generator().manual_seed(9))
    half = n // world
    sl = slice(rank * half, (rank + 1) * half)
    # every rank permutes its own half with the same shared-seed permutation
    perms = [torch.randperm(half, generator=torch.Generator().manual_seed(100 + ep)) for ep in range(cfg['update_epochs'])]
    steps = ppo.ppo_update(adv[sl], returns[sl], values[sl], logp[sl], actions[sl], obs[sl], permutation=lambda ep: perms[ep])
    torch.save({'steps': steps, 'sd': ppo.agent.state_dict()}, os.path.join(tmp, f'rank{rank}.pt'))
    dist.destroy_process_group()


def test_data_parallel_update_equals_single_process(tmp_path):
    """Env sharding across 2 ranks + all-reduced gradients and global minibatch
    statistics == one process on the concatenated batch."""
    import torch.multiprocessing as mp
    from self_play_racing_b200 import configs
    cfg = configs.self_play_config(num_envs=8, num_steps=32, update_epochs=2, num_minibatches=4, kl_target=1e9)
    n, world = cfg['batch_size'], 2
    port = 29500 + os.getpid() % 2000
    mp.spawn(_dp_worker, args=(world, port, str(tmp_path), cfg, n), nprocs=world, join=True)
    r0 = torch.load(tmp_path / 'rank0.pt')
    r1 = torch.load(tmp_path / 'rank1.pt')
    for a, b in zip(r0['sd'].values(), r1['sd'].values()):
        assert torch.equal(a, b)                       # ranks stay in lock-step
    # single process: global minibatch k = union of the ranks' local minibatches k
    half = n // world
    perms = [torch.randperm(half, generator=torch.Generator().manual_seed(100 + ep)) for ep in range(cfg['update_epochs'])]
    mb = half // cfg['num_minibatches']

    def global_perm(ep):
        chunks = []
        for s in range(0, half, mb):
            chunks += [perms[ep][s:s + mb], perms[ep][s:s + mb] + half]
        return torch.cat(chunks)
    one = _fake_ppo(cfg)
    obs, actions, adv, values, returns = _batch(n)
    with torch.no_grad():
        _, logp, _, _ = one.agent.get_action_and_value(obs, actions)
    logp = logp + 0.05 * torch.randn(n, generator=torch.Generator().manual_seed(9))
    steps = one.ppo_update(adv, returns, values, logp, actions, obs, permutation=global_perm)
    assert steps == r0['steps'] == cfg['update_epochs'] * cfg['num_minibatches']
    for (k, a), b in zip(one.agent.state_dict().items(), r0['sd'].values()):
        torch.testing.assert_close(a, b, rtol=1e-4, atol=2e-6, msg=k)


def test_episode_stats_dict_is_lazy_and_complete():
    """infos["episode"] of the vector env: 'r' / 'l' as given, 't' derived on first access from the episode length
    and the step clock (an episode of length l ending at step s was reset during step s - l)."""
    from self_play_racing_b200.environment.vec_env import _EpisodeStats
    times = np.arange(4096) * 0.5
    mask = np.array([False, True, False, True])
    length = np.array([0, 3, 0, 7], np.int32)
    d = _EpisodeStats(np.array([0.0, 2.5, 0.0, -1.0]), length, mask, 10, times)
    assert set(d) == {'r', 'l', 't'} and dict.__getitem__(d, 't') is None          # nothing computed yet
    np.testing.assert_allclose(d['t'], [0.0, 1.5, 0.0, 3.5])
    np.testing.assert_allclose(dict(d.items())['t'], [0.0, 1.5, 0.0, 3.5])
    assert d.get('t') is d['t'] and d.get('x', 5) == 5 and d.copy()['l'] is length
