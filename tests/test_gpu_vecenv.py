"""GPU tests of the reference-facing host layer: the Gymnasium faces
(RacingEnv / MultiRacingEnv / SelfPlayWrapper / BatchedRacingVecEnv) and the
device-resident PPO loop."""
import os

import numpy as np
import pytest

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu

from oracle import racing_oracle as O  # noqa: E402


@pytest.fixture(scope='module')
def pkg():
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    import self_play_racing_b200.environment as env
    import self_play_racing_b200.agent as agent
    import self_play_racing_b200.configs as configs
    return env, agent, configs


def test_single_env_object_matches_reference_golden(pkg, golden):
    """RacingEnv(num_sensors=11) used the way evaluate.py uses it (reset/step/info)."""
    env_mod, _, _ = pkg
    g = golden('single_default_10k')
    env = env_mod.RacingEnv(num_sensors=11)
    assert env.observation_space.shape == (15,) and env.action_space.shape == (2,)
    obs, info = env.reset()
    np.testing.assert_allclose(obs, g['obs0'], atol=1e-6)
    assert info['progress'] == 0.0 and not info['crashed']
    for k in range(120):
        obs, r, te, tr, info = env.step(g['actions'][k])
        np.testing.assert_allclose(obs, g['obs'][k], atol=1e-6)
        assert abs(r - g['reward'][k]) < 1e-9 and te == g['terminated'][k] and tr == g['truncated'][k]
        assert set(info) >= {'position', 'speed', 'progress', 'crashed', 'finished', 'reward', 'progress_delta'}
        np.testing.assert_allclose(info['position'], g['state'][k, :2], atol=1e-9)
        if te:
            break
    assert te and info['crashed']
    trk = env.track
    np.testing.assert_allclose(trk.waypoints, g['waypoints'], atol=1e-11)
    assert trk.left_boundary.shape == trk.right_boundary.shape == trk.waypoints.shape
    np.testing.assert_allclose(env.car.get_corners().mean(0), [env.car.x, env.car.y], atol=1e-9)
    env.close()


def test_multi_env_object_and_selfplay_wrapper(pkg, golden):
    env_mod, agent_mod, _ = pkg
    g = golden('multi2_default')
    env = env_mod.MultiRacingEnv(num_agents=2, num_sensors=11)
    assert env.observation_space['0'].shape == (19,)
    np.random.seed(107)  # tools/make_golden.py seeds 100 + seed before the first reset
    np.random.seed(101)
    obs, infos = env.reset()
    np.testing.assert_allclose(np.stack([obs['0'], obs['1']]), g['obs0'], atol=1e-6)
    for k in range(60):
        obs, rew, dones, trunc, infos = env.step({'0': g['actions'][k, 0], '1': g['actions'][k, 1]})
        np.testing.assert_allclose(np.stack([obs['0'], obs['1']]), g['obs'][k], atol=1e-6)
        np.testing.assert_allclose([rew['0'], rew['1']], g['reward'][k], atol=1e-9)
        assert dones['__all__'] == (g['terminated'][k] or g['truncated'][k]) and dones['0'] == g['terminated'][k]
        if dones['__all__']:
            assert [infos['0']['placement'], infos['1']['placement']] == list(g['placement'][k])
            break
    env.close()
    # wrapper: learner view of car 0, random opponent in Box([-1,0],[1,1]), then a frozen policy
    w = env_mod.SelfPlayWrapper(env_mod.MultiRacingEnv(num_agents=2, num_sensors=11), 0)
    assert w.observation_space.shape == (19,)
    o, info = w.reset()
    assert o.shape == (19,)
    o, r, done, trunc, info = w.step(np.array([0.1, 0.5], np.float32))
    assert o.shape == (19,) and isinstance(r, float) and isinstance(done, bool)
    torch.manual_seed(0)
    pol = agent_mod.Agent(w.observation_space, w.action_space).cuda()
    w.set_opponent(pol)
    o, r, done, trunc, info = w.step(np.array([0.0, 1.0], np.float32))
    assert np.isfinite(o).all()
    w.close()


def test_vec_env_gym_face_matches_reference_vector_golden(pkg, golden):
    """BatchedRacingVecEnv built from env_fn thunks == SyncVectorEnv(RecordEpisodeStatistics(RacingEnv))."""
    env_mod, _, _ = pkg
    g = golden('vector_single4')
    cps = np.split(g['pool'], np.cumsum(g['pool_sizes'])[:-1])
    widths = [int(w) for w in g['widths']]
    fns = [(lambda i=i: env_mod.RacingEnv(num_sensors=11, track_pool=cps, track_id=i, track_width=widths[i]))
           for i in range(4)]
    # host_chunks: rk_step_host (one C call per step, chunked over internal streams); 0 = torch-level path,
    # optionally with the torch-level chunked pipeline
    for query, chunks, host_chunks in (('exact', 1, 1), ('culled', 1, 0), ('culled', 3, 0), ('culled', 1, 3)):
        vec = env_mod.BatchedRacingVecEnv(fns, query=query, pipeline_chunks=chunks)
        vec.host_chunks = host_chunks
        vec._host_io.n_chunks = max(host_chunks, 1)
        assert vec.single_observation_space.shape == (15,) and vec.single_action_space.shape == (2,)
        obs, infos = vec.reset()
        assert obs.dtype == np.float32 and infos == {}
        np.testing.assert_allclose(obs, g['obs0'], atol=1e-6)
        for k in range(len(g['actions'])):
            obs, rew, term, trunc, infos = vec.step(g['actions'][k])
            assert rew.dtype == np.float64 and term.dtype == np.bool_
            np.testing.assert_allclose(obs, g['obs'][k], atol=1e-6)
            np.testing.assert_allclose(rew, g['reward'][k], atol=1e-9)
            np.testing.assert_array_equal(term, g['terminated'][k])
            np.testing.assert_array_equal(trunc, g['truncated'][k])
            if g['ep_mask'][k].any():
                m = g['ep_mask'][k]
                np.testing.assert_array_equal(infos['_episode'], m)
                np.testing.assert_allclose(infos['episode']['r'][m], g['ep_r'][k][m], atol=1e-9)
                np.testing.assert_array_equal(infos['episode']['l'][m], g['ep_l'][k][m])
            else:
                assert 'episode' not in infos
        vec.close()


def test_selfplay_vec_env_vs_oracle(pkg):
    """SelfPlayWrapper semantics batched: learner = car 0, opponent = fused MLP
    inference; the oracle is fed the actions the device actually used."""
    env_mod, agent_mod, _ = pkg
    rs = np.random.RandomState(4)
    cps = [O.gen_random_track(12, 60, 15, 0.4, 0.5, rng=rs), O.gen_random_track(10, 55, 12, 0.3, 0.4, rng=rs)]
    widths = [8, 7]
    E = 96
    fns = [(lambda i=i: env_mod.SelfPlayWrapper(env_mod.MultiRacingEnv(
        num_agents=2, num_sensors=11, track_pool=cps, track_id=i % 2, track_width=widths), 0)) for i in range(E)]
    vec = env_mod.BatchedRacingVecEnv(fns, seed=3, query='culled')
    assert vec.selfplay and vec.be.num_tracks == 2
    tracks = O.make_pool(cps, widths)
    orc = O.OracleVecEnv(tracks, np.arange(E) % 2, kind='multi', num_agents=2, num_sensors=11, seed=0)
    so = orc._draw_start_order(E)
    vec.be.reset(start_slot=torch.from_numpy(so.astype(np.int32)).cuda())
    oobs, _ = orc.reset(start_order=so)
    torch.manual_seed(5)
    pol = agent_mod.Agent(vec.single_observation_space, vec.single_action_space)
    pol.log_std.data.fill_(-0.3)
    ended = 0
    for k in range(150):
        if k == 40:
            vec.set_opponent(pol)  # random opponent before, frozen snapshot after
        a0 = rs.uniform(-1, 1, size=(E, 2)).astype(np.float32)
        a0[:, 1] = np.abs(a0[:, 1])
        so = orc._draw_start_order(E)
        obs, rew, term, trunc, infos = vec.step(a0, start_slot=so)
        used = vec.be.actions.cpu().numpy()              # [A, E, 2]: what the kernel consumed
        np.testing.assert_array_equal(used[0], a0)
        if k < 40:
            assert (used[1, :, 1] >= 0).all() and (np.abs(used[1]) <= 1).all()
        oobs, orew, ote, otr, _ = orc.step(np.transpose(used, (1, 0, 2)), start_order=so)
        v_obs, v_rew, v_done, v_trunc = O.OracleVecEnv.selfplay_view(oobs, orew, ote, otr)
        np.testing.assert_allclose(obs, v_obs, atol=1e-6, err_msg=f'step {k}')
        np.testing.assert_allclose(rew, v_rew, atol=1e-9)
        np.testing.assert_array_equal(term, v_done)
        np.testing.assert_array_equal(trunc, v_trunc)
        ended += int(v_done.sum())
    assert ended > 10
    vec.close()


def test_selfplay_ppo_runs_on_device(pkg, tmp_path):
    """Two updates of the full self-play loop (rollout + GAE + update + snapshot
    + checkpoint) at a tiny size: finite losses, parameters move, checkpoint
    round-trips with the reference's keys."""
    env_mod, agent_mod, configs = pkg
    np.random.seed(1)
    pool = env_mod.gen_tracks(num_tracks=8, seed=1)
    widths = [int(np.random.randint(6, 10)) for _ in range(8)]
    cfg = configs.self_play_config(num_envs=256, num_steps=32, total_timesteps=256 * 32 * 3, snapshot_freq=1,
                                   pool_size=2, update_epochs=2, num_minibatches=4)

    def env_fn(i):
        return env_mod.MultiRacingEnv(num_agents=2, num_sensors=11, track_pool=pool, track_id=i % 8, track_width=widths)
    trainer = agent_mod.SelfPlayPPO(env_fn, cfg, device='cuda', checkpoint_dir=str(tmp_path))
    before = torch.cat([p.detach().flatten().clone() for p in trainer.agent.parameters()])
    logs = []
    info = trainer.train(log=logs.append)
    after = torch.cat([p.detach().flatten() for p in trainer.agent.parameters()])
    assert torch.isfinite(after).all() and not torch.equal(before, after)
    assert len(trainer.opponent_pool) == 2 and len(logs) == 3
    assert set(info) == {'steps', 'rewards', 'opponent_pool_size'}
    path = trainer.save_checkpoint(2, 3 * cfg['batch_size'], info)
    ck = torch.load(path, weights_only=False)
    assert set(ck) == {'update', 'global_step', 'agent_state_dict', 'optimizer_state_dict', 'opponent_pool',
                       'config', 'training_info'}
    assert len(ck['agent_state_dict']) == 13
    t2 = agent_mod.SelfPlayPPO(env_fn, cfg, device='cuda', checkpoint_dir=str(tmp_path))
    upd, gs, _ = t2.load_checkpoint(path)
    assert upd == 2 and len(t2.opponent_pool) == 2
    for a, b in zip(t2.agent.state_dict().values(), trainer.agent.state_dict().values()):
        assert torch.equal(a, b)
    t2.envs.close()


def test_reference_style_checkpoint_resumes_on_device(pkg, tmp_path):
    """A checkpoint as the reference writes it (self_play_ppo.py:155-166: plain torch Adam --
    python-float lr, non-capturable, host `step`) loads into the CUDA trainer, the
    device-resident update keeps working on the loaded optimizer state, and the state
    written back has the reference's layout."""
    env_mod, agent_mod, configs = pkg
    cfg = configs.self_play_config(num_envs=256, num_steps=16, total_timesteps=256 * 16 * 2, update_epochs=2,
                                   num_minibatches=2, snapshot_freq=1, pool_size=2)
    vec = env_mod.BatchedRacingVecEnv.synthetic('multi', 256, n_tracks=4, num_agents=2, selfplay=True, seed=0)
    tr = agent_mod.SelfPlayPPO(vec, cfg, device='cuda')
    # reference-style optimizer state: CPU agent, plain Adam, three steps
    ref_agent = agent_mod.Agent(vec.single_observation_space, vec.single_action_space)
    ref_opt = torch.optim.Adam(ref_agent.parameters(), lr=2.5e-4, eps=1e-5)
    for _ in range(3):
        ref_opt.zero_grad()
        _, lp, _, v = ref_agent.get_action_and_value(torch.rand(32, 19) * 2 - 1)
        (lp.mean() + v.mean()).backward()
        ref_opt.step()
    path = tmp_path / 'ref_style.pth'
    torch.save({'update': 7, 'global_step': 7 * 256 * 16, 'agent_state_dict': ref_agent.state_dict(),
                'optimizer_state_dict': ref_opt.state_dict(), 'opponent_pool': [ref_agent.state_dict()],
                'config': cfg, 'training_info': {'steps': [], 'rewards': [], 'opponent_pool_size': []}}, path)
    upd, gstep, _ = tr.load_checkpoint(str(path))
    assert (upd, gstep) == (7, 7 * 256 * 16) and len(tr.opponent_pool) == 1
    p0 = next(tr.agent.parameters())
    st = tr.optimizer.state[p0]
    assert st['step'].is_cuda and float(st['step']) == 3.0 and st['exp_avg'].is_cuda
    assert isinstance(tr.optimizer.param_groups[0]['lr'], torch.Tensor) and tr.optimizer.param_groups[0]['lr'].is_cuda
    buf = tr.alloc_buffers()
    buf['obs'][0].copy_(tr._reset_all())
    tr.update_opponent()
    tr._anneal(7, 100)
    tr.collect_rollout(buf)
    before = p0.detach().clone()
    steps = tr._learn_from(buf)
    assert steps == 4 and tr._graphed.fused_mlp and tr._graphed.adam is not None
    assert float(tr.optimizer.state[p0]['step']) == 7.0            # the kernel advanced torch's own step tensors
    assert not torch.equal(before, p0)
    sd = tr.optimizer.state_dict()
    assert set(sd['state'][0]) >= {'step', 'exp_avg', 'exp_avg_sq'} and len(sd['state']) == 12
    vec.close()


def test_single_ppo_learns_something(pkg):
    """A short single-agent PPO run: mean episode return improves over the
    first updates (a sanity check of rollout/GAE/update wiring, not a claim)."""
    env_mod, agent_mod, configs = pkg
    cfg = configs.base_config(num_envs=1024, num_steps=64, total_timesteps=1024 * 64 * 12)
    trainer = agent_mod.PPO(lambda i: env_mod.RacingEnv(num_sensors=11), cfg, device='cuda')
    info = trainer.train(log=lambda s: None)
    r = info['rewards']
    assert len(r) >= 8 and np.isfinite(r).all()
    assert np.mean(r[-3:]) > np.mean(r[:3])
    trainer.envs.close()


def test_graphed_update_equals_eager_update(pkg):
    """The CUDA-graph replay of the minibatch step (default on the GPU) gives the
    same parameters as the eager step-by-step update, and honours the KL stop."""
    env_mod, agent_mod, configs = pkg
    results = []
    for graphed, fused, fused_mlp, fused_adam, tc in ((False, False, False, False, False), (True, False, False, False, False),
                                                      (True, True, False, False, False), (True, True, True, False, False),
                                                      (True, True, True, True, False), (True, True, True, True, 1), (True, True, True, True, 2)):
        cfg = configs.base_config(num_envs=64, num_steps=64, update_epochs=2, num_minibatches=4, kl_target=1e9,
                                  cuda_graph_update=graphed, fused_update_kernels=fused,
                                  fused_mlp_update=fused_mlp, fused_adam_step=fused_adam, tensor_core_update=tc)
        vec = env_mod.BatchedRacingVecEnv.synthetic('single', 64, n_tracks=4, seed=0)
        tr = agent_mod.PPO(vec, cfg, device='cuda')
        g = torch.Generator(device='cuda').manual_seed(0)
        n = cfg['batch_size']
        obs = torch.rand(n, 15, device='cuda', generator=g) * 2 - 1
        act = torch.rand(n, 2, device='cuda', generator=g) * 2 - 1
        adv = torch.randn(n, device='cuda', generator=g)
        val = torch.randn(n, device='cuda', generator=g)
        with torch.no_grad():
            _, logp, _, _ = tr.agent.get_action_and_value(obs, act)
        logp = logp + 0.05 * torch.randn(n, device='cuda', generator=g)
        perms = [torch.randperm(n, device='cuda', generator=g) for _ in range(cfg['update_epochs'])]
        tr._anneal(3, 10)  # exercises the in-place learning-rate tensor
        steps = tr.ppo_update(adv, val + adv, val, logp, act, obs, permutation=lambda ep: perms[ep])
        assert steps == 8
        results.append([p.detach().clone() for p in tr.agent.parameters()])
        # KL stop: nothing is applied when the very first minibatch exceeds the target
        tr.config['kl_target'] = 0.015
        before = [p.detach().clone() for p in tr.agent.parameters()]
        assert tr.ppo_update(adv, val + adv, val, logp + 1.0, act, obs, permutation=lambda ep: perms[ep]) == 0
        for a, b in zip(before, tr.agent.parameters()):
            assert torch.equal(a, b)
        vec.close()
    # eager autograd == graphed autograd == graphed + fused loss-gradient kernel == one-kernel forward/loss/backward
    # == the same with clip + Adam + KL stop as one kernel (no per-minibatch host sync)
    # == the same with the per-sample products on the tensor cores (TF32 x 3 passes through TMEM)
    # == the same with the weight gradients on the tensor cores too (the default)
    for other in results[1:]:
        for a, b in zip(results[0], other):
            torch.testing.assert_close(a, b, rtol=1e-4, atol=2e-6)


@pytest.mark.parametrize('n,obs_dim,use_idx', [(1000, 19, True), (128, 15, False), (5000, 19, True), (77, 7, True), (3000, 20, True)])
def test_fused_minibatch_gradient_matches_autograd(pkg, n, obs_dim, use_idx):
    """rk_ppo_minibatch_grad (forward + PPO loss + backward of both MLPs in one
    kernel, rows gathered through the minibatch indices) against torch autograd of
    the reference's loss expression (agent/ppo.py:173-204): fp32 autograd within
    rtol 1e-3, and no farther from a float64 evaluation than fp32 autograd is."""
    _, agent_mod, configs = pkg
    import copy
    from self_play_racing_b200 import spaces
    from self_play_racing_b200.backend import PpoMinibatchGrad
    torch.manual_seed(n)
    agent = agent_mod.Agent(spaces.Box(-1, 1, (obs_dim,)), spaces.Box(-1, 1, (2,))).cuda()
    with torch.no_grad():
        for p in agent.parameters():
            p.add_(0.2 * torch.randn_like(p))      # away from the 0.01-gain output layer: every term matters
        agent.log_std.fill_(-0.7)
    g = torch.Generator(device='cuda').manual_seed(1)
    B = 3 * n
    obs = torch.rand(B, obs_dim, device='cuda', generator=g) * 2 - 1
    act = torch.rand(B, 2, device='cuda', generator=g) * 2 - 1
    adv = torch.randn(B, device='cuda', generator=g) * 3 + 0.5
    val = torch.randn(B, device='cuda', generator=g)
    ret = val + 0.3 * torch.randn(B, device='cuda', generator=g)
    with torch.no_grad():
        _, logp, _, _ = agent.get_action_and_value(obs, act)
    logp = logp + 0.15 * torch.randn(B, device='cuda', generator=g)   # ratios on both sides of the clip range
    idx = torch.randperm(B, device='cuda', generator=g)[:n] if use_idx else None
    clip, vf = 0.2, 0.5

    def reference(net, dtype):
        sel = (lambda t: t[idx]) if use_idx else (lambda t: t[:n])
        o, a, lp, ad, rt, vl = (sel(t).to(dtype) for t in (obs, act, logp, adv, ret, val))
        net.zero_grad()
        _, new_lp, _, new_v = net.get_action_and_value(o, a)
        logratio = new_lp - lp
        ratio = logratio.exp()
        adn = (ad - ad.mean()) / (ad.std() + 1e-8)
        pg = torch.max(-adn * ratio, -adn * torch.clamp(ratio, 1 - clip, 1 + clip)).mean()
        new_v = new_v.flatten()
        vclip = vl + torch.clamp(new_v - vl, -clip, clip)
        vloss = 0.5 * torch.max((new_v - rt) ** 2, (vclip - rt) ** 2).mean()
        (pg + vf * vloss).backward()
        return torch.cat([p.grad.reshape(-1) for p in net.parameters()]), float((-logratio).detach().sum())

    g32, kl32 = reference(agent, torch.float32)
    g64, kl64 = reference(copy.deepcopy(agent).double(), torch.float64)
    fused = PpoMinibatchGrad(list(agent.parameters()), agent.log_std, obs_dim, clip, vf)
    src = (obs, act, logp, adv, ret, val) if use_idx else tuple(t[:n].contiguous() for t in (obs, act, logp, adv, ret, val))
    fused.stats(idx, src[3])
    flat, kl = fused(idx, *src)
    torch.cuda.synchronize()
    assert flat.shape == g32.shape
    torch.testing.assert_close(flat, g32, rtol=1e-3, atol=2e-6)
    err_fused = float((flat.double() - g64).abs().max())
    err_torch = float((g32.double() - g64).abs().max())
    assert err_fused <= 4 * err_torch + 1e-7, (err_fused, err_torch)
    assert abs(float(kl) - kl64) <= 1e-5 * max(1.0, abs(kl64)) + 1e-3
    # replay: deterministic (fixed-order reductions)
    first = flat.clone()
    fused(idx, *src)
    assert torch.equal(first, fused.flat_grad)
    # observation rows padded to 16 bytes (the 128-bit gather path): bit-identical gradient
    dp = (obs_dim + 3) & ~3
    padded = torch.full((src[0].shape[0], dp), 7.0, device='cuda')
    padded[:, :obs_dim] = src[0]
    fused(idx, padded[:, :obs_dim], *src[1:])
    assert torch.equal(first, fused.flat_grad)


@pytest.mark.parametrize('variant', [1, 2])
@pytest.mark.parametrize('n,obs_dim', [(128, 19), (1000, 19), (4133, 15), (77, 7), (40000, 19), (3000, 20)])
def test_tensor_core_gradient_matches_fma_kernel(pkg, n, obs_dim, variant):
    """The tcgen05 variants of rk_ppo_minibatch_grad (1: layer 1, layer 2 and dH1 as TF32 x 3-pass
    tensor-core products chained through TMEM; 2: the weight gradients dW2 / dW1 and their biases as
    M = 64 tensor-core products over shared-memory operand tiles as well) against the fp32 FMA kernel
    on the same inputs: relative difference of the whole gradient <= 2e-5 of its largest element, same
    KL sum, deterministic on replay.  (40000 rows: several tiles per CTA, i.e. accumulation in TMEM.)"""
    _, agent_mod, _ = pkg
    from self_play_racing_b200 import spaces
    from self_play_racing_b200.backend import PpoMinibatchGrad
    torch.manual_seed(n)
    agent = agent_mod.Agent(spaces.Box(-1, 1, (obs_dim,)), spaces.Box(-1, 1, (2,))).cuda()
    with torch.no_grad():
        for p in agent.parameters():
            p.add_(0.2 * torch.randn_like(p))
        agent.log_std.fill_(-0.7)
    g = torch.Generator(device='cuda').manual_seed(2)
    B = 2 * n
    obs = torch.rand(B, obs_dim, device='cuda', generator=g) * 2 - 1
    act = torch.rand(B, 2, device='cuda', generator=g) * 2 - 1
    adv = torch.randn(B, device='cuda', generator=g)
    val = torch.randn(B, device='cuda', generator=g)
    ret = val + 0.3 * torch.randn(B, device='cuda', generator=g)
    with torch.no_grad():
        _, logp, _, _ = agent.get_action_and_value(obs, act)
    logp_new = logp
    logp = logp + 0.15 * torch.randn(B, device='cuda', generator=g)
    # The clipped surrogate is discontinuous in the probability ratio at 1 -+ clip: a row within rounding distance of a
    # clip edge contributes its whole gradient or nothing depending on the last bit of logp (seen at 40,000 rows: ONE row
    # flips between the kernels and moves the actor gradient by 6e-4 of its largest element, identically in both
    # tensor-core variants).  Rows that close to an edge are moved off it; the comparison is about the arithmetic.
    ratio = (logp_new - logp).exp()
    near = ((ratio - 0.8).abs() < 2e-3) | ((ratio - 1.2).abs() < 2e-3)
    logp = torch.where(near, logp + 0.02, logp)
    idx = torch.randperm(B, device='cuda', generator=g)[:n]
    params = list(agent.parameters())
    fma = PpoMinibatchGrad(params, agent.log_std, obs_dim, 0.2, 0.5)
    fma.stats(idx, adv)
    g0, k0 = fma(idx, obs, act, logp, adv, ret, val)
    g0, k0 = g0.clone(), float(k0)
    tc = PpoMinibatchGrad(params, agent.log_std, obs_dim, 0.2, 0.5, tensor_cores=variant)
    tc.stats(idx, adv)
    g1, k1 = tc(idx, obs, act, logp, adv, ret, val)
    g1, k1 = g1.clone(), float(k1)
    assert torch.isfinite(g1).all()
    assert float((g1 - g0).abs().max()) <= 2e-5 * float(g0.abs().max())
    # (the KL SUM: the tensor core accumulates with truncation, which biases every output by a fraction of an fp32 ulp
    #  -- measured 5e-7 per row on log-ratios of order 0.1; the early stop compares kl_sum / n with 0.015)
    assert abs(k1 - k0) <= 1e-3 * max(1.0, abs(k0)) + 2e-6 * n
    tc(idx, obs, act, logp, adv, ret, val)
    assert torch.equal(g1, tc.flat_grad)


@pytest.mark.parametrize('n', [1, 2, 5, 64, 1000, 65536, 4194304, 3000001])
def test_device_permutation_is_a_permutation(pkg, n):
    """rk_random_permutation: every index exactly once, reproducible per (seed,
    counter), different across counters, and not the identity / not sorted."""
    from self_play_racing_b200.backend import random_permutation
    a = random_permutation(n, 7, 1, device='cuda')
    assert a.dtype == torch.int64 and a.shape == (n,)
    assert torch.equal(torch.sort(a).values, torch.arange(n, device='cuda'))
    assert torch.equal(a, random_permutation(n, 7, 1, device='cuda'))
    if n >= 64:
        b = random_permutation(n, 7, 2, device='cuda')
        assert not torch.equal(a, b)
        fixed = float((a == torch.arange(n, device='cuda')).float().mean())
        assert fixed < 0.2
        # no long-range order: neighbours are uncorrelated
        x = a[:-1].double() / n - 0.5
        y = a[1:].double() / n - 0.5
        assert abs(float((x * y).mean()) * 12) < 0.2 if n < 1000 else abs(float((x * y).mean()) * 12) < 0.05


def test_opponent_pool_one_launch_equals_per_opponent_launches(pkg):
    """rk_policy_act_pool: blocks of 256 envs driven by different pool members in ONE
    launch give exactly the actions of one rk_policy_act launch per member on its blocks;
    BatchedRacingVecEnv.set_opponents routes the self-play step through it."""
    env_mod, agent_mod, _ = pkg
    from self_play_racing_b200 import spaces
    from self_play_racing_b200.backend import flatten_agent, policy_act, policy_act_pool
    torch.manual_seed(5)
    agents = [agent_mod.Agent(spaces.Box(-1, 1, (19,)), spaces.Box(-1, 1, (2,))).cuda() for _ in range(3)]
    for a in agents:
        with torch.no_grad():
            for p in a.parameters():
                p.add_(0.3 * torch.randn_like(p))
    flats = [flatten_agent(a.state_dict()).cuda() for a in agents]
    pool = torch.stack(flats).contiguous()
    B, block = 2000, 512            # 4 blocks, the last one ragged
    ids = torch.tensor([2, 0, 1, 2], dtype=torch.int32, device='cuda')
    obs = torch.rand(B, 19, device='cuda') * 2 - 1
    out = torch.zeros(B, 2, device='cuda')
    mean = torch.zeros(B, 2, device='cuda')
    policy_act_pool(pool, ids, block, obs, out, seed=11, counter=3, mean=mean)
    ref, ref_mean = torch.zeros_like(out), torch.zeros_like(mean)
    for k in range(3):
        full, fmean = torch.zeros_like(out), torch.zeros_like(mean)
        policy_act(flats[k], obs, full, seed=11, counter=3, mean=fmean)
        for b in range(4):
            if int(ids[b]) == k:
                ref[b * block:(b + 1) * block] = full[b * block:(b + 1) * block]
                ref_mean[b * block:(b + 1) * block] = fmean[b * block:(b + 1) * block]
    assert torch.equal(out, ref) and torch.equal(mean, ref_mean)
    for k in range(3):   # and the means are the torch Agents' means
        sel = torch.cat([torch.arange(b * block, min((b + 1) * block, B)) for b in range(4) if int(ids[b]) == k]).cuda()
        with torch.no_grad():
            mu = agents[k].actor_mu(obs[sel])
        # (tcgen05 inference: 3-term TF32 products, >= 21 bits per factor, and the tensor core's truncating fp32
        #  accumulation; these agents carry 0.3-sigma weight noise, i.e. pre-activations of several units)
        torch.testing.assert_close(mean[sel], mu, rtol=0, atol=4e-6)
    # through the vector env: the opponent's actions differ between blocks with different members
    vec = env_mod.BatchedRacingVecEnv.synthetic('multi', 1024, n_tracks=4, num_agents=2, selfplay=True, seed=0)
    vec.set_opponents(agents, block_policy=[0, 1, 2, 0], block_len=256)
    vec.reset()
    o, r, te, tr, _ = vec.step(np.zeros((1024, 2), dtype=np.float32))
    assert o.shape == (1024, 19) and np.isfinite(o).all()
    act1 = vec.be.actions[1].cpu().numpy()
    assert not np.allclose(act1[:256], act1[256:512])
    with pytest.raises(ValueError):
        vec.set_opponents(agents, block_policy=[0, 1, 5, 0], block_len=256)
    vec.close()


def test_batched_evaluation_protocol(pkg):
    """evaluate.py's tracks x runs protocol as one batch: result keys of
    evaluate.py:51-64 / utils/metrics.py, reproducible, and -- with a (nearly)
    deterministic policy -- equal to the per-environment loop of utils/metrics.py
    run against the single-env facade."""
    env_mod, agent_mod, _ = pkg
    from self_play_racing_b200.evaluation import evaluate_batched
    np.random.seed(42)
    pool = env_mod.gen_tracks(num_tracks=3, seed=42)
    widths = [int(np.random.RandomState(42 + i).randint(4, 10)) for i in range(3)]
    torch.manual_seed(3)
    proto = env_mod.RacingEnv(num_sensors=11)
    agent = agent_mod.Agent(proto.observation_space, proto.action_space).cuda()
    agent.log_std.data.fill_(-18.0)  # sigma ~ 1.5e-8: sampling is numerically deterministic
    res = evaluate_batched('single', agent, pool, widths, num_tracks=3, num_runs=2, max_steps=80)
    assert set(res) == {'num_episodes', 'num_successful', 'success_rate', 'crash_rate', 'avg_steps', 'avg_reward',
                        'avg_progress', 'avg_speed', 'avg_distance', 'avg_steps_per_progress', 'all_episodes'}
    assert res['num_episodes'] == 6 and 0 <= res['crash_rate'] <= 1
    res2 = evaluate_batched('single', agent, pool, widths, num_tracks=3, num_runs=2, max_steps=80)
    assert res['all_episodes'] == res2['all_episodes']
    # the reference's per-env loop (utils/metrics.py:39-78) against the facade
    for t in range(3):
        for r in range(2):
            env = env_mod.RacingEnv(num_sensors=11, track_pool=pool, track_id=t, track_width=widths[r])
            obs, _ = env.reset()
            total, prev, dist = 0.0, None, 0.0
            for step in range(80):
                with torch.no_grad():
                    a = agent.get_action_and_value(torch.from_numpy(obs).float().unsqueeze(0).cuda())[0].cpu().numpy()[0]
                obs, rew, term, trunc, info = env.step(a)
                total += rew
                if prev is not None:
                    dist += float(np.hypot(info['position'][0] - prev[0], info['position'][1] - prev[1]))
                prev = info['position']
                if term or trunc:
                    break
            m = res['all_episodes'][t * 2 + r]
            assert m['steps'] == step + 1 and m['crashed'] == info['crashed'] and m['finished'] == info['finished']
            assert abs(m['total_reward'] - total) < 1e-2 * max(1.0, abs(total))
            assert abs(m['total_distance'] - dist) < 1e-3 * max(1.0, dist)
            assert abs(m['progress'] - info['progress']) < 1e-9
            env.close()
    # 2-car protocol: both cars driven by the same policy
    magent = agent_mod.Agent(env_mod.MultiRacingEnv(2, 11).observation_space['0'], proto.action_space).cuda()
    magent.log_std.data.fill_(-0.3)
    mres = evaluate_batched('multi', magent, pool, [8, 9], num_tracks=3, num_runs=2, max_steps=400)
    assert mres['num_episodes'] == 6
    for m in mres['all_episodes']:
        assert 1 <= m['steps'] <= 400 and ('placement' in m) and 0 <= m['progress'] <= 1


# ------------------------------------------------ the Gymnasium face at benchmark size (zero-copy host rows)
@pytest.mark.parametrize('kind,E', [('single', 131072), ('single', 65536 + 37), ('multi', 65536)])
def test_gym_face_at_scale_equals_device_face(pkg, kind, E):
    """The path bench.py times as `e2e` -- rk_step_host with host_chunks = 4, the step kernel reading the
    actions from and writing car 0's rows into pinned host memory -- against the device-resident arrays of
    the SAME step, and against a second batch stepped through the plain device face.  131,072 single-car
    envs keep 32 environments per warp (the `nr_sh` scratch of the zero-copy rows is indexed by 32 / A);
    65,536 two-car self-play envs are the benchmark's own configuration."""
    env_mod, agent_mod, _ = pkg
    selfplay = kind == 'multi'
    mk = lambda: env_mod.BatchedRacingVecEnv.synthetic(kind, E, n_tracks=16, num_agents=2, num_sensors=11,
                                                       selfplay=selfplay, seed=77, copy=True)
    vec, ref = mk(), mk()
    assert vec.host_chunks == 4 and vec._host_io.reserved0 == 7
    ref.host_chunks = 0                       # torch-level path: plain copies, one launch
    if selfplay:
        torch.manual_seed(3)
        opp = agent_mod.Agent(vec.single_observation_space, vec.single_action_space)
        vec.set_opponent(opp); ref.set_opponent(opp)
    o1, _ = vec.reset(); o2, _ = ref.reset()
    np.testing.assert_array_equal(o1, o2)
    rs = np.random.RandomState(0)
    n_done = 0
    pinned = vec.pinned_action_buffers(2)    # every other step hands over a page-locked array: read in place, no staging copy
    for k in range(90):                      # random drivers leave the track within 50-90 steps: auto-resets are covered
        a = rs.uniform(-1, 1, size=(E, 2)).astype(np.float32)
        a[:, 1] = np.abs(a[:, 1])
        if k % 2 == 0:
            buf = pinned[(k // 2) % 2]
            buf[...] = a
            obs, rew, term, trunc, infos = vec.step(buf)
            assert vec._host_io.actions == vec._pinned_acts[buf.ctypes.data].data_ptr()
        else:
            obs, rew, term, trunc, infos = vec.step(a)
            assert vec._host_io.actions == vec._h_actions.data_ptr()
        be = vec.be
        # host rows == the device arrays the same kernel wrote
        np.testing.assert_array_equal(obs, be.obs[0].cpu().numpy())
        np.testing.assert_array_equal(rew, be.reward64[0].cpu().numpy())
        np.testing.assert_array_equal(be.actions[0].cpu().numpy(), a)          # actions were written through
        dterm, dtrunc = be.terminated.cpu().numpy().astype(bool), be.truncated.cpu().numpy().astype(bool)
        np.testing.assert_array_equal(term, (dterm | dtrunc) if selfplay else dterm)
        np.testing.assert_array_equal(trunc, dtrunc)
        if not selfplay:
            # ... and == an independent batch stepped through the plain path (the opponent's Philox counters differ
            # between the chunked and the one-launch path, so the two-car batches are not comparable step by step)
            obs2, rew2, term2, trunc2, infos2 = ref.step(a)
            np.testing.assert_array_equal(obs, obs2)
            np.testing.assert_array_equal(rew, rew2)
            np.testing.assert_array_equal(term, term2)
            assert ('episode' in infos) == ('episode' in infos2)
            if 'episode' in infos:
                np.testing.assert_array_equal(infos['_episode'], infos2['_episode'])
                np.testing.assert_array_equal(infos['episode']['r'], infos2['episode']['r'])
                np.testing.assert_array_equal(infos['episode']['l'], infos2['episode']['l'])
                assert set(infos['episode']) == {'r', 'l', 't'}
        n_done += int(term.sum())
    # (single-car ray readings are not clamped at the sensor range -- SURVEY quirk 2 -- so only the 2-car rows are <= 1)
    assert np.isfinite(obs).all() and np.isfinite(rew).all() and (selfplay is False or np.abs(obs).max() <= 1.0)
    assert n_done > 0
    vec.close(); ref.close()


def test_vec_env_reference_style_attribute_loop_and_seed(pkg):
    """agent/ppo.py:256-258 writes speed_weight through `envs.envs[i]` for every i: the proxy list is built once;
    reset(seed=...) re-keys the start-grid shuffle and is reproducible; bad opponent block lengths are rejected
    where they are given."""
    env_mod, agent_mod, _ = pkg
    vec = env_mod.BatchedRacingVecEnv.synthetic('multi', 2048, n_tracks=4, selfplay=True, seed=5)
    assert vec.envs is vec.envs and len(vec.envs) == 2048
    for i in range(vec.num_envs):
        setattr(vec.envs[i], 'speed_weight', 11.0)
    assert vec.speed_weight == 11.0
    o1, _ = vec.reset(seed=123)
    s1 = vec.be.get_state()['car_f64'][..., :2].copy()
    vec.reset(seed=124)
    s2 = vec.be.get_state()['car_f64'][..., :2].copy()
    vec.reset(seed=123)
    s3 = vec.be.get_state()['car_f64'][..., :2].copy()
    assert not np.array_equal(s1, s2)
    np.testing.assert_array_equal(s1, s3)
    with pytest.raises(ValueError, match='options'):
        vec.reset(options={'x': 1})
    pol = agent_mod.Agent(vec.single_observation_space, vec.single_action_space)
    with pytest.raises(ValueError, match='multiple of 256'):
        vec.set_opponents([pol, pol], block_len=100)
    vec.close()


@pytest.mark.parametrize('mode', ['selfplay_random', 'selfplay_snapshot', 'selfplay_pool', 'single'])
def test_native_rollout_equals_python_loop(pkg, mode):
    """rk_rollout (one C call per T-step rollout: the learner's and the opponent's inference in one launch + the step
    kernel, 2 launches per step) against the per-step Python loop (3 launches per step) on twin trainers: every rollout
    buffer is bit-identical over two consecutive rollouts, including the stale slot 0 after update_opponent's reset
    (SURVEY quirk 10)."""
    env_mod, agent_mod, configs = pkg
    np.random.seed(1)
    pool = env_mod.gen_tracks(num_tracks=8, seed=1)
    widths = [int(np.random.randint(6, 10)) for _ in range(8)]
    E, T = 768, 112
    out = []
    for native in (True, False):
        if mode == 'single':
            cfg = configs.base_config(num_envs=E, num_steps=T, total_timesteps=10 ** 9, native_rollout=native)
            env_fn = lambda i: env_mod.RacingEnv(num_sensors=11, track_pool=pool, track_id=i % 8, track_width=widths)
            tr = agent_mod.PPO(env_fn, cfg, device='cuda')
        else:
            cfg = configs.self_play_config(num_envs=E, num_steps=T, total_timesteps=10 ** 9, native_rollout=native,
                                           opponents_per_update='pool' if mode == 'selfplay_pool' else 'one')
            env_fn = lambda i: env_mod.MultiRacingEnv(num_agents=2, num_sensors=11, track_pool=pool, track_id=i % 8,
                                                      track_width=widths)
            tr = agent_mod.SelfPlayPPO(env_fn, cfg, device='cuda')
            if mode != 'selfplay_random':
                torch.manual_seed(5)
                for k in range(3):
                    snap = tr.snapshot_agent()
                    for p_ in snap.parameters():
                        p_.data.add_(0.05 * torch.randn_like(p_))
                    tr.opponent_pool.append(snap)
        buf = tr.alloc_buffers()
        buf['obs'][0].copy_(tr._reset_all())
        snaps = []
        for it in range(2):
            if mode != 'single':
                np.random.seed(11 + it)
                tr.update_opponent()
            stats = tr.collect_rollout(buf)
            snaps.append(({k: v.clone() for k, v in buf.items()}, stats))
            buf['obs'][0].copy_(buf['obs'][T]); buf['dones'][0].copy_(buf['dones'][T])
        assert ('native' in tr.rollout_mode) == native
        out.append(snaps)
        tr.envs.close()
    for (b1, s1), (b2, s2) in zip(*out):
        assert s1[0] == s2[0] and np.allclose(s1[1:], s2[1:], rtol=1e-12)   # (episode sums are atomic adds: order-dependent ulps)
        for k in b1:
            assert torch.equal(b1[k], b2[k]), k
        assert float(b1['dones'].sum()) > 0 and torch.isfinite(b1['values']).all()


def test_rollout_buffers_match_reference_collect_rollout(pkg, golden):
    """The reference's own SelfPlayPPO.collect_rollout (agent/ppo.py:97-132) over SyncVectorEnv(RecordEpisodeStatistics(
    SelfPlayWrapper(MultiRacingEnv))) -- recorded by tools/make_golden.py record_rollout_buffers, two rollouts with
    update_opponent() in between -- against the device rollout buffers: slot t of obs / dones holds next_obs / next_done
    BEFORE step t, rewards / values / logprobs belong to step t, the rebuilt envs start fresh while slot 0 carries the
    stale next_obs, episode statistics are those of RecordEpisodeStatistics.  The learner's and the
    opponent's recorded actions and the recorded start slots are injected."""
    env_mod, agent_mod, _ = pkg
    g = golden('rollout_selfplay16')
    cps = np.split(g['pool'], np.cumsum(g['pool_sizes'])[:-1])
    widths = [float(w) for w in g['widths']]
    T, E = g['actions0'].shape[:2]
    fns = [(lambda i=i: env_mod.SelfPlayWrapper(env_mod.MultiRacingEnv(num_agents=2, num_sensors=11, track_pool=cps,
                                                                      track_id=i % len(cps), track_width=widths), 0))
           for i in range(E)]
    vec = env_mod.BatchedRacingVecEnv(fns, query='culled')
    be, dev = vec.be, vec.be.device
    sd = {k[3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith('sd.')}
    agent = agent_mod.Agent(vec.single_observation_space, vec.single_action_space).to(dev)
    agent.load_state_dict(sd)
    D = be.D
    buf = dict(obs=torch.zeros(T + 1, 2, E, D, device=dev), actions=torch.zeros(T, 2, E, 2, device=dev),
               dones=torch.zeros(T + 1, E, device=dev), rewards=torch.zeros(T, 2, E, device=dev))
    t_ = lambda a, dt=torch.float32: torch.from_numpy(np.ascontiguousarray(a)).to(dev, dt)
    for it in range(2):
        # update_opponent(): the envs are rebuilt (fresh reset with new grid slots), slot 0 keeps the carried next_obs
        vec.reset_device(start_slot=t_(g[f'slots_init{it}'], torch.int32))
        be.ep_stats.zero_()
        buf['obs'][0, 0].copy_(t_(g[f'obs{it}'][0]))
        buf['dones'][0].copy_(t_(g[f'dones{it}'][0]))
        buf['actions'][:, 0].copy_(t_(g[f'actions{it}']))
        opp, slots = t_(g[f'opp_actions{it}']), t_(g[f'slots{it}'], torch.int32)
        for t in range(T):
            vec.step_into(buf['actions'][t], buf['obs'][t + 1], buf['rewards'][t], buf['dones'][t + 1],
                          start_slot=slots[t], opponent_actions=opp[t])
        obs = buf['obs'][:, 0].cpu().numpy()
        np.testing.assert_allclose(obs[1:T], g[f'obs{it}'][1:], rtol=0, atol=1e-6)
        np.testing.assert_allclose(obs[T], g[f'next_obs{it}'], rtol=0, atol=1e-6)
        np.testing.assert_array_equal(buf['dones'][1:T].cpu().numpy(), g[f'dones{it}'][1:])
        np.testing.assert_array_equal(buf['dones'][T].cpu().numpy().astype(bool), g[f'next_done{it}'])
        np.testing.assert_allclose(buf['rewards'][:, 0].cpu().numpy(), g[f'rewards{it}'], rtol=1e-6, atol=1e-5)  # float32 buffers
        # a reset shows as reward 0 one step after a done (NEXT_STEP), exactly where the reference reset
        np.testing.assert_array_equal((buf['dones'][:T] > 0).cpu().numpy() & (np.arange(T)[:, None] > 0), g[f'reset{it}'] & (np.arange(T)[:, None] > 0))
        n, sr, sl = len(g[f'ep_r{it}']), g[f'ep_r{it}'].sum(), g[f'ep_l{it}'].sum()
        st = be.ep_stats.cpu().numpy()
        assert int(st[2]) == n and abs(st[0] - sr) < 1e-6 and int(st[1]) == int(sl)
        # values / log-probs of the recorded (obs, action) pairs under the same weights (torch path of the host layer)
        with torch.no_grad():
            o = t_(g[f'obs{it}']).reshape(T * E, D)
            _, lp, _, v = agent.get_action_and_value(o, t_(g[f'actions{it}']).reshape(T * E, 2))
        np.testing.assert_allclose(v.reshape(T, E).cpu().numpy(), g[f'values{it}'], rtol=0, atol=2e-5)
        np.testing.assert_allclose(lp.reshape(T, E).cpu().numpy(), g[f'logprobs{it}'], rtol=0, atol=2e-4)
        assert g[f'dones{it}'].sum() >= 5
    vec.close()


def test_batched_evaluation_matches_reference_metrics(pkg, golden):
    """evaluate.py's protocol: utils/metrics.py eval_single_agent / eval_multi_agent run on the UNMODIFIED reference envs
    (tools/make_golden.py record_eval) against evaluate_batched on the device.  The policy is a hand-wired, deterministic
    Agent (log_std = -100: sampling returns the mean) that drives complete laps; both sides evaluate the same MLP in
    float32 by different code, so actions agree to ~1e-6 and the episodes stay together: flags / placement exact, steps
    within 1, totals within 2e-3 relative."""
    env_mod, agent_mod, _ = pkg
    from self_play_racing_b200.evaluation import evaluate_batched
    g = golden('eval_protocol')
    cps = np.split(g['pool'], np.cumsum(g['pool_sizes'])[:-1])
    run_w = [float(w) for w in g['run_widths']]
    for kind, key, sdk, nt in (('single', 'single', 'sd1.', 3), ('multi', 'multi', 'sd2.', 2)):
        proto = env_mod.RacingEnv(num_sensors=11) if kind == 'single' else env_mod.MultiRacingEnv(num_agents=2, num_sensors=11)
        osp = proto.observation_space if kind == 'single' else proto.observation_space['0']
        asp = proto.action_space if kind == 'single' else proto.action_space['0']
        agent = agent_mod.Agent(osp, asp)
        agent.load_state_dict({k[len(sdk):]: torch.from_numpy(v) for k, v in g.items() if k.startswith(sdk)})
        kw = {}
        if kind == 'multi':
            kw['start_slot'] = g['multi_slots']
        res = evaluate_batched(kind, agent, cps[:nt], run_w, num_tracks=nt, num_runs=len(run_w), **kw)
        ref = g[key]
        assert res['num_episodes'] == len(ref) and res['num_successful'] == int(ref[:, 3].sum()) == len(ref)
        for m, r in zip(res['all_episodes'], ref):
            assert m['finished'] == bool(r[3]) and m['crashed'] == bool(r[4])
            assert abs(m['steps'] - r[1]) <= 1
            assert abs(m['progress'] - r[2]) < 1e-9
            assert abs(m['total_reward'] - r[0]) <= 2e-3 * abs(r[0])
            assert abs(m['total_distance'] - r[6]) <= 2e-3 * r[6] and abs(m['speed'] - r[5]) < 2e-2
            if kind == 'multi':
                assert m['placement'] == int(r[7])
        assert abs(res['avg_steps'] - ref[:, 1].mean()) <= 1 and res['success_rate'] == 1.0 and res['crash_rate'] == 0.0
