#!/usr/bin/env python
"""Benchmark of the batched racing step path (BASELINE.json metric: agent
env-steps/s at 1/2/4/8 B200 with 65,536 envs; self-play PPO SPS).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

One "step" = one pass of the hot path over all E environments of a rank: the
opponent-snapshot MLP inference kernel + the fused environment step kernel
(2-car workload), or the step kernel alone (single-car workload).  Rank 0 prints
ONE JSON line.  The top-level numbers are weak-scaled (65,536 envs PER GPU); the
`strong` block repeats them for the metric's literal configuration (65,536 envs
in TOTAL, split evenly over the ranks).  See DESIGN.md "Measurement".
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE configs[2]: the configuration the metric and the 1e9 target are quoted on
    'multi2_selfplay_65536': dict(kind='multi', A=2, R=11, E=65536, tracks=16, factor=30, selfplay=True),
    # BASELINE configs[1]
    'single_65536': dict(kind='single', A=1, R=11, E=65536, tracks=16, factor=30, selfplay=False),
    # BASELINE configs[4]: 1,048,576 envs x 4 cars x 64 rays (S via --factor; widths >= 9 so 4 cars fit abreast)
    'sweep_4car_64ray': dict(kind='multi', A=4, R=64, E=1048576, tracks=16, factor=30, selfplay=False, width_lo=9.0),
}


def alg_bytes_per_agent_step(D):
    """state r/w (fp64 here) + action + obs + reward + flags, shared tracks."""
    return 2 * (5 * 8 + 4 + 16) + 8 + 4 * D + 4 + 8 + 3


def alg_flops_per_agent_step(R, S, A, N):
    """F = 13 R (S + 4(A-1)) + 25 N + 200 (reference formulation, brute force)."""
    return 13 * R * (S + 4 * (A - 1)) + 25 * N + 200


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons of one GPU through NVML: a synchronous read when the timed region
    starts and when it ends (so that a region of a few milliseconds is never `samples: 0`) plus a
    background thread sampling every 5 ms in between."""
    REASONS = (('hw_slowdown', 0x8), ('hw_thermal_slowdown', 0x40), ('sw_thermal_slowdown', 0x20), ('sw_power_cap', 0x4))

    def __init__(self, index):
        super().__init__(daemon=True)
        self.rows, self._halt, self.h, self.nv = [], threading.Event(), None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get('CUDA_VISIBLE_DEVICES')
            phys = int(vis.split(',')[index]) if vis and all(v.strip().isdigit() for v in vis.split(',')) else index
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def read(self):
        if self.h is None:
            return
        try:
            nv = self.nv
            sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            get = getattr(nv, 'nvmlDeviceGetCurrentClocksEventReasons', None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            self.rows.append((sm, int(get(self.h))))
        except Exception:  # noqa: BLE001
            pass

    def run(self):
        while not self._halt.is_set():
            self.read()
            time.sleep(0.005)

    def begin(self):
        self.read()
        self.start()

    def stop(self):
        self._halt.set()
        if self.is_alive():
            self.join(timeout=1.0)
        self.read()
        if not self.rows:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0, 'error': getattr(self, 'err', 'no samples')}
        sm = [r[0] for r in self.rows]
        bits = 0
        for r in self.rows:
            bits |= r[1]
        return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': self.max_mhz, 'sm_min_mhz': float(min(sm)),
                'reasons': [n for n, b in self.REASONS if bits & b], 'samples': len(sm), 'source': 'nvml, 5 ms period'}


# ------------------------------------------------------------------ CPU legs
def _oracle_shard(args):
    """Time `steps` oracle steps of `E` envs in this process (after 1 warm-up step)."""
    kind, A, R, E, steps, seed = args
    from oracle import racing_oracle as O
    rs = np.random.RandomState(seed)
    cps = [O.gen_random_track(rs.randint(10, 15), rs.randint(50, 80), rs.randint(10, 20), rs.uniform(0.2, 0.7),
                              rs.uniform(0.2, 0.7), rng=rs) for _ in range(4)]
    tracks = O.make_pool(cps, [8.0, 7.0, 9.0, 8.0])
    env = O.OracleVecEnv(tracks, np.arange(E) % 4, kind=kind, num_agents=A, num_sensors=R, seed=seed)
    env.reset()

    def act():
        a = rs.uniform(-1, 1, size=(E, env.A, 2)).astype(np.float32)
        a[..., 1] = np.abs(a[..., 1])
        return a
    env.step(act())
    t0 = time.perf_counter()
    for _ in range(steps):
        env.step(act())
    return time.perf_counter() - t0


def _have_ref():
    from oracle import make_ref
    return make_ref.import_ref()


def _ref_make_env(wl, seed):
    """One UNMODIFIED reference environment of the workload (oracle/_ref): RacingEnv(11 sensors) or
    SelfPlayWrapper(MultiRacingEnv(2, 11)) with a frozen Agent snapshot as the opponent, on a procedural
    track drawn like train.py's pool (gen_random_track, widths 6-9)."""
    import torch
    from environment.track import gen_random_track
    from environment.racing_env import RacingEnv
    from environment.multi_racing_env import MultiRacingEnv
    from environment.wrappers import SelfPlayWrapper
    from agent.ppo import Agent
    torch.set_num_threads(1)
    rs = np.random.RandomState(seed)
    pool = [gen_random_track(rs.randint(10, 15), rs.randint(50, 80), rs.randint(10, 20), rs.uniform(0.2, 0.7),
                             rs.uniform(0.2, 0.7), seed=int(rs.randint(1 << 30))) for _ in range(4)]
    widths = [8, 7, 9, 8]
    np.random.seed(seed)
    if wl['kind'] == 'single':
        return RacingEnv(num_sensors=wl['R'], track_pool=pool, track_id=seed % 4, track_width=widths)
    env = MultiRacingEnv(num_agents=wl['A'], num_sensors=wl['R'], track_pool=pool, track_id=seed % 4, track_width=widths)
    if not wl['selfplay']:
        return env
    env = SelfPlayWrapper(env, 0)
    env.device = torch.device('cpu')                     # the CPU arm: the opponent snapshot runs on the host as well
    torch.manual_seed(1)
    opp = Agent(env.observation_space, env.action_space)
    opp.log_std.data.fill_(-0.3)
    env.set_opponent(opp)
    return env


def _ref_env_steps(env, wl, rs, n):
    """n steps with uniform random learner actions, reset on done; returns agent-steps done."""
    A = wl['A']
    for _ in range(n):
        if wl['kind'] == 'multi' and not wl['selfplay']:
            acts = {str(i): rs.uniform(-1, 1, 2).astype(np.float32) for i in range(A)}
            _, _, dones, trunc, _ = env.step(acts)
            done = dones['__all__']
        else:
            a = rs.uniform(-1, 1, 2).astype(np.float32)
            a[1] = abs(a[1])
            _, _, te, tr, _ = env.step(a)
            done = te or tr
        if done:
            env.reset()
    return n * A


def cpu_baseline(wl, seconds=12.0):
    """Rank 0, N=1 only: a bounded single-core sample of the same workload.  With oracle/_ref present this is the
    UNMODIFIED reference (`kind: "reference"`: per-environment Python, the reference's execution model); the batched
    numpy port (oracle/racing_oracle.py) is timed beside it, or alone when the copy is absent."""
    port_envs, port_steps = 64, 12
    dt = _oracle_shard((wl['kind'], wl['A'], wl['R'], port_envs, port_steps, 0))
    port = port_envs * wl['A'] * port_steps / dt
    out = {'value': port, 'unit': 'agent-steps/s', 'cores': 1, 'kind': 'port',
           'sample': f'{port_envs} envs x {port_steps} steps of the {wl["kind"]} workload, oracle/racing_oracle.py '
                     f'(batched numpy restatement)'}
    if _have_ref():
        env = _ref_make_env(wl, 0)
        env.reset()
        rs = np.random.RandomState(0)
        _ref_env_steps(env, wl, rs, 20)
        t0, units, n = time.perf_counter(), 0, 0
        while time.perf_counter() - t0 < seconds:
            units += _ref_env_steps(env, wl, rs, 50)
            n += 50
        dt = time.perf_counter() - t0
        out = {'value': units / dt, 'unit': 'agent-steps/s', 'cores': 1, 'kind': 'reference',
               'sample': f'{n} steps of ONE unmodified reference env of the {wl["kind"]} workload (oracle/_ref, '
                         f'{"SelfPlayWrapper + frozen Agent opponent on the host, " if wl["selfplay"] else ""}'
                         f'uniform random actions, reset on done), {dt:.1f} s',
               'port': {'value': port, 'unit': 'agent-steps/s', 'cores': 1,
                        'sample': f'{port_envs} envs x {port_steps} steps, oracle/racing_oracle.py (batched numpy port)'}}
    return out


def ppo_cpu_baseline(num_steps=48):
    """BASELINE.md 3.2 on a bounded sample: the UNMODIFIED reference's SelfPlayPPO loop (oracle/_ref) on the host --
    16 envs (the reference's num_envs), `num_steps` steps per rollout instead of 2048, 10 epochs x 16 minibatches,
    one warm-up iteration then one timed iteration: update_opponent + collect_rollout + GAE + ppo_update."""
    if not _have_ref():
        return None
    import contextlib
    import io
    import torch
    if torch.cuda.is_available():
        # the reference's SelfPlayWrapper puts the opponent's observations on "cuda if available" whatever device
        # the trainer was given (environment/wrappers.py:17): the host-only baseline therefore runs the unmodified
        # code in a child process that sees no GPU
        import subprocess
        env = dict(os.environ, CUDA_VISIBLE_DEVICES='')
        for k in ('RANK', 'WORLD_SIZE', 'LOCAL_RANK'):
            env.pop(k, None)
        out = subprocess.run([sys.executable, os.path.abspath(__file__), '--ppo-cpu-baseline', str(num_steps)],
                             env=env, capture_output=True, text=True, timeout=900)
        lines = [ln for ln in out.stdout.splitlines() if ln.startswith('{')]
        return json.loads(lines[-1]) if lines else {'error': (out.stderr or 'no output')[-400:]}
    from agent.self_play_ppo import SelfPlayPPO
    from configs.self_play_config import hyperparams_config
    from environment.multi_racing_env import MultiRacingEnv
    from environment.track import gen_tracks
    torch.set_num_threads(max(1, min(8, len(os.sched_getaffinity(0)))))
    cfg = hyperparams_config()
    cfg.update(num_steps=num_steps, cuda=False)
    cfg['batch_size'] = cfg['num_steps'] * cfg['num_envs']
    cfg['minibatch_size'] = cfg['batch_size'] // cfg['num_minibatches']
    np.random.seed(1)
    pool = gen_tracks(num_tracks=16, seed=1)                       # train.py:29-30
    widths = [int(np.random.randint(6, 10)) for _ in range(16)]
    env_fn = lambda i: MultiRacingEnv(num_agents=2, num_sensors=11, track_pool=pool, track_id=i % 16, track_width=widths)
    with contextlib.redirect_stdout(io.StringIO()):
        tr = SelfPlayPPO(env_fn, cfg, device='cpu')
        tr.opponent_pool = [tr.snapshot_agent() for _ in range(cfg['pool_size'])]
        c = cfg
        z = lambda *s: torch.zeros(*s)
        obs, actions = z(c['num_steps'], c['num_envs'], 19), z(c['num_steps'], c['num_envs'], 2)
        logprobs, dones, rewards, values = (z(c['num_steps'], c['num_envs']) for _ in range(4))
        init_obs, _ = tr.envs.reset()
        next_obs, next_done = torch.from_numpy(init_obs), torch.zeros(c['num_envs'], dtype=torch.bool)
        times = []
        for it in range(2):
            t0 = time.perf_counter()
            tr.update_opponent()
            out = tr.collect_rollout(obs, actions, logprobs, dones, rewards, values, next_obs, next_done)
            obs, actions, logprobs, dones, rewards, values, next_obs, next_done, _ = out
            t1 = time.perf_counter()
            with torch.no_grad():
                next_value = tr.agent.get_value(next_obs).flatten()
            adv, ret = tr.compute_advantages(rewards, dones, values, next_value, next_done)
            tr.config['kl_target'] = 1e9                         # every optimizer step is paid for, as in the GPU arm
            tr.ppo_update(adv, ret, values, logprobs, actions, obs)
            t2 = time.perf_counter()
            times.append((t1 - t0, t2 - t1))
    roll, upd = times[-1]
    return {'sps': c['batch_size'] / (roll + upd), 'unit': 'learner transitions/s', 'kind': 'reference',
            'rollout_s': roll, 'gae_plus_update_s': upd, 'cores': torch.get_num_threads(),
            'sample': f'unmodified reference SelfPlayPPO (oracle/_ref) on the host: 16 envs x {num_steps} steps '
                      f'(reference: 16 x 2048), 10 epochs x 16 minibatches, pool of 5, second of two iterations'}


_REF_ENV = None


def _ref_init(wl, E, seed_base, use_ref):
    """Pool initializer: every worker process builds its own environment(s) once."""
    global _REF_ENV
    seed = seed_base + os.getpid() % 1000
    rs = np.random.RandomState(seed)
    if use_ref and _have_ref():
        env = _ref_make_env(wl, seed)
        env.reset()
        _REF_ENV = ('ref', env, rs)
        return
    from oracle import racing_oracle as O
    cps = [O.gen_random_track(rs.randint(10, 15), rs.randint(50, 80), rs.randint(10, 20), rs.uniform(0.2, 0.7),
                              rs.uniform(0.2, 0.7), rng=rs) for _ in range(4)]
    env = O.OracleVecEnv(O.make_pool(cps, [8.0, 7.0, 9.0, 8.0]), np.arange(E) % 4, kind=wl['kind'], num_agents=wl['A'],
                         num_sensors=wl['R'], seed=seed)
    env.reset()
    _REF_ENV = ('port', env, rs)


def _ref_step(arg):
    inner, wl = arg
    kind, env, rs = _REF_ENV
    if kind == 'ref':
        return _ref_env_steps(env, wl, rs, inner)
    for _ in range(inner):
        a = rs.uniform(-1, 1, size=(env.E, env.A, 2)).astype(np.float32)
        a[..., 1] = np.abs(a[..., 1])
        env.step(a)
    return env.E * env.A * inner


def _cpu_quota():
    try:
        q, per = open('/sys/fs/cgroup/cpu.max').read().split()
        return None if q == 'max' else float(q) / float(per)
    except Exception:  # noqa: BLE001
        return None


def run_reference_arm(args, wl, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on all host cores.  With oracle/_ref
    present (a copy of the unmodified pure-Python reference made by oracle/make_ref.py in the build container) every
    worker process steps ONE reference environment of the workload -- the embarrassingly parallel CPU deployment of
    BASELINE.md section 3 -- `kind: "reference"`; `--ref-kind port` (or a missing copy) times the batched numpy port
    instead."""
    if rank != 0:
        return
    import multiprocessing as mp
    cores = len(os.sched_getaffinity(0))
    use_ref = args.ref_kind != 'port' and _have_ref()
    E_proc, inner = (1, 40) if use_ref else (32, 2)
    ctx = mp.get_context('fork')
    with ctx.Pool(cores, initializer=_ref_init, initargs=(wl, E_proc, 100, use_ref)) as pool:
        for _ in range(max(args.warmup, 1)):
            pool.map(_ref_step, [(4 if use_ref else 1, wl)] * cores, chunksize=1)
        t0 = time.perf_counter()
        units = 0
        for _ in range(args.steps):
            units += sum(pool.map(_ref_step, [(inner, wl)] * cores, chunksize=1))
        wall = time.perf_counter() - t0
    val = units / wall
    kind = 'reference' if use_ref else 'port'
    what = ('ONE unmodified reference env (oracle/_ref' + (', SelfPlayWrapper + frozen Agent opponent' if wl['selfplay'] else '') + ')') \
        if use_ref else f'{E_proc} envs of the batched numpy port (oracle/racing_oracle.py)'
    sample = f'each step = {cores} processes x {inner} env-steps of {what}, {wl["kind"]} workload'
    line = {'impl': 'reference', 'metric': 'agent_env_steps_per_sec', 'value': val, 'unit': 'agent-steps/s',
            'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': 1e3 * wall / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': args.workload, 'sample': sample, 'cgroup_cpu_quota': _cpu_quota(),
                       'note': 'a bounded CPU sample of the same workload kind (the GPU arm steps 65,536 such envs per GPU)'},
            'cpu_baseline': {'value': val, 'unit': 'agent-steps/s', 'cores': cores, 'kind': kind, 'sample': sample},
            'e2e': {'value': val, 'unit': 'agent-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    if args.ppo_updates > 0 and wl['selfplay'] and use_ref:
        line['ppo'] = ppo_cpu_baseline()
    print(json.dumps(line))


# ------------------------------------------------------------------ GPU arm
def make_vec(wl, E, dev, rank, query):
    import torch
    from self_play_racing_b200.backend import flatten_agent
    from self_play_racing_b200.environment.vec_env import BatchedRacingVecEnv
    from self_play_racing_b200.agent.ppo import Agent
    vec = BatchedRacingVecEnv.synthetic(wl['kind'], E, n_tracks=wl['tracks'], num_agents=wl['A'], num_sensors=wl['R'],
                                        selfplay=wl['selfplay'], device=dev, query=query, seed=1000 + rank,
                                        copy=False, factor=wl['factor'], width_lo=wl.get('width_lo', 6.0))
    if wl['selfplay']:
        # opponent snapshot: orthogonal-init Agent, manual_seed(1), log_std -0.3
        torch.manual_seed(1)
        opp = Agent(vec.single_observation_space, vec.single_action_space)
        opp.log_std.data.fill_(-0.3)
        vec.set_opponent(flatten_agent(opp.state_dict()).to(dev))
    return vec


def measure_step(args, wl, vec, E, dev, rank, world, flush, sample_clocks):
    """value (device-resident), the step kernel alone, and e2e (Gymnasium face, host numpy in / out) for one batch."""
    import torch
    import torch.distributed as dist
    from self_play_racing_b200 import _lib
    be, A = vec.be, wl['A']
    g = torch.Generator(device=dev)
    g.manual_seed(7 + rank)
    n_act = 8       # learner / all-car actions resident in HBM: 8 pre-drawn uniform tensors cycled through
    if wl['selfplay']:
        act_pool = torch.rand(n_act, E, 2, device=dev, generator=g) * 2 - 1
    else:
        act_pool = torch.rand(n_act, *be.actions.shape, device=dev, generator=g) * 2 - 1
    act_pool[..., 1] = act_pool[..., 1].abs()

    def one_step(k):
        if wl['selfplay']:
            vec.step_device(act_pool[k % n_act])
        else:
            be.actions.copy_(act_pool[k % n_act])
            be.step()

    vec.reset_device()
    for k in range(args.warmup):
        one_step(k)
    torch.cuda.synchronize(dev)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sampler = ClockSampler(dev.index) if sample_clocks else None
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    if sampler:
        sampler.begin()
    l0 = _lib.launch_count()
    for k in range(args.steps):
        flush.zero_()
        ev[k][0].record()
        one_step(k)
        ev[k][1].record()
    torch.cuda.synchronize(dev)
    launches = _lib.launch_count() - l0
    if world > 1:
        dist.barrier()
    total_ms = float(sum(a.elapsed_time(b) for a, b in ev))
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())

    # ---- the step kernel alone (the dominant kernel), for the roofline line
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for k in range(args.steps):
        if wl['selfplay']:
            be.actions[0].copy_(act_pool[k % n_act])
            vec._opponent_act()
        else:
            be.actions.copy_(act_pool[k % n_act])
        flush.zero_()
        kev[k][0].record()
        be.step()
        kev[k][1].record()
    torch.cuda.synchronize(dev)
    kms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    clocks = sampler.stop() if sampler else None

    # ---- end to end through the Gymnasium face: host numpy in, host numpy out
    rs = np.random.RandomState(3 + rank)
    # the step's inputs live in pinned host memory (four pre-drawn action arrays cycled through, as a host-side
    # policy would fill them): BatchedRacingVecEnv.pinned_action_buffers(); the kernel reads them over PCIe
    host_actions = vec.pinned_action_buffers(4) if hasattr(vec, 'pinned_action_buffers') else [np.empty((E, 2), np.float32) for _ in range(4)]
    for a in host_actions:
        a[...] = rs.uniform(-1, 1, size=(E, 2)).astype(np.float32)
        a[:, 1] = np.abs(a[:, 1])
    e2e = None
    if wl['kind'] == 'single' or wl['selfplay']:
        vec.reset()
        for k in range(min(args.warmup, 10)):
            vec.step(host_actions[k % 4])
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for k in range(args.steps):
            vec.step(host_actions[k % 4])
        torch.cuda.synchronize(dev)
        wall = time.perf_counter() - t0
        tw = torch.tensor([wall], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tw, op=dist.ReduceOp.MAX)
        zero_copy = getattr(vec, 'host_chunks', 0) > 0 and int(os.environ.get('RK_B200_ZEROCOPY_OBS', '7')) & 1
        e2e = {'value': E * A * args.steps * world / float(tw.item()), 'unit': 'agent-steps/s',
               'h2d_bytes_per_step': int(vec.h2d_bytes_per_step), 'd2h_bytes_per_step': int(vec.d2h_bytes_per_step),
               'ms_per_step': 1e3 * float(tw.item()) / args.steps,
               'transfer': ('the step kernel reads the actions from and writes the results into the pinned host buffers '
                            '(zero-copy over PCIe, same bytes)' if zero_copy else 'cudaMemcpyAsync from/to pinned host buffers')}
    return {'value': E * A * args.steps * world / (total_ms_max * 1e-3), 'ms_per_step': total_ms_max / args.steps,
            'kernel_ms': kms, 'launches': int(launches), 'e2e': e2e, 'clocks': clocks}


def time_allreduce(dev, world, n_floats=11076, reps=160):
    """The update's only collective on its own: `reps` NCCL all-reduces of the flat gradient (+ KL slot), captured
    in one CUDA graph like the epoch graph, device-timed; microseconds per all-reduce, max over ranks."""
    import torch
    import torch.distributed as dist
    if world <= 1:
        return None
    buf = torch.zeros(n_floats, device=dev)
    for _ in range(5):
        dist.all_reduce(buf)
    torch.cuda.synchronize(dev)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for _ in range(reps):
            dist.all_reduce(buf)
    graph.replay()
    torch.cuda.synchronize(dev)
    out = []
    for _ in range(5):
        dist.barrier()
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        graph.replay()
        b.record()
        torch.cuda.synchronize(dev)
        out.append(a.elapsed_time(b) * 1e3 / reps)
    t = torch.tensor([float(np.median(out)), float(min(out))], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return {'bytes': 4 * n_floats, 'in_graph_reps': reps, 'us_median': float(t[0]), 'us_best': float(t[1])}


def time_ppo(args, wl, vec, E, dev, rank, world, with_grad_kernel=True):
    """Full self-play PPO iterations on the same batch of environments: device-resident
    rollout (policy + opponent inference + env step), GAE kernel, clipped-surrogate update
    (10 epochs x 16 minibatches, KL early stop disabled so that every optimizer step is
    paid for), NCCL all-reduce of gradients and minibatch statistics when world > 1.
    SPS = learner transitions per second = world * E * T / seconds per iteration."""
    import torch
    import torch.distributed as dist
    from self_play_racing_b200 import configs
    from self_play_racing_b200.agent import SelfPlayPPO
    T = args.ppo_steps
    cfg = configs.self_play_config(num_envs=E, num_steps=T, total_timesteps=10 ** 12, kl_target=1e9,
                                   update_matmul_precision=args.ppo_precision)
    trainer = SelfPlayPPO(vec, cfg, device=str(dev))
    trainer.opponent_pool = [trainer.snapshot_agent() for _ in range(cfg['pool_size'])]  # pool of 5 snapshots
    buf = trainer.alloc_buffers()
    buf['obs'][0].copy_(trainer._reset_all())
    ev = lambda: torch.cuda.Event(enable_timing=True)
    times = []
    for it in range(args.ppo_updates + 2):  # the first two iterations are warm-up (graph captures)
        trainer.update_opponent()
        trainer._anneal(it, 100)
        e0, e1, e2 = ev(), ev(), ev()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        e0.record()
        trainer.collect_rollout(buf)
        e1.record()
        steps = trainer._learn_from(buf)
        e2.record()
        torch.cuda.synchronize(dev)
        if it > 1:
            times.append((e0.elapsed_time(e1), e1.elapsed_time(e2), steps))
    # the update's dominant kernel on its own: rk_ppo_minibatch_grad over the trainer's padded observation
    # buffer, CUDA events on the launching stream, L2 flushed between calls; both implementations of the entry
    # point (fp32 FMA kernel / tcgen05 TF32 x 3-pass kernel) on the same inputs
    grad_kernel = None
    g = getattr(trainer, '_graphed', None)
    if with_grad_kernel and g is not None and getattr(g, 'fused_mlp', False) and getattr(g, 'obs_pad', None) is not None:
        from self_play_racing_b200.backend import PpoMinibatchGrad
        n_rows, mb = g.obs_pad.shape[0], g.mb
        D = buf['obs'].shape[-1]
        gen = torch.Generator(device=dev).manual_seed(0)
        act = torch.rand(n_rows, 2, device=dev, generator=gen) * 2 - 1
        lp, adv, ret, val = (torch.randn(n_rows, device=dev, generator=gen) for _ in range(4))
        perm = torch.randperm(n_rows, device=dev, generator=gen)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        evs = [(ev(), ev()) for _ in range(min(16, n_rows // mb))]
        params = list(trainer.agent.parameters())

        def time_impl(tensor_cores):
            k = PpoMinibatchGrad(params, trainer.agent.log_std, D, cfg['clip_coef'], cfg['vf_coef'], tensor_cores=tensor_cores)
            for rep in range(2):
                for i, (a, b) in enumerate(evs):
                    idx = perm[i * mb:(i + 1) * mb]
                    flush.zero_()
                    k.stats(idx, adv)
                    a.record()
                    k(idx, g.obs_pad[:, :D], act, lp, adv, ret, val)
                    b.record()
                torch.cuda.synchronize(dev)
            return sorted(a.elapsed_time(b) * 1e3 for a, b in evs)[len(evs) // 2], k.flat_grad.clone()
        us_fma, grad_fma = time_impl(0)
        us_tc1, _ = time_impl(1)
        us_tc, grad_tc = time_impl(2)
        used_tc = int(getattr(g, 'tensor_cores', 0))
        us = {0: us_fma, 1: us_tc1, 2: us_tc}[used_tc]
        fma_per_row = 2 * (2 * D * 64 + 3 * 64 * 64)   # both networks: forward 2 products, backward dH1 + dW2 + dW1
        grad_kernel = {'kernel': ('rk::ppo_mlp_grad_kernel', 'rk::ppo_mlp_grad_tc_kernel', 'rk::ppo_mlp_grad_tc2_kernel')[used_tc] +
                                 ' + rk::ppo_grad_reduce_kernel',
                       'rows_per_minibatch': mb, 'us_per_minibatch': us, 'fma_per_row': fma_per_row,
                       'tflops': 2.0 * fma_per_row * mb / (us * 1e-6) / 1e12,
                       'fp32_fma_kernel_us': us_fma, 'tcgen05_per_sample_products_only_us': us_tc1, 'tcgen05_kernel_us': us_tc,
                       'max_rel_diff_between_them': float((grad_tc - grad_fma).abs().max() /
                                                          grad_fma.abs().max().clamp_min(1e-30))}
        del flush
    roll = float(np.mean([t[0] for t in times]))
    upd = float(np.mean([t[1] for t in times]))
    per_it = [t[0] + t[1] for t in times]
    tot = torch.tensor([roll + upd], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    total_ms = float(tot.item())
    return {'sps': world * E * T / (total_ms * 1e-3), 'unit': 'learner transitions/s', 'envs_per_gpu': E, 'T': T,
            'timed_iterations': len(times), 'iteration_ms_min_max': [float(min(per_it)), float(max(per_it))],
            'rollout_ms': roll, 'gae_plus_update_ms': upd, 'optimizer_steps': int(times[-1][2]),
            'update_epochs': cfg['update_epochs'], 'num_minibatches': cfg['num_minibatches'],
            'kl_early_stop': 'disabled for timing', 'opponent_pool': cfg['pool_size'],
            'update_matmul_precision': args.ppo_precision,
            'rollout_launch_mode': getattr(trainer, 'rollout_mode', 'eager'),
            'update_products': ('fp32 FMA', 'per-sample products on tcgen05, tf32 x 3 terms (fp32 emulation, fp32 accumulate)',
                                'all 64-wide products incl. the weight gradients on tcgen05, tf32 x 3 terms (fp32 emulation, '
                                'fp32 accumulate)')[int(getattr(getattr(trainer, '_graphed', None), 'tensor_cores', 0))],
            'rollout_agent_steps_per_s': world * E * 2 * T / (roll * 1e-3), 'grad_kernel': grad_kernel}


def fp_peaks(torch, dev):
    """Live FMA-throughput micro-benchmarks: torch has no such kernel, so this
    uses a tiny dependent-chain FMA kernel compiled into librk_b200.so."""
    from self_play_racing_b200 import _lib
    lib = _lib.load()
    if not hasattr(lib, 'rk_fma_peak'):
        return None
    import ctypes as C
    lib.rk_fma_peak.restype = C.c_double
    lib.rk_fma_peak.argtypes = [C.c_int32, C.c_int32]
    return {'fp32_tflops': lib.rk_fma_peak(0, 4096), 'fp64_tflops': lib.rk_fma_peak(1, 2048)}


def pin_rank_to_cores(local, world):
    """One process per GPU, each on its own slice of the host cores: the Gymnasium face is host-paced (a numpy ->
    pinned copy, one C call, a stream synchronisation per step) and eight ranks on one shared core set slow each
    other down."""
    try:
        cores = sorted(os.sched_getaffinity(0))
        if world <= 1 or len(cores) < 2 * world or os.environ.get('RK_B200_NO_PIN') == '1':
            return None
        per = len(cores) // world
        mine = cores[local * per:(local + 1) * per]
        os.sched_setaffinity(0, mine)
        return [mine[0], mine[-1]]
    except Exception:  # noqa: BLE001
        return None


def main():
    if len(sys.argv) >= 2 and sys.argv[1] == '--ppo-cpu-baseline':   # child of ppo_cpu_baseline(): no GPU visible
        print(json.dumps(ppo_cpu_baseline(int(sys.argv[2]) if len(sys.argv) > 2 else 48)))
        return
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=500)
    ap.add_argument('--warmup', type=int, default=20)
    ap.add_argument('--workload', default='multi2_selfplay_65536', choices=sorted(WORKLOADS))
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--ref-kind', default='auto', choices=['auto', 'port'],
                    help='reference arm: the unmodified reference from oracle/_ref when present (auto) or the numpy port')
    ap.add_argument('--envs', type=int, default=None)
    ap.add_argument('--tracks', type=int, default=None, help='size of the procedural track pool (default 16)')
    ap.add_argument('--factor', type=int, default=None, help='waypoints per control point (reference: 30); S = 2 * n_ctrl * factor')
    ap.add_argument('--query', default='culled', choices=['culled', 'exact', 'grid'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-strong', action='store_true', help='skip the strong-scaled block (65,536 envs in total)')
    ap.add_argument('--ppo-updates', type=int, default=5,
                    help='also time N full self-play PPO iterations (rollout + GAE + update); 0 skips')
    ap.add_argument('--ppo-steps', type=int, default=64, help='rollout length T of the PPO timing')
    ap.add_argument('--ppo-precision', default='fp32', choices=['fp32', 'tf32'],
                    help='matmul precision of the PPO update (rollout kernels are always fp32/fp64)')
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.envs:
        wl['E'] = args.envs
    if args.tracks:
        wl['tracks'] = args.tracks
    if args.factor:
        wl['factor'] = args.factor
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        run_reference_arm(args, wl, rank, world)
        return

    pinned = pin_rank_to_cores(local, world)
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device; there is no CPU path (use --impl reference for the CPU arm)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)

    E, A, R = wl['E'], wl['A'], wl['R']
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    vec = make_vec(wl, E, dev, rank, args.query)
    D = vec.be.D
    weak = measure_step(args, wl, vec, E, dev, rank, world, flush, sample_clocks=True)
    ppo = None
    if args.ppo_updates > 0 and wl['selfplay']:
        ppo = time_ppo(args, wl, vec, E, dev, rank, world)
        ar = time_allreduce(dev, world)
        if ar:
            ppo['allreduce'] = ar
    vec.close()
    del vec

    # ---- the metric's literal configuration: 65,536 envs in TOTAL, split evenly over the ranks
    strong = None
    if not args.no_strong and E % world == 0:
        if world == 1:
            strong = {'envs_per_gpu': E, 'envs_total': E, 'note': 'N = 1: identical to the weak-scaled numbers of this line',
                      'value': weak['value'], 'ms_per_step': weak['ms_per_step'], 'kernel_ms': weak['kernel_ms'],
                      'e2e': weak['e2e'], 'ppo': ({k: ppo[k] for k in ('sps', 'rollout_ms', 'gae_plus_update_ms', 'T')} if ppo else None)}
        else:
            Es = E // world
            svec = make_vec(wl, Es, dev, rank, args.query)
            s = measure_step(args, wl, svec, Es, dev, rank, world, flush, sample_clocks=False)
            sppo = None
            if args.ppo_updates > 0 and wl['selfplay']:
                sppo = time_ppo(args, wl, svec, Es, dev, rank, world, with_grad_kernel=False)
            strong = {'envs_per_gpu': Es, 'envs_total': E, 'scaling': 'strong', 'value': s['value'],
                      'unit': 'agent-steps/s', 'ms_per_step': s['ms_per_step'], 'kernel_ms': s['kernel_ms'],
                      'e2e': s['e2e'], 'envs_per_warp': int(os.environ.get('RK_B200_EPW', 0)) or None,
                      'ppo': ({k: sppo[k] for k in ('sps', 'unit', 'envs_per_gpu', 'T', 'rollout_ms', 'gae_plus_update_ms',
                                                    'timed_iterations', 'iteration_ms_min_max', 'optimizer_steps')} if sppo else None)}
            svec.close()
            del svec

    if rank == 0:
        peaks_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
        if os.path.exists(peaks_path):
            hbm_peak, peak_src = json.load(open(peaks_path))['hbm_gbs'], 'measured (MEASURED_PEAKS.json hbm_gbs)'
        else:
            hbm_peak, peak_src = 6650.0, 'fallback (B200_PROFILING.md)'
        kms = weak['kernel_ms']
        agent_steps = E * A
        bytes_step = alg_bytes_per_agent_step(D) * agent_steps
        n_mean = 12 * wl['factor']  # n_ctrl in [10, 15) -> mean 12 control points
        flop_as = alg_flops_per_agent_step(R, 2 * n_mean, A, n_mean)
        eff_tflops = flop_as * agent_steps / (kms * 1e-3) / 1e12
        ach = bytes_step / (kms * 1e-3) / 1e9
        # per-launch DRAM traffic and issue statistics of the step kernel from the committed ncu capture of this command
        ncu = None
        for name in ('r02_step_kernel_ncu.json', 'r01_step_kernel_ncu.json'):
            path = os.path.join(ROOT, 'profiles', name)
            if os.path.exists(path) and E == WORKLOADS[args.workload]['E']:
                rec = json.load(open(path)).get(args.workload)
                if rec:
                    ncu = dict(rec, source='profiles/' + name)
                    break
        traffic = (ncu['dram_bytes_read'] + ncu['dram_bytes_write']) if ncu else None
        fp = fp_peaks(torch, dev)
        fp32_peak = fp['fp32_tflops'] if fp and fp.get('fp32_tflops') else 74.4
        # the step kernel is bound by the FP32/FP64 issue pipes (arithmetic intensity ~500 FLOP/B), so the
        # roofline line leads with that pipe; the HBM line the contract asks for follows as `hbm`
        roof = {'bound': 'fp32_pipe', 'achieved': eff_tflops, 'peak': fp32_peak, 'unit': 'TFLOP/s',
                'frac': eff_tflops / fp32_peak, 'traffic': traffic,
                'peak_source': ('builder-measured FFMA peak of this GPU (rk_fma_peak, dependent-chain FMA kernel in librk_b200.so); '
                                'MEASURED_PEAKS.json holds no FP32 figure' if fp else 'nominal 148 SMs x 128 lanes x 2 x 1.965 GHz'),
                'achieved_is': 'EFFECTIVE rate: brute-force FLOPs of the reference formulation (SURVEY 8d F) / kernel time; '
                               'the kernel culls, so this is not a pipe utilisation -- see `ncu` for the measured issue statistics',
                'kernel': 'rk::step_kernel', 'kernel_ms': kms, 'brute_force_flop_per_agent_step': flop_as,
                'measured_peaks': fp, 'ncu': ncu,
                'hbm': {'bound': 'hbm', 'achieved': ach, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': ach / hbm_peak,
                        'traffic': traffic, 'peak_source': peak_src,
                        'algorithmic_bytes_per_agent_step': alg_bytes_per_agent_step(D)}}
        if ppo and ppo.get('grad_kernel') and fp and fp.get('fp32_tflops'):
            ppo['grad_kernel']['frac_of_fp32_peak'] = ppo['grad_kernel']['tflops'] / fp['fp32_tflops']
        line = {'metric': 'agent_env_steps_per_sec', 'value': weak['value'],
                'unit': 'agent-steps/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
                'ms_per_step': weak['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': 'f64 state + f32 candidate search', 'data': 'synthetic',
                'config': {'workload': args.workload, 'envs_per_gpu': E, 'cars_per_env': A, 'rays': R,
                           'tracks': wl['tracks'], 'waypoints_per_track': f"{10 * wl['factor']}-{14 * wl['factor']}", 'query': args.query,
                           'autoreset': 'next_step', 'actions': 'uniform random, resident in HBM',
                           'opponent': 'frozen MLP snapshot (fused inference kernel)' if wl['selfplay'] else None,
                           'l2': 'flushed between timed steps (256 MiB memset outside the event pair)',
                           'rank_core_pinning': pinned},
                'roofline': roof, 'clocks': weak['clocks'], 'gpu_launches': weak['launches'], 'e2e': weak['e2e'],
                'ppo': ppo, 'strong': strong}
        if not args.no_cpu_baseline and world == 1:
            line['cpu_baseline'] = cpu_baseline(wl)
            if ppo is not None:
                ppo['cpu_baseline'] = ppo_cpu_baseline()
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
