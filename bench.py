#!/usr/bin/env python
"""Benchmark of the batched racing step path (BASELINE.json metric: agent
env-steps/s, 65,536 envs per GPU).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

One "step" = one pass of the hot path over all E environments of a rank: the
opponent-snapshot MLP inference kernel + the fused environment step kernel
(2-car workload), or the step kernel alone (single-car workload).  Rank 0 prints
ONE JSON line.  See DESIGN.md "Measurement" for the definitions used here.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
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
    # BASELINE configs[4] (one point of the sweep; the rest via --envs/--agents/--rays/--factor)
    'sweep_4car_64ray': dict(kind='multi', A=4, R=64, E=262144, tracks=16, factor=30, selfplay=False, width_lo=9.0),
}


def alg_bytes_per_agent_step(D):
    """state r/w (fp64 here) + action + obs + reward + flags, shared tracks."""
    return 2 * (5 * 8 + 4 + 16) + 8 + 4 * D + 4 + 8 + 3


def alg_flops_per_agent_step(R, S, A, N):
    """F = 13 R (S + 4(A-1)) + 25 N + 200 (reference formulation, brute force)."""
    return 13 * R * (S + 4 * (A - 1)) + 25 * N + 200


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop = index, [], threading.Event()

    def run(self):
        try:
            self.p = subprocess.Popen(['nvidia-smi', f'--id={self.index}', f'--query-gpu={self.Q}',
                                       '--format=csv,noheader,nounits', '-lms', '100'],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.p.stdout:
                self.rows.append([c.strip() for c in line.split(',')])
                if self._stop.is_set():
                    break
        except Exception:
            pass

    def stop(self):
        self._stop.set()
        try:
            self.p.terminate()
        except Exception:
            pass
        sm = [float(r[0]) for r in self.rows if r and r[0].replace('.', '').isdigit()]
        reasons = set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            for k, n in enumerate(names):
                if len(r) > 3 + k and r[3 + k].lower().startswith('active'):
                    reasons.add(n)
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace('.', '').isdigit()]
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': mx[0] if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


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


def cpu_baseline(wl, budget_envs=64, steps=12):
    """Single-process oracle port on a bounded sample of the same workload."""
    dt = _oracle_shard((wl['kind'], wl['A'], wl['R'], budget_envs, steps, 0))
    return {'value': budget_envs * wl['A'] * steps / dt, 'unit': 'agent-steps/s', 'cores': 1, 'kind': 'port',
            'sample': f'{budget_envs} envs x {steps} steps of the {wl["kind"]} workload, oracle/racing_oracle.py '
                      f'(batched numpy restatement; the unmodified per-env reference ran 617 steps/s single / '
                      f'240 agent-steps/s 2-car on one core at survey time)'}


_REF_ENV = None


def _ref_init(kind, A, R, E, seed_base):
    """Pool initializer: every worker process builds its own oracle batch once."""
    global _REF_ENV
    from oracle import racing_oracle as O
    seed = seed_base + os.getpid() % 1000
    rs = np.random.RandomState(seed)
    cps = [O.gen_random_track(rs.randint(10, 15), rs.randint(50, 80), rs.randint(10, 20), rs.uniform(0.2, 0.7),
                              rs.uniform(0.2, 0.7), rng=rs) for _ in range(4)]
    env = O.OracleVecEnv(O.make_pool(cps, [8.0, 7.0, 9.0, 8.0]), np.arange(E) % 4, kind=kind, num_agents=A,
                         num_sensors=R, seed=seed)
    env.reset()
    _REF_ENV = (env, rs)


def _ref_step(inner):
    env, rs = _REF_ENV
    for _ in range(inner):
        a = rs.uniform(-1, 1, size=(env.E, env.A, 2)).astype(np.float32)
        a[..., 1] = np.abs(a[..., 1])
        env.step(a)
    return env.E * env.A * inner


def _cpu_quota():
    try:
        q, per = open('/sys/fs/cgroup/cpu.max').read().split()
        return None if q == 'max' else float(q) / float(per)
    except Exception:
        return None


def run_reference_arm(args, wl, rank, world):
    """--impl reference: the reference's CPU path on all host cores.  The reference is
    Python under /root/reference and does not exist on the GPU box, so this times the
    oracle port (oracle/racing_oracle.py), one independent process per core, each
    stepping its own batch of environments -- the embarrassingly parallel CPU
    deployment of BASELINE.md section 3."""
    if rank != 0:
        return
    import multiprocessing as mp
    cores = len(os.sched_getaffinity(0))
    E_proc, inner = 32, 2
    ctx = mp.get_context('fork')
    with ctx.Pool(cores, initializer=_ref_init, initargs=(wl['kind'], wl['A'], wl['R'], E_proc, 100)) as pool:
        for _ in range(max(args.warmup, 1)):
            pool.map(_ref_step, [1] * cores, chunksize=1)
        t0 = time.perf_counter()
        units = 0
        for _ in range(args.steps):
            units += sum(pool.map(_ref_step, [inner] * cores, chunksize=1))
        wall = time.perf_counter() - t0
    val = units / wall
    sample = f'each step = {cores} processes x {E_proc} envs x {inner} env-steps of the {wl["kind"]} workload'
    line = {'impl': 'reference', 'metric': 'agent_env_steps_per_sec', 'value': val, 'unit': 'agent-steps/s',
            'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': 1e3 * wall / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': args.workload, 'sample': sample, 'cgroup_cpu_quota': _cpu_quota()},
            'cpu_baseline': {'value': val, 'unit': 'agent-steps/s', 'cores': cores, 'kind': 'port', 'sample': sample},
            'e2e': {'value': val, 'unit': 'agent-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------ GPU arm
def time_ppo(args, wl, vec, dev, rank, world):
    """Full self-play PPO iterations on the same batch of environments: device-resident
    rollout (policy + opponent inference + env step), GAE kernel, clipped-surrogate update
    (10 epochs x 16 minibatches, KL early stop disabled so that every optimizer step is
    paid for), NCCL all-reduce of gradients and minibatch statistics when world > 1.
    SPS = learner transitions per second = world * E * T / seconds per iteration."""
    import torch
    import torch.distributed as dist
    from self_play_racing_b200 import configs
    from self_play_racing_b200.agent import SelfPlayPPO
    E, T = wl['E'], args.ppo_steps
    cfg = configs.self_play_config(num_envs=E, num_steps=T, total_timesteps=10 ** 12, kl_target=1e9,
                                   update_matmul_precision=args.ppo_precision)
    trainer = SelfPlayPPO(vec, cfg, device=str(dev))
    trainer.opponent_pool = [trainer.snapshot_agent() for _ in range(cfg['pool_size'])]  # pool of 5 snapshots
    buf = trainer.alloc_buffers()
    buf['obs'][0].copy_(trainer._reset_all())
    ev = lambda: torch.cuda.Event(enable_timing=True)
    times = []
    for it in range(args.ppo_updates + 1):  # first iteration is warm-up
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
        if it > 0:
            times.append((e0.elapsed_time(e1), e1.elapsed_time(e2), steps))
    # the update's dominant kernel on its own: rk_ppo_minibatch_grad over the trainer's padded observation
    # buffer, CUDA events on the launching stream, L2 flushed between calls; both implementations of the entry
    # point (fp32 FMA kernel / tcgen05 TF32 x 3-pass kernel) on the same inputs
    grad_kernel = None
    g = getattr(trainer, '_graphed', None)
    if g is not None and getattr(g, 'fused_mlp', False) and getattr(g, 'obs_pad', None) is not None:
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
        us_fma, grad_fma = time_impl(False)
        us_tc, grad_tc = time_impl(True)
        used_tc = bool(getattr(g, 'tensor_cores', False))
        us = us_tc if used_tc else us_fma
        fma_per_row = 2 * (2 * D * 64 + 3 * 64 * 64)   # both networks: forward 2 products, backward dH1 + dW2 + dW1
        grad_kernel = {'kernel': ('rk::ppo_mlp_grad_tc_kernel' if used_tc else 'rk::ppo_mlp_grad_kernel') +
                                 ' + rk::ppo_grad_reduce_kernel',
                       'rows_per_minibatch': mb, 'us_per_minibatch': us, 'fma_per_row': fma_per_row,
                       'tflops': 2.0 * fma_per_row * mb / (us * 1e-6) / 1e12,
                       'fp32_fma_kernel_us': us_fma, 'tcgen05_kernel_us': us_tc,
                       'max_rel_diff_between_them': float((grad_tc - grad_fma).abs().max() /
                                                          grad_fma.abs().max().clamp_min(1e-30))}
        del flush
    roll = float(np.mean([t[0] for t in times]))
    upd = float(np.mean([t[1] for t in times]))
    tot = torch.tensor([roll + upd], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    total_ms = float(tot.item())
    return {'sps': world * E * T / (total_ms * 1e-3), 'unit': 'learner transitions/s', 'envs_per_gpu': E, 'T': T,
            'rollout_ms': roll, 'gae_plus_update_ms': upd, 'optimizer_steps': int(times[-1][2]),
            'update_epochs': cfg['update_epochs'], 'num_minibatches': cfg['num_minibatches'],
            'kl_early_stop': 'disabled for timing', 'opponent_pool': cfg['pool_size'],
            'update_matmul_precision': args.ppo_precision,
            'update_products': ('tcgen05 tf32 x 3 passes (fp32 emulation, fp32 accumulate)'
                                if getattr(getattr(trainer, '_graphed', None), 'tensor_cores', False) else 'fp32 FMA'),
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=500)
    ap.add_argument('--warmup', type=int, default=20)
    ap.add_argument('--workload', default='multi2_selfplay_65536', choices=sorted(WORKLOADS))
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--envs', type=int, default=None)
    ap.add_argument('--tracks', type=int, default=None, help='size of the procedural track pool (default 16)')
    ap.add_argument('--factor', type=int, default=None, help='waypoints per control point (reference: 30); S = 2 * n_ctrl * factor')
    ap.add_argument('--query', default='culled', choices=['culled', 'exact'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--ppo-updates', type=int, default=2,
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

    import torch
    import torch.distributed as dist
    from self_play_racing_b200 import _lib
    from self_play_racing_b200.backend import flatten_agent
    from self_play_racing_b200.environment.vec_env import BatchedRacingVecEnv
    from self_play_racing_b200.agent.ppo import Agent

    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device; there is no CPU path (use --impl reference for the CPU arm)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)

    E, A, R = wl['E'], wl['A'], wl['R']
    vec = BatchedRacingVecEnv.synthetic(wl['kind'], E, n_tracks=wl['tracks'], num_agents=A, num_sensors=R,
                                        selfplay=wl['selfplay'], device=dev, query=args.query, seed=1000 + rank,
                                        copy=False, factor=wl['factor'], width_lo=wl.get('width_lo', 6.0))
    be = vec.be
    D = be.D
    if wl['selfplay']:
        # opponent snapshot: orthogonal-init Agent, manual_seed(1), log_std -0.3
        torch.manual_seed(1)
        opp = Agent(vec.single_observation_space, vec.single_action_space)
        opp.log_std.data.fill_(-0.3)
        vec.set_opponent(flatten_agent(opp.state_dict()).to(dev))
    # learner / all-car actions resident in HBM: 8 pre-drawn uniform tensors cycled through
    g = torch.Generator(device=dev)
    g.manual_seed(7 + rank)
    n_act = 8
    if wl['selfplay']:
        act_pool = torch.rand(n_act, E, 2, device=dev, generator=g) * 2 - 1
    else:
        act_pool = torch.rand(n_act, *be.actions.shape, device=dev, generator=g) * 2 - 1
    act_pool[..., 1] = act_pool[..., 1].abs()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

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

    # ---- device-resident throughput: K steps, each bracketed by CUDA events, L2 flushed between
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sampler = ClockSampler(local)
    sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
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
    ms = np.array([a.elapsed_time(b) for a, b in ev])
    total_ms = float(ms.sum())
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
    clocks = sampler.stop()

    # ---- end to end through the Gymnasium face: host numpy in, host numpy out
    rs = np.random.RandomState(3 + rank)
    host_actions = [rs.uniform(-1, 1, size=(E, 2)).astype(np.float32) for _ in range(4)]
    for a in host_actions:
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
            obs, rew, term, trunc, infos = vec.step(host_actions[k % 4])
        torch.cuda.synchronize(dev)
        wall = time.perf_counter() - t0
        tw = torch.tensor([wall], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tw, op=dist.ReduceOp.MAX)
        e2e = {'value': E * A * args.steps * world / float(tw.item()), 'unit': 'agent-steps/s',
               'h2d_bytes_per_step': int(vec.h2d_bytes_per_step), 'd2h_bytes_per_step': int(vec.d2h_bytes_per_step),
               'ms_per_step': 1e3 * float(tw.item()) / args.steps,
               'transfer': ('the step kernel reads the actions from and writes the results into the pinned host buffers '
                            '(zero-copy over PCIe, same bytes)' if getattr(vec, 'host_chunks', 0) > 0 and
                            int(os.environ.get('RK_B200_ZEROCOPY_OBS', '7')) & 1 else 'cudaMemcpyAsync from/to pinned host buffers')}

    ppo = None
    if args.ppo_updates > 0 and wl['selfplay']:
        ppo = time_ppo(args, wl, vec, dev, rank, world)

    if rank == 0:
        peaks_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
        if os.path.exists(peaks_path):
            hbm_peak, peak_src = json.load(open(peaks_path))['hbm_gbs'], 'measured (MEASURED_PEAKS.json hbm_gbs)'
        else:
            hbm_peak, peak_src = 6650.0, 'fallback (B200_PROFILING.md)'
        agent_steps = E * A
        bytes_step = alg_bytes_per_agent_step(D) * agent_steps
        n_mean = 12 * wl['factor']  # n_ctrl in [10, 15) -> mean 12 control points
        flops_step = alg_flops_per_agent_step(R, 2 * n_mean, A, n_mean) * agent_steps
        ach = bytes_step / (kms * 1e-3) / 1e9
        traffic = None  # dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed ncu capture
        ncu_path = os.path.join(ROOT, 'profiles', 'r01_step_kernel_ncu.json')
        if os.path.exists(ncu_path) and E == WORKLOADS[args.workload]['E']:
            rec = json.load(open(ncu_path)).get(args.workload)
            if rec:
                traffic = rec['dram_bytes_read'] + rec['dram_bytes_write']
        roof = {'bound': 'hbm', 'achieved': ach, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': ach / hbm_peak,
                'traffic': traffic, 'peak_source': peak_src, 'kernel': 'rk::step_kernel', 'kernel_ms': kms,
                'algorithmic_bytes_per_agent_step': alg_bytes_per_agent_step(D),
                'note': 'the step kernel is FP32/FP64-pipe bound, not HBM bound (SURVEY 8d); see fp_pipe'}
        fp = fp_peaks(torch, dev)
        roof['fp_pipe'] = {'brute_force_flop_per_agent_step': alg_flops_per_agent_step(R, 2 * n_mean, A, n_mean),
                           'effective_tflops': flops_step / (kms * 1e-3) / 1e12, 'measured_peaks': fp}
        if fp and fp.get('fp32_tflops'):
            roof['fp_pipe']['frac_of_fp32_peak'] = roof['fp_pipe']['effective_tflops'] / fp['fp32_tflops']
        if ppo and ppo.get('grad_kernel') and fp and fp.get('fp32_tflops'):
            ppo['grad_kernel']['frac_of_fp32_peak'] = ppo['grad_kernel']['tflops'] / fp['fp32_tflops']
        line = {'metric': 'agent_env_steps_per_sec', 'value': agent_steps * args.steps * world / (total_ms_max * 1e-3),
                'unit': 'agent-steps/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
                'ms_per_step': total_ms_max / args.steps, 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': 'f64 state + f32 candidate search', 'data': 'synthetic',
                'config': {'workload': args.workload, 'envs_per_gpu': E, 'cars_per_env': A, 'rays': R,
                           'tracks': wl['tracks'], 'waypoints_per_track': f"{10 * wl['factor']}-{14 * wl['factor']}", 'query': args.query,
                           'autoreset': 'next_step', 'actions': 'uniform random, resident in HBM',
                           'opponent': 'frozen MLP snapshot (fused inference kernel)' if wl['selfplay'] else None,
                           'l2': 'flushed between timed steps (256 MiB memset outside the event pair)'},
                'roofline': roof, 'clocks': clocks, 'gpu_launches': int(launches), 'e2e': e2e, 'ppo': ppo}
        if not args.no_cpu_baseline:
            line['cpu_baseline'] = cpu_baseline(wl)
        print(json.dumps(line))
    if ppo is None:
        vec.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
