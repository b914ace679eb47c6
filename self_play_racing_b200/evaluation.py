"""Batched evaluation: the protocol of the reference's evaluate.py (tracks x runs
episodes, `evaluate.py:10-64,66-120,173-183`) and utils/metrics.py
(`eval_single_agent` :39-78, `eval_multi_agent` :80-150) as ONE device batch --
every (track, run) episode is an environment of the same launch -- returning the
reference's result dictionaries so that `display_comparison` works unchanged.

The per-environment loops of utils/metrics.py also run as they are against the
`RacingEnv` / `MultiRacingEnv` facades of this package; this module is the fast path.
"""
from __future__ import annotations

import numpy as np
import torch

from .backend import RacingBackend, flatten_agent, policy_act


def _aggregate(all_metrics):
    """evaluate.py:40-64 (same keys, same formulas)."""
    ok = [m for m in all_metrics if m['finished']]
    eff = [m['steps'] / m['progress'] for m in all_metrics if m['progress'] > 0.01]
    mean = lambda k: float(np.mean([m[k] for m in ok])) if ok else 0
    return {
        'num_episodes': len(all_metrics), 'num_successful': len(ok),
        'success_rate': len(ok) / len(all_metrics),
        'crash_rate': sum(m['crashed'] for m in all_metrics) / len(all_metrics),
        'avg_steps': mean('steps'), 'avg_reward': mean('total_reward'), 'avg_progress': mean('progress'),
        'avg_speed': mean('speed'), 'avg_distance': mean('total_distance'),
        'avg_steps_per_progress': float(np.mean(eff)) if eff else float('nan'),
        'all_episodes': all_metrics,
    }


def evaluate_batched(kind, agent, track_pool, track_widths, num_tracks=20, num_runs=10, max_steps=None,
                     num_sensors=11, device=None, seed=0, query='culled', start_slot=None):
    """All num_tracks x num_runs episodes of `evaluate_single_agent_overall` /
    `evaluate_multi_agent_overall` at once.  Episode (t, r) runs on track t with
    width track_widths[r] (the reference indexes widths by run);
    in the 2-car evaluation BOTH cars are driven by `agent` (utils/metrics.py:94-106).
    Actions are sampled (no deterministic mode in the reference, quirk 13)."""
    multi = kind == 'multi'
    A = 2 if multi else 1
    max_steps = max_steps or (3000 if multi else 2000)
    E = num_tracks * num_runs
    be = RacingBackend(E, kind=kind, num_agents=A, num_sensors=num_sensors, device=device, autoreset='disabled',
                       query=query, seed=seed, agent_major=True, want_info=True)
    # one device track per (track, width) pair that occurs
    keys, cps, widths, e2t = {}, [], [], np.zeros(E, dtype=np.int32)
    for t in range(num_tracks):
        for r in range(num_runs):
            k = (t, float(track_widths[r]))
            if k not in keys:
                keys[k] = len(cps)
                cps.append(np.asarray(track_pool[t], dtype=np.float64))
                widths.append(float(track_widths[r]))
            e2t[t * num_runs + r] = keys[k]
    be.set_tracks_from_control_points(cps, widths, env_to_track=e2t)
    dev = be.device
    params = flatten_agent(agent.state_dict()).to(dev)
    if start_slot is not None:   # the grid slots of each episode's reset (int [E, A]); default: the backend's Philox shuffle
        start_slot = torch.as_tensor(np.ascontiguousarray(start_slot, dtype=np.int32)).to(dev)
    obs = be.reset(start_slot=start_slot)                        # [A, E, D]
    alive = torch.ones(E, dtype=torch.bool, device=dev)
    total_reward = torch.zeros(A, E, dtype=torch.float64, device=dev)
    distance = torch.zeros(A, E, dtype=torch.float64, device=dev)
    steps = torch.zeros(E, dtype=torch.int32, device=dev)
    final_f = torch.zeros(A, E, 5, dtype=torch.float64, device=dev)
    final_i = torch.zeros(A, E, 4, dtype=torch.int32, device=dev)
    prev_pos = None
    obs_flat, act_flat = obs.view(A * E, -1), be.actions.view(A * E, 2)
    for k in range(max_steps):
        policy_act(params, obs_flat, act_flat, seed=seed, counter=k + 1)   # every car of every episode, one launch
        be.step()
        a2 = alive[None, :]
        total_reward += torch.where(a2, be.reward64, torch.zeros_like(be.reward64))
        pos = be.info_f64[..., :2]
        if prev_pos is not None:
            distance += torch.where(a2, (pos - prev_pos).norm(dim=-1), torch.zeros_like(distance))
        prev_pos = pos.clone()
        steps += alive.to(torch.int32)
        done_now = alive & be.done.bool()
        final_f = torch.where((alive[None, :, None]), be.info_f64, final_f)   # keep the last live step's info
        final_i = torch.where((alive[None, :, None]), be.info_i32, final_i)
        alive = alive & ~done_now
        if k % 64 == 63 and not bool(alive.any()):
            break
    f, i = final_f.cpu().numpy(), final_i.cpu().numpy()
    tr, ds, st = total_reward.cpu().numpy(), distance.cpu().numpy(), steps.cpu().numpy()
    be.close()
    out = []
    for e in range(E):
        c = 0
        if multi and not i[0, e, 1] and i[1, e, 1]:
            c = 1  # utils/metrics.py:124-135: report the car that finished, car 0 otherwise
        m = {'total_reward': float(tr[c, e]), 'steps': int(st[e]), 'progress': float(f[c, e, 3]),
             'finished': bool(i[c, e, 1]), 'crashed': bool(i[c, e, 0]), 'speed': float(f[c, e, 2]),
             'total_distance': float(ds[c, e]),
             'distance_per_step': float(ds[c, e] / st[e]) if st[e] > 1 else 0}
        if multi:
            m['placement'] = int(i[c, e, 2]) if i[c, e, 2] else None
        out.append(m)
    return _aggregate(out)
