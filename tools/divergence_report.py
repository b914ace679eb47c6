"""Free-running divergence of the CUDA path from the reference (golden
trajectories) and from the oracle, per step: max |difference| of the float64
car state, float32 observations and float64 rewards, plus the number of discrete
mismatches.  Writes profiles/r01_divergence.txt.  Run on the GPU box."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from oracle import racing_oracle as O
from self_play_racing_b200 import backend as B

lines = []
def emit(s):
    print(s); lines.append(s)

def load(name):
    with np.load(os.path.join(ROOT, 'tests', 'golden', name + '.npz')) as z:
        return {k: z[k] for k in z.files}

emit('# divergence of the CUDA path over free-running trajectories (identical actions, no teacher forcing)')
emit('# columns: case, query mode, steps, episodes, max|obs diff| (float32), max|reward diff|, max|state diff| (float64), discrete mismatches')
for name in ['single_default_10k', 'single_proc1', 'multi2_default', 'multi2_proc1', 'multi3_proc2']:
    g = load(name)
    multi = name.startswith('multi')
    A = g['actions'].shape[1] if multi else 1
    for query in ('exact', 'culled'):
        be = B.RacingBackend(1, kind='multi' if multi else 'single', num_agents=A, num_sensors=11, query=query)
        if multi:
            trk = O.TrackTables(g['control_points'], float(g['width']))
            be.set_tracks_from_waypoints([trk.waypoints], [float(g['width'])])
            be.reset(start_slot=torch.from_numpy(g['start_order0'].astype(np.int32)[None]).cuda())
        else:
            be.set_tracks_from_waypoints([g['waypoints']], [float(g['width'])])
            be.reset()
        n = len(g['actions'])
        acts = torch.from_numpy(g['actions']).cuda().reshape(n, A, 2)
        slots = torch.from_numpy(g['start_order'].astype(np.int32)).cuda() if multi else None
        d_obs = d_rew = d_st = 0.0
        mism = 0
        ended = g['terminated'] | g['truncated']
        for k in range(n):
            be.actions[0].copy_(acts[k])
            be.step(start_slot=slots[k:k + 1] if multi else None)
            obs = be.obs[0].cpu().numpy().reshape(A, -1)
            d_obs = max(d_obs, float(np.abs(obs - g['obs'][k].reshape(A, -1)).max()))
            d_rew = max(d_rew, float(np.abs(be.reward64[0].cpu().numpy() - np.atleast_1d(g['reward'][k])).max()))
            mism += int(bool(be.terminated[0]) != bool(g['terminated'][k])) + int(bool(be.truncated[0]) != bool(g['truncated'][k]))
            if k % 50 == 0 or k == n - 1:
                st = be.get_state()['car_f64'][0, :, :5]
                d_st = max(d_st, float(np.abs(st - g['state'][k].reshape(A, 5)).max()))
        emit(f'{name:20s} {query:7s} {n:6d} {int(ended.sum()):4d}  {d_obs:.3e}  {d_rew:.3e}  {d_st:.3e}  {mism}')
        be.close()

emit('')
emit('# batched, 256 envs x 400 steps over 8 procedural tracks against the live oracle (2-car)')
rs = np.random.RandomState(0)
cps = [O.gen_random_track(rs.randint(10, 15), rs.randint(50, 80), rs.randint(10, 20), rs.uniform(0.2, 0.7), rs.uniform(0.2, 0.7), rng=rs) for _ in range(8)]
widths = [float(rs.randint(6, 10)) for _ in range(8)]
tracks = O.make_pool(cps, widths)
E = 256
for query in ('exact', 'culled'):
    orc = O.OracleVecEnv(tracks, np.arange(E) % 8, kind='multi', num_agents=2, num_sensors=11, seed=1)
    be = B.RacingBackend(E, kind='multi', num_agents=2, num_sensors=11, query=query)
    be.set_tracks_from_waypoints([t.waypoints for t in tracks], widths, env_to_track=np.arange(E) % 8)
    so = orc._draw_start_order(E)
    orc.reset(start_order=so)
    be.reset(start_slot=torch.from_numpy(so.astype(np.int32)).cuda())
    per_step = []
    mism = 0
    rs2 = np.random.RandomState(3)
    for k in range(400):
        a = rs2.uniform(-1, 1, size=(E, 2, 2)).astype(np.float32); a[..., 1] = np.abs(a[..., 1])
        so = orc._draw_start_order(E)
        oobs, orew, ote, otr, _ = orc.step(a, start_order=so)
        be.actions.copy_(torch.from_numpy(a))
        be.step(start_slot=torch.from_numpy(so.astype(np.int32)).cuda())
        mism += int((be.terminated.cpu().numpy().astype(bool) != ote).sum())
        per_step.append((float(np.abs(be.obs.cpu().numpy() - oobs).max()), float(np.abs(be.reward64.cpu().numpy() - orew).max())))
    st = be.get_state()['car_f64'][..., :5]
    ost = np.stack([orc.x, orc.y, orc.angle, orc.vx, orc.vy], axis=2)
    po = np.array(per_step)
    emit(f'{query:7s} max|obs| by step quartile: ' + ' '.join(f'{po[i:i+100, 0].max():.2e}' for i in range(0, 400, 100)) +
         f'  max|reward| {po[:, 1].max():.2e}  final max|state| {np.abs(st - ost).max():.2e}  termination mismatches {mism}')
    be.close()
open(os.path.join(ROOT, 'gpurun_out', 'divergence.txt'), 'w').write('\n'.join(lines) + '\n')
