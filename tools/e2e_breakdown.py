"""Where does the Gymnasium-face step time go?  Host-side timers around the phases of vec.step()."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from self_play_racing_b200.environment.vec_env import BatchedRacingVecEnv
from self_play_racing_b200.agent.ppo import Agent
from self_play_racing_b200.backend import flatten_agent
import ctypes as C
from self_play_racing_b200 import _lib

E = 65536
vec = BatchedRacingVecEnv.synthetic('multi', E, n_tracks=16, num_agents=2, selfplay=True, copy=False)
torch.manual_seed(1)
opp = Agent(vec.single_observation_space, vec.single_action_space)
vec.set_opponent(flatten_agent(opp.state_dict()).cuda())
vec.reset()
rs = np.random.RandomState(0)
acts = [rs.uniform(-1, 1, size=(E, 2)).astype(np.float32) for _ in range(4)]
for k in range(20):
    vec.step(acts[k % 4])
be = vec.be
T = {'prep': 0.0, 'call': 0.0, 'post': 0.0}
n = 300
for k in range(n):
    t0 = time.perf_counter()
    vec._h_actions.numpy()[...] = acts[k % 4]
    t1 = time.perf_counter()
    vec._step_host(None)
    t2 = time.perf_counter()
    term = vec._np_terminated | vec._np_truncated
    any_ep = vec._np_ep_mask.any()
    t3 = time.perf_counter()
    T['prep'] += t1 - t0; T['call'] += t2 - t1; T['post'] += t3 - t2
print({k: round(v / n * 1e6, 1) for k, v in T.items()}, 'us per step')
# the C call alone with different chunk counts
for ch in (1, 2, 4, 8):
    vec._host_io.n_chunks = ch
    for _ in range(10): vec._step_host(None)
    t0 = time.perf_counter()
    for _ in range(200): vec._step_host(None)
    print('chunks', ch, 'rk_step_host us', round((time.perf_counter() - t0) / 200 * 1e6, 1))
# device-only step for reference
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(200):
    vec._opponent_act(); be.step()
torch.cuda.synchronize(); print('device-only us', round((time.perf_counter() - t0) / 200 * 1e6, 1))
