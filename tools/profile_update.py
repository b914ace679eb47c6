"""Where does the PPO update's time go?  torch.profiler over a few minibatch steps."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from self_play_racing_b200 import configs
from self_play_racing_b200.agent import SelfPlayPPO
from self_play_racing_b200.environment.vec_env import BatchedRacingVecEnv

E, T = int(sys.argv[1]) if len(sys.argv) > 1 else 65536, 64
vec = BatchedRacingVecEnv.synthetic('multi', E, n_tracks=16, num_agents=2, selfplay=True, copy=False)
cfg = configs.self_play_config(num_envs=E, num_steps=T, total_timesteps=10 ** 12, kl_target=1e9, update_epochs=1)
tr = SelfPlayPPO(vec, cfg, device='cuda')
buf = tr.alloc_buffers()
buf['obs'][0].copy_(tr._reset_all())
tr.update_opponent(); tr._anneal(0, 100)
tr.collect_rollout(buf); tr._learn_from(buf)
torch.cuda.synchronize()
tr.collect_rollout(buf)
torch.cuda.synchronize()
t0 = time.perf_counter()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    tr._learn_from(buf)
    torch.cuda.synchronize()
print('wall ms for 16 minibatch steps', 1e3 * (time.perf_counter() - t0))
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=25, max_name_column_width=60))
