"""Small end-to-end case for compute-sanitizer (memcheck / racecheck): both env
kinds, both query modes, resets, the policy and GAE kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from self_play_racing_b200 import backend as B

torch.manual_seed(0)
for kind, A in (('single', 1), ('multi', 2), ('multi', 3)):
    for query in ('exact', 'culled'):
        be = B.RacingBackend(50, kind=kind, num_agents=A, num_sensors=11, query=query, seed=1)
        be.generate_tracks(seed=2, n_tracks=5)
        be.reset()
        for k in range(30):
            be.actions.uniform_(-1, 1)
            be.actions[..., 1].abs_()
            be.step()
        torch.cuda.synchronize()
        print(kind, A, query, 'episodes ended so far:', float(be.ep_stats[2]))
        be.close()
from self_play_racing_b200.agent.ppo import Agent
from self_play_racing_b200 import spaces
import numpy as np
ag = Agent(spaces.Box(-1, 1, (19,)), spaces.Box(np.array([-1., 0.]), np.array([1., 1.]), (2,)))
params = B.flatten_agent(ag.state_dict()).cuda()
obs = torch.rand(300, 19, device='cuda'); act = torch.zeros(300, 2, device='cuda')
lp = torch.zeros(300, device='cuda'); val = torch.zeros(300, device='cuda')
B.policy_act(params, obs, act, 1, 2, logprob=lp, value=val)
B.policy_act(None, None, act, 1, 3)
r = torch.randn(16, 300, device='cuda'); v = torch.randn(16, 300, device='cuda'); d = torch.zeros(16, 300, device='cuda')
B.gae(r, v, d, torch.randn(300, device='cuda'), torch.zeros(300, device='cuda'), 0.99, 0.97)
torch.cuda.synchronize()
print('sanitize case done')
