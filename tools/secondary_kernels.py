"""CUDA-event timings of the rollout-side kernels at the benchmark size, with their
algorithmic-byte / FLOP rates."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from self_play_racing_b200 import backend as B
from self_play_racing_b200.agent.ppo import Agent
from self_play_racing_b200 import spaces
import numpy as np

def timeit(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / n * 1e3  # us

E, T, D = 65536, 64, 19
r = torch.randn(T, E, device='cuda'); v = torch.randn(T, E, device='cuda'); d = torch.zeros(T, E, device='cuda')
nv = torch.randn(E, device='cuda'); nd = torch.zeros(E, device='cuda'); adv = torch.empty_like(r); ret = torch.empty_like(r)
us = timeit(lambda: B.gae(r, v, d, nv, nd, 0.99, 0.97, adv=adv, ret=ret))
print(f'gae_kernel T={T} E={E}: {us:.1f} us, {T*E*20/us/1e3:.0f} GB/s algorithmic (20 B/transition)')
ag = Agent(spaces.Box(-1, 1, (D,)), spaces.Box(np.array([-1., 0.]), np.array([1., 1.]), (2,)))
params = B.flatten_agent(ag.state_dict()).cuda()
obs = torch.rand(E, D, device='cuda') * 2 - 1; act = torch.zeros(E, 2, device='cuda'); lp = torch.zeros(E, device='cuda'); val = torch.zeros(E, device='cuda')
us = timeit(lambda: B.policy_act(params, obs, act, 1, 2))
print(f'policy_act_kernel actor only B={E}: {us:.1f} us, {E*10880/us/1e6:.1f} TFLOP/s fp32')
us = timeit(lambda: B.policy_act(params, obs, act, 1, 2, logprob=lp, value=val))
print(f'policy_act_kernel actor+critic B={E}: {us:.1f} us, {E*21632/us/1e6:.1f} TFLOP/s fp32')
n = 262144
bo = torch.rand(T * E, D, device='cuda'); ba = torch.rand(T * E, 2, device='cuda'); v1 = [torch.rand(T * E, device='cuda') for _ in range(4)]
idx = torch.randperm(T * E, device='cuda')[:n]
dst = (torch.empty(n, D, device='cuda'), torch.empty(n, 2, device='cuda')) + tuple(torch.empty(n, device='cuda') for _ in range(4))
us = timeit(lambda: B.gather_minibatch(idx, (bo, ba) + tuple(v1), dst))
print(f'gather_minibatch_kernel n={n}: {us:.1f} us, {n*(D+6)*8/us/1e3:.0f} GB/s (read+write)')
