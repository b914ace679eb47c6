"""Times the fused PPO minibatch-gradient kernel (rk_ppo_minibatch_grad) alone:
262,144-row minibatches drawn by a random permutation from a 64 x 65,536 rollout
buffer (the bench's PPO shape), CUDA events around each call, L2 flushed between."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from self_play_racing_b200 import spaces
from self_play_racing_b200.agent.ppo import Agent
from self_play_racing_b200.backend import PpoMinibatchGrad
from self_play_racing_b200 import _lib

D, B, n = 19, 64 * 65536, 262144
torch.manual_seed(0)
agent = Agent(spaces.Box(-1, 1, (D,)), spaces.Box(-1, 1, (2,))).cuda()
g = torch.Generator(device='cuda').manual_seed(0)
obs = torch.rand(B, D, device='cuda', generator=g) * 2 - 1
act = torch.rand(B, 2, device='cuda', generator=g) * 2 - 1
adv, val, ret, logp = (torch.randn(B, device='cuda', generator=g) for _ in range(4))
perm = torch.randperm(B, device='cuda', generator=g)
tc = int(os.environ.get('RK_TC', '0'))   # 0 FMA kernel, 1 per-sample products on tcgen05, 2 weight gradients too
fused = PpoMinibatchGrad(list(agent.parameters()), agent.log_std, D, 0.2, 0.5, tensor_cores=tc)
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(16)]
for rep in range(2):
    l0 = _lib.launch_count()
    for k in range(16):
        idx = perm[k * n:(k + 1) * n]
        flush.zero_()
        fused.stats(idx, adv)
        ev[k][0].record()
        fused(idx, obs, act, logp, adv, ret, val)
        ev[k][1].record()
    torch.cuda.synchronize()
ms = sorted(a.elapsed_time(b) for a, b in ev)
med = ms[len(ms) // 2]
flop = 2.0 * n * 2 * (2 * D * 64 + 3 * 64 * 64)   # per net: forward 2 products, backward dH1 + dW2 + dW1 (no dX)
print(f'rk_ppo_minibatch_grad{" [tensor cores, variant %d]" % tc if tc else ""} n={n} D={D}: median {med * 1e3:.1f} us (min {ms[0] * 1e3:.1f}), '
      f'{flop / med / 1e9:.1f} TFLOP/s fp32 (hidden layers), {n / med / 1e3:.1f} M rows/s; launches per call '
      f'{(_lib.launch_count() - l0) // 16 - 1}')
