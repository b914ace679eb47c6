import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
VARIANT = int(os.environ.get('RK_TC', '1'))
import torch
from self_play_racing_b200 import spaces
from self_play_racing_b200.agent.ppo import Agent
from self_play_racing_b200.backend import PpoMinibatchGrad
torch.manual_seed(0)
D = 19
agent = Agent(spaces.Box(-1, 1, (D,)), spaces.Box(-1, 1, (2,))).cuda()
with torch.no_grad():
    for p in agent.parameters(): p.add_(0.2 * torch.randn_like(p))
    agent.log_std.fill_(-0.7)
g = torch.Generator(device='cuda').manual_seed(1)
for n in (128, 1000, 5000, 40000):
    B = 2 * n
    obs = torch.rand(B, D, device='cuda', generator=g) * 2 - 1
    act = torch.rand(B, 2, device='cuda', generator=g) * 2 - 1
    adv = torch.randn(B, device='cuda', generator=g); val = torch.randn(B, device='cuda', generator=g)
    ret = val + 0.3 * torch.randn(B, device='cuda', generator=g)
    with torch.no_grad(): _, logp, _, _ = agent.get_action_and_value(obs, act)
    logp = logp + 0.15 * torch.randn(B, device='cuda', generator=g)
    idx = torch.randperm(B, device='cuda', generator=g)[:n]
    ref = PpoMinibatchGrad(list(agent.parameters()), agent.log_std, D, 0.2, 0.5)
    ref.stats(idx, adv); f0, k0 = ref(idx, obs, act, logp, adv, ret, val); f0 = f0.clone(); k0 = float(k0)
    tc = PpoMinibatchGrad(list(agent.parameters()), agent.log_std, D, 0.2, 0.5, tensor_cores=VARIANT)
    tc.stats(idx, adv); f1, k1 = tc(idx, obs, act, logp, adv, ret, val)
    torch.cuda.synchronize()
    d = (f1 - f0).abs()
    rel = float(d.max() / f0.abs().max())
    gmax = f0.abs().max()
    blocks = {'W1': (0, 64 * D), 'b1': (64 * D, 64 * D + 64), 'W2': (64 * D + 64, 64 * D + 64 + 4096), 'b2': (64 * D + 64 + 4096, 64 * D + 128 + 4096)}
    per = ' '.join(f'{k} {float(d[a:b].max()):.1e}' for k, (a, b) in blocks.items())
    print(f'variant {VARIANT} n={n}: [{per}] max |tc - fma| = {float(d.max()):.3e} (max |grad| {float(f0.abs().max()):.3e}, rel {rel:.2e}); kl {k0:.6f} vs {float(k1):.6f}; nan={bool(torch.isnan(f1).any())}')
