"""Multi-GPU check of the device-resident PPO update (run under torchrun, one rank per GPU):

  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_update_check.py

Every rank updates on its shard of one synthetic batch through the fused path
(rk_ppo_minibatch_grad + one NCCL all-reduce per minibatch + rk_ppo_adam_step).  Rank 0 then
repeats the update in ONE process with eager autograd on the concatenated batch, with the
permutation that makes global minibatch k the union of the ranks' local minibatches k, and
compares parameters (env sharding + all-reduced gradients == one big batch)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from self_play_racing_b200 import configs, spaces
from self_play_racing_b200.agent.ppo import PPO, Agent


def make_ppo(cfg, dev, world, rank, seed=3):
    ppo = PPO.__new__(PPO)
    ppo.config, ppo.device, ppo.world, ppo.rank, ppo._perm_gen, ppo._graphed = cfg, dev, world, rank, None, None
    torch.manual_seed(seed)
    ppo.agent = Agent(spaces.Box(-1, 1, (19,)), spaces.Box(-1, 1, (2,))).to(dev)
    ppo.agent.log_std.data.fill_(-0.5)
    ppo.optimizer = ppo._make_optimizer()
    return ppo


def batch(n, dev, seed=0):
    g = torch.Generator().manual_seed(seed)
    obs = torch.rand(n, 19, generator=g) * 2 - 1
    act = torch.rand(n, 2, generator=g) * 2 - 1
    adv, val = torch.randn(n, generator=g), torch.randn(n, generator=g)
    noise = 0.05 * torch.randn(n, generator=g)
    return [t.to(dev) for t in (obs, act, adv, val, val + adv, noise)]


def main():
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    dev = torch.device('cuda', int(os.environ.get('LOCAL_RANK', 0)))
    torch.cuda.set_device(dev)
    dist.init_process_group('nccl', device_id=dev)
    half, nmb, epochs = 8192, 4, 2
    n = half * world
    for kl_target, expect in ((1e9, epochs * nmb), (1e-9, 0)):
        cfg = configs.self_play_config(num_envs=8, num_steps=32, update_epochs=epochs, num_minibatches=nmb, kl_target=kl_target)
        obs, act, adv, val, ret, noise = batch(n, dev)
        ppo = make_ppo(cfg, dev, world, rank)
        with torch.no_grad():
            _, logp, _, _ = ppo.agent.get_action_and_value(obs, act)
        logp = logp + noise
        perms = [torch.randperm(half, generator=torch.Generator().manual_seed(100 + ep)).to(dev) for ep in range(epochs)]
        sl = slice(rank * half, (rank + 1) * half)
        before = [p.detach().clone() for p in ppo.agent.parameters()]
        steps = ppo.ppo_update(adv[sl], ret[sl], val[sl], logp[sl], act[sl], obs[sl], permutation=lambda ep: perms[ep])
        assert ppo._graphed.fused_mlp and ppo._graphed.adam is not None, 'fused path not taken'
        mine = torch.cat([p.detach().reshape(-1) for p in ppo.agent.parameters()])
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        for other in gathered[1:]:
            assert torch.equal(gathered[0], other), 'ranks diverged'
        assert steps == expect, (steps, expect)
        if expect == 0:
            assert all(torch.equal(a, b) for a, b in zip(before, ppo.agent.parameters())), 'KL stop applied a step'
        if rank == 0 and expect:
            mb = half // nmb

            def global_perm(ep):
                chunks = []
                for s in range(0, half, mb):
                    chunks += [perms[ep][s:s + mb] + r * half for r in range(world)]
                return torch.cat(chunks)
            ecfg = dict(cfg, cuda_graph_update=False)
            one = make_ppo(ecfg, dev, 1, 0)
            with torch.no_grad():
                _, logp1, _, _ = one.agent.get_action_and_value(obs, act)
            s1 = one.ppo_update(adv, ret, val, logp1 + noise, act, obs, permutation=global_perm)
            assert s1 == steps
            ref = torch.cat([p.detach().reshape(-1) for p in one.agent.parameters()])
            torch.testing.assert_close(mine, ref, rtol=1e-4, atol=2e-6)
            print(f'world {world}: fused data-parallel update == single-process eager update on the union '
                  f'({steps} steps, max |diff| {float((mine - ref).abs().max()):.2e}); KL stop next')
    if rank == 0:
        print('KL stop: no step applied on any rank; OK')
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
