"""`Agent` and `PPO` with the reference's constructor and method surface
(agent/ppo.py:11-293), re-designed so that the rollout never leaves the device:

  * the vector env is ONE `BatchedRacingVecEnv` (one fused kernel per step);
  * the per-step policy forward + sampling + log-prob + value is one fused
    kernel (`rk_policy_act`) reading the observation the step kernel just wrote;
  * the step kernel writes observation, reward and done straight into the
    [T, ...] rollout buffers (no per-step copies, no host synchronisation);
  * GAE is one backward-scan kernel (`rk_gae`);
  * the clipped-surrogate update stays PyTorch autograd on the device, with an
    optional NCCL all-reduce of the flat gradient and of the three global
    minibatch statistics when torch.distributed is initialised (environments
    shard across GPUs; nothing else is communicated).
"""
from __future__ import annotations

import random

import numpy as np
import torch
import torch.nn as nn
import torch.optim as optim

from ..backend import (PPO_MAX_OBS_DIM, PpoAdamStep, PpoMinibatchGrad, random_permutation, flatten_agent, gae as gae_kernel, gather_minibatch,
                       policy_act, ppo_loss_grad)
from ..environment.vec_env import BatchedRacingVecEnv


class _SplitKLinearFn(torch.autograd.Function):
    """y = x W^T + b whose weight/bias gradients are reduced in row chunks.

    A PPO minibatch here has ~10^5 rows and 64 columns, so dW = dY^T X is a
    64x64 output with a 10^5-long reduction: a plain GEMM maps it onto one or two
    thread blocks (measured 279 us per layer, profiles/r01_ppo_update_profile.txt).
    Splitting the rows into chunks turns it into a batched GEMM that fills the
    GPU, followed by a tiny sum over chunks."""
    CHUNK = 2048

    @staticmethod
    def forward(ctx, x, w, b):
        ctx.save_for_backward(x, w)
        return torch.addmm(b, x, w.t())

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dx = dy @ w if ctx.needs_input_grad[0] else None
        n, m = x.shape[0], _SplitKLinearFn.CHUNK
        c = n // m
        xa, dya = x[:c * m].view(c, m, -1), dy[:c * m].view(c, m, -1)
        dw = torch.bmm(dya.transpose(1, 2), xa).sum(0)
        db = dya.sum(1).sum(0)
        if c * m < n:
            dw = dw + dy[c * m:].t() @ x[c * m:]
            db = db + dy[c * m:].sum(0)
        return dx, dw, db


class _Linear(nn.Linear):
    """nn.Linear (same parameters, same state_dict keys) that switches to the
    chunked-reduction backward for tall training batches."""

    def forward(self, x):
        if x.dim() == 2 and x.shape[0] >= 4 * _SplitKLinearFn.CHUNK and torch.is_grad_enabled() and self.weight.requires_grad:
            return _SplitKLinearFn.apply(x, self.weight, self.bias)
        return super().forward(x)


def _ortho(layer, std=np.sqrt(2), bias=0.0):
    torch.nn.init.orthogonal_(layer.weight, std)
    torch.nn.init.constant_(layer.bias, bias)
    return layer


class Agent(nn.Module):
    """Two separate 64-64 tanh MLPs and a state-independent log_std *buffer*;
    state_dict keys match the reference (agent/ppo.py:11-37) so its checkpoints
    load here and vice versa."""

    def __init__(self, obs_space, action_space):
        super().__init__()
        obs_dim = int(np.array(obs_space.shape).prod())
        action_dim = action_space.shape[0]
        self.actor_mu = nn.Sequential(
            _ortho(_Linear(obs_dim, 64)), nn.Tanh(),
            _ortho(_Linear(64, 64)), nn.Tanh(),
            _ortho(_Linear(64, action_dim), std=0.01), nn.Tanh())
        self.register_buffer('log_std', torch.zeros(action_dim))
        self.critic = nn.Sequential(
            _ortho(_Linear(obs_dim, 64)), nn.Tanh(),
            _ortho(_Linear(64, 64)), nn.Tanh(),
            _ortho(_Linear(64, 1), std=1.0))

    def get_value(self, obs):
        return self.critic(obs)

    def get_action_and_value(self, obs, action=None):
        mu = self.actor_mu(obs)
        # validate_args=False: the argument checks synchronise with the host, which CUDA-graph capture forbids
        dist = torch.distributions.Normal(mu, torch.exp(self.log_std).expand_as(mu), validate_args=False)
        if action is None:
            action = torch.clamp(dist.sample(), -1.0, 1.0)
        return action, dist.log_prob(action).sum(-1), dist.entropy().sum(-1), self.critic(obs)


class _GraphedMinibatch:
    """Static buffers + two captured CUDA graphs for one PPO minibatch step."""

    def __init__(self, ppo, mb, obs_dim):
        dev, c = ppo.device, ppo.config
        self.mb = mb
        z = lambda *shape, dt=torch.float32: torch.zeros(*shape, device=dev, dtype=dt)
        self.obs, self.act = z(mb, obs_dim), z(mb, 2)
        self.old_logp, self.adv, self.ret, self.val = z(mb), z(mb), z(mb), z(mb)
        self.adv_mean, self.adv_std = z(()), torch.ones((), device=dev)
        self.kl_sum = z((), dt=torch.float64)
        params = [p for p in ppo.agent.parameters()]
        self.flat_grad = z(sum(p.numel() for p in params))
        world = ppo.world
        opt = ppo.optimizer

        fused = c.get('fused_update_kernels', True)
        # 'fused_mlp_update': forward + loss + backward of both MLPs as ONE kernel reading the rollout
        # buffers through the minibatch indices (rk_ppo_minibatch_grad); only clip + Adam stay a graph
        self.fused_mlp = bool(fused and c.get('fused_mlp_update', True) and obs_dim <= PPO_MAX_OBS_DIM
                              and len(params) == 12)
        self.dmu, self.dv = z(mb, 2), z(mb)
        agent = ppo.agent
        if self.fused_mlp:
            # 'tensor_core_update' (default 2): every 64-wide product of the update -- the per-sample ones chained through
            # TMEM and the weight gradients over shared-memory operand tiles -- runs on tcgen05 tensor cores as TF32 x
            # 3-term fp32 emulation (27 % faster than the FMA kernel and equal to it to 1e-6); 1 keeps the
            # weight gradients on the CUDA cores, False / 0 selects the pure fp32 FMA kernel
            tc = c.get('tensor_core_update', 2)
            self.tensor_cores = (1 if tc else 0) if isinstance(tc, bool) else int(tc)
            self.grad = PpoMinibatchGrad(params, agent.log_std, obs_dim, c['clip_coef'], c['vf_coef'],
                                         tensor_cores=self.tensor_cores)
            self.flat_grad, self.kl_sum = self.grad.flat_grad, self.grad.kl_sum
            for p, gview in zip(params, self.grad.grad_views()):
                p.grad = gview                      # the kernel writes the gradients where Adam reads them

        def fwd_bwd():
            opt.zero_grad(set_to_none=False)
            if fused:
                # network outputs -> one kernel for d(loss)/d(mu, v) and the KL sum -> autograd through the MLPs only
                mu = agent.actor_mu(self.obs)
                v = agent.critic(self.obs).flatten()
                self.kl_sum.zero_()
                ppo_loss_grad(mu.detach(), v.detach(), self.act, self.old_logp, self.adv, self.ret, self.val,
                              agent.log_std, self.adv_mean, self.adv_std, c['clip_coef'], c['vf_coef'],
                              self.dmu, self.dv, self.kl_sum)
                torch.autograd.backward([mu, v], [self.dmu, self.dv])
            else:
                loss, kl = ppo._losses(self.obs, self.act, self.old_logp, self.adv, self.ret, self.val,
                                       self.adv_mean, self.adv_std)
                self.kl_sum.copy_(kl)
                loss.backward()
            if world > 1:
                torch.cat([p.grad.reshape(-1) for p in params], out=self.flat_grad)

        def clip_step():
            if self.fused_mlp:
                if world > 1:
                    self.flat_grad.div_(world)
            elif world > 1:
                off = 0
                for p in params:
                    p.grad.copy_(self.flat_grad[off:off + p.numel()].view_as(p)).div_(world)
                    off += p.numel()
            nn.utils.clip_grad_norm_(params, c['max_grad_norm'], foreach=True)
            opt.step()

        # warm-up on a side stream (allocates grads and Adam state), then restore the
        # parameters and the optimizer state IN PLACE so that training is unaffected
        saved_p = [p.detach().clone() for p in params]
        had_state = {id(p): {k: (v.clone() if torch.is_tensor(v) else v) for k, v in opt.state[p].items()}
                     for p in params if p in opt.state and len(opt.state[p])}
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(3):
                if not self.fused_mlp:
                    fwd_bwd()
                clip_step()
        torch.cuda.current_stream(dev).wait_stream(side)
        with torch.no_grad():
            for p, q in zip(params, saved_p):
                p.copy_(q)
                for k, v in opt.state[p].items():
                    if torch.is_tensor(v):
                        v.copy_(had_state[id(p)][k]) if id(p) in had_state else v.zero_()
        self.fwd_bwd, self.clip_step = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        self.adam = None
        if self.fused_mlp:
            with torch.cuda.graph(self.clip_step):
                clip_step()
            if c.get('fused_adam_step', True):
                # clip + Adam + KL stop as one kernel on the optimizer's own state: no per-minibatch host sync
                self.adam = PpoAdamStep(opt, params, self.flat_grad, self.kl_sum, c['max_grad_norm'], c['kl_target'],
                                        world=world, kl_sum_f32=self.grad.kl_f32 if world > 1 else None)
            return
        with torch.cuda.graph(self.fwd_bwd):
            fwd_bwd()
        with torch.cuda.graph(self.clip_step, pool=self.fwd_bwd.pool()):
            clip_step()


def _dist_ready():
    return torch.distributed.is_available() and torch.distributed.is_initialized()


class PPO:
    LOG_STD_RANGE = (-0.5, -1.6)  # agent/ppo.py:250-253

    def __init__(self, env_fn, config, device='cuda', query='culled'):
        self.config = config
        self.device = torch.device(device)
        self.query = query
        self.world = torch.distributed.get_world_size() if _dist_ready() else 1
        self.rank = torch.distributed.get_rank() if _dist_ready() else 0
        n_batch = config['num_envs'] * config['num_steps']
        if n_batch % config['num_minibatches'] != 0:
            # the reference's `range(0, batch_size, minibatch_size)` (agent/ppo.py:169) would run a final, smaller
            # minibatch; the captured update works on equal minibatches, so such a shape is rejected, not truncated
            raise ValueError(f"num_envs * num_steps = {n_batch} must be divisible by num_minibatches = "
                             f"{config['num_minibatches']} (equal minibatches only)")
        self.envs = self._make_vec_env(env_fn)
        random.seed(config['seed'])
        np.random.seed(config['seed'])
        torch.manual_seed(config['seed'])
        self.agent = Agent(self.envs.single_observation_space, self.envs.single_action_space).to(self.device)
        self.optimizer = self._make_optimizer()
        self._act_counter = 0
        self._perm_gen = None
        self._graphed = None

    def _make_optimizer(self):
        """Adam(lr 3e-4, eps 1e-5) as agent/ppo.py:83.  On CUDA it is created
        capturable with a tensor learning rate so that the minibatch step can be
        replayed as a CUDA graph while the learning rate is annealed in place."""
        c = self.config
        if self.device.type == 'cuda':
            lr = torch.tensor(float(c['learning_rate']), device=self.device)
            return optim.Adam(self.agent.parameters(), lr=lr, eps=1e-5, capturable=True, fused=True)
        return optim.Adam(self.agent.parameters(), lr=c['learning_rate'], eps=1e-5)

    def load_optimizer_state(self, state_dict):
        """optimizer.load_state_dict that keeps the device-resident update valid: the loaded
        state replaces the optimizer's tensors, so the cached graphs / kernel argument blocks
        are dropped, and a checkpoint written by the reference (plain Adam: python-float lr,
        non-capturable, `step` on the host) is converted to the capturable form used here."""
        self.optimizer.load_state_dict(state_dict)
        self._graphed = None
        if self.device.type != 'cuda':
            return
        for group in self.optimizer.param_groups:
            lr = group['lr']
            group['lr'] = torch.tensor(float(lr), device=self.device) if not isinstance(lr, torch.Tensor) \
                else lr.to(self.device, torch.float32).reshape(())
            group['capturable'], group['fused'], group['foreach'] = True, True, False
            for p in group['params']:
                st = self.optimizer.state.get(p)
                if st and 'step' in st:
                    step = st['step']
                    st['step'] = (step if isinstance(step, torch.Tensor) else torch.tensor(float(step))).to(
                        self.device, torch.float32).reshape(())
                    st['exp_avg'] = st['exp_avg'].to(self.device, torch.float32).contiguous()
                    st['exp_avg_sq'] = st['exp_avg_sq'].to(self.device, torch.float32).contiguous()

    def _set_lr(self, value):
        g = self.optimizer.param_groups[0]
        if isinstance(g['lr'], torch.Tensor):
            g['lr'].fill_(float(value))
        else:
            g['lr'] = float(value)

    # ---- env construction (agent/ppo.py:70,85-95) ------------------------------
    def _make_env(self, env_fn, seed, env_idx):
        return lambda: env_fn(env_idx)

    def _make_vec_env(self, env_fn):
        c = self.config
        if isinstance(env_fn, BatchedRacingVecEnv):
            return env_fn
        return BatchedRacingVecEnv([self._make_env(env_fn, c['seed'] + i, i) for i in range(c['num_envs'])],
                                   device=self.device if self.device.type == 'cuda' else None, query=self.query,
                                   seed=c['seed'] + 7919 * self.rank)

    # ---- rollout (agent/ppo.py:97-132), device resident ---------------------------
    def alloc_buffers(self):
        """[T+1] observation/done slots so the step kernel can write step t's
        successor in place; `obs[:T]`, `dones[:T]` are the reference's buffers."""
        c, be = self.config, self.envs.be
        T, E, A, D, dev = c['num_steps'], c['num_envs'], be.A, be.D, be.device
        return dict(obs=torch.zeros(T + 1, A, E, D, device=dev), actions=torch.zeros(T, A, E, 2, device=dev),
                    logprobs=torch.zeros(T, E, device=dev), dones=torch.zeros(T + 1, E, device=dev),
                    rewards=torch.zeros(T, A, E, device=dev), values=torch.zeros(T, E, device=dev))

    def collect_rollout(self, buf):
        """Fills buf in place; slot 0 of obs/dones must hold next_obs/next_done
        of the previous rollout.  Returns (episodes, mean_return, mean_length)."""
        c, envs = self.config, self.envs
        be = envs.be
        T = c['num_steps']
        params = flatten_agent(self.agent.state_dict()).to(be.device)
        be.ep_stats.zero_()
        seed = c['seed'] * 2654435761 + self.rank
        self.rollout_mode = 'native loop (rk_rollout: 2 launches per step)' if c.get('native_rollout', True) else \
            'python loop (3 launches per step)'
        if c.get('native_rollout', True):
            # the whole loop below as one C call: launches are issued back to back from native code, learner and
            # opponent inference share one grid ('native_rollout': False keeps the per-step Python loop)
            envs.rollout_into(buf, params, seed, self._act_counter + 1, T)
            self._act_counter += T
            T = 0
        with torch.no_grad():
            for t in range(T):
                self._act_counter += 1
                policy_act(params, buf['obs'][t, 0], buf['actions'][t, 0], seed=seed,
                           counter=self._act_counter, logprob=buf['logprobs'][t], value=buf['values'][t])
                envs.step_into(buf['actions'][t], buf['obs'][t + 1], buf['rewards'][t], buf['dones'][t + 1])
        n, ret, length = (float(v) for v in (be.ep_stats[2], be.ep_stats[0], be.ep_stats[1]))  # one sync per rollout
        return int(n), (ret / n if n else 0.0), (length / n if n else 0.0)

    # ---- GAE (agent/ppo.py:134-154) ---------------------------------------------------
    def compute_advantages(self, rewards, dones, values, next_value, next_done):
        c = self.config
        if rewards.is_cuda:
            return gae_kernel(rewards, values, dones, next_value, next_done, c['gamma'], c['gae_lambda'])
        raise RuntimeError('compute_advantages runs on the device only (rk_gae); there is no CPU path')

    # ---- update (agent/ppo.py:156-209) ------------------------------------------------
    def _all_reduce(self, t):
        if self.world > 1:
            torch.distributed.all_reduce(t)
        return t

    def _permutation(self, n, device):
        """Shared-seed device permutation (the reference shuffles on the host
        with np.random, agent/ppo.py:168); identical on every rank by construction.
        On CUDA it comes from rk_random_permutation (Feistel network, no sort: 4M
        indices in a few microseconds where torch.randperm sorts for 0.4 ms)."""
        if device.type == 'cuda':
            self._perm_count = getattr(self, '_perm_count', 0) + 1
            return random_permutation(n, self.config['seed'] + 12345, self._perm_count, device=device)
        if self._perm_gen is None or self._perm_gen.device != device:
            self._perm_gen = torch.Generator(device=device)
            self._perm_gen.manual_seed(self.config['seed'] + 12345)
        return torch.randperm(n, device=device, generator=self._perm_gen)

    def _losses(self, mb_obs, mb_act, old_logp, mb_adv, ret, v_old, adv_mean, adv_std):
        """Clipped surrogate + clipped value loss + entropy bonus of one minibatch
        (agent/ppo.py:173-204).  Returns (loss, sum(old_logp - new_logp))."""
        c = self.config
        _, new_logp, entropy, new_v = self.agent.get_action_and_value(mb_obs, mb_act)
        logratio = new_logp - old_logp
        ratio = logratio.exp()
        adv = (mb_adv - adv_mean) / (adv_std + 1e-8)
        pg_loss = torch.max(-adv * ratio, -adv * torch.clamp(ratio, 1 - c['clip_coef'], 1 + c['clip_coef'])).mean()
        new_v = new_v.flatten()
        v_clip = v_old + torch.clamp(new_v - v_old, -c['clip_coef'], c['clip_coef'])
        v_loss = 0.5 * torch.max((new_v - ret) ** 2, (v_clip - ret) ** 2).mean()
        loss = pg_loss + c['ent_coef'] * (-entropy.mean()) + c['vf_coef'] * v_loss
        return loss, (-logratio).detach().sum()

    def ppo_update(self, advantages, returns, values, logprobs, actions, obs, permutation=None):
        """Flat [B, ...] tensors of THIS rank's share of the batch.  Returns the
        number of optimizer steps taken (the KL early stop aborts the update)."""
        c = self.config
        if obs.is_cuda and c.get('cuda_graph_update', True):
            # 'update_matmul_precision': 'fp32' (default, what the reference computes on a GPU) or
            # 'tf32' (tensor-core GEMMs in the update only; rollout kernels are unaffected)
            prev = torch.backends.cuda.matmul.allow_tf32
            torch.backends.cuda.matmul.allow_tf32 = c.get('update_matmul_precision', 'fp32') == 'tf32'
            try:
                return self._ppo_update_graphed(advantages, returns, values, logprobs, actions, obs, permutation)
            finally:
                torch.backends.cuda.matmul.allow_tf32 = prev
        b_obs = obs.reshape(-1, obs.shape[-1])
        b_actions = actions.reshape(-1, actions.shape[-1])
        b_logprobs, b_adv = logprobs.reshape(-1), advantages.reshape(-1)
        b_returns, b_values = returns.reshape(-1), values.reshape(-1)
        n_local = b_obs.shape[0]
        mb_local = max(n_local // c['num_minibatches'], 1)
        params = [p for p in self.agent.parameters()]
        n_steps = 0
        for epoch in range(c['update_epochs']):
            perm = permutation(epoch) if permutation is not None else self._permutation(n_local, b_obs.device)
            for start in range(0, n_local - mb_local + 1, mb_local):
                idx = perm[start:start + mb_local]
                mb_obs, mb_act = b_obs[idx], b_actions[idx]
                _, new_logp, entropy, new_v = self.agent.get_action_and_value(mb_obs, mb_act)
                old_logp = b_logprobs[idx]
                logratio = new_logp - old_logp
                ratio = logratio.exp()
                mb_adv = b_adv[idx]
                with torch.no_grad():
                    # global-minibatch statistics: KL stop (ppo.py:178-182) and advantage
                    # normalisation with the unbiased std (ppo.py:187) over all ranks' samples
                    stats = torch.stack([(-logratio).sum(), mb_adv.sum(), (mb_adv * mb_adv).sum(),
                                         torch.tensor(float(mb_local), device=mb_adv.device)]).double()
                    self._all_reduce(stats)
                    n = stats[3]
                    approx_kl = stats[0] / n
                    mean = stats[1] / n
                    std = ((stats[2] - n * mean * mean) / (n - 1)).clamp_min(0).sqrt()
                    if approx_kl > c['kl_target']:
                        if self.rank == 0:
                            print(f'  Early stopping at epoch {epoch + 1} due to KL divergence: {float(approx_kl):.4f}')
                        return n_steps
                adv = (mb_adv - mean.float()) / (std.float() + 1e-8)
                pg_loss = torch.max(-adv * ratio, -adv * torch.clamp(ratio, 1 - c['clip_coef'], 1 + c['clip_coef'])).mean()
                new_v = new_v.flatten()
                v_old, ret = b_values[idx], b_returns[idx]
                v_clip = v_old + torch.clamp(new_v - v_old, -c['clip_coef'], c['clip_coef'])
                v_loss = 0.5 * torch.max((new_v - ret) ** 2, (v_clip - ret) ** 2).mean()
                loss = pg_loss + c['ent_coef'] * (-entropy.mean()) + c['vf_coef'] * v_loss
                self.optimizer.zero_grad(set_to_none=True)
                loss.backward()
                if self.world > 1:  # one flat all-reduce of the 11,075-float gradient
                    flat = torch.cat([p.grad.reshape(-1) for p in params])
                    self._all_reduce(flat).div_(self.world)
                    off = 0
                    for p in params:
                        p.grad.copy_(flat[off:off + p.numel()].view_as(p))
                        off += p.numel()
                nn.utils.clip_grad_norm_(params, c['max_grad_norm'])
                self.optimizer.step()
                n_steps += 1
        return n_steps


    # ---- the same update with the minibatch step replayed as two CUDA graphs ----------------
    def _ppo_update_graphed(self, advantages, returns, values, logprobs, actions, obs, permutation=None):
        """ppo_update for CUDA tensors.  Per minibatch: gather into static buffers and
        global advantage statistics (eager, a handful of launches), graph 1 = forward +
        losses + backward, then the KL early-stop test of agent/ppo.py:178-182 on the
        host (the only synchronisation; the step is NOT applied when it fires, as in
        the reference), graph 2 = gradient clipping + Adam.  With world > 1 the flat
        gradient and the KL sum are all-reduced between the two graphs."""
        c = self.config
        g = self._graphed
        d_obs = obs.shape[-1]
        n_rows = obs.numel() // d_obs
        mb_rows = max(n_rows // c['num_minibatches'], 1)
        if g is None or g.mb != mb_rows or g.obs.shape[1] != d_obs:
            g = self._graphed = _GraphedMinibatch(self, mb_rows, d_obs)
        if g.fused_mlp:
            # rows padded to a multiple of 4 floats (the kernel gathers them with 128-bit loads), filled
            # straight from the rollout buffer's view
            if getattr(g, 'obs_pad', None) is None or g.obs_pad.shape[0] != n_rows:
                g.obs_pad = torch.zeros(n_rows, (d_obs + 3) & ~3, device=obs.device)
            b_obs = g.obs_pad[:, :d_obs]
            b_obs.unflatten(0, obs.shape[:-1]).copy_(obs)
        else:
            b_obs = obs.reshape(-1, d_obs).contiguous()
        b_actions = actions.reshape(-1, actions.shape[-1]).contiguous()
        b_logprobs, b_adv = logprobs.reshape(-1).contiguous(), advantages.reshape(-1).contiguous()
        b_returns, b_values = returns.reshape(-1).contiguous(), values.reshape(-1).contiguous()
        n_local = b_obs.shape[0]
        mb = max(n_local // c['num_minibatches'], 1)
        fused = c.get('fused_update_kernels', True)
        n_steps = 0
        n_glob = float(mb * self.world)
        if g.fused_mlp and g.adam is not None:
            # The whole minibatch step stays on the device; the KL stop latches there and is read once per epoch.
            # One epoch (statistics, 16 x [gradient, all-reduce, clip + Adam]) is captured ONCE as a CUDA graph
            # over static buffers and replayed with a fresh permutation per epoch, so the update's speed does not
            # depend on how fast the host can issue ~50 launches per epoch ('graph_update_epoch': False disables).
            g.adam.reset()
            dev = b_obs.device
            st = getattr(g, 'epoch_static', None)
            if st is None or st['perm'].numel() != n_local:
                st = g.epoch_static = {'perm': torch.empty(n_local, dtype=torch.int64, device=dev),
                                       'act': torch.empty_like(b_actions), 'logp': torch.empty_like(b_logprobs),
                                       'adv': torch.empty_like(b_adv), 'ret': torch.empty_like(b_returns),
                                       'val': torch.empty_like(b_values)}
                g.epoch_graph, g.epoch_key = None, None
            for key, src in (('act', b_actions), ('logp', b_logprobs), ('adv', b_adv), ('ret', b_returns), ('val', b_values)):
                st[key].copy_(src)

            def run_epoch():
                self._all_reduce(g.grad.stats_epoch(st['perm'], mb, st['adv']))
                for k, start in enumerate(range(0, n_local - mb + 1, mb)):
                    g.grad.use_stats(k)
                    g.grad(st['perm'][start:start + mb], b_obs, st['act'], st['logp'], st['adv'], st['ret'], st['val'],
                           n_global=n_glob)
                    self._all_reduce(g.grad.grad_and_kl)
                    g.adam(n_glob, kl_target=c['kl_target'])
            use_graph = bool(c.get('graph_update_epoch', True))
            key = (n_local, mb, float(c['kl_target']), b_obs.data_ptr(), self.world)
            for epoch in range(c['update_epochs']):
                if permutation is not None:
                    st['perm'].copy_(permutation(epoch))
                else:
                    self._perm_count = getattr(self, '_perm_count', 0) + 1
                    random_permutation(n_local, c['seed'] + 12345, self._perm_count, out=st['perm'])
                if not use_graph:
                    run_epoch()
                elif g.epoch_graph is None or g.epoch_key != key:
                    run_epoch()                                   # this epoch runs eagerly (and warms everything up) ...
                    torch.cuda.synchronize(dev)
                    stopped = int(g.adam.state[0])
                    if not stopped:                               # ... then the same sequence is captured for the next ones
                        saved_state = g.adam.state.clone()
                        graph = torch.cuda.CUDAGraph()
                        snap = [p.detach().clone() for p in self.agent.parameters()]
                        osnap = [{k2: v.clone() for k2, v in self.optimizer.state[p].items() if torch.is_tensor(v)}
                                 for p in self.agent.parameters()]
                        with torch.cuda.graph(graph):
                            run_epoch()
                        # capture does not execute, but be explicit: parameters, optimizer state and flags are as before
                        with torch.no_grad():
                            for p, q, o in zip(self.agent.parameters(), snap, osnap):
                                p.copy_(q)
                                for k2, v in o.items():
                                    self.optimizer.state[p][k2].copy_(v)
                        g.adam.state.copy_(saved_state)
                        g.epoch_graph, g.epoch_key = graph, key
                else:
                    g.epoch_graph.replay()
                stopped, n_steps = (int(v) for v in g.adam.state.tolist()[:2])   # one host sync per epoch
                if stopped:
                    if self.rank == 0:
                        print(f'  Early stopping at epoch {epoch + 1} due to KL divergence: {float(g.adam.kl_at_stop):.4f}')
                    break
            return n_steps
        for epoch in range(c['update_epochs']):
            perm = permutation(epoch) if permutation is not None else self._permutation(n_local, b_obs.device)
            for start in range(0, n_local - mb + 1, mb):
                idx = perm[start:start + mb]
                if g.fused_mlp:
                    part = g.grad.stats(idx, b_adv)
                    self._all_reduce(part)
                    g.grad(idx, b_obs, b_actions, b_logprobs, b_adv, b_returns, b_values, n_global=n_glob)
                    if self.world > 1:
                        self._all_reduce(g.kl_sum)
                        self._all_reduce(g.flat_grad)
                    approx_kl = float(g.kl_sum) / n_glob  # host sync: the early-stop test needs the value
                    if approx_kl > c['kl_target']:
                        if self.rank == 0:
                            print(f'  Early stopping at epoch {epoch + 1} due to KL divergence: {approx_kl:.4f}')
                        return n_steps
                    g.clip_step.replay()
                    n_steps += 1
                    continue
                if fused:
                    gather_minibatch(idx, (b_obs, b_actions, b_logprobs, b_adv, b_returns, b_values),
                                     (g.obs, g.act, g.old_logp, g.adv, g.ret, g.val))
                else:
                    torch.index_select(b_obs, 0, idx, out=g.obs)
                    torch.index_select(b_actions, 0, idx, out=g.act)
                    torch.index_select(b_logprobs, 0, idx, out=g.old_logp)
                    torch.index_select(b_adv, 0, idx, out=g.adv)
                    torch.index_select(b_returns, 0, idx, out=g.ret)
                    torch.index_select(b_values, 0, idx, out=g.val)
                a64 = g.adv.double()
                stats = torch.stack([a64.sum(), (a64 * a64).sum()])
                self._all_reduce(stats)
                mean = stats[0] / n_glob
                g.adv_mean.copy_(mean)
                g.adv_std.copy_(((stats[1] - n_glob * mean * mean) / (n_glob - 1)).clamp_min(0).sqrt())
                g.fwd_bwd.replay()
                if self.world > 1:
                    self._all_reduce(g.kl_sum)
                    self._all_reduce(g.flat_grad)
                approx_kl = float(g.kl_sum) / n_glob  # host sync: the early-stop test needs the value
                if approx_kl > c['kl_target']:
                    if self.rank == 0:
                        print(f'  Early stopping at epoch {epoch + 1} due to KL divergence: {approx_kl:.4f}')
                    return n_steps
                g.clip_step.replay()
                n_steps += 1
        return n_steps

    # ---- training loop (agent/ppo.py:211-287) ---------------------------------------------
    def _anneal(self, update, num_updates):
        c = self.config
        frac = max(0.0, 1.0 - update / num_updates)
        self._set_lr(frac * c['learning_rate'])
        lo, hi = self.LOG_STD_RANGE
        self.agent.log_std.data.fill_(frac * lo + (1 - frac) * hi)
        return frac

    def _learn_from(self, buf):
        """GAE + update from a filled rollout buffer; returns optimizer steps taken."""
        c = self.config
        T = c['num_steps']
        obs, dones = buf['obs'][:T, 0], buf['dones'][:T]
        next_obs, next_done = buf['obs'][T, 0], buf['dones'][T]
        rewards, actions = buf['rewards'][:, 0], buf['actions'][:, 0]
        with torch.no_grad():
            next_value = self.agent.get_value(next_obs).flatten()
        adv, ret = self.compute_advantages(rewards.contiguous(), dones.contiguous(), buf['values'], next_value, next_done)
        steps = self.ppo_update(adv, ret, buf['values'], buf['logprobs'], actions, obs)
        # carry next_obs / next_done into slot 0 of the next rollout (ppo.py:104-106)
        buf['obs'][0].copy_(buf['obs'][T])
        buf['dones'][0].copy_(buf['dones'][T])
        return steps

    def train(self, log=print):
        c = self.config
        buf = self.alloc_buffers()
        buf['obs'][0].copy_(self._reset_all())
        num_updates = c['total_timesteps'] // (c['batch_size'] * self.world)
        global_step = 0
        training_info = {'steps': [], 'rewards': []}
        for update in range(num_updates):
            frac = self._anneal(update, num_updates)
            if c.get('anneal_speed_weight', False):
                # ppo.py:256-258 writes speed_weight 8 -> 14 onto the outermost gymnasium
                # wrapper, which gymnasium >= 1.0 does not forward to RacingEnv (SURVEY quirk
                # 11): off by default to match the pinned dependency's behaviour.
                self.envs.set_speed_weight(8.0 + (1 - frac) * 6.0)
            n_ep, mean_r, mean_l = self.collect_rollout(buf)
            self._learn_from(buf)
            global_step += c['batch_size'] * self.world
            if n_ep:
                training_info['steps'].append(global_step)
                training_info['rewards'].append(float(mean_r))
                log(f'Update {update + 1}/{num_updates} | Step {global_step} | Episodes: {n_ep} | '
                    f'Mean Reward: {mean_r:.2f} | Mean Length: {mean_l:.2f}')
            else:
                log(f'Update {update + 1}/{num_updates} | Step {global_step} | No episodes completed this rollout')
        self.training_info = training_info
        return training_info

    def _reset_all(self):
        self.envs.reset_device()
        return self.envs.be.obs

    def save(self, path):
        torch.save(self.agent.state_dict(), path)

    def load(self, path):
        self.agent.load_state_dict(torch.load(path, map_location=self.device))
