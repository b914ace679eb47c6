"""`SelfPlayPPO` (reference agent/self_play_ppo.py:9-186): PPO against a FIFO
pool of frozen snapshots of itself.  The pool lives on the device as stacked
flat parameter vectors; the opponent's inference for all environments is one
fused kernel per step (see BatchedRacingVecEnv.step_into)."""
from __future__ import annotations

import copy
import os

import numpy as np
import torch

from ..environment.wrappers import SelfPlayWrapper
from .ppo import PPO, Agent


class SelfPlayPPO(PPO):
    LOG_STD_RANGE = (-0.3, -1.2)  # self_play_ppo.py:136-139

    def __init__(self, env_fn, config, device='cuda', query='culled', checkpoint_dir='models'):
        self.env_fn = env_fn
        self.opponent_pool = []   # frozen Agent snapshots, oldest first (self_play_ppo.py:12)
        self.curr_opponent = None
        self.snapshot_freq = config['snapshot_freq']
        self.pool_size = config['pool_size']
        self.checkpoint_dir = checkpoint_dir
        super().__init__(env_fn, config, device, query=query)

    def _make_env(self, env_fn, seed, env_idx=0):
        def thunk():
            env = SelfPlayWrapper(self.env_fn(env_idx), 0)  # self_play_ppo.py:21-22
            env.set_opponent(self.curr_opponent)
            return env
        return thunk

    def snapshot_agent(self):
        """A frozen deep copy of the learner, log_std buffer included (self_play_ppo.py:31-38)."""
        snap = Agent(self.envs.single_observation_space, self.envs.single_action_space).to(self.device)
        snap.load_state_dict(copy.deepcopy(self.agent.state_dict()))
        snap.eval()
        for p in snap.parameters():
            p.requires_grad = False
        return snap

    def select_opponent(self):
        if not self.opponent_pool:
            return None
        return self.opponent_pool[np.random.choice(len(self.opponent_pool))]  # self_play_ppo.py:40-44

    def update_opponent(self):
        """One opponent per update for all environments; the reference closes and
        rebuilds every env here (self_play_ppo.py:46-50), i.e. every rollout
        starts from fresh start-line states while the learner's carried
        next_obs stays stale for step 0 -- reproduced by
        resetting the batch without touching the rollout buffer's slot 0."""
        if self.config.get('opponents_per_update', 'one') == 'pool' and len(self.opponent_pool) > 1:
            # superset of the reference: every block of 256 envs draws its own pool member,
            # all of them served by one inference launch
            self.curr_opponent = self.opponent_pool[-1]
            self.envs.set_opponents(self.opponent_pool, seed=int(np.random.randint(0, 2 ** 31 - 1)))
        else:
            self.curr_opponent = self.select_opponent()
            self.envs.set_opponent(self.curr_opponent)
        self.envs.reset_device()

    def load_checkpoint(self, checkpoint_path):
        ck = torch.load(checkpoint_path, map_location=self.device, weights_only=False)
        self.agent.load_state_dict(ck['agent_state_dict'])
        self.load_optimizer_state(ck['optimizer_state_dict'])
        self.opponent_pool = []
        for sd in ck['opponent_pool']:
            opp = Agent(self.envs.single_observation_space, self.envs.single_action_space).to(self.device)
            opp.load_state_dict(sd)
            opp.eval()
            for p in opp.parameters():
                p.requires_grad = False
            self.opponent_pool.append(opp)
        info = ck.get('training_info', {'steps': [], 'rewards': [], 'opponent_pool_size': []})
        return ck['update'], ck['global_step'], info

    def save_checkpoint(self, update, global_step, training_info):
        """Same dict layout as self_play_ppo.py:154-167."""
        os.makedirs(self.checkpoint_dir, exist_ok=True)
        path = os.path.join(self.checkpoint_dir, f'checkpoint_update_{update}.pth')
        torch.save({'update': update, 'global_step': global_step, 'agent_state_dict': self.agent.state_dict(),
                    'optimizer_state_dict': self.optimizer.state_dict(),
                    'opponent_pool': [o.state_dict() for o in self.opponent_pool],
                    'config': self.config, 'training_info': training_info}, path)
        return path

    def train(self, resume_from=None, log=print):
        c = self.config
        buf = self.alloc_buffers()
        buf['obs'][0].copy_(self._reset_all())
        num_updates = c['total_timesteps'] // (c['batch_size'] * self.world)
        if resume_from:
            start_update, global_step, training_info = self.load_checkpoint(resume_from)
            start_update += 1
            log(f'RESUMING TRAINING from update {start_update}/{num_updates} (global step {global_step}, '
                f'pool {len(self.opponent_pool)})')
        else:
            start_update, global_step = 0, 0
            training_info = {'steps': [], 'rewards': [], 'opponent_pool_size': []}
        for update in range(start_update, num_updates):
            if update > 0 and update % self.snapshot_freq == 0:  # self_play_ppo.py:115-122
                self.opponent_pool.append(self.snapshot_agent())
                if len(self.opponent_pool) > self.pool_size:
                    self.opponent_pool.pop(0)
            self.update_opponent()
            self._anneal(update, num_updates)
            n_ep, mean_r, mean_l = self.collect_rollout(buf)
            self._learn_from(buf)
            global_step += c['batch_size'] * self.world
            if update > 0 and update % 10 == 0 and self.rank == 0:
                self.save_checkpoint(update, global_step, training_info)
            if n_ep:
                training_info['steps'].append(global_step)
                training_info['rewards'].append(float(mean_r))
                training_info['opponent_pool_size'].append(len(self.opponent_pool))
                log(f'Update {update + 1}/{num_updates} | Step {global_step} | Episodes: {n_ep} | '
                    f'Mean Reward: {mean_r:.2f} | Mean Length: {mean_l:.2f} | Pool Size: {len(self.opponent_pool)}')
            else:
                log(f'Update {update + 1}/{num_updates} | Step {global_step} | No episodes completed this rollout')
        self.training_info = training_info
        self.envs.close()
        return training_info
