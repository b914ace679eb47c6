"""`SelfPlayWrapper` (reference environment/wrappers.py:5-62): the single-agent
view of a two-car env, the other car driven by a frozen policy snapshot or, when
no opponent is set, by uniform samples of its action space.

Wrapped around one `MultiRacingEnv` it behaves like the reference object (used
by evaluate-style loops).  `BatchedRacingVecEnv` recognises it and runs the
opponent's inference for all environments in one fused kernel instead."""
from __future__ import annotations

from .. import spaces


class SelfPlayWrapper(spaces.Wrapper):
    def __init__(self, env, agent_idx=0):
        super().__init__(env)
        self.agent_idx = agent_idx
        self.opponent_idx = 1 if agent_idx == 0 else 0
        self.action_space = env.action_space[f'{agent_idx}']
        self.observation_space = env.observation_space[f'{agent_idx}']
        self.opponent_policy = None
        self.opponent_action_space = env.action_space[f'{self.opponent_idx}']
        self.last_obs_dict = None

    def set_opponent(self, opponent_policy):
        self.opponent_policy = opponent_policy

    def reset(self, **kwargs):
        obs_dict, info_dict = self.env.reset(**kwargs)
        self.last_obs_dict = obs_dict
        return obs_dict[f'{self.agent_idx}'], info_dict[f'{self.agent_idx}']

    def step(self, action):
        if self.opponent_policy is None:
            opponent_action = self.opponent_action_space.sample()
        else:
            import torch
            dev = next(self.opponent_policy.parameters()).device
            opp_obs = torch.from_numpy(self.last_obs_dict[f'{self.opponent_idx}']).float().unsqueeze(0).to(dev)
            with torch.no_grad():
                opponent_action = self.opponent_policy.get_action_and_value(opp_obs)[0].squeeze(0).cpu().numpy()
        obs_dict, reward_dict, done_dict, truncated, info_dict = self.env.step(
            {f'{self.agent_idx}': action, f'{self.opponent_idx}': opponent_action})
        self.last_obs_dict = obs_dict
        k = f'{self.agent_idx}'
        return obs_dict[k], reward_dict[k], done_dict['__all__'], truncated, info_dict[k]

    @property
    def speed_weight(self):
        return self.env.speed_weight

    @speed_weight.setter
    def speed_weight(self, value):
        self.env.speed_weight = value
