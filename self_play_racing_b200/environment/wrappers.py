"""Single-agent view of a two-car race for self-play training.

Mirror of the reference's `SelfPlayWrapper` (environment/wrappers.py:5-62): the
learner drives one car, the other car is driven by a frozen policy snapshot, or
by uniform samples of its action box while no snapshot has been set.  Public
surface kept: `SelfPlayWrapper(env, agent_idx=0)`, `set_opponent(policy)`,
`reset(**kw)`, `step(action) -> (obs, reward, done, truncated, info)` where
`done` is the env's `dones["__all__"]`, the `action_space` / `observation_space`
of the learner's car, and the `speed_weight` pass-through.

Around one `MultiRacingEnv` it serves evaluate-style loops; the opponent's
forward pass then runs through the same fused inference kernel the batched path
uses (`rk_policy_act`, batch of one, reading the opponent's observation where the
step kernel left it on the device).  `BatchedRacingVecEnv` recognises wrapped
envs and applies the same semantics to the whole batch in one launch.
"""
from __future__ import annotations

import itertools

from .. import spaces

_launch_ids = itertools.count(1)


class SelfPlayWrapper(spaces.Wrapper):
    def __init__(self, env, agent_idx=0):
        super().__init__(env)
        me, them = int(agent_idx), 1 - int(agent_idx) if agent_idx in (0, 1) else 0
        self.agent_idx, self.opponent_idx = me, them
        self._me, self._them = str(me), str(them)
        self.action_space = env.action_space[self._me]
        self.observation_space = env.observation_space[self._me]
        self.opponent_action_space = env.action_space[self._them]
        self.opponent_policy = None
        self._packed_opponent = None   # device copy of the snapshot in the kernel's layout
        self.last_obs_dict = None

    # -- opponent ------------------------------------------------------------
    def set_opponent(self, opponent_policy):
        """None -> random opponent; otherwise anything with a `state_dict()` of an Agent."""
        self.opponent_policy = opponent_policy
        self._packed_opponent = None

    def _opponent_action(self):
        if self.opponent_policy is None:
            return self.opponent_action_space.sample()          # wrappers.py:30-32
        be = getattr(self.env, '_be', None)
        if be is None:                                          # foreign env object: plain torch forward
            import torch
            dev = next(self.opponent_policy.parameters()).device
            x = torch.as_tensor(self.last_obs_dict[self._them], dtype=torch.float32, device=dev)[None]
            with torch.no_grad():
                return self.opponent_policy.get_action_and_value(x)[0][0].cpu().numpy()
        from ..backend import flatten_agent, policy_act
        if self._packed_opponent is None:
            self._packed_opponent = flatten_agent(self.opponent_policy.state_dict()).to(be.device)
        out = be.actions[0, self.opponent_idx:self.opponent_idx + 1]
        policy_act(self._packed_opponent, be.obs[0, self.opponent_idx:self.opponent_idx + 1], out,
                   seed=0x0bb0, counter=next(_launch_ids))
        return out[0].cpu().numpy()

    # -- gymnasium API -------------------------------------------------------
    def reset(self, **kwargs):
        self.last_obs_dict, infos = self.env.reset(**kwargs)
        return self.last_obs_dict[self._me], infos[self._me]

    def step(self, action):
        joint = {self._me: action, self._them: self._opponent_action()}
        self.last_obs_dict, rewards, dones, truncated, infos = self.env.step(joint)
        return self.last_obs_dict[self._me], rewards[self._me], dones['__all__'], truncated, infos[self._me]

    # -- RacingEnv.speed_weight pass-through (wrappers.py:57-63) ----------------
    speed_weight = property(lambda self: self.env.speed_weight,
                            lambda self, value: setattr(self.env, 'speed_weight', value))
