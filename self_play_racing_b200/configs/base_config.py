"""Single-agent PPO schedule (reference configs/base_config.py): 16 envs x 2048 steps by default."""
from ._common import build


def hyperparams_config(num_envs=16, num_steps=2048, **overrides):
    return build(1, num_envs, num_steps, overrides)


def b200_config(num_envs=65536, num_steps=64, **overrides):
    """The same schedule re-shaped for one B200: many environments, short rollouts."""
    return build(1, num_envs, num_steps, overrides)
