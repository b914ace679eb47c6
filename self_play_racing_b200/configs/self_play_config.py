"""Self-play PPO schedule (reference configs/self_play_config.py): adds snapshot_freq and pool_size."""
from ._common import build


def hyperparams_config(num_envs=16, num_steps=2048, **overrides):
    return build(2, num_envs, num_steps, overrides)


def b200_config(num_envs=65536, num_steps=64, **overrides):
    return build(2, num_envs, num_steps, overrides)
