from .ppo import Agent, PPO  # noqa: F401
from .self_play_ppo import SelfPlayPPO  # noqa: F401
