"""self_play_racing_b200 -- B200-native batched backend for the racing step and
rollout path of LucasHJin/self-play-racing."""
from . import _lib  # noqa: F401

__all__ = ['_lib']
