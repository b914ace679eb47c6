from .base_config import hyperparams_config as base_config  # noqa: F401
from .self_play_config import hyperparams_config as self_play_config  # noqa: F401
