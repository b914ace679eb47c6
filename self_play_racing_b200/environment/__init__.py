"""Host-side mirror of the reference's `environment` package: the same class
names, constructor arguments and return conventions, backed by the CUDA
kernels of librk_b200.so."""
from .track import Track, gen_random_track, gen_tracks  # noqa: F401
from .racing_env import RacingEnv  # noqa: F401
from .multi_racing_env import MultiRacingEnv  # noqa: F401
from .wrappers import SelfPlayWrapper  # noqa: F401
from .vec_env import BatchedRacingVecEnv  # noqa: F401
