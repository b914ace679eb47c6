"""Make the reference's own scripts (`train.py`, `evaluate.py`) run on this
backend without editing them: register this package's mirrors under the module
names those scripts import (`environment.*`, `agent.*`, `configs.*`)."""
from __future__ import annotations

import importlib
import sys

_MIRRORS = {
    'environment': ['track', 'racing_env', 'multi_racing_env', 'wrappers', 'vec_env'],
    'agent': ['ppo', 'self_play_ppo'],
    'configs': ['base_config', 'self_play_config'],
}


def install_as_reference_modules():
    """After this call `from environment.multi_racing_env import MultiRacingEnv`,
    `from agent.self_play_ppo import SelfPlayPPO`, ... resolve to the CUDA-backed
    classes.  Returns the list of module names registered."""
    done = []
    for pkg, subs in _MIRRORS.items():
        mod = importlib.import_module(f'self_play_racing_b200.{pkg}')
        sys.modules[pkg] = mod
        done.append(pkg)
        for sub in subs:
            sys.modules[f'{pkg}.{sub}'] = importlib.import_module(f'self_play_racing_b200.{pkg}.{sub}')
            done.append(f'{pkg}.{sub}')
    return done
