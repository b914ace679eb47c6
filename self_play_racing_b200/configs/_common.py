"""Shared construction of the hyper-parameter dictionaries.  The KEYS and VALUES
are the reference's (configs/base_config.py:1-28, configs/self_play_config.py:1-32)
because its trainers read them by name; `num_envs` / `num_steps` / any key can be
overridden to re-shape the schedule for a GPU-sized batch."""

# (key, single-agent value, self-play value); None = key absent in that config
_TABLE = (
    ('total_timesteps', 5_000_000, 3_000_000),
    ('learning_rate', 3e-4, 3e-4),
    ('gamma', 0.99, 0.99),
    ('gae_lambda', 0.95, 0.97),
    ('clip_coef', 0.2, 0.2),
    ('ent_coef', 0.01, 0.02),
    ('vf_coef', 0.5, 0.5),
    ('update_epochs', 10, 10),
    ('num_minibatches', 16, 16),
    ('max_grad_norm', 0.5, 0.5),
    ('kl_target', 0.015, 0.015),
    ('snapshot_freq', None, 15),
    ('pool_size', None, 5),
    ('seed', 1, 1),
    ('cuda', True, True),
    ('torch_deterministic', True, True),
)


def build(column, num_envs, num_steps, overrides):
    cfg = {'num_envs': num_envs, 'num_steps': num_steps}
    cfg.update({row[0]: row[column] for row in _TABLE if row[column] is not None})
    cfg.update(overrides)
    cfg['batch_size'] = cfg['num_steps'] * cfg['num_envs']                     # derived as in the reference
    cfg['minibatch_size'] = cfg['batch_size'] // cfg['num_minibatches']
    return cfg
