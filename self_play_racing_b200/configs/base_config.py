"""Single-agent PPO hyper-parameters: same keys and values as the reference's
configs/base_config.py:1-28, plus the B200 batch-shape overrides."""


def hyperparams_config(num_envs=16, num_steps=2048, **overrides):
    config = dict(
        total_timesteps=5_000_000, num_envs=num_envs, num_steps=num_steps, learning_rate=3e-4,
        gamma=0.99, gae_lambda=0.95, clip_coef=0.2, ent_coef=0.01, vf_coef=0.5,
        update_epochs=10, num_minibatches=16, max_grad_norm=0.5, kl_target=0.015,
        seed=1, cuda=True, torch_deterministic=True,
    )
    config.update(overrides)
    config['batch_size'] = config['num_steps'] * config['num_envs']
    config['minibatch_size'] = config['batch_size'] // config['num_minibatches']
    return config


def b200_config(num_envs=65536, num_steps=64, **overrides):
    """The reference schedule (16 envs x 2048 steps) re-shaped for one B200:
    many environments, short rollouts (SURVEY.md section 7 'hyper-parameter shape change')."""
    return hyperparams_config(num_envs=num_envs, num_steps=num_steps, **overrides)
