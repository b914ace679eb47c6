"""Self-play PPO hyper-parameters: same keys and values as the reference's
configs/self_play_config.py:1-32."""


def hyperparams_config(num_envs=16, num_steps=2048, **overrides):
    config = dict(
        total_timesteps=3_000_000, num_envs=num_envs, num_steps=num_steps, learning_rate=3e-4,
        gamma=0.99, gae_lambda=0.97, clip_coef=0.2, ent_coef=0.02, vf_coef=0.5,
        update_epochs=10, num_minibatches=16, max_grad_norm=0.5, kl_target=0.015,
        snapshot_freq=15, pool_size=5,
        seed=1, cuda=True, torch_deterministic=True,
    )
    config.update(overrides)
    config['batch_size'] = config['num_steps'] * config['num_envs']
    config['minibatch_size'] = config['batch_size'] // config['num_minibatches']
    return config


def b200_config(num_envs=65536, num_steps=64, **overrides):
    return hyperparams_config(num_envs=num_envs, num_steps=num_steps, **overrides)
