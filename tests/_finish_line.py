"""Shared helpers of the finish-line tests: push the hand-built states of
tests/golden/injected_*.npz (recorded from the unmodified reference by
tools/make_golden.py) into the oracle or into the CUDA backend (rk_set_state)."""
import numpy as np

F_CRASHED, F_FINISHED, F_CP25, F_CP50, F_CP75, F_HAS_CRASHED = 1, 2, 4, 8, 16, 32   # csrc/rk_types.cuh


def inject_oracle_single(env, g):
    f, i = g['init_f'], g['init_i']
    n = env.tracks[0].num_waypoints
    env.x[:, 0], env.y[:, 0], env.angle[:, 0], env.vx[:, 0], env.vy[:, 0] = f.T
    env.progress_idx[:, 0] = i[:, 0]
    env.progress[:, 0] = i[:, 0] / n
    env.last_progress[:, 0] = i[:, 1] / n
    env.checkpoints[:, 0, :] = i[:, 2:5].astype(bool)
    env.steps[:] = i[:, 5]


def inject_oracle_multi(env, g):
    f, i = g['init_f'], g['init_i']
    n = env.tracks[0].num_waypoints
    env.x[:], env.y[:], env.angle[:], env.vx[:], env.vy[:] = (f[..., k] for k in range(5))
    env.progress_idx[:] = i[..., 0]
    env.progress[:] = i[..., 0] / n
    env.last_progress[:] = i[..., 1] / n
    env.checkpoints[:] = i[..., 2:5].astype(bool)
    env.has_crashed[:] = i[..., 5].astype(bool)
    env.crashed[:] = i[..., 6].astype(bool)
    env.finished_step[:] = i[..., 7]
    env.steps[:] = g['init_steps']


def backend_state_single(g):
    """(car_f64 [S,1,6], car_i32 [S,1,4], env_i32 [S,3]) for RacingBackend.set_state."""
    f, i = g['init_f'], g['init_i']
    S = len(f)
    car_f = np.zeros((S, 1, 6)); car_f[:, 0, :5] = f
    flags = i[:, 2] * F_CP25 + i[:, 3] * F_CP50 + i[:, 4] * F_CP75
    car_i = np.stack([i[:, 0], i[:, 1], flags, np.zeros(S, np.int64)], axis=1)[:, None, :].astype(np.int32)
    env_i = np.stack([i[:, 5], np.zeros(S, np.int64), np.zeros(S, np.int64)], axis=1).astype(np.int32)
    return car_f, car_i, env_i


def backend_state_multi(g):
    f, i = g['init_f'], g['init_i']
    S, A = f.shape[:2]
    car_f = np.zeros((S, A, 6)); car_f[..., :5] = f
    flags = (i[..., 2] * F_CP25 + i[..., 3] * F_CP50 + i[..., 4] * F_CP75 + i[..., 5] * F_HAS_CRASHED +
             i[..., 6] * F_CRASHED)
    car_i = np.stack([i[..., 0], i[..., 1], flags, i[..., 7]], axis=2).astype(np.int32)
    env_i = np.stack([g['init_steps'], np.zeros(S, np.int64), np.zeros(S, np.int64)], axis=1).astype(np.int32)
    return car_f, car_i, env_i
