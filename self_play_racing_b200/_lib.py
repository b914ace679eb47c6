"""ctypes binding of librk_b200.so (include/racing_b200.h).

There is no CPU fallback: if the shared library is missing the import raises,
and rk_create refuses to run without a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('RK_B200_LIB') or os.path.join(_HERE, 'librk_b200.so')  # env override: A/B kernel builds

RK_ENV_SINGLE, RK_ENV_MULTI = 0, 1
RK_AUTORESET_NEXT_STEP, RK_AUTORESET_SAME_STEP, RK_AUTORESET_DISABLED = 0, 1, 2
RK_QUERY_EXACT_F64, RK_QUERY_CULLED, RK_QUERY_GRID = 0, 1, 2
RK_MAX_AGENTS, RK_MAX_SENSORS = 8, 64
RK_LAYOUT_ENV_MAJOR, RK_LAYOUT_AGENT_MAJOR = 0, 1

AUTORESET = {'next_step': RK_AUTORESET_NEXT_STEP, 'same_step': RK_AUTORESET_SAME_STEP,
             'disabled': RK_AUTORESET_DISABLED}
QUERY = {'exact': RK_QUERY_EXACT_F64, 'culled': RK_QUERY_CULLED, 'grid': RK_QUERY_GRID}


class RkConfig(C.Structure):
    _fields_ = [('struct_size', C.c_int32), ('device', C.c_int32), ('num_envs', C.c_int32),
                ('num_agents', C.c_int32), ('num_sensors', C.c_int32), ('env_kind', C.c_int32),
                ('autoreset_mode', C.c_int32), ('query_mode', C.c_int32),
                ('max_episode_steps', C.c_int32), ('reserved0', C.c_int32),
                ('speed_weight', C.c_double), ('seed', C.c_uint64)]


class RkStepIO(C.Structure):
    _fields_ = [('struct_size', C.c_int32), ('layout', C.c_int32),
                ('actions', C.c_void_p), ('start_slot', C.c_void_p), ('obs', C.c_void_p),
                ('reward_f32', C.c_void_p), ('reward_f64', C.c_void_p),
                ('terminated', C.c_void_p), ('truncated', C.c_void_p), ('done', C.c_void_p),
                ('done_f32', C.c_void_p), ('ep_mask', C.c_void_p), ('ep_return', C.c_void_p),
                ('ep_length', C.c_void_p), ('info_f64', C.c_void_p), ('info_i32', C.c_void_p),
                ('ep_stats', C.c_void_p), ('env_begin', C.c_int32), ('env_count', C.c_int32)]


class RkPpoGradIO(C.Structure):
    _fields_ = [('struct_size', C.c_int32), ('obs_dim', C.c_int32), ('n', C.c_int32), ('obs_stride', C.c_int32),
                ('n_global', C.c_double), ('params', C.c_void_p * 12), ('log_std', C.c_void_p),
                ('obs', C.c_void_p), ('act', C.c_void_p), ('old_logp', C.c_void_p), ('adv', C.c_void_p),
                ('ret', C.c_void_p), ('val', C.c_void_p), ('idx', C.c_void_p), ('adv_part', C.c_void_p),
                ('clip_coef', C.c_float), ('vf_coef', C.c_float), ('workspace', C.c_void_p),
                ('workspace_bytes', C.c_uint64), ('flat_grad', C.c_void_p), ('kl_sum', C.c_void_p),
                ('kl_sum_f32', C.c_void_p), ('tensor_cores', C.c_int32), ('reserved1', C.c_int32)]


class RkAdamIO(C.Structure):
    _fields_ = [('struct_size', C.c_int32), ('world', C.c_int32), ('params', C.c_void_p * 12),
                ('exp_avg', C.c_void_p * 12), ('exp_avg_sq', C.c_void_p * 12), ('step', C.c_void_p * 12),
                ('numel', C.c_int32 * 12), ('flat_grad', C.c_void_p), ('lr', C.c_void_p),
                ('beta1', C.c_float), ('beta2', C.c_float), ('eps', C.c_float), ('max_grad_norm', C.c_float),
                ('kl_target', C.c_float), ('reserved0', C.c_int32), ('kl_sum', C.c_void_p), ('kl_sum_f32', C.c_void_p),
                ('n_global', C.c_double), ('state', C.c_void_p), ('kl_at_stop', C.c_void_p)]


class RkHostIO(C.Structure):
    _fields_ = [('struct_size', C.c_int32), ('n_chunks', C.c_int32), ('actions', C.c_void_p), ('obs', C.c_void_p),
                ('arena_host', C.c_void_p), ('arena_dev', C.c_void_p), ('arena_bytes', C.c_int64),
                ('selfplay', C.c_int32), ('reserved0', C.c_int32), ('opponent_params', C.c_void_p),
                ('seed', C.c_uint64), ('counter', C.c_uint64)]


class RkRolloutIO(C.Structure):
    _fields_ = [('struct_size', C.c_int32), ('T', C.c_int32), ('selfplay', C.c_int32), ('block_len', C.c_int32),
                ('learner_params', C.c_void_p), ('opponent_params', C.c_void_p), ('block_policy', C.c_void_p),
                ('pool_stride', C.c_int64), ('opponent_obs0', C.c_void_p),
                ('learner_seed', C.c_uint64), ('learner_counter0', C.c_uint64),
                ('opponent_seed', C.c_uint64), ('opponent_counter0', C.c_uint64),
                ('obs', C.c_void_p), ('actions', C.c_void_p), ('logprobs', C.c_void_p), ('values', C.c_void_p),
                ('rewards', C.c_void_p), ('dones', C.c_void_p)]


# name -> (restype, argtypes); every symbol include/racing_b200.h declares
SIGNATURES = {
    'rk_create': (C.c_int, [C.POINTER(RkConfig), C.POINTER(C.c_void_p)]),
    'rk_destroy': (C.c_int, [C.c_void_p]),
    'rk_last_error': (C.c_char_p, [C.c_void_p]),
    'rk_abi_version': (C.c_int, []),
    'rk_launch_count': (C.c_uint64, []),
    'rk_set_tracks_from_control_points': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                                    C.c_int32, C.c_void_p]),
    'rk_set_tracks_from_waypoints': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                               C.c_void_p]),
    'rk_generate_tracks': (C.c_int, [C.c_void_p, C.c_uint64, C.c_int32, C.c_int32, C.c_double, C.c_int32,
                                     C.c_void_p]),
    'rk_num_tracks': (C.c_int, [C.c_void_p]),
    'rk_get_track': (C.c_int, [C.c_void_p, C.c_int32] + [C.c_void_p] * 7),
    'rk_reset': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    'rk_step': (C.c_int, [C.c_void_p, C.POINTER(RkStepIO), C.c_void_p]),
    'rk_step_host': (C.c_int, [C.c_void_p, C.POINTER(RkStepIO), C.POINTER(RkHostIO), C.c_void_p]),
    'rk_rollout': (C.c_int, [C.c_void_p, C.POINTER(RkStepIO), C.POINTER(RkRolloutIO), C.c_void_p]),
    'rk_set_speed_weight': (C.c_int, [C.c_void_p, C.c_double]),
    'rk_set_seed': (C.c_int, [C.c_void_p, C.c_uint64]),
    'rk_get_state': (C.c_int, [C.c_void_p] * 5),
    'rk_set_state': (C.c_int, [C.c_void_p] * 5),
    'rk_observe': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    'rk_gae': (C.c_int, [C.c_void_p] * 5 + [C.c_float, C.c_float, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                            C.c_void_p]),
    'rk_policy_act': (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.c_uint64, C.c_uint64,
                                C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    'rk_policy_param_count': (C.c_int, [C.c_int32]),
    'rk_fma_peak': (C.c_double, [C.c_int32, C.c_int32]),
    'rk_gather_minibatch': (C.c_int, [C.c_void_p, C.c_int32, C.c_int32] + [C.c_void_p] * 13),
    'rk_ppo_adv_stats': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    'rk_ppo_grad_workspace_bytes': (C.c_uint64, []),
    'rk_ppo_minibatch_grad': (C.c_int, [C.POINTER(RkPpoGradIO), C.c_void_p]),
    'rk_policy_act_pool': (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int64,
                                     C.c_int32, C.c_uint64, C.c_uint64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p]),
    'rk_random_permutation': (C.c_int, [C.c_uint64, C.c_uint64, C.c_int64, C.c_void_p, C.c_void_p]),
    'rk_ppo_adam_step': (C.c_int, [C.POINTER(RkAdamIO), C.c_void_p]),
    'rk_ppo_loss_grad': (C.c_int, [C.c_void_p] * 10 + [C.c_int32, C.c_float, C.c_float, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p]),
}

_lib = None


def load():
    """Load the shared library once; raise loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f'{LIB_PATH} is missing: build it with `python -c "import __graft_entry__ as g; g.build()"` '
            f'or `make -C self_play_racing_b200/csrc`. This backend has no CPU fallback.')
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.rk_abi_version() != 1:
        raise RuntimeError('librk_b200.so ABI version mismatch')
    _lib = lib
    return lib


def last_error(handle=None) -> str:
    msg = load().rk_last_error(handle)
    return msg.decode() if msg else ''


def check(rc, handle=None, what=''):
    if rc != 0:
        raise RuntimeError(f'{what or "librk_b200"} failed: {last_error(handle)}')


def launch_count() -> int:
    return int(load().rk_launch_count())
