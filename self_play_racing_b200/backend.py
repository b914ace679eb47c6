"""Device-resident batched racing simulator: a thin object over the C ABI.

PyTorch is used only for device memory and streams; every computation on the
step path is a kernel of librk_b200.so.  Host-side mirror of the reference's
vectorised execution: E x (RacingEnv | MultiRacingEnv) stepped in one launch
(reference call site: SyncVectorEnv.step in agent/ppo.py:114).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


def _ptr(t):
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def _np_ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class RacingBackend:
    """E environments of one kind on one GPU.

    kind='single' -> RacingEnv rules (A=1, 120 deg cone, obs R+4);
    kind='multi'  -> MultiRacingEnv rules (A cars, 180 deg cone, obs R+4+4(A-1)).
    Output tensors are allocated once and overwritten by every step.
    """

    def __init__(self, num_envs, kind='single', num_agents=1, num_sensors=11, device=None,
                 autoreset='next_step', query='exact', speed_weight=8.0, seed=0,
                 max_episode_steps=3000, want_info=True, agent_major=False):
        if not torch.cuda.is_available():
            raise RuntimeError('self_play_racing_b200 needs a CUDA device (sm_100a); there is no CPU path')
        self.lib = _lib.load()
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.index is None:
            self.device = torch.device('cuda', torch.cuda.current_device())
        self.kind = kind
        self.E = int(num_envs)
        self.A = 1 if kind == 'single' else int(num_agents)
        self.R = int(num_sensors)
        self.D = self.R + 4 if kind == 'single' else self.R + 4 + 4 * (self.A - 1)
        cfg = _lib.RkConfig(struct_size=C.sizeof(_lib.RkConfig), device=self.device.index, num_envs=self.E,
                            num_agents=self.A, num_sensors=self.R,
                            env_kind=_lib.RK_ENV_SINGLE if kind == 'single' else _lib.RK_ENV_MULTI,
                            autoreset_mode=_lib.AUTORESET[autoreset], query_mode=_lib.QUERY[query],
                            max_episode_steps=int(max_episode_steps), reserved0=0,
                            speed_weight=float(speed_weight), seed=int(seed))
        h = C.c_void_p()
        _lib.check(self.lib.rk_create(C.byref(cfg), C.byref(h)), None, 'rk_create')
        self.h = h
        E, A, D, dev = self.E, self.A, self.D, self.device
        f32, f64, u8, i32 = torch.float32, torch.float64, torch.uint8, torch.int32
        # per-car arrays are [E,A,...], or [A,E,...] when agent_major (car a of
        # every env contiguous: the SelfPlayWrapper view needs no gather)
        self.agent_major = bool(agent_major)
        self.layout = _lib.RK_LAYOUT_AGENT_MAJOR if agent_major else _lib.RK_LAYOUT_ENV_MAJOR
        lead = (A, E) if agent_major else (E, A)
        self.actions = torch.zeros(*lead, 2, dtype=f32, device=dev)
        self.obs = torch.zeros(*lead, D, dtype=f32, device=dev)
        self.reward = torch.zeros(*lead, dtype=f32, device=dev)
        self.done = torch.zeros(E, dtype=u8, device=dev)
        self.done_f32 = torch.zeros(E, dtype=f32, device=dev)
        # Host-visible results of car 0 live back to back in one arena so the
        # Gymnasium-facing step can fetch them with a single device->host copy:
        # [ep_return f64 | ep_length i32 | terminated | truncated | ep_mask | pad | reward64 (all cars)]
        sizes = [8 * E, 4 * E, E, E, E]
        offs = np.concatenate([[0], np.cumsum(sizes)])
        r64_off = int((offs[-1] + 7) // 8 * 8)
        self.arena = torch.zeros(r64_off + 8 * A * E, dtype=u8, device=dev)
        cut = lambda o, n, dt: self.arena[int(o):int(o) + n].view(dt)
        self.ep_return = cut(offs[0], 8 * E, f64)
        self.ep_length = cut(offs[1], 4 * E, i32)
        self.terminated = cut(offs[2], E, u8)
        self.truncated = cut(offs[3], E, u8)
        self.ep_mask = cut(offs[4], E, u8)
        self.reward64 = cut(r64_off, 8 * A * E, f64).view(*lead)
        self.arena_host_bytes = r64_off + 8 * E if agent_major or A == 1 else r64_off + 8 * A * E
        self.arena_offsets = dict(ep_return=0, ep_length=int(offs[1]), terminated=int(offs[2]),
                                  truncated=int(offs[3]), ep_mask=int(offs[4]), reward64=r64_off)
        self.ep_stats = torch.zeros(3, dtype=f64, device=dev)  # sum return, sum length, episodes
        self.info_f64 = torch.zeros(*lead, 5, dtype=f64, device=dev) if want_info else None
        self.info_i32 = torch.zeros(*lead, 4, dtype=i32, device=dev) if want_info else None
        self._io = _lib.RkStepIO(struct_size=C.sizeof(_lib.RkStepIO))
        self._bind_io()

    def _bind_io(self):
        io = self._io
        io.layout = self.layout
        io.ep_stats = self.ep_stats.data_ptr()
        io.actions = self.actions.data_ptr()
        io.start_slot = None
        io.obs = self.obs.data_ptr()
        io.reward_f32 = self.reward.data_ptr()
        io.reward_f64 = self.reward64.data_ptr()
        io.terminated = self.terminated.data_ptr()
        io.truncated = self.truncated.data_ptr()
        io.done = self.done.data_ptr()
        io.done_f32 = self.done_f32.data_ptr()
        io.ep_mask = self.ep_mask.data_ptr()
        io.ep_return = self.ep_return.data_ptr()
        io.ep_length = self.ep_length.data_ptr()
        io.info_f64 = self.info_f64.data_ptr() if self.info_f64 is not None else None
        io.info_i32 = self.info_i32.data_ptr() if self.info_i32 is not None else None

    # ---- lifetime ---------------------------------------------------------
    def close(self):
        if getattr(self, 'h', None):
            torch.cuda.synchronize(self.device)
            self.lib.rk_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ---- tracks -----------------------------------------------------------
    def _e2t(self, env_to_track):
        if env_to_track is None:
            return None
        a = np.ascontiguousarray(env_to_track, dtype=np.int32)
        assert a.shape == (self.E,)
        return a

    def set_tracks_from_control_points(self, control_points, widths, env_to_track=None, factor=30):
        """Device-side Track.__init__ (track.py:61-148) for a pool of control-point arrays."""
        n = np.array([len(c) for c in control_points], dtype=np.int32)
        xy = np.ascontiguousarray(np.concatenate([np.asarray(c, dtype=np.float64) for c in control_points]))
        w = np.ascontiguousarray(np.broadcast_to(np.asarray(widths, dtype=np.float64), (len(n),)))
        e2t = self._e2t(env_to_track)
        _lib.check(self.lib.rk_set_tracks_from_control_points(self.h, _np_ptr(xy), _np_ptr(n), _np_ptr(w), len(n),
                                                              int(factor), _np_ptr(e2t)), self.h, 'set_tracks')

    def set_tracks_from_waypoints(self, waypoints, widths, env_to_track=None):
        n = np.array([len(c) for c in waypoints], dtype=np.int32)
        xy = np.ascontiguousarray(np.concatenate([np.asarray(c, dtype=np.float64) for c in waypoints]))
        w = np.ascontiguousarray(np.broadcast_to(np.asarray(widths, dtype=np.float64), (len(n),)))
        e2t = self._e2t(env_to_track)
        _lib.check(self.lib.rk_set_tracks_from_waypoints(self.h, _np_ptr(xy), _np_ptr(n), _np_ptr(w), len(n),
                                                         _np_ptr(e2t)), self.h, 'set_tracks')

    def generate_tracks(self, seed, n_tracks, factor=30, width_lo=6.0, width_mod=4, env_to_track=None):
        e2t = self._e2t(env_to_track)
        _lib.check(self.lib.rk_generate_tracks(self.h, int(seed), int(n_tracks), int(factor), float(width_lo),
                                               int(width_mod), _np_ptr(e2t)), self.h, 'generate_tracks')

    @property
    def num_tracks(self):
        return int(self.lib.rk_num_tracks(self.h))

    def get_track(self, track_id):
        """dict with waypoints, normals, left/right boundary, width, ... (float64, host)."""
        meta = np.zeros(6)
        _lib.check(self.lib.rk_get_track(self.h, int(track_id), _np_ptr(meta), None, None, None, None, None, None),
                   self.h, 'get_track')
        n = int(meta[0])
        wp, nrm, left, right = (np.zeros((n, 2)) for _ in range(4))
        ctrl = np.zeros((256, 2))
        nc = C.c_int32(0)
        _lib.check(self.lib.rk_get_track(self.h, int(track_id), _np_ptr(meta), _np_ptr(wp), _np_ptr(nrm),
                                         _np_ptr(left), _np_ptr(right), _np_ptr(ctrl), C.byref(nc)),
                   self.h, 'get_track')
        return dict(waypoints=wp, normals=nrm, left_boundary=left, right_boundary=right,
                    control_points=ctrl[:nc.value].copy(), track_width=meta[1], max_track_distance=meta[2],
                    start_pos=(meta[3], meta[4], meta[5]))

    # ---- env --------------------------------------------------------------
    def reset(self, mask=None, start_slot=None):
        """Reset all (or masked) environments; returns the obs tensor [E,A,D]."""
        _lib.check(self.lib.rk_reset(self.h, _ptr(mask), _ptr(start_slot), _ptr(self.obs), self.layout,
                                     self._stream()), self.h, 'rk_reset')
        return self.obs

    def step(self, start_slot=None, env_range=None):
        """One step of all environments (or of env_range = (begin, end)) on
        self.actions; outputs land in the pre-allocated tensors (obs, reward,
        terminated, truncated, done, ep_*)."""
        self._io.start_slot = start_slot.data_ptr() if start_slot is not None else None
        self._io.env_begin, self._io.env_count = (env_range[0], env_range[1] - env_range[0]) if env_range else (0, 0)
        _lib.check(self.lib.rk_step(self.h, C.byref(self._io), self._stream()), self.h, 'rk_step')

    def observe(self):
        _lib.check(self.lib.rk_observe(self.h, _ptr(self.obs), self.layout, self._stream()), self.h, 'rk_observe')
        return self.obs

    def set_speed_weight(self, w):
        _lib.check(self.lib.rk_set_speed_weight(self.h, float(w)), self.h, 'set_speed_weight')

    def set_seed(self, seed):
        _lib.check(self.lib.rk_set_seed(self.h, int(seed) & (2 ** 64 - 1)), self.h, 'set_seed')

    def get_state(self):
        E, A = self.E, self.A
        car_f = np.zeros((E, A, 6))
        car_i = np.zeros((E, A, 4), dtype=np.int32)
        env_i = np.zeros((E, 3), dtype=np.int32)
        env_f = np.zeros(E)
        _lib.check(self.lib.rk_get_state(self.h, _np_ptr(car_f), _np_ptr(car_i), _np_ptr(env_i), _np_ptr(env_f)),
                   self.h, 'get_state')
        return dict(car_f64=car_f, car_i32=car_i, env_i32=env_i, env_f64=env_f)

    def set_state(self, car_f64=None, car_i32=None, env_i32=None, env_f64=None):
        cf = None if car_f64 is None else np.ascontiguousarray(car_f64, dtype=np.float64)
        ci = None if car_i32 is None else np.ascontiguousarray(car_i32, dtype=np.int32)
        ei = None if env_i32 is None else np.ascontiguousarray(env_i32, dtype=np.int32)
        ef = None if env_f64 is None else np.ascontiguousarray(env_f64, dtype=np.float64)
        _lib.check(self.lib.rk_set_state(self.h, _np_ptr(cf), _np_ptr(ci), _np_ptr(ei), _np_ptr(ef)),
                   self.h, 'set_state')


# ---- rollout-side kernels ---------------------------------------------------
def gae(rewards, values, dones, next_value, next_done, gamma, lam, adv=None, ret=None):
    """PPO.compute_advantages (agent/ppo.py:134-154) as one kernel.  All [T,E]
    float32 CUDA tensors; next_value [E]; next_done [E] (bool or float)."""
    lib = _lib.load()
    T, E = rewards.shape
    nd = next_done.to(torch.float32)
    rewards, values, dones, next_value = (t.contiguous() for t in (rewards, values, dones, next_value))
    adv = torch.empty_like(rewards) if adv is None else adv
    ret = torch.empty_like(rewards) if ret is None else ret
    stream = C.c_void_p(torch.cuda.current_stream(rewards.device).cuda_stream)
    _lib.check(lib.rk_gae(_ptr(rewards), _ptr(values), _ptr(dones), _ptr(next_value), _ptr(nd), float(gamma),
                          float(lam), T, E, _ptr(adv), _ptr(ret), stream), None, 'rk_gae')
    return adv, ret


PARAM_ORDER = ['actor_mu.0.weight', 'actor_mu.0.bias', 'actor_mu.2.weight', 'actor_mu.2.bias',
               'actor_mu.4.weight', 'actor_mu.4.bias', 'log_std',
               'critic.0.weight', 'critic.0.bias', 'critic.2.weight', 'critic.2.bias',
               'critic.4.weight', 'critic.4.bias']


_TRANSPOSED = {'actor_mu.0.weight', 'actor_mu.2.weight', 'critic.0.weight', 'critic.2.weight'}


def flatten_agent(state_dict, out=None):
    """Agent state_dict (agent/ppo.py:11-37) -> the packed float32 block
    rk_policy_act expects (see racing_b200.h): PARAM_ORDER, hidden-layer weights
    transposed to [in][out], zero padded to a multiple of 4 floats."""
    parts = []
    for k in PARAM_ORDER:
        v = state_dict[k].detach().to(torch.float32)
        parts.append((v.t().contiguous() if k in _TRANSPOSED else v).reshape(-1))
    flat = torch.cat(parts)
    pad = (-flat.numel()) % 4
    if pad:
        flat = torch.cat([flat, flat.new_zeros(pad)])
    if out is not None:
        out.copy_(flat)
        return out
    return flat.contiguous()


def policy_act(params, obs, action_out, seed, counter, logprob=None, value=None, mean=None):
    """Fused Agent.get_action_and_value(obs) (action=None).  obs: [B, D] view of
    a CUDA float32 tensor whose rows are obs.stride(0) apart (last dim
    contiguous); action_out: [B, 2] view likewise.  params=None -> uniform
    Box([-1,0],[1,1]) actions (SelfPlayWrapper with an empty pool)."""
    lib = _lib.load()
    B = action_out.shape[0]
    assert action_out.stride(-1) == 1
    if params is not None:
        assert obs.stride(-1) == 1 and obs.dtype == torch.float32
        obs_ptr, obs_stride, obs_dim = _ptr(obs), obs.stride(0), obs.shape[-1]
    else:
        obs_ptr, obs_stride, obs_dim = None, 0, 0
    stream = C.c_void_p(torch.cuda.current_stream(action_out.device).cuda_stream)
    _lib.check(lib.rk_policy_act(_ptr(params), obs_dim, obs_ptr, obs_stride, B, int(seed), int(counter),
                                 _ptr(action_out), action_out.stride(0), _ptr(logprob), _ptr(value), _ptr(mean),
                                 stream), None, 'rk_policy_act')


POLICY_BLOCK = 256   # RK_POLICY_BLOCK


def policy_act_pool(params_pool, block_policy, block_len, obs, action_out, seed, counter, logprob=None, value=None,
                    mean=None):
    """policy_act against a pool: params_pool [P, n_packed] (rows = flatten_agent blocks, row stride
    a multiple of 4 floats), block_policy int32 [ceil(B / block_len)] = pool row driving each block of
    `block_len` consecutive samples (a multiple of POLICY_BLOCK)."""
    lib = _lib.load()
    B = action_out.shape[0]
    assert action_out.stride(-1) == 1 and obs.stride(-1) == 1 and obs.dtype == torch.float32
    assert params_pool.dim() == 2 and params_pool.stride(1) == 1 and block_policy.dtype == torch.int32
    assert block_policy.numel() * block_len >= B
    stream = C.c_void_p(torch.cuda.current_stream(action_out.device).cuda_stream)
    _lib.check(lib.rk_policy_act_pool(_ptr(params_pool), params_pool.stride(0), _ptr(block_policy), int(block_len),
                                      obs.shape[-1], _ptr(obs), obs.stride(0), B, int(seed), int(counter),
                                      _ptr(action_out), action_out.stride(0), _ptr(logprob), _ptr(value), _ptr(mean),
                                      stream), None, 'rk_policy_act_pool')


def gather_minibatch(idx, src, dst):
    """dst[k] = src[idx[k]] for the six per-sample arrays of a PPO minibatch, one
    launch.  src/dst: (obs [.,D], actions [.,2], logprobs, advantages, returns,
    values), contiguous float32 CUDA tensors; idx: int64 [n]."""
    lib = _lib.load()
    obs = src[0]
    stream = C.c_void_p(torch.cuda.current_stream(obs.device).cuda_stream)
    _lib.check(lib.rk_gather_minibatch(_ptr(idx), idx.numel(), obs.shape[1], *[_ptr(t) for t in src],
                                       *[_ptr(t) for t in dst], stream), None, 'rk_gather_minibatch')


def ppo_loss_grad(mu, v, act, old_logp, adv, ret, v_old, log_std, adv_mean, adv_std, clip_coef, vf_coef, dmu, dv,
                  kl_sum):
    """Fused d(loss)/d(mu, v) of the PPO minibatch loss + KL-sum accumulation."""
    lib = _lib.load()
    stream = C.c_void_p(torch.cuda.current_stream(mu.device).cuda_stream)
    _lib.check(lib.rk_ppo_loss_grad(_ptr(mu), _ptr(v), _ptr(act), _ptr(old_logp), _ptr(adv), _ptr(ret), _ptr(v_old),
                                    _ptr(log_std), _ptr(adv_mean), _ptr(adv_std), mu.shape[0], float(clip_coef),
                                    float(vf_coef), _ptr(dmu), _ptr(dv), _ptr(kl_sum), stream), None,
               'rk_ppo_loss_grad')


ADV_STAT_BLOCKS = 128    # RK_ADV_STAT_BLOCKS
PPO_MAX_OBS_DIM = 20     # RK_PPO_MAX_OBS_DIM


def ppo_adv_stats(idx, adv, part):
    """part[b] = (sum, sum of squares) in float64 of block b's share of adv[idx]
    (idx None: adv itself); part: float64 [ADV_STAT_BLOCKS, 2]."""
    lib = _lib.load()
    n = adv.numel() if idx is None else idx.numel()
    stream = C.c_void_p(torch.cuda.current_stream(adv.device).cuda_stream)
    _lib.check(lib.rk_ppo_adv_stats(_ptr(idx), _ptr(adv), n, _ptr(part), stream), None, 'rk_ppo_adv_stats')


class PpoMinibatchGrad:
    """Forward + PPO loss + backward of the Agent's two MLPs as one kernel
    (rk_ppo_minibatch_grad).  Holds the argument block, the scratch buffer and the
    flat gradient; `params` are the twelve weight/bias tensors in
    `Agent.parameters()` order (actor_mu.{0,2,4}, critic.{0,2,4}), read in place."""

    def __init__(self, params, log_std, obs_dim, clip_coef, vf_coef, tensor_cores=False):
        lib = _lib.load()
        params = list(params)
        if len(params) != 12 or obs_dim > PPO_MAX_OBS_DIM:
            raise ValueError('PpoMinibatchGrad: needs the 12 Agent parameters and obs_dim <= %d' % PPO_MAX_OBS_DIM)
        dev = params[0].device
        self._keep = (params, log_std)
        self.device = dev
        n_total = sum(p.numel() for p in params)
        # gradient and (float32) KL sum share one buffer: a single all-reduce carries both across ranks
        self.grad_and_kl = torch.zeros(n_total + 1, device=dev)
        self.flat_grad, self.kl_f32 = self.grad_and_kl[:n_total], self.grad_and_kl[n_total:]
        self.kl_sum = torch.zeros((), device=dev, dtype=torch.float64)
        self.adv_part = torch.zeros(ADV_STAT_BLOCKS, 2, device=dev, dtype=torch.float64)
        self._adv_all = None
        nbytes = int(lib.rk_ppo_grad_workspace_bytes())
        self.workspace = torch.empty(nbytes, device=dev, dtype=torch.uint8)
        io = _lib.RkPpoGradIO()
        io.struct_size = C.sizeof(_lib.RkPpoGradIO)
        io.obs_dim = int(obs_dim)
        for k, p in enumerate(params):
            if not (p.is_cuda and p.is_contiguous() and p.dtype == torch.float32):
                raise ValueError('PpoMinibatchGrad: parameters must be contiguous float32 CUDA tensors')
            io.params[k] = p.data_ptr()
        io.log_std = log_std.data_ptr()
        io.adv_part = self.adv_part.data_ptr()
        io.clip_coef, io.vf_coef = float(clip_coef), float(vf_coef)
        io.workspace, io.workspace_bytes = self.workspace.data_ptr(), nbytes
        io.flat_grad, io.kl_sum = self.flat_grad.data_ptr(), self.kl_sum.data_ptr()
        io.kl_sum_f32 = self.kl_f32.data_ptr()
        # 0: fp32 FMA kernel; 1 / True: per-sample products on tcgen05; 2: the weight gradients on tcgen05 as well
        io.tensor_cores = int(tensor_cores) if not isinstance(tensor_cores, bool) else (1 if tensor_cores else 0)
        self.tensor_cores = io.tensor_cores
        self.io = io

    def grad_views(self):
        """Views of flat_grad shaped like the parameters (to be installed as .grad)."""
        out, off = [], 0
        for p in self._keep[0]:
            out.append(self.flat_grad[off:off + p.numel()].view_as(p))
            off += p.numel()
        return out

    def stats(self, idx, adv):
        ppo_adv_stats(idx, adv, self.adv_part)
        self.io.adv_part = self.adv_part.data_ptr()
        return self.adv_part

    def stats_epoch(self, perm, mb, adv):
        """Advantage statistics of ALL minibatches perm[k*mb:(k+1)*mb] of an epoch: float64
        [n_minibatches, ADV_STAT_BLOCKS, 2] (to be all-reduced once); select with use_stats(k)."""
        nmb = perm.numel() // mb
        if self._adv_all is None or self._adv_all.shape[0] != nmb:
            self._adv_all = torch.zeros(nmb, ADV_STAT_BLOCKS, 2, device=self.device, dtype=torch.float64)
        for k in range(nmb):
            ppo_adv_stats(perm[k * mb:(k + 1) * mb], adv, self._adv_all[k])
        return self._adv_all

    def use_stats(self, k):
        self.io.adv_part = self._adv_all[k].data_ptr()

    def __call__(self, idx, obs, act, old_logp, adv, ret, val, n_global=None):
        """flat_grad, kl_sum <- gradient of the minibatch rows idx (None: all rows);
        the advantage statistics must already be in self.adv_part (`stats`)."""
        lib = _lib.load()
        io = self.io
        io.n = int(obs.shape[0] if idx is None else idx.numel())
        io.obs_stride = int(obs.stride(0))    # rows may be padded (e.g. to 20 floats: 128-bit gathers)
        io.n_global = float(n_global if n_global is not None else io.n)
        io.obs, io.act, io.old_logp = obs.data_ptr(), act.data_ptr(), old_logp.data_ptr()
        io.adv, io.ret, io.val = adv.data_ptr(), ret.data_ptr(), val.data_ptr()
        io.idx = None if idx is None else idx.data_ptr()
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(lib.rk_ppo_minibatch_grad(C.byref(io), stream), None, 'rk_ppo_minibatch_grad')
        return self.flat_grad, self.kl_sum


class PpoAdamStep:
    """Gradient clipping + Adam + the KL early stop of one minibatch as one kernel
    (rk_ppo_adam_step), in place on a torch.optim.Adam's own state tensors.
    `state` is a device int32[2] = (stopped, steps applied); a latched stop turns
    every later call into a no-op until `reset()`."""

    def __init__(self, optimizer, params, flat_grad, kl_sum, max_grad_norm, kl_target, world=1, kl_sum_f32=None):
        _lib.load()
        group = optimizer.param_groups[0]
        if group.get('weight_decay', 0) or group.get('amsgrad', False) or group.get('maximize', False):
            raise ValueError('PpoAdamStep: plain Adam only (no weight decay / amsgrad / maximize)')
        if not isinstance(group['lr'], torch.Tensor) or not group['lr'].is_cuda:
            raise ValueError('PpoAdamStep: needs a device-tensor learning rate (capturable Adam)')
        params = list(params)
        dev = params[0].device
        self.device = dev
        self.state = torch.zeros(4, dtype=torch.int32, device=dev)
        self.kl_at_stop = torch.zeros((), device=dev)
        io = _lib.RkAdamIO()
        io.struct_size = C.sizeof(_lib.RkAdamIO)
        io.world = int(world)
        keep = [flat_grad, kl_sum, group['lr']]
        for k, p in enumerate(params):
            st = optimizer.state[p]
            if not all(key in st for key in ('exp_avg', 'exp_avg_sq', 'step')) or not st['step'].is_cuda:
                raise ValueError('PpoAdamStep: optimizer state must exist on the device (run one step first)')
            io.params[k], io.numel[k] = p.data_ptr(), p.numel()
            io.exp_avg[k], io.exp_avg_sq[k], io.step[k] = (st['exp_avg'].data_ptr(), st['exp_avg_sq'].data_ptr(),
                                                         st['step'].data_ptr())
            keep += [p, st['exp_avg'], st['exp_avg_sq'], st['step']]
        io.flat_grad, io.lr, io.kl_sum = flat_grad.data_ptr(), group['lr'].data_ptr(), kl_sum.data_ptr()
        io.beta1, io.beta2 = float(group['betas'][0]), float(group['betas'][1])
        io.eps, io.max_grad_norm, io.kl_target = float(group['eps']), float(max_grad_norm), float(kl_target)
        io.state, io.kl_at_stop = self.state.data_ptr(), self.kl_at_stop.data_ptr()
        io.kl_sum_f32 = None if kl_sum_f32 is None else kl_sum_f32.data_ptr()
        keep.append(kl_sum_f32)
        self.io, self._keep = io, keep

    def reset(self):
        self.state.zero_()

    def __call__(self, n_global, kl_target=None):
        if kl_target is not None:
            self.io.kl_target = float(kl_target)
        self.io.n_global = float(n_global)
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(_lib.load().rk_ppo_adam_step(C.byref(self.io), stream), None, 'rk_ppo_adam_step')


def random_permutation(n, seed, counter, out=None, device=None):
    """int64 [n]: seeded pseudo-random permutation of 0..n-1 generated on the device
    without sorting (rk_random_permutation)."""
    lib = _lib.load()
    if out is None:
        out = torch.empty(n, dtype=torch.int64, device=device)
    stream = C.c_void_p(torch.cuda.current_stream(out.device).cuda_stream)
    _lib.check(lib.rk_random_permutation(int(seed) & (2 ** 64 - 1), int(counter), int(n), _ptr(out), stream), None,
               'rk_random_permutation')
    return out
