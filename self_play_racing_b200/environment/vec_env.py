"""`BatchedRacingVecEnv`: the vector-env surface the reference's PPO drives
(`gym.vector.SyncVectorEnv` of `RecordEpisodeStatistics`-wrapped envs; call
sites agent/ppo.py:70,88,114-130,230 and agent/self_play_ppo.py:19-29,49-50),
backed by ONE device-resident batch and one fused kernel per step.

Two faces:
  * Gymnasium face -- `reset()` / `step(actions)` with HOST numpy arrays, the
    exact return convention of SyncVectorEnv (obs float32 (E,D), reward float64
    (E,), terminated/truncated bool (E,), infos with "episode"/"_episode" only
    on steps where an episode ended; NEXT_STEP auto-reset).
  * device face -- `reset_device()` / `step_device(actions)` on CUDA tensors
    with no host synchronisation, used by the device-resident rollout.

Auto-reset follows gymnasium 1.x NEXT_STEP (the reference's default).  `autoreset='same_step'` resets inside the
terminal step and returns the reset observation; the terminal observation is then NOT available (`infos` carries no
`final_obs` / `final_info`), which is why it is not the default.

Environments that are `SelfPlayWrapper(MultiRacingEnv)` get the wrapper's
semantics (environment/wrappers.py:29-55): car 0 is the learner, car 1 is driven
by the frozen opponent snapshot (or uniform random actions when none is set),
whose inference is one fused kernel over all environments.
"""
from __future__ import annotations

import os
import time

import numpy as np
import torch

from .. import spaces
from ..backend import RacingBackend, flatten_agent, policy_act, policy_act_pool
from .multi_racing_env import MultiRacingEnv
from .racing_env import RacingEnv
from .track import Track
from .wrappers import SelfPlayWrapper


class _EnvProxy:
    """Element of `vec.envs`: attribute writes such as the speed_weight
    annealing of agent/ppo.py:256-258 reach the batch."""

    def __init__(self, vec, idx, spec):
        object.__setattr__(self, '_vec', vec)
        object.__setattr__(self, '_idx', idx)
        object.__setattr__(self, '_spec', spec)

    def __setattr__(self, name, value):
        if name == 'speed_weight':
            self._vec.set_speed_weight(value)
        else:
            object.__setattr__(self, name, value)

    def __getattr__(self, name):
        if name == 'speed_weight':
            return self._vec.speed_weight
        if name == 'track':
            return self._vec.track_of(self._idx)
        return getattr(self._spec, name)


class _EpisodeStats(dict):
    """infos["episode"] of a vector step: {'r', 'l', 't'} arrays over all environments, valid where infos["_episode"].
    'r' and 'l' are filled by the step kernel; 't' is computed when it is first read (an episode of length l that ends at
    step s was reset during step s - l, whose wall time the vector env remembers), so steps whose caller never looks at
    't' -- the reference's PPO loop reads only 'r' and 'l', agent/ppo.py:123-130 -- pay nothing for it."""

    def __init__(self, r, l, mask, step_no, step_times):
        super().__init__(r=r, l=l, t=None)
        self._lazy = (mask, step_no, step_times, step_times[step_no % len(step_times)])

    def _t(self):
        t = dict.__getitem__(self, 't')
        if t is None:
            mask, step_no, times, now = self._lazy
            idx = np.flatnonzero(mask)
            t = np.zeros(len(mask))
            t[idx] = np.round(now - times[(step_no - dict.__getitem__(self, 'l')[idx]) % len(times)], 6)
            dict.__setitem__(self, 't', t)
        return t

    def __getitem__(self, key):
        return self._t() if key == 't' else dict.__getitem__(self, key)

    def get(self, key, default=None):
        return self[key] if key in self else default

    def items(self):
        self._t()
        return dict.items(self)

    def values(self):
        self._t()
        return dict.values(self)

    def copy(self):
        self._t()
        return dict(self)


class BatchedRacingVecEnv:
    def __init__(self, env_fns, device=None, query='culled', autoreset='next_step', seed=0, copy=True,
                 want_info=False, pipeline_chunks=None):
        self.pipeline_chunks = pipeline_chunks
        envs = [fn() if callable(fn) else fn for fn in env_fns]
        if not envs:
            raise ValueError('need at least one environment')
        self.selfplay = isinstance(envs[0], SelfPlayWrapper)
        base = [e.env if isinstance(e, SelfPlayWrapper) else e for e in envs]
        first = base[0]
        if not isinstance(first, (RacingEnv, MultiRacingEnv)):
            raise TypeError(f'unsupported environment type {type(first).__name__}')
        for b in base:
            if type(b) is not type(first) or b.num_sensors != first.num_sensors or b.num_agents != first.num_agents:
                raise ValueError('all environments of a batch must share class, num_sensors and num_agents')
        if self.selfplay and first.num_agents != 2:
            raise ValueError('SelfPlayWrapper drives exactly one opponent (num_agents must be 2)')
        self._specs = envs
        self.kind = first.KIND
        self.num_envs = len(envs)
        self.num_agents = first.num_agents
        self.copy = copy
        # de-duplicate tracks: one device table per distinct (control points, width)
        keys, cps, widths, e2t = {}, [], [], np.zeros(self.num_envs, dtype=np.int32)
        for i, b in enumerate(base):
            k = (b.control_points.tobytes(), b.track_width)
            if k not in keys:
                keys[k] = len(cps)
                cps.append(b.control_points)
                widths.append(b.track_width)
            e2t[i] = keys[k]
        self.env_to_track = e2t
        sw = first.speed_weight if self.kind == 'single' else 8.0
        self.be = RacingBackend(self.num_envs, kind=self.kind, num_agents=self.num_agents,
                                num_sensors=first.num_sensors, device=device, autoreset=autoreset, query=query,
                                speed_weight=sw, seed=seed, want_info=want_info, agent_major=True)
        self.be.set_tracks_from_control_points(cps, widths, env_to_track=e2t)
        self._finish_init(first, seed)

    # -- alternative constructors for large synthetic batches -------------------
    @classmethod
    def synthetic(cls, kind, num_envs, n_tracks=16, num_agents=2, num_sensors=11, selfplay=True, device=None,
                  query='culled', autoreset='next_step', seed=0, copy=True, factor=30, width_lo=6.0, width_mod=4,
                  want_info=False, pipeline_chunks=None):
        """E environments over a device-generated procedural pool (BASELINE
        configs 2-5): env e runs on track e % n_tracks, widths width_lo + (t % width_mod)."""
        self = cls.__new__(cls)
        self.pipeline_chunks = pipeline_chunks
        self.kind = kind
        self.num_envs = int(num_envs)
        self.num_agents = 1 if kind == 'single' else int(num_agents)
        self.selfplay = bool(selfplay) and kind == 'multi' and self.num_agents == 2
        self.copy = copy
        self._specs = None
        # env e runs on track e % n_tracks; RK_B200_BLOCKED=1 assigns contiguous
        # blocks of environments to each track instead (what the opt-in staged launch prefers)
        if os.environ.get('RK_B200_BLOCKED') == '1':
            self.env_to_track = (np.arange(self.num_envs, dtype=np.int64) * n_tracks // self.num_envs).astype(np.int32)
        else:
            self.env_to_track = (np.arange(self.num_envs) % n_tracks).astype(np.int32)
        self.be = RacingBackend(self.num_envs, kind=kind, num_agents=self.num_agents, num_sensors=num_sensors,
                                device=device, autoreset=autoreset, query=query, seed=seed, want_info=want_info,
                                agent_major=True)
        self.be.generate_tracks(seed, n_tracks, factor=factor, width_lo=width_lo, width_mod=width_mod,
                                env_to_track=self.env_to_track)
        proto = RacingEnv(num_sensors=num_sensors) if kind == 'single' else \
            MultiRacingEnv(num_agents=self.num_agents, num_sensors=num_sensors)
        self._finish_init(proto, seed)
        return self

    def _finish_init(self, proto, seed):
        be = self.be
        if self.kind == 'single':
            self.single_observation_space = proto.observation_space
            self.single_action_space = proto.action_space
        else:
            self.single_observation_space = proto.observation_space['0']
            self.single_action_space = proto.action_space['0']
        self.observation_space = self.single_observation_space
        self.action_space = self.single_action_space
        E, D = self.num_envs, be.D
        self.seed = int(seed)
        self._opp_params = None     # flattened frozen opponent (device float32) or None -> random
        self._opp_pool = None       # (stacked params [P, n], block -> pool row, block length) after set_opponents
        self._opp_counter = 0
        self._obs_cur = be.obs
        self._tracks = {}
        # pinned host staging for the Gymnasium face
        self._h_actions = torch.zeros(E, 2, dtype=torch.float32).pin_memory()
        self._h_obs = torch.zeros(E, D, dtype=torch.float32).pin_memory()
        self._h_arena = torch.zeros(be.arena_host_bytes, dtype=torch.uint8).pin_memory()
        a, o = self._h_arena.numpy(), be.arena_offsets
        self._np_ep_return = a[o['ep_return']:o['ep_return'] + 8 * E].view(np.float64)
        self._np_ep_length = a[o['ep_length']:o['ep_length'] + 4 * E].view(np.int32)
        self._np_terminated = a[o['terminated']:o['terminated'] + E].view(np.bool_)
        self._np_truncated = a[o['truncated']:o['truncated'] + E].view(np.bool_)
        self._np_ep_mask = a[o['ep_mask']:o['ep_mask'] + E].view(np.bool_)
        self._np_reward = a[o['reward64']:o['reward64'] + 8 * E].view(np.float64)
        # Optional Gymnasium-face pipeline: the batch is stepped in chunks on side streams so that the
        # device->host copy of one chunk's observations overlaps the kernel of the next.  Measured on
        # B200 (profiles/r01_e2e_pipeline.log) the extra launches cost more than the overlap wins at
        # 65,536 envs, so the default is a single chunk.
        if self.pipeline_chunks is None:
            self.pipeline_chunks = int(os.environ.get('RK_B200_PIPELINE_CHUNKS', 1))
        self._streams = [torch.cuda.Stream(device=be.device) for _ in range(self.pipeline_chunks)] \
            if self.pipeline_chunks > 1 else []
        # default Gymnasium-face path: one C call per step with host buffers (rk_step_host), the batch cut
        # into `host_chunks` ranges so that copies overlap kernels; 0 falls back to the torch-level path
        import ctypes as C
        from .. import _lib
        self.host_chunks = int(os.environ.get('RK_B200_HOST_CHUNKS', 4 if E >= 16384 else 1))
        self._host_io = _lib.RkHostIO(struct_size=C.sizeof(_lib.RkHostIO), n_chunks=max(self.host_chunks, 1),
                                      actions=self._h_actions.data_ptr(), obs=self._h_obs.data_ptr(),
                                      arena_host=self._h_arena.data_ptr(), arena_dev=be.arena.data_ptr(),
                                      arena_bytes=be.arena_host_bytes, selfplay=1 if self.selfplay else 0,
                                      reserved0=int(os.environ.get('RK_B200_ZEROCOPY_OBS', '7')), opponent_params=None, seed=self.seed ^ 0x5eed0bb, counter=0)
        self._step_no, self._step_times = 0, np.full(4096, time.perf_counter())   # wall time of the last 4096 steps (episodes last <= 3000)
        # numpy views and ctypes references of the per-step call, built once (the Gymnasium face is host-paced)
        self._np_actions, self._np_obs = self._h_actions.numpy(), self._h_obs.numpy()
        self._pinned_acts = {}      # address -> pinned tensor handed out by pinned_action_buffers()
        self._io_ref, self._host_ref = C.byref(be._io), C.byref(self._host_io)
        self.h2d_bytes_per_step = self._h_actions.numel() * 4
        self.d2h_bytes_per_step = self._h_obs.numel() * 4 + be.arena_host_bytes

    # ---- reference-facing attributes ------------------------------------------
    @property
    def envs(self):
        # built once: reference-style loops `for i in range(num_envs): setattr(vec.envs[i], ...)`
        # (agent/ppo.py:256-258) would otherwise create E proxies per access
        if getattr(self, '_proxies', None) is None:
            specs = self._specs or [None] * self.num_envs
            self._proxies = [_EnvProxy(self, i, s) for i, s in enumerate(specs)]
        return self._proxies

    @property
    def speed_weight(self):
        return getattr(self, '_speed_weight', 8.0)

    def set_speed_weight(self, value):
        self._speed_weight = float(value)
        self.be.set_speed_weight(value)

    def track_of(self, env_idx):
        t = int(self.env_to_track[env_idx])
        if t not in self._tracks:
            self._tracks[t] = Track(self.be.get_track(t))
        return self._tracks[t]

    def set_opponent(self, opponent_policy):
        """SelfPlayWrapper.set_opponent for every environment at once.  Accepts
        an Agent module, a state_dict, an already flattened parameter vector, or
        None (uniform random opponent, wrappers.py:30-32)."""
        self._opp_pool = None
        if opponent_policy is None:
            self._opp_params = None
            return
        if isinstance(opponent_policy, torch.Tensor):
            flat = opponent_policy
        else:
            sd = opponent_policy.state_dict() if hasattr(opponent_policy, 'state_dict') else opponent_policy
            flat = flatten_agent(sd)
        self._opp_params = flat.to(self.be.device, torch.float32).contiguous()

    def set_opponents(self, policies, block_policy=None, block_len=256, seed=None):
        """Several frozen opponents at once (a superset of the reference's one
        opponent per update): the environments are split into blocks
        of `block_len` consecutive envs and block k plays against
        policies[block_policy[k]] (default: drawn uniformly, as select_opponent
        does per update, self_play_ppo.py:40-44).  One inference launch serves
        all of them (rk_policy_act_pool)."""
        flats = []
        for pol in policies:
            if isinstance(pol, torch.Tensor):
                flats.append(pol)
            else:
                flats.append(flatten_agent(pol.state_dict() if hasattr(pol, 'state_dict') else pol))
        pool = torch.stack([f.to(self.be.device, torch.float32) for f in flats]).contiguous()
        from ..backend import POLICY_BLOCK
        if block_len <= 0 or block_len % POLICY_BLOCK != 0:
            raise ValueError(f'set_opponents: block_len must be a positive multiple of {POLICY_BLOCK} '
                             f'(one inference CTA serves {POLICY_BLOCK} consecutive environments), got {block_len}')
        n_blocks = (self.num_envs + block_len - 1) // block_len
        if block_policy is None:
            rs = np.random.RandomState(self.seed if seed is None else seed)
            block_policy = rs.randint(0, len(flats), size=n_blocks)
        block_policy = torch.as_tensor(np.asarray(block_policy, dtype=np.int32)).to(self.be.device)
        if block_policy.numel() != n_blocks or int(block_policy.max()) >= len(flats) or int(block_policy.min()) < 0:
            raise ValueError('set_opponents: block_policy needs one valid pool index per block of envs')
        self._opp_params = pool[0]
        self._opp_pool = (pool, block_policy, int(block_len))

    # ---- device face -------------------------------------------------------------
    @property
    def obs_device(self):
        """Learner observation [E, D] (car 0), a view of the kernel's output."""
        return self.be.obs[0]

    def reset_device(self, start_slot=None):
        self.be.reset(start_slot=start_slot)
        self._obs_cur = self.be.obs
        return self.be.obs[0]

    def _opponent_act(self, obs=None, actions=None):
        """SelfPlayWrapper.step's opponent half (wrappers.py:30-39) for every env:
        car 1's action from the frozen snapshot on car 1's latest observation."""
        be = self.be
        obs = be.obs if obs is None else obs
        actions = be.actions if actions is None else actions
        self._opp_counter += 1
        if getattr(self, '_opp_pool', None) is not None:
            pool, block_policy, block_len = self._opp_pool
            policy_act_pool(pool, block_policy, block_len, obs[1], actions[1], seed=self.seed ^ 0x5eed0bb,
                            counter=self._opp_counter)
            return
        policy_act(self._opp_params, obs[1] if self._opp_params is not None else None, actions[1],
                   seed=self.seed ^ 0x5eed0bb, counter=self._opp_counter)

    def step_into(self, actions, obs_out, reward_out, done_out, start_slot=None, opponent_actions=None):
        """Zero-copy rollout step: `actions` [A,E,2] already holds the learner's
        action in actions[0]; the kernel writes the successor observation
        [A,E,D], reward [A,E] and done [E] (float32) straight into the caller's
        rollout-buffer slots.  No host synchronisation.  `opponent_actions` [E,2]
        replaces the opponent's inference (replay of recorded trajectories)."""
        be, io = self.be, self.be._io
        if self.selfplay and opponent_actions is not None:
            actions[1].copy_(opponent_actions)
        elif self.selfplay:
            self._opponent_act(self._obs_cur, actions)
        io.actions, io.obs = actions.data_ptr(), obs_out.data_ptr()
        io.reward_f32, io.done_f32 = reward_out.data_ptr(), done_out.data_ptr()
        try:
            be.step(start_slot=start_slot)
        finally:
            be._bind_io()
        self._obs_cur = obs_out

    def rollout_into(self, buf, learner_params, learner_seed, learner_counter0, T):
        """PPO.collect_rollout's loop (agent/ppo.py:104-120) as ONE native call (rk_rollout): per step one inference
        launch (learner + frozen opponent together) and the step kernel, written straight into the [T(+1), ...] rollout
        buffers `buf` (PPO.alloc_buffers).  Bit-identical to T x (policy_act, step_into); no host synchronisation."""
        import ctypes as C
        from .. import _lib
        be = self.be
        r = _lib.RkRolloutIO(struct_size=C.sizeof(_lib.RkRolloutIO), T=int(T), selfplay=1 if self.selfplay else 0)
        r.learner_params = learner_params.data_ptr()
        r.learner_seed, r.learner_counter0 = int(learner_seed) & (2 ** 64 - 1), int(learner_counter0)
        keep = [learner_params]
        if self.selfplay:
            pool = getattr(self, '_opp_pool', None)
            if pool is not None:
                params_pool, block_policy, block_len = pool
                r.opponent_params, r.block_policy = params_pool.data_ptr(), block_policy.data_ptr()
                r.block_len, r.pool_stride = int(block_len), int(params_pool.stride(0))
                keep += [params_pool, block_policy]
            elif self._opp_params is not None:
                r.opponent_params = self._opp_params.data_ptr()
                keep.append(self._opp_params)
            r.opponent_seed, r.opponent_counter0 = (self.seed ^ 0x5eed0bb) & (2 ** 64 - 1), self._opp_counter + 1
            self._opp_counter += int(T)
            # car 1's observation for step 0 is the environment's CURRENT one (e.g. fresh after update_opponent's
            # reset), not slot 0 of the buffer, which carries the learner's previous next_obs
            if self._obs_cur.data_ptr() != buf['obs'][0].data_ptr():
                r.opponent_obs0 = self._obs_cur[1].data_ptr()
                keep.append(self._obs_cur)
        for name in ('obs', 'actions', 'logprobs', 'values', 'rewards', 'dones'):
            t = buf[name]
            if not (t.is_cuda and t.is_contiguous() and t.dtype == torch.float32):
                raise ValueError(f'rollout_into: buffer {name!r} must be a contiguous float32 CUDA tensor')
            setattr(r, name, t.data_ptr())
        be._io.start_slot = None
        be._io.env_begin, be._io.env_count = 0, 0
        _lib.check(be.lib.rk_rollout(be.h, C.byref(be._io), C.byref(r), be._stream()), be.h, 'rk_rollout')
        self._obs_cur = buf['obs'][T]
        del keep

    def step_device(self, actions, start_slot=None):
        """actions: CUDA float32 [E, 2] (learner).  Returns views (obs [E,D],
        reward float32 [E], done float32 [E]) valid until the next step; no
        host synchronisation.  `done` = terminated | truncated, which is what
        SelfPlayWrapper reports as `terminated` and PPO uses as next_done."""
        be = self.be
        if actions.data_ptr() != be.actions[0].data_ptr():
            be.actions[0].copy_(actions)
        if self.selfplay:
            self._opponent_act()
        be.step(start_slot=start_slot)
        self._obs_cur = be.obs
        return be.obs[0], be.reward[0], be.done_f32

    def pinned_action_buffers(self, n=1):
        """`n` float32 [num_envs, 2] numpy arrays in page-locked host memory.  An array from here that is passed to
        `step` is read by the step kernel in place (zero-copy over PCIe): the 0.5 MB numpy -> staging copy that any
        other array costs per step disappears.  The caller (a host-side policy) writes its actions into them."""
        out = []
        for _ in range(int(n)):
            t = torch.zeros(self.num_envs, 2, dtype=torch.float32).pin_memory()
            a = t.numpy()
            self._pinned_acts[a.ctypes.data] = t
            out.append(a)
        return out

    # ---- Gymnasium face ------------------------------------------------------------
    def reset(self, seed=None, options=None):
        """SyncVectorEnv.reset.  `seed` re-keys the Philox stream of the start-grid shuffles (the
        reference's envs ignore their seed -- RacingEnv.reset only forwards it to gym.Env, and the grid
        shuffle reads the global np.random stream -- so any value is accepted and none changes the
        single-car envs); `options` must be empty."""
        if options:
            raise ValueError('BatchedRacingVecEnv.reset: options are not supported')
        if seed is not None:
            if isinstance(seed, (list, tuple, np.ndarray)):
                seed = int(np.asarray(seed).reshape(-1)[0])
            self.be.set_seed(int(seed))
        obs = self.reset_device()
        self._h_obs.copy_(obs, non_blocking=True)
        torch.cuda.current_stream(self.be.device).synchronize()
        self._step_no = 0
        self._step_times[0] = time.perf_counter()
        out = self._h_obs.numpy()
        return (out.copy() if self.copy else out), {}

    def step(self, actions, start_slot=None):
        """SyncVectorEnv.step.  `start_slot` (int [E,A], optional, not part of the
        gymnasium signature) injects the grid slots used by auto-resets this
        step; by default they come from the backend's Philox stream."""
        be = self.be
        h_act = None
        if self._pinned_acts and isinstance(actions, np.ndarray) and actions.dtype == np.float32 \
                and actions.flags.c_contiguous and actions.size == 2 * self.num_envs:
            h_act = self._pinned_acts.get(actions.ctypes.data)    # one of pinned_action_buffers(): used in place
        if h_act is None:
            np.copyto(self._np_actions, np.asarray(actions, dtype=np.float32).reshape(self.num_envs, 2))
            h_act = self._h_actions
        if self.host_chunks > 0 and self._opp_pool is None:
            self._host_io.actions = h_act.data_ptr()
            self._step_host(start_slot)
        elif self.pipeline_chunks > 1 and start_slot is None and self._opp_pool is None:
            if h_act is not self._h_actions:
                self._h_actions.copy_(h_act)
            self._step_pipelined()
        else:
            be.actions[0].copy_(h_act, non_blocking=True)
            if self.selfplay:
                self._opponent_act()
            if start_slot is not None:
                start_slot = torch.as_tensor(np.ascontiguousarray(start_slot, dtype=np.int32)).to(be.device)
            be.step(start_slot=start_slot)
            self._h_obs.copy_(be.obs[0], non_blocking=True)
            self._h_arena.copy_(be.arena[:be.arena_host_bytes], non_blocking=True)
            torch.cuda.current_stream(be.device).synchronize()
        self._obs_cur = be.obs
        obs, rew = self._np_obs, self._np_reward
        term, trunc = self._np_terminated, self._np_truncated
        if self.selfplay:  # wrappers.py:52: the wrapper reports dones["__all__"] as `terminated`
            term = term | trunc
        infos = {}
        self._step_no += 1
        self._step_times[self._step_no % len(self._step_times)] = time.perf_counter()
        if self._np_ep_mask.any():
            # RecordEpisodeStatistics' keys: return, length and elapsed wall time of the episode.  'r' and 'l' come from
            # the device; 't' (host clock, from the environment's previous reset, as the wrapper measures it) is derived
            # on first access from the episode length and the wall time of the step that reset the environment.
            mask = self._np_ep_mask.copy() if self.copy else self._np_ep_mask
            r = self._np_ep_return.copy() if self.copy else self._np_ep_return
            l = self._np_ep_length.copy() if self.copy else self._np_ep_length
            infos['episode'] = _EpisodeStats(r, l, mask, self._step_no, self._step_times)
            infos['_episode'] = mask
        if self.copy:
            return obs.copy(), rew.copy(), term.copy(), trunc.copy(), infos
        return obs, rew, term, trunc, infos

    def _step_host(self, start_slot=None):
        # The whole Gymnasium-face step as ONE C call (rk_step_host): chunked host->device copy of the
        # actions, opponent inference, step kernel and device->host copy of the observations on the
        # library's internal streams, then the small per-env results; returns when the host buffers are ready.
        from .. import _lib
        be = self.be
        self._opp_counter += 1
        if start_slot is not None:
            start_slot = torch.as_tensor(np.ascontiguousarray(start_slot, dtype=np.int32)).to(be.device)
        be._io.start_slot = start_slot.data_ptr() if start_slot is not None else None
        be._io.env_begin, be._io.env_count = 0, 0
        h = self._host_io
        h.counter = self._opp_counter
        h.opponent_params = self._opp_params.data_ptr() if self._opp_params is not None else None
        rc = be.lib.rk_step_host(be.h, self._io_ref, self._host_ref, be._stream())
        if rc:
            _lib.check(rc, be.h, 'rk_step_host')

    def _step_pipelined(self):
        # One logical step as `pipeline_chunks` range launches on side streams: chunk i's
        # host->device actions, opponent inference, step kernel and device->host observation
        # copy are ordered on stream i; the small per-env results follow once on the last one.
        be, E = self.be, self.num_envs
        cur = torch.cuda.current_stream(be.device)
        n = self.pipeline_chunks
        bounds = [E * i // n for i in range(n + 1)]
        self._opp_counter += 1
        for i, s in enumerate(self._streams):
            lo, hi = bounds[i], bounds[i + 1]
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                be.actions[0, lo:hi].copy_(self._h_actions[lo:hi], non_blocking=True)
                if self.selfplay:
                    policy_act(self._opp_params, be.obs[1, lo:hi] if self._opp_params is not None else None,
                               be.actions[1, lo:hi], seed=self.seed ^ 0x5eed0bb, counter=self._opp_counter * 64 + i)
                be.step(env_range=(lo, hi))
                self._h_obs[lo:hi].copy_(be.obs[0, lo:hi], non_blocking=True)
        last = self._streams[-1]
        for s in self._streams[:-1]:
            last.wait_stream(s)
        with torch.cuda.stream(last):
            self._h_arena.copy_(be.arena[:be.arena_host_bytes], non_blocking=True)
        last.synchronize()
        cur.wait_stream(last)

    def close(self):
        if getattr(self, 'be', None) is not None:
            self.be.close()
            self.be = None
