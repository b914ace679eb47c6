/*
 * racing_b200.h -- C ABI of the B200-native batched racing backend.
 *
 * The reference (LucasHJin/self-play-racing) is pure Python and has no FFI; the
 * boundary its hot path sits behind is the Gymnasium surface.
 * Each entry point below names the reference interface it replaces
 * (paths relative to the reference root).  INTEGRATION.md shows the ctypes
 * binding a maintainer of the reference would add.
 *
 * Conventions
 *  - every function returns 0 on success, non-zero on failure and never
 *    throws; rk_last_error() describes the last failure on that handle (or the
 *    last create failure when the handle is NULL);
 *  - "dev" pointers are DEVICE pointers owned by the caller (e.g. torch
 *    tensors); "host" pointers are ordinary host memory;
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default
 *    stream); all step/reset/gae/policy calls are asynchronous on it;
 *  - one handle per GPU; a handle is not thread-safe, distinct handles are;
 *  - E = num_envs, A = num_agents, R = num_sensors,
 *    D = obs_dim = R + 4 (single) or R + 4 + 4*(A-1) (multi).
 */
#ifndef RACING_B200_H
#define RACING_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define RK_API __attribute__((visibility("default")))
#else
#define RK_API
#endif

#define RK_ABI_VERSION 1
#define RK_MAX_AGENTS 8
#define RK_MAX_SENSORS 64

typedef struct rk_env_s* rk_handle;

/* env_kind: which reference class' rules apply */
enum { RK_ENV_SINGLE = 0,   /* environment/racing_env.py  RacingEnv          */
       RK_ENV_MULTI = 1 };  /* environment/multi_racing_env.py MultiRacingEnv */
/* autoreset_mode: gymnasium 1.x SyncVectorEnv semantics (agent/ppo.py:70,114) */
enum { RK_AUTORESET_NEXT_STEP = 0, RK_AUTORESET_SAME_STEP = 1, RK_AUTORESET_DISABLED = 2 };
/* query_mode: how waypoint-argmin and raycast candidates are found.  Every mode decides the discrete events
 * (progress index, wall test) exactly as the reference's float64 arithmetic does.
 *   EXACT_F64: float64 brute force over the whole tables (the in-library reference); ray distances are the
 *              reference's correctly rounded float64 quotients.
 *   CULLED:    warp-per-environment; candidates found in fp32 over bounding-circle chunks and an angular sweep.
 *   GRID:      thread-per-car waypoint search and thread-per-ray traversal of a per-track uniform grid over the
 *              boundary segments.
 * CULLED and GRID pick each ray's nearest segment in fp32 (two candidates within fp32 rounding of each other are
 * both re-evaluated) and compute its distance in float64 with a Newton reciprocal (<= 2 ulp): readings agree with
 * the reference to ~1e-15 relative, far inside the 1e-6 the observations are compared at. */
enum { RK_QUERY_EXACT_F64 = 0, RK_QUERY_CULLED = 1, RK_QUERY_GRID = 2 };

typedef struct rk_config {
    int32_t struct_size;        /* sizeof(rk_config), for ABI checking                */
    int32_t device;             /* CUDA device ordinal                                 */
    int32_t num_envs;           /* E                                                   */
    int32_t num_agents;         /* A (must be 1 for RK_ENV_SINGLE)                     */
    int32_t num_sensors;        /* R, rays per car (racing_env.py:9, multi:9)          */
    int32_t env_kind;           /* RK_ENV_*                                            */
    int32_t autoreset_mode;     /* RK_AUTORESET_*                                      */
    int32_t query_mode;         /* RK_QUERY_*                                          */
    int32_t max_episode_steps;  /* <=0 -> 3000 (racing_env.py:162)                     */
    int32_t reserved0;
    double speed_weight;        /* single env speed bonus weight (racing_env.py:9,140) */
    uint64_t seed;              /* Philox key for start-slot shuffles, random opponent */
} rk_config;

/* ---- lifetime ------------------------------------------------------------ */
RK_API int rk_create(const rk_config* cfg, rk_handle* out);
RK_API int rk_destroy(rk_handle h);
RK_API const char* rk_last_error(rk_handle h);
RK_API int rk_abi_version(void);
/* kernels launched by this library since load (bench.py's gpu_launches) */
RK_API uint64_t rk_launch_count(void);

/* ---- tracks: replaces Track.__init__ (environment/track.py:61-148) ------- */
/* Build the pool on the device from control points: periodic cubic spline
 * (the algorithm of scipy CubicSpline(bc_type='periodic'), track.py:100-115),
 * normals (:117-124), boundaries (:93-94), segment table (:126-148), bbox
 * diagonal (:82-91), start pose (:154-157).  host_ctrl_xy holds the tracks'
 * control points back to back as (x, y) pairs; waypoints per track =
 * n_ctrl[t] * factor (the reference uses factor 30).  host_env_to_track may be
 * NULL (env e uses track e % n_tracks). */
RK_API int rk_set_tracks_from_control_points(rk_handle h, const double* host_ctrl_xy, const int32_t* host_n_ctrl,
                                      const double* host_widths, int32_t n_tracks, int32_t factor,
                                      const int32_t* host_env_to_track);
/* Same, but from ready-made waypoints (float64, e.g. taken from the reference's
 * Track.waypoints) so that discrete events are bit-comparable with the
 * reference: only IEEE-exact operations separate waypoints from the tables. */
RK_API int rk_set_tracks_from_waypoints(rk_handle h, const double* host_wp_xy, const int32_t* host_n_wp,
                                 const double* host_widths, int32_t n_tracks,
                                 const int32_t* host_env_to_track);
/* Procedural pool generated on the device with the distributions of
 * gen_tracks/gen_random_track (track.py:4-56) from a Philox stream; widths are
 * width_lo + (t % width_mod).  For synthetic benchmark configurations. */
RK_API int rk_generate_tracks(rk_handle h, uint64_t seed, int32_t n_tracks, int32_t factor,
                       double width_lo, int32_t width_mod, const int32_t* host_env_to_track);
RK_API int rk_num_tracks(rk_handle h);
/* Export one track (what utils/visualization.py reads from env.track).  Any
 * pointer may be NULL; host_meta receives {n_wp, width, max_track_distance,
 * start_x, start_y, start_angle}.  Arrays are (x, y) pairs, n_wp entries. */
RK_API int rk_get_track(rk_handle h, int32_t track_id, double* host_meta6, double* host_waypoints,
                 double* host_normals, double* host_left, double* host_right, double* host_ctrl,
                 int32_t* n_ctrl_out);

/* ---- env: replaces RacingEnv/MultiRacingEnv.reset/step + SyncVectorEnv ---- */
/* reset (racing_env.py:86-102, multi_racing_env.py:118-153).  dev_mask: uint8
 * [E] or NULL (all).  dev_start_slot: int32 [E,A] slot of each car on the grid
 * (multi_racing_env.py:127-138) or NULL (Philox shuffle).  dev_obs: float32
 * [E,A,D] ([A,E,D] if layout is RK_LAYOUT_AGENT_MAJOR) or NULL. */
RK_API int rk_reset(rk_handle h, const uint8_t* dev_mask, const int32_t* dev_start_slot, float* dev_obs,
                    int32_t layout, void* stream);

/* layout of the per-car arrays (actions, obs, reward_*, info_*) */
enum { RK_LAYOUT_ENV_MAJOR = 0,     /* [E,A,...]: one row per environment              */
       RK_LAYOUT_AGENT_MAJOR = 1 }; /* [A,E,...]: car a of every env is one contiguous
                                       block (the SelfPlayWrapper view of car 0 / car 1) */

typedef struct rk_step_io {
    int32_t struct_size;
    int32_t layout;              /* RK_LAYOUT_*                                                 */
    const float* actions;        /* in  [E,A,2] steer, throttle (racing_env.py:104-107)        */
    const int32_t* start_slot;   /* in  [E,A] or NULL: slots used by auto-resets this step     */
    float* obs;                  /* out [E,A,D]                                                 */
    float* reward_f32;           /* out [E,A] or NULL                                           */
    double* reward_f64;          /* out [E,A] or NULL (SyncVectorEnv returns float64)           */
    uint8_t* terminated;         /* out [E]  crashed|finished / any finished|all crashed        */
    uint8_t* truncated;          /* out [E]  steps >= max_episode_steps                         */
    uint8_t* done;               /* out [E] or NULL: terminated|truncated (dones["__all__"])    */
    float* done_f32;             /* out [E] or NULL: same as float (PPO next_done)              */
    /* RecordEpisodeStatistics (agent/ppo.py:88,123-130): valid where ep_mask != 0 */
    uint8_t* ep_mask;            /* out [E] or NULL */
    double* ep_return;           /* out [E] or NULL */
    int32_t* ep_length;          /* out [E] or NULL */
    /* per-car info of this step (racing_env.py:77-84,156-159): may be NULL     */
    double* info_f64;            /* out [E,A,5]: x, y, speed, progress (1.0 if finished), progress_delta */
    int32_t* info_i32;           /* out [E,A,4]: crashed, finished, placement (0 unless ended), progress_idx */
    /* running totals over finished episodes, accumulated atomically: {sum of
     * returns, sum of lengths, count} -- what PPO.train averages (agent/ppo.py:272-276) */
    double* ep_stats;            /* inout [3] or NULL */
    /* Step only environments [env_begin, env_begin + env_count) (env_count <= 0: all).  Lets a
     * caller split one logical step into chunks on several streams so that the device->host
     * copy of one chunk overlaps the kernel of the next (BatchedRacingVecEnv.step does). */
    int32_t env_begin, env_count;
} rk_step_io;
RK_API int rk_step(rk_handle h, const rk_step_io* io, void* stream);

/* One step with HOST buffers -- the call SyncVectorEnv.step(actions) -> (obs, rewards, ...)
 * maps to (agent/ppo.py:114-120).  `io` names the caller's DEVICE buffers exactly as for
 * rk_step and must use RK_LAYOUT_AGENT_MAJOR; `host` names the host side.  The batch is cut
 * into n_chunks environment ranges, each on its own internal stream: host->device copy of the
 * learner's actions, (self-play) the opponent's inference, the step kernel over that range,
 * device->host copy of car 0's observations -- so the copies of one chunk overlap the kernels
 * of the others; the small per-environment results (arena_*) follow in one copy.  Returns
 * when every host output is complete.  Host buffers should be page-locked. */
typedef struct rk_host_io {
    int32_t struct_size;
    int32_t n_chunks;             /* 1..8                                                        */
    const float* actions;         /* host in  [E,2]: car 0's actions                             */
    float* obs;                   /* host out [E,D]: car 0's observations                        */
    void* arena_host;             /* host out: arena_bytes copied from arena_dev (or NULL)       */
    const void* arena_dev;        /* device: contiguous block holding the small per-env results  */
    int64_t arena_bytes;
    int32_t selfplay;             /* != 0: car 1 is driven here (SelfPlayWrapper, wrappers.py:29-45) */
    int32_t reserved0;            /* bit 0: zero-copy observations -- if `obs` is pinned (mapped) host memory the step
                                   * kernel writes car 0's complete rows straight into it (one coalesced store per
                                   * environment) and no device->host copy of the observations follows; culled
                                   * queries, num_agents <= 2, num_agents * num_sensors <= 32; ignored otherwise.
                                   * bit 1 (with bit 0): the same for the small per-environment results -- stores
                                   * that fall into arena_dev are mirrored into a pinned arena_host, no copy follows;
                                   * bit 2 (with bit 0): car 0's actions are read by the kernel from a pinned
                                   * `actions` buffer (and written through to the device array), no copy precedes */
    const float* opponent_params; /* device, packed Agent (see rk_policy_act) or NULL = uniform Box samples */
    uint64_t seed, counter;       /* Philox stream of the opponent's sampling                    */
} rk_host_io;
RK_API int rk_step_host(rk_handle h, const rk_step_io* io, const rk_host_io* host, void* caller_stream);

/* A whole T-step rollout in ONE call: PPO.collect_rollout (agent/ppo.py:97-132) with the vector env's step inside,
 * for a handle with RK_LAYOUT_AGENT_MAJOR buffers.  Per step t the library enqueues (a) one inference launch that
 * carries the learner's Agent.get_action_and_value on obs[t][0] -> actions[t][0], logprobs[t], values[t] and, for
 * self-play (2 cars, environment/wrappers.py:29-45), the frozen opponent's action on car 1's latest observation ->
 * actions[t][1], and (b) the fused step kernel writing obs[t+1], rewards[t], dones[t+1] in place.  All 2T launches are
 * issued back to back from native code on `stream`; nothing synchronises with the host.  Philox counters of step t are
 * counter0 + t.  `base` supplies the step's remaining outputs (terminated, truncated, ep_stats, ...) exactly as for
 * rk_step; its actions / obs / reward_f32 / done_f32 are overridden per step. */
typedef struct rk_rollout_io {
    int32_t struct_size;
    int32_t T;                       /* steps                                                               */
    int32_t selfplay;                /* != 0: car 1 is driven by the opponent (requires num_agents == 2)    */
    int32_t block_len;               /* opponent pool: consecutive envs per pool row (multiple of RK_POLICY_BLOCK) */
    const float* learner_params;     /* device, packed Agent (see rk_policy_act)                            */
    const float* opponent_params;    /* device, packed Agent / pool base, or NULL = uniform Box samples     */
    const int32_t* block_policy;     /* device int32 [ceil(E / block_len)] or NULL (single opponent)        */
    int64_t pool_stride;             /* floats between pool rows                                            */
    const float* opponent_obs0;      /* device [E,D]: car 1's observation for step 0, or NULL = obs[0][1]   */
    uint64_t learner_seed, learner_counter0;
    uint64_t opponent_seed, opponent_counter0;
    float* obs;                      /* device [T+1,A,E,D]: slot 0 holds the learner's current observation  */
    float* actions;                  /* device [T,A,E,2]                                                    */
    float* logprobs;                 /* device [T,E]                                                        */
    float* values;                   /* device [T,E]                                                        */
    float* rewards;                  /* device [T,A,E] float32                                              */
    float* dones;                    /* device [T+1,E] float32: slot t+1 = done after step t                */
} rk_rollout_io;
RK_API int rk_rollout(rk_handle h, const rk_step_io* base, const rk_rollout_io* r, void* stream);

/* RacingEnv.speed_weight (racing_env.py:26; annealed by agent/ppo.py:256-258) */
RK_API int rk_set_speed_weight(rk_handle h, double speed_weight);
/* Re-key and restart the Philox stream of the start-grid shuffles (rk_config::seed; the per-environment reset
 * counters return to zero, so the same seed reproduces the same grids); what `reset(seed=...)` of the vector
 * env forwards (gymnasium call site agent/ppo.py:230; the reference's envs draw the shuffle from the global
 * np.random stream, multi_racing_env.py:127-128). */
RK_API int rk_set_seed(rk_handle h, uint64_t seed);

/* parity harness: raw state.  host_car_f64 [E,A,6] = x, y, angle, vx, vy,
 * last_steering; host_car_i32 [E,A,4] = progress_idx, last_progress_idx, flags
 * (bit0 crashed, 1 finished, 2..4 checkpoints, 5 has_crashed), finished_step;
 * host_env_i32 [E,3] = steps, needs_reset, ep_length; host_env_f64 [E] =
 * ep_return. */
RK_API int rk_get_state(rk_handle h, double* host_car_f64, int32_t* host_car_i32, int32_t* host_env_i32, double* host_env_f64);
RK_API int rk_set_state(rk_handle h, const double* host_car_f64, const int32_t* host_car_i32, const int32_t* host_env_i32, const double* host_env_f64);
/* recompute observations from the current state (racing_env.py:55-75) */
RK_API int rk_observe(rk_handle h, float* dev_obs, int32_t layout, void* stream);

/* ---- rollout-side kernels -------------------------------------------------- */
/* GAE backward scan: PPO.compute_advantages (agent/ppo.py:134-154).  All
 * pointers device float32; rewards/values/dones/adv/ret are [T,E]; next_value,
 * next_done_f32 are [E]. */
RK_API int rk_gae(const float* rewards, const float* values, const float* dones, const float* next_value,
           const float* next_done_f32, float gamma, float lam, int32_t T, int32_t E,
           float* adv, float* ret, void* stream);

/* Fused policy inference: Agent.get_action_and_value with action=None
 * (agent/ppo.py:43-56) for a batch of B observations of width obs_dim.
 * params: the Agent state_dict packed for the kernel (float32, 16-byte aligned,
 * rk_policy_param_count(obs_dim) values, zero padded).  Hidden-layer weights are
 * stored TRANSPOSED, [in][out] row-major (torch stores [out][in]):
 *   actor_mu.0.weight^T, actor_mu.0.bias, actor_mu.2.weight^T, actor_mu.2.bias,
 *   actor_mu.4.weight [2][64], actor_mu.4.bias, log_std,
 *   critic.0.weight^T, critic.0.bias, critic.2.weight^T, critic.2.bias,
 *   critic.4.weight [64], critic.4.bias.  obs rows are
 * obs_stride floats apart, action rows act_stride floats apart (so a car's
 * slice of an [E,A,*] tensor can be read/written in place).  Normal samples
 * come from Philox(seed, counter).  logprob [B], value [B] and mean [B,2] (the
 * pre-noise tanh output) may each be NULL (opponent inference,
 * environment/wrappers.py:35-39, needs only the action).  If params is NULL,
 * actions are uniform in Box([-1,0],[1,1]) -- the pool-empty opponent
 * (wrappers.py:30-32). */
RK_API int rk_policy_act(const float* params, int32_t obs_dim, const float* obs, int64_t obs_stride, int32_t B,
                  uint64_t seed, uint64_t counter, float* action, int64_t act_stride,
                  float* logprob, float* value, float* mean, void* stream);
/* The same against a POOL of stacked packed blocks (`pool_stride` floats apart, a multiple of 4):
 * samples are split into consecutive blocks of `block_len` (a multiple of RK_POLICY_BLOCK) and block k
 * is driven by policy block_policy[k] (device int32 [ceil(B / block_len)]).  One launch plays every
 * environment against its own snapshot of SelfPlayPPO's opponent pool (self_play_ppo.py:12,40-44) --
 * a superset of the reference's one-opponent-per-update rule. */
#define RK_POLICY_BLOCK 256
RK_API int rk_policy_act_pool(const float* params_pool, int64_t pool_stride, const int32_t* block_policy,
                              int32_t block_len, int32_t obs_dim, const float* obs, int64_t obs_stride, int32_t B,
                              uint64_t seed, uint64_t counter, float* action, int64_t act_stride,
                              float* logprob, float* value, float* mean, void* stream);
/* number of float32 values in the packed Agent block for a given obs_dim (action_dim = 2) */
RK_API int rk_policy_param_count(int32_t obs_dim);

/* ---- PPO update helpers: PPO.ppo_update (agent/ppo.py:156-209) -------------------- */
/* Gather one minibatch: dst[k] = src[idx[k]] for obs [.,obs_dim], actions [.,2], old
 * log-probs, advantages, returns, values (ppo.py:170-195's fancy indexing).  idx: int64 [n]. */
RK_API int rk_gather_minibatch(const int64_t* idx, int32_t n, int32_t obs_dim, const float* obs, const float* act,
                               const float* logp, const float* adv, const float* ret, const float* val,
                               float* o_obs, float* o_act, float* o_logp, float* o_adv, float* o_ret, float* o_val,
                               void* stream);
/* Gradients of the clipped-surrogate + clipped-value loss (ppo.py:173-204) with respect to
 * the network outputs mu [n,2] and v [n], with torch.maximum / torch.clamp tie conventions;
 * the entropy bonus has no gradient because log_std is a buffer.  adv_mean/adv_std are the
 * (global) minibatch statistics as device scalars; *kl_sum += sum(logp_old - logp_new). */
RK_API int rk_ppo_loss_grad(const float* mu, const float* v, const float* act, const float* old_logp, const float* adv,
                            const float* ret, const float* v_old, const float* log_std, const float* adv_mean,
                            const float* adv_std, int32_t n, float clip_coef, float vf_coef, float* dmu, float* dv,
                            double* kl_sum, void* stream);

/* ---- fused minibatch gradient: forward + loss + backward of PPO.ppo_update --------
 * (agent/ppo.py:170-206 up to and including loss.backward(); gradient clipping and
 * Adam stay with the caller's optimizer).  The Agent's two 64-64 tanh MLPs
 * (agent/ppo.py:11-37) are evaluated, differentiated and their weight gradients
 * reduced over the minibatch by ONE kernel, fp32 throughout; nothing of size
 * [minibatch, 64] ever reaches HBM.  obs_dim <= RK_PPO_MAX_OBS_DIM. */
#define RK_ADV_STAT_BLOCKS 128
#define RK_PPO_MAX_OBS_DIM 20
/* partial (sum, sum of squares) in float64 of the minibatch's advantages adv[idx[k]],
 * k < n (idx NULL: adv[k]); part: double [RK_ADV_STAT_BLOCKS][2].  With several ranks
 * the caller all-reduces `part` so that every rank normalises with the statistics of
 * the GLOBAL minibatch (ppo.py:187 `mb_adv.mean()`, `mb_adv.std()` unbiased). */
RK_API int rk_ppo_adv_stats(const int64_t* idx, const float* adv, int32_t n, double* part, void* stream);

typedef struct rk_ppo_grad_io {
    int32_t struct_size;       /* sizeof(rk_ppo_grad_io) */
    int32_t obs_dim;
    int32_t n;                 /* rows of this rank's minibatch */
    int32_t obs_stride;        /* floats between observation rows; 0 = obs_dim.  A stride that is a multiple
                                * of 4 (rows padded to 16 bytes, padding readable) is gathered with 128-bit loads */
    double n_global;           /* rows of the global minibatch (n * world size) */
    /* parameters in torch layout [out][in]: actor_mu.{0,2,4}.{weight,bias} then critic.{0,2,4}.{weight,bias} */
    const float* params[12];
    const float* log_std;      /* [2] */
    /* the flat rollout buffers (agent/ppo.py:158-165) and the minibatch's row indices into them */
    const float* obs;          /* [B, obs_stride] */
    const float* act;          /* [B, 2] */
    const float* old_logp;     /* [B] */
    const float* adv;          /* [B] raw advantages; normalised inside with `adv_part` */
    const float* ret;          /* [B] */
    const float* val;          /* [B] */
    const int64_t* idx;        /* [n] or NULL for rows 0..n-1 */
    const double* adv_part;    /* [RK_ADV_STAT_BLOCKS][2] from rk_ppo_adv_stats (summed over ranks) */
    float clip_coef, vf_coef;
    void* workspace;           /* device scratch of rk_ppo_grad_workspace_bytes() bytes */
    uint64_t workspace_bytes;
    float* flat_grad;          /* out: d loss / d params, concatenated in the order of `params` */
    double* kl_sum;            /* out: sum over the n rows of (logp_old - logp_new)  (ppo.py:178-182) */
    float* kl_sum_f32;         /* optional out: the same as float32 -- e.g. the slot right after flat_grad, so
                                * that ONE all-reduce carries the gradient and the KL sum across ranks */
    int32_t tensor_cores;      /* 0: fp32 FMA kernel; 1: the per-sample products on tcgen05 tensor cores (TF32 x 3
                                * split, fp32-grade accuracy, accumulators and chained operands in TMEM); 2: the weight
                                * gradients dW2 / dW1 and their biases as tcgen05 products as well (operand tiles in
                                * shared memory, accumulators resident in TMEM across the CTA's tiles) */
    int32_t reserved1;
} rk_ppo_grad_io;
RK_API uint64_t rk_ppo_grad_workspace_bytes(void);
RK_API int rk_ppo_minibatch_grad(const rk_ppo_grad_io* io, void* stream);

/* Gradient clipping (nn.utils.clip_grad_norm_, agent/ppo.py:205), the Adam step (ppo.py:206;
 * torch.optim.Adam without weight decay / amsgrad, learning rate read from a device scalar) and the
 * KL early stop (ppo.py:178-182) of one minibatch as ONE kernel, in place on the caller's optimizer
 * state.  If kl_sum / n_global > kl_target the step is not applied and state[0] latches to 1: every
 * later call is a no-op until the caller clears it -- the host no longer has to synchronise once per
 * minibatch to take the decision.  state[1] counts the steps applied.  flat_grad is divided by `world`
 * first (the mean over ranks of an all-reduced sum). */
typedef struct rk_adam_io {
    int32_t struct_size;       /* sizeof(rk_adam_io) */
    int32_t world;
    float* params[12];         /* the tensors of rk_ppo_grad_io.params, updated in place */
    float* exp_avg[12];        /* torch.optim.Adam state, updated in place */
    float* exp_avg_sq[12];
    float* step[12];           /* device float32 scalars (capturable Adam); each is advanced by 1 */
    int32_t numel[12];
    const float* flat_grad;    /* from rk_ppo_minibatch_grad (summed over ranks) */
    const float* lr;           /* device scalar */
    float beta1, beta2, eps, max_grad_norm, kl_target;
    int32_t reserved0;
    const double* kl_sum;      /* from rk_ppo_minibatch_grad (summed over ranks) */
    const float* kl_sum_f32;   /* optional: if not NULL it is used instead of kl_sum */
    double n_global;
    int32_t* state;            /* device int32[4], zeroed by the caller: {stopped, steps applied, scratch, -} */
    float* kl_at_stop;         /* device scalar: approx_kl that latched the stop */
} rk_adam_io;
RK_API int rk_ppo_adam_step(const rk_adam_io* io, void* stream);

/* out[k], k < n: a seeded pseudo-random permutation of 0..n-1 (Feistel network + cycle walking, no
 * sort) -- the minibatch shuffle of agent/ppo.py:168 (`np.random.shuffle(b_inds)`); a different
 * (seed, counter) gives a different permutation, the same pair the same one on every rank. */
RK_API int rk_random_permutation(uint64_t seed, uint64_t counter, int64_t n, int64_t* out, void* stream);

/* ---- measurement aid --------------------------------------------------------- */
/* Sustained FMA throughput of the current device in TFLOP/s (use_fp64 = 0: fp32 FFMA, 1: fp64 DFMA,
 * 2: packed fp32 FFMA2, 3: legacy tensor path mma.sync m16n8k8 TF32): the non-tensor roofline denominator bench.py reports the step
 * kernel against (SURVEY.md 8d asks the builder to measure it). */
RK_API double rk_fma_peak(int32_t use_fp64, int32_t iters);

#ifdef __cplusplus
}
#endif
#endif /* RACING_B200_H */
