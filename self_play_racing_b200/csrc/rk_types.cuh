// Internal device-side types shared by the kernels of librk_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/racing_b200.h"

namespace rk {

constexpr int kChunk = 16;          // waypoints per bounding-circle chunk
constexpr int kRaySegs = 15;        // boundary segments per chunk (16 points: one half-warp, lane j+1 holds lane j's end point)
constexpr int kMaxKnots = 130;      // n_ctrl + 1 <= kMaxKnots
constexpr float kGridMargin = 0.02f;  // a segment is listed in every cell its bounding box, grown by this, overlaps
constexpr int kListMax = 480;       // chunks of one kind per track: the culled queries keep per-warp chunk lists of 512 entries
#ifndef RK_WARPS
#define RK_WARPS 4
#endif
constexpr int kWarpsPerCta = RK_WARPS;

// flags word of a car
enum : int { F_CRASHED = 1, F_FINISHED = 2, F_CP25 = 4, F_CP50 = 8, F_CP75 = 16, F_HAS_CRASHED = 32,
             F_RAYCACHE = 64 };  // internal: EnvState::wall_cache holds this (crashed, hence frozen) car's wall distances

// One record per track.  Everything float64 is what the reference keeps in
// Track (environment/track.py:61-98); org/chunk fields serve the culled query
// path only.
struct TrackMeta {
    int32_t n_wp;        // N = n_ctrl * factor            (track.py:110)
    int32_t wp_off;      // offset into per-waypoint arrays; segment arrays start at 2*wp_off
    int32_t n_wchunk;    // ceil(N / kChunk) waypoint chunks
    int32_t wchunk_off;
    int32_t n_bchunk;    // 2 * ceil(N / kRaySegs) boundary chunks (left side first)
    int32_t bchunk_off;
    int32_t bpt_off;     // offset of this track's fp32 boundary rows: 2 rows of N+1 points (point 0 repeated)
    int32_t pad0;
    int32_t n_ctrl;
    int32_t ctrl_off;
    double width;               // track.py:77-80
    double max_track_distance;  // track.py:82-91
    double start_x, start_y, start_angle;  // track.py:154-157
    double start_nx, start_ny;  // normals[0] (multi_racing_env.py:123)
    double org_x, org_y;        // bbox centre: origin of the fp32 tables
    // uniform grid over the boundary segments (RK_QUERY_GRID): cell (ix, iy) covers
    // [gx0 + ix*gcell, gx0 + (ix+1)*gcell) x [gy0 + iy*gcell, ...) in fp32 table coordinates
    float gx0, gy0, gcell, ginv;
    int32_t gnx, gny;
    int32_t gcell_off, glist_off;  // offsets into TrackPool::gcell / glist
};

struct TrackPool {
    const TrackMeta* meta;
    const int32_t* env_to_track;  // [E]
    // float64 tables, the reference's arithmetic
    const double *wx, *wy, *nx, *ny;      // [sum N]   waypoints, unit normals
    const double *sx, *sy, *v2x, *v2y;    // [sum 2N]  segment starts and vectors (track.py:134-148)
    // fp32 tables relative to (org_x, org_y): candidate search only
    const float2* wpt;     // [sum N]  waypoints
    const float2* bpt;     // [sum 2(N+1)] boundary points, per track: left row then right row, each closed
    const float4* wchunk;  // (cx, cy, r, -) bounding circle of kChunk waypoints
    const float4* bchunk;  // (cx, cy, r, -) bounding circle of kRaySegs boundary segments
    // RK_QUERY_GRID: per track a uniform grid whose cells list the boundary segments that come within kGridMargin
    const float4* bseg;      // [sum 2N] (px, py, vx, vy): segment start and vector, fp32, same numbering as sx/sy
    const uint32_t* gcell;   // per cell: first list entry << 10 | number of entries
    const uint16_t* glist;   // segment ids, cell after cell
};

struct EnvState {
    // per car, index e*A + a
    double *x, *y, *ang, *vx, *vy;
    float* last_steer;
    int32_t *pidx, *lpidx, *flags, *fstep;
    // per env
    int32_t* steps;
    uint8_t* needs_reset;
    double* ep_return;
    int32_t* ep_length;
    uint32_t* reset_count;
    // per car (RK_QUERY_GRID, R <= 15): this car's ray indices ordered by the previous step's readings, longest first,
    // 4 bits each, bits 60..63 = 0xF when valid.  A scheduling hint only -- results do not depend on it.
    unsigned long long* ray_order;
    // per (car, ray) (RK_QUERY_CULLED, multi): the wall distance of each ray of a CRASHED car.  Crashed cars stay frozen
    // (car.py:51-52) but are observed until the episode ends (multi_racing_env.py:236-249): their walls do not move, so
    // the sweep runs once, in the step of the crash, and later steps only add the other cars' edges.  Valid while the
    // car's flags carry F_RAYCACHE (cleared by every reset and by rk_set_state).
    float* wall_cache;
};

struct StepParams {
    TrackPool trk;
    EnvState st;
    rk_step_io io;
    const double* sensor_angles;  // [R] np.linspace(-half, half, R)
    const double *sensor_cos, *sensor_sin;  // [R] cos / sin of the above
    float inv_dphi;    // (R-1) / (2H): ray-index units per radian (H = half cone: pi/3 single, pi/2 multi)
    float cone_half, cone_sin, cone_cos;  // H plus a small margin, and its sine / cosine
    // staged launch (environments grouped by track, tables staged in shared memory); null = plain launch
    const int32_t* group_env;    // [n_ctas * kWarpsPerCta * epw] environment ids, -1 padded
    const int32_t* group_count;  // [n_ctas * kWarpsPerCta] environments in each warp's group
    const int32_t* cta_track;    // [n_ctas]
    int32_t n_ctas, stage_bytes;
    int32_t epw;       // environments per warp (<= 32 / A): fewer means more warps for the cooperative queries
    int32_t list_cap;  // entries of the per-warp chunk list (largest chunk list of the pool + 16, a multiple of 32; <= 512)
    int32_t n_shells;  // distance shells of the ray sweep: (-inf, shell[0]], (shell[0], shell[1]], ...
    float shell[4];
    int32_t E, A, R, D;
    int32_t env_begin, env_end;  // environments stepped by this launch (plain launch only)
    int32_t autoreset, max_steps;
    double speed_weight;
    uint64_t seed;
    // reset-only
    const uint8_t* reset_mask;
    // optional zero-copy of the small per-environment results: stores that land in [arena_lo, arena_hi) are mirrored
    // at +arena_delta bytes (the caller's pinned host arena, same layout)
    const char *arena_lo, *arena_hi;
    long long arena_delta;
    const float* act_host0;  // optional: car 0's actions [E, 2] read straight from mapped pinned host memory
    float* obs_host0;  // optional: car 0's observation block [E, D] in mapped pinned HOST memory; each warp then also
                       // writes the complete row there with one coalesced store, so that the rows cross PCIe
                       // while the kernel is still running (culled queries, A <= 2, A*R <= 32 only)
    int32_t lane_argmin;  // culled / grid: 1 = every car searches its nearest waypoints on its own lane (throughput: many cars
                          // per warp), 0 = one car at a time with the whole warp (latency: small batches, few cars per warp)
    int32_t mode;  // 0 = step, 1 = reset(mask), 2 = observe only
};

// ---- Philox4x32-10 (Salmon et al. 2011), counter-based RNG ----------------
__host__ __device__ inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int i = 0; i < 10; ++i) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}
__host__ __device__ inline float u01(uint32_t r) { return ((r >> 8) + 0.5f) * (1.0f / 16777216.0f); }
__host__ __device__ inline double u01d(uint32_t hi, uint32_t lo) {
    return ((((uint64_t)hi << 21) ^ (uint64_t)(lo >> 11)) + 0.5) * (1.0 / 9007199254740992.0);
}

// launch bookkeeping (rk_launch_count)
void count_launch(int n = 1);
// true exactly once per (kernel slot, current device): function attributes such as the dynamic shared-memory
// limit are per device, so a process that drives several GPUs must set them on each
bool first_use_on_device(int slot);

// host-side launchers implemented in the .cu files
int launch_step(const StepParams& p, int query_mode, int env_kind, cudaStream_t stream);
int launch_gae(const float* rewards, const float* values, const float* dones, const float* next_value,
               const float* next_done, float gamma, float lam, int T, int E, float* adv, float* ret,
               cudaStream_t stream);
int launch_policy_act(const float* params, int obs_dim, const float* obs, int64_t obs_stride, int B,
                      uint64_t seed, uint64_t counter, float* action, int64_t act_stride, float* logprob,
                      float* value, float* mean, cudaStream_t stream, const int32_t* block_policy = nullptr,
                      int block_len = 0, int64_t pool_stride = 0);
int policy_param_count(int obs_dim);
// one inference job of policy_act_kernel; a launch carries one or two (blockIdx.y)
struct PolicyJob {
    const float* params;         // packed block (or pool base); NULL = uniform Box actions
    const float* obs;
    int64_t obs_stride;
    int B;
    uint64_t seed, counter;
    float* action;
    int64_t act_stride;
    float *logprob, *value, *mean;
    const int32_t* block_policy;
    int block_len;
    int64_t pool_stride;
};
struct PolicyJobs { PolicyJob job[2]; };
int launch_policy_jobs(const PolicyJobs& jobs, int n_jobs, int obs_dim, cudaStream_t stream);
int launch_gather_minibatch(const int64_t* idx, int n, int obs_dim, const float* obs, const float* act,
                            const float* logp, const float* adv, const float* ret, const float* val, float* o_obs,
                            float* o_act, float* o_logp, float* o_adv, float* o_ret, float* o_val, cudaStream_t stream);
int launch_ppo_loss_grad(const float* mu, const float* v, const float* act, const float* old_logp, const float* adv,
                         const float* ret, const float* v_old, const float* log_std, const float* adv_mean,
                         const float* adv_std, int n, float clip, float vf_coef, float* dmu, float* dv,
                         double* kl_sum, cudaStream_t stream);

// fused PPO minibatch gradient (rk_train.cu)
constexpr int kAdvBlocks = 128;  // partial sums produced by launch_adv_stats ([kAdvBlocks][2] doubles)
struct PpoGradIO {
    int obs_dim, n, obs_stride;
    double n_global;
    const float* params[12];  // actor W1,b1,W2,b2,W3,b3 then critic, torch layouts [out][in]
    const float* log_std;
    const float *obs, *act, *old_logp, *adv, *ret, *val;
    const int64_t* idx;
    const double* adv_part;
    float clip, vf_coef;
    void* workspace;
    size_t workspace_bytes;
    float* flat_grad;
    double* kl_sum;
    float* kl_sum_f32;
    int tensor_cores;
};
struct PpoAdamIO {
    float* params[12];
    float* exp_avg[12];
    float* exp_avg_sq[12];
    float* step[12];
    int numel[12];
    const float* flat_grad;
    const float* lr;
    float beta1, beta2, eps, max_norm, kl_target;
    int world;
    const double* kl_sum;
    const float* kl_sum_f32;
    double n_global;
    int* state;
    float* kl_at_stop;
};
int launch_clip_adam(const PpoAdamIO& io, cudaStream_t stream);
int launch_permutation(uint64_t seed, uint64_t counter, int64_t n, int64_t* out, cudaStream_t stream);
size_t ppo_grad_workspace_bytes();
int launch_adv_stats(const int64_t* idx, const float* adv, int n, double* part, cudaStream_t stream);
int launch_ppo_minibatch_grad(const PpoGradIO& io, cudaStream_t stream);

}  // namespace rk
