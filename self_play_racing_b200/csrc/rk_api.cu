// C ABI of librk_b200.so (declared in include/racing_b200.h).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <string>
#include <vector>

#include "rk_types.cuh"

namespace rk {

struct PoolBuffers;  // rk_track.cu
int build_pool(PoolBuffers& pb, int n_tracks, const int32_t* n_ctrl, const int32_t* ctrl_off, size_t total_ctrl,
               const double* host_ctrl, const int32_t* n_wp, const double* host_wp, const double* widths,
               const int32_t* host_env_to_track, int E, char* err, size_t errlen);
int generate_control_points(uint64_t seed, int n_tracks, std::vector<double>& ctrl, std::vector<int32_t>& n_ctrl,
                            char* err, size_t errlen);
PoolBuffers* pool_new();
void pool_delete(PoolBuffers* p);
TrackPool pool_view(const PoolBuffers* p);
int pool_num_tracks(const PoolBuffers* p);
const TrackMeta* pool_host_meta(const PoolBuffers* p, int t);
const double* pool_ctrl(const PoolBuffers* p);

static std::atomic<uint64_t> g_launches{0};
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

static std::atomic<uint64_t> g_attr_done[8];   // bit d of slot s: attribute set on device d
bool first_use_on_device(int slot) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev > 63) return true;
    const uint64_t bit = 1ull << dev;
    return (g_attr_done[slot & 7].fetch_or(bit, std::memory_order_relaxed) & bit) == 0;
}

}  // namespace rk

using namespace rk;

struct rk_env_s {
    rk_config cfg;
    int D = 0;
    PoolBuffers* pool = nullptr;
    EnvState st{};
    double* sensor_angles = nullptr;  // [3R]: angles, cos, sin
    int epw = 1;                       // environments per warp of the step kernel
    // staged launch plan (environments grouped by track); null when not applicable
    int32_t *group_env = nullptr, *group_count = nullptr, *cta_track = nullptr;
    int n_ctas = 0, stage_bytes = 0;
    int list_cap = 512;                // per-warp chunk-list entries the step kernel reserves (sized by set_tracks)
    // internal streams / events of rk_step_host
    cudaStream_t hstream[8] = {nullptr};
    cudaEvent_t hevent[9] = {nullptr};
    std::vector<void*> owned;
    char err[512] = {0};
};

static char g_create_err[512] = {0};

#define H_CUDA(h, call)                                                                        \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            snprintf((h)->err, sizeof((h)->err), "%s failed: %s", #call, cudaGetErrorString(e_)); \
            return 1;                                                                          \
        }                                                                                      \
    } while (0)

template <typename T>
static int dev_alloc(rk_env_s* h, T** p, size_t n) {
    H_CUDA(h, cudaMalloc((void**)p, n * sizeof(T)));
    H_CUDA(h, cudaMemset(*p, 0, n * sizeof(T)));
    h->owned.push_back(*p);
    return 0;
}

extern "C" {

int rk_abi_version(void) { return RK_ABI_VERSION; }
uint64_t rk_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

const char* rk_last_error(rk_handle h) { return h ? h->err : g_create_err; }

int rk_create(const rk_config* cfg, rk_handle* out) {
    if (!cfg || !out) {
        snprintf(g_create_err, sizeof(g_create_err), "rk_create: null argument");
        return 1;
    }
    *out = nullptr;
    if (cfg->struct_size != (int32_t)sizeof(rk_config)) {
        snprintf(g_create_err, sizeof(g_create_err), "rk_create: struct_size %d != %zu (ABI mismatch)",
                 cfg->struct_size, sizeof(rk_config));
        return 1;
    }
    const bool single = cfg->env_kind == RK_ENV_SINGLE;
    if (cfg->num_envs <= 0 || cfg->num_agents <= 0 || cfg->num_agents > RK_MAX_AGENTS || cfg->num_sensors <= 0 ||
        cfg->num_sensors > RK_MAX_SENSORS || (single && cfg->num_agents != 1) ||
        (cfg->env_kind != RK_ENV_SINGLE && cfg->env_kind != RK_ENV_MULTI) || cfg->autoreset_mode < 0 ||
        cfg->autoreset_mode > RK_AUTORESET_DISABLED || cfg->query_mode < 0 || cfg->query_mode > RK_QUERY_GRID) {
        snprintf(g_create_err, sizeof(g_create_err),
                 "rk_create: invalid config (E=%d A=%d R=%d kind=%d autoreset=%d query=%d)", cfg->num_envs,
                 cfg->num_agents, cfg->num_sensors, cfg->env_kind, cfg->autoreset_mode, cfg->query_mode);
        return 1;
    }
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || cfg->device < 0 || cfg->device >= ndev) {
        snprintf(g_create_err, sizeof(g_create_err), "rk_create: no usable CUDA device %d (%s); there is no CPU path",
                 cfg->device, ce != cudaSuccess ? cudaGetErrorString(ce) : "ordinal out of range");
        return 1;
    }
    rk_env_s* h = new rk_env_s();
    h->cfg = *cfg;
    if (h->cfg.max_episode_steps <= 0) h->cfg.max_episode_steps = 3000;
    const int E = cfg->num_envs, A = cfg->num_agents, R = cfg->num_sensors;
    h->D = single ? R + 4 : R + 4 + 4 * (A - 1);
    h->pool = pool_new();
    int rc = 0;
    if (cudaSetDevice(cfg->device) != cudaSuccess) rc = 1;
    const size_t C = (size_t)E * A;
    rc = rc || dev_alloc(h, &h->st.x, C) || dev_alloc(h, &h->st.y, C) || dev_alloc(h, &h->st.ang, C) ||
         dev_alloc(h, &h->st.vx, C) || dev_alloc(h, &h->st.vy, C) || dev_alloc(h, &h->st.last_steer, C) ||
         dev_alloc(h, &h->st.pidx, C) || dev_alloc(h, &h->st.lpidx, C) || dev_alloc(h, &h->st.flags, C) ||
         dev_alloc(h, &h->st.fstep, C) || dev_alloc(h, &h->st.steps, (size_t)E) ||
         dev_alloc(h, &h->st.needs_reset, (size_t)E) || dev_alloc(h, &h->st.ep_return, (size_t)E) ||
         dev_alloc(h, &h->st.ep_length, (size_t)E) || dev_alloc(h, &h->st.reset_count, (size_t)E) ||
         dev_alloc(h, &h->st.ray_order, C) || dev_alloc(h, &h->st.wall_cache, C * (size_t)R) ||
         dev_alloc(h, &h->sensor_angles, (size_t)3 * R);
    if (rc) {
        snprintf(g_create_err, sizeof(g_create_err), "rk_create: %s", h->err[0] ? h->err : "cudaSetDevice failed");
        rk_destroy(h);
        return 1;
    }
    // np.linspace(-half, half, R): racing_env.py:45 (120 deg) / multi_racing_env.py:50 (180 deg)
    std::vector<double> ang(3 * R);
    const double half = single ? M_PI / 3 : M_PI / 2;
    const double start = -half, stop = half;
    if (R == 1) {
        ang[0] = start;
    } else {
        volatile double step = (stop - start) / (double)(R - 1);
        for (int k = 0; k < R; ++k) {
            volatile double prod = (double)k * step;  // numpy: arange(R) * step + start, no fused multiply-add
            ang[k] = prod + start;
        }
        ang[R - 1] = stop;
    }
    for (int k = 0; k < R; ++k) {
        ang[R + k] = cos(ang[k]);
        ang[2 * R + k] = sin(ang[k]);
    }
    cudaMemcpy(h->sensor_angles, ang.data(), 3 * R * sizeof(double), cudaMemcpyHostToDevice);
    {
        // environments per warp: as many as fit (one car per lane) unless that leaves the GPU short of
        // warps; measured (profiles/r01_epw_sweep.log): one full wave of ~27 warps per SM is the
        // sweet spot, more environments per warp beyond that only serialises the ray queries
        // (RK_B200_EPW overrides for tuning)
        int epw = 32 / A, sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cfg->device);
        while (epw > 1 && (E + epw - 1) / epw < 26 * sms) epw >>= 1;
        if (const char* env = getenv("RK_B200_EPW")) epw = atoi(env);
        h->epw = epw < 1 ? 1 : (epw > 32 / A ? 32 / A : epw);
    }
    *out = h;
    return 0;
}

int rk_destroy(rk_handle h) {
    if (!h) return 0;
    cudaSetDevice(h->cfg.device);
    for (cudaStream_t st : h->hstream)
        if (st) cudaStreamDestroy(st);
    for (cudaEvent_t ev : h->hevent)
        if (ev) cudaEventDestroy(ev);
    for (void* p : h->owned) cudaFree(p);
    for (void* p : {(void*)h->group_env, (void*)h->group_count, (void*)h->cta_track})
        if (p) cudaFree(p);
    if (h->pool) pool_delete(h->pool);
    delete h;
    return 0;
}

// Group the environments by track so that every CTA of the step kernel works on ONE
// track and can stage its search tables in shared memory.  Each track's environments
// are cut into warp groups of `epw`; a track's group count is padded to whole CTAs.
// Opt-in (RK_B200_STAGED=1): measured on B200 the staged launch is within 1 % of the plain
// one for pools of 16..1024 tracks (profiles/r01_staged_vs_plain.log) -- once a warp's
// environments share a track the tables are L1 hits anyway -- so the plain launch stays
// the default.  Also skipped when the padding would waste more than a quarter of the warps
// (e.g. a distinct track per environment).
static int plan_staged_launch(rk_handle h, const int32_t* e2t_in) {
    for (int32_t** p : {&h->group_env, &h->group_count, &h->cta_track}) {
        if (*p) cudaFree(*p);
        *p = nullptr;
    }
    h->n_ctas = 0;
    const char* env = getenv("RK_B200_STAGED");
    if (!env || atoi(env) == 0 || h->cfg.query_mode != RK_QUERY_CULLED) return 0;
    const int E = h->cfg.num_envs, epw = h->epw, nt = pool_num_tracks(h->pool);
    std::vector<std::vector<int32_t>> per_track(nt);
    for (int e = 0; e < E; ++e) per_track[e2t_in ? e2t_in[e] : e % nt].push_back(e);
    std::vector<int32_t> genv, gcount, ctrack;
    int stage = 0;
    for (int t = 0; t < nt; ++t) {
        const auto& v = per_track[t];
        if (v.empty()) continue;
        const TrackMeta& m = *pool_host_meta(h->pool, t);
        stage = std::max(stage, 16 * (m.n_wp + 1) + 16 * m.n_bchunk + 16 * m.n_wchunk);
        const int groups = ((int)v.size() + epw - 1) / epw;
        const int ctas = (groups + kWarpsPerCta - 1) / kWarpsPerCta;
        for (int c = 0; c < ctas; ++c) {
            ctrack.push_back(t);
            for (int w = 0; w < kWarpsPerCta; ++w) {
                const int g = c * kWarpsPerCta + w;
                int cnt = 0;
                for (int k = 0; k < epw; ++k) {
                    const size_t i = (size_t)g * epw + k;
                    const bool ok = g < groups && i < v.size();
                    genv.push_back(ok ? v[i] : -1);
                    cnt += ok;
                }
                gcount.push_back(cnt);
            }
        }
    }
    const size_t plain_warps = ((size_t)E + epw - 1) / epw;
    if (gcount.size() > plain_warps + plain_warps / 4 + kWarpsPerCta) return 0;  // too much padding: plain launch
    H_CUDA(h, cudaMalloc(&h->group_env, genv.size() * sizeof(int32_t)));
    H_CUDA(h, cudaMalloc(&h->group_count, gcount.size() * sizeof(int32_t)));
    H_CUDA(h, cudaMalloc(&h->cta_track, ctrack.size() * sizeof(int32_t)));
    H_CUDA(h, cudaMemcpy(h->group_env, genv.data(), genv.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    H_CUDA(h, cudaMemcpy(h->group_count, gcount.data(), gcount.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    H_CUDA(h, cudaMemcpy(h->cta_track, ctrack.data(), ctrack.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    h->n_ctas = (int)ctrack.size();
    h->stage_bytes = (stage + 15) / 16 * 16;
    return 0;
}

static int set_tracks_common(rk_handle h, const double* ctrl, const int32_t* n_ctrl, const double* wp,
                             const int32_t* n_wp_in, const double* widths, int n_tracks, int factor,
                             const int32_t* e2t) {
    if (!h) return 1;
    if (n_tracks <= 0 || !widths || (!ctrl && !wp)) {
        snprintf(h->err, sizeof(h->err), "set_tracks: invalid arguments");
        return 1;
    }
    cudaSetDevice(h->cfg.device);
    std::vector<int32_t> n_wp(n_tracks), off(n_tracks);
    size_t total_ctrl = 0;
    for (int t = 0; t < n_tracks; ++t) {
        if (ctrl) {
            if (factor <= 0) {
                snprintf(h->err, sizeof(h->err), "set_tracks: factor must be positive");
                return 1;
            }
            off[t] = (int32_t)total_ctrl;
            total_ctrl += n_ctrl[t];
            n_wp[t] = n_ctrl[t] * factor;
        } else {
            n_wp[t] = n_wp_in[t];
        }
    }
    if (build_pool(*h->pool, n_tracks, ctrl ? n_ctrl : nullptr, ctrl ? off.data() : nullptr, total_ctrl, ctrl,
                   n_wp.data(), wp, widths, e2t, h->cfg.num_envs, h->err, sizeof(h->err)))
        return 1;
    if (h->cfg.query_mode != RK_QUERY_EXACT_F64)
        for (int t = 0; t < n_tracks; ++t) {
            const TrackMeta& m = *pool_host_meta(h->pool, t);
            if (m.n_bchunk > kListMax || m.n_wchunk > kListMax) {
                snprintf(h->err, sizeof(h->err),
                         "set_tracks: track %d has %d waypoints; the culled / grid query modes handle at most %d per track "
                         "(use RK_QUERY_EXACT_F64 or fewer waypoints per control point)", t, m.n_wp, kListMax / 2 * kRaySegs);
                return 1;
            }
        }
    h->list_cap = 512;
    if (const char* env = getenv("RK_B200_LIST_CAP")) {
        // Opt-in: size the step kernel's per-warp chunk list for the pool's largest track (+ 16 entries of scratch).  At
        // 65,536 two-car envs that takes 7 CTAs x 26 KB of shared memory below the 164 KB carve-out (L1 60 -> 92 KB):
        // step kernel 0.2685 -> 0.2644 ms device-resident, but the Gymnasium face gets SLOWER (0.333 -> 0.344 ms; the
        // inference kernel needs the 196 KB carve-out, and alternating configurations cost more than the larger L1
        // wins when the two kernels are launched per step from the host) -- hence off by default.
        if (strcmp(env, "auto") == 0) {
            int most = 0;
            for (int t = 0; t < n_tracks; ++t) {
                const TrackMeta& m = *pool_host_meta(h->pool, t);
                most = std::max(most, std::max((int)m.n_bchunk, (int)m.n_wchunk));
            }
            h->list_cap = std::min(512, (most + 16 + 31) / 32 * 32);
        }
    }
    return plan_staged_launch(h, e2t);
}

int rk_set_tracks_from_control_points(rk_handle h, const double* host_ctrl_xy, const int32_t* host_n_ctrl,
                                      const double* host_widths, int32_t n_tracks, int32_t factor,
                                      const int32_t* host_env_to_track) {
    if (h && (!host_ctrl_xy || !host_n_ctrl)) {
        snprintf(h->err, sizeof(h->err), "set_tracks_from_control_points: null argument");
        return 1;
    }
    return set_tracks_common(h, host_ctrl_xy, host_n_ctrl, nullptr, nullptr, host_widths, n_tracks, factor,
                             host_env_to_track);
}

int rk_set_tracks_from_waypoints(rk_handle h, const double* host_wp_xy, const int32_t* host_n_wp,
                                 const double* host_widths, int32_t n_tracks, const int32_t* host_env_to_track) {
    if (h && (!host_wp_xy || !host_n_wp)) {
        snprintf(h->err, sizeof(h->err), "set_tracks_from_waypoints: null argument");
        return 1;
    }
    return set_tracks_common(h, nullptr, nullptr, host_wp_xy, host_n_wp, host_widths, n_tracks, 0, host_env_to_track);
}

int rk_generate_tracks(rk_handle h, uint64_t seed, int32_t n_tracks, int32_t factor, double width_lo,
                       int32_t width_mod, const int32_t* host_env_to_track) {
    if (!h) return 1;
    if (n_tracks <= 0 || factor <= 0) {
        snprintf(h->err, sizeof(h->err), "generate_tracks: invalid arguments");
        return 1;
    }
    cudaSetDevice(h->cfg.device);
    std::vector<double> ctrl;
    std::vector<int32_t> n_ctrl;
    if (generate_control_points(seed, n_tracks, ctrl, n_ctrl, h->err, sizeof(h->err))) return 1;
    std::vector<double> widths(n_tracks);
    for (int t = 0; t < n_tracks; ++t) widths[t] = width_lo + (width_mod > 0 ? t % width_mod : 0);
    return set_tracks_common(h, ctrl.data(), n_ctrl.data(), nullptr, nullptr, widths.data(), n_tracks, factor,
                             host_env_to_track);
}

int rk_num_tracks(rk_handle h) { return h ? pool_num_tracks(h->pool) : 0; }

int rk_get_track(rk_handle h, int32_t track_id, double* meta6, double* wp, double* nrm, double* left, double* right,
                 double* ctrl, int32_t* n_ctrl_out) {
    if (!h) return 1;
    if (track_id < 0 || track_id >= pool_num_tracks(h->pool)) {
        snprintf(h->err, sizeof(h->err), "get_track: id %d out of range", track_id);
        return 1;
    }
    cudaSetDevice(h->cfg.device);
    const TrackMeta& m = *pool_host_meta(h->pool, track_id);
    const TrackPool v = pool_view(h->pool);
    const int N = m.n_wp;
    if (meta6) {
        meta6[0] = N; meta6[1] = m.width; meta6[2] = m.max_track_distance;
        meta6[3] = m.start_x; meta6[4] = m.start_y; meta6[5] = m.start_angle;
    }
    std::vector<double> a(N), b(N);
    auto fetch = [&](const double* dx, const double* dy, size_t off, double* out) -> int {
        if (!out) return 0;
        H_CUDA(h, cudaMemcpy(a.data(), dx + off, N * sizeof(double), cudaMemcpyDeviceToHost));
        H_CUDA(h, cudaMemcpy(b.data(), dy + off, N * sizeof(double), cudaMemcpyDeviceToHost));
        for (int i = 0; i < N; ++i) {
            out[2 * i] = a[i];
            out[2 * i + 1] = b[i];
        }
        return 0;
    };
    if (fetch(v.wx, v.wy, m.wp_off, wp) || fetch(v.nx, v.ny, m.wp_off, nrm) ||
        fetch(v.sx, v.sy, 2 * (size_t)m.wp_off, left) || fetch(v.sx, v.sy, 2 * (size_t)m.wp_off + N, right))
        return 1;
    if (n_ctrl_out) *n_ctrl_out = m.n_ctrl;
    if (ctrl && m.n_ctrl > 0 && pool_ctrl(h->pool))
        H_CUDA(h, cudaMemcpy(ctrl, pool_ctrl(h->pool) + 2 * (size_t)m.ctrl_off, 2 * (size_t)m.n_ctrl * sizeof(double),
                             cudaMemcpyDeviceToHost));
    return 0;
}

static int fill_params(rk_handle h, StepParams& p, const char* who) {
    if (pool_num_tracks(h->pool) == 0) {
        snprintf(h->err, sizeof(h->err), "%s: no tracks set (call rk_set_tracks_* or rk_generate_tracks first)", who);
        return 1;
    }
    memset(&p, 0, sizeof(p));
    p.trk = pool_view(h->pool);
    p.st = h->st;
    p.sensor_angles = h->sensor_angles;
    p.sensor_cos = h->sensor_angles + h->cfg.num_sensors;
    p.sensor_sin = h->sensor_angles + 2 * h->cfg.num_sensors;
    {
        const double half = (h->cfg.env_kind == RK_ENV_SINGLE) ? M_PI / 3 : M_PI / 2;
        const int R = h->cfg.num_sensors;
        p.inv_dphi = (R > 1) ? (float)((R - 1) / (2.0 * half)) : 1.0f;
        p.cone_half = (float)half;
        p.cone_sin = (float)sin(half + 2e-3);
        p.cone_cos = (float)cos(half + 2e-3);
        // distance shells of the ray sweep (last one unbounded); RK_B200_SHELLS="a,b" overrides for tuning
        const bool single = h->cfg.env_kind == RK_ENV_SINGLE;
        // measured (profiles/r01_shell_sweep.log): the clamped multi env is fastest in one pass,
        // the unclamped single env with one split at 30 units
        float sh[4] = {single ? 30.f : INFINITY, INFINITY, INFINITY, INFINITY};
        int ns = single ? 2 : 1;
        if (const char* env = getenv("RK_B200_SHELLS")) {
            ns = 0;
            for (const char* q = env; *q && ns < 3;) {
                sh[ns++] = (float)atof(q);
                const char* c = strchr(q, ',');
                if (!c) break;
                q = c + 1;
            }
            sh[ns++] = INFINITY;
        }
        p.epw = h->epw;
        p.group_env = h->group_env;
        p.group_count = h->group_count;
        p.cta_track = h->cta_track;
        p.n_ctas = h->n_ctas;
        p.stage_bytes = h->stage_bytes;
        p.n_shells = ns;
        for (int i = 0; i < 4; ++i) p.shell[i] = (i < ns - 1) ? sh[i] : INFINITY;
    }
    p.E = h->cfg.num_envs; p.A = h->cfg.num_agents; p.R = h->cfg.num_sensors; p.D = h->D;
    p.list_cap = h->list_cap;
    p.lane_argmin = h->epw * h->cfg.num_agents >= 16;   // measured: profiles/r02_small_batch.txt (RK_B200_LANE_ARGMIN overrides)
    if (const char* env = getenv("RK_B200_LANE_ARGMIN")) p.lane_argmin = atoi(env) != 0;
    p.env_begin = 0;
    p.env_end = h->cfg.num_envs;
    p.autoreset = h->cfg.autoreset_mode;
    p.max_steps = h->cfg.max_episode_steps;
    p.speed_weight = h->cfg.speed_weight;
    p.seed = h->cfg.seed;
    return 0;
}

int rk_reset(rk_handle h, const uint8_t* dev_mask, const int32_t* dev_start_slot, float* dev_obs, int32_t layout,
             void* stream) {
    if (!h) return 1;
    StepParams p;
    if (fill_params(h, p, "rk_reset")) return 1;
    p.mode = 1;
    p.reset_mask = dev_mask;
    p.io.start_slot = dev_start_slot;
    p.io.obs = dev_obs;
    p.io.layout = layout;
    if (launch_step(p, h->cfg.query_mode, h->cfg.env_kind, (cudaStream_t)stream)) {
        snprintf(h->err, sizeof(h->err), "rk_reset: launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        return 1;
    }
    return 0;
}

int rk_step(rk_handle h, const rk_step_io* io, void* stream) {
    if (!h) return 1;
    if (!io || io->struct_size != (int32_t)sizeof(rk_step_io)) {
        snprintf(h->err, sizeof(h->err), "rk_step: bad rk_step_io (struct_size mismatch)");
        return 1;
    }
    if (!io->actions || !io->terminated || !io->truncated) {
        snprintf(h->err, sizeof(h->err), "rk_step: actions, terminated and truncated are required");
        return 1;
    }
    StepParams p;
    if (fill_params(h, p, "rk_step")) return 1;
    p.mode = 0;
    p.io = *io;
    if (io->env_count > 0) {
        if (io->env_begin < 0 || io->env_begin + io->env_count > h->cfg.num_envs) {
            snprintf(h->err, sizeof(h->err), "rk_step: environment range [%d, %d) out of bounds", io->env_begin,
                     io->env_begin + io->env_count);
            return 1;
        }
        p.env_begin = io->env_begin;
        p.env_end = io->env_begin + io->env_count;
        p.group_env = nullptr;  // a sub-range is always a plain launch
    }
    if (launch_step(p, h->cfg.query_mode, h->cfg.env_kind, (cudaStream_t)stream)) {
        snprintf(h->err, sizeof(h->err), "rk_step: launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        return 1;
    }
    return 0;
}

static int step_host_impl(rk_handle h, const rk_step_io* io, const rk_host_io* host, void* caller_stream);

int rk_step_host(rk_handle h, const rk_step_io* io, const rk_host_io* host, void* caller_stream) {
    if (!h) return 1;
    const int rc = step_host_impl(h, io, host, caller_stream);
    if (rc) {
        // an error after work was queued on the internal streams: drain them, so that the caller may free or
        // reuse its buffers as soon as the call has returned (the error text in h->err is kept)
        for (cudaStream_t st : h->hstream)
            if (st) cudaStreamSynchronize(st);
        cudaGetLastError();
    }
    return rc;
}

static int step_host_impl(rk_handle h, const rk_step_io* io, const rk_host_io* host, void* caller_stream) {
    if (!io || io->struct_size != (int32_t)sizeof(rk_step_io) || !host ||
        host->struct_size != (int32_t)sizeof(rk_host_io)) {
        snprintf(h->err, sizeof(h->err), "rk_step_host: bad io structs (struct_size mismatch)");
        return 1;
    }
    const int E = h->cfg.num_envs, A = h->cfg.num_agents, D = h->D;
    if (!io->actions || !io->obs || !io->terminated || !io->truncated || !host->actions || !host->obs ||
        io->layout != RK_LAYOUT_AGENT_MAJOR || (host->selfplay && A != 2)) {
        snprintf(h->err, sizeof(h->err),
                 "rk_step_host: needs device actions/obs/terminated/truncated, host actions/obs, agent-major "
                 "layout (and 2 cars for self-play)");
        return 1;
    }
    StepParams p;
    if (fill_params(h, p, "rk_step_host")) return 1;
    p.mode = 0;
    p.io = *io;
    p.group_env = nullptr;  // range launches are plain launches
    // zero-copy observations (host->reserved0 bit 0): if the caller's host buffer is pinned (mapped into the device's
    // address space) the step kernel also writes car 0's complete rows straight into it, one coalesced store per
    // environment, so that they cross PCIe while the kernel is still running and no device->host copy follows
    bool zero_copy = false;
    if ((host->reserved0 & 1) && h->cfg.query_mode != RK_QUERY_EXACT_F64 && A <= 2 && A * h->cfg.num_sensors <= 32 && D <= 32) {
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, host->obs) == cudaSuccess && attr.type == cudaMemoryTypeHost && attr.devicePointer) {
            p.obs_host0 = static_cast<float*>(attr.devicePointer);
            zero_copy = true;
        } else {
            cudaGetLastError();
        }
    }
    // the small per-environment results (rewards, flags, episode statistics) take the same route when the host
    // arena is pinned: the kernel mirrors every store that falls into the device arena at the same offset there
    // ... and car 0's actions are read by the kernel straight from the caller's pinned buffer (bit 2)
    bool zero_copy_act = false;
    if (zero_copy && (host->reserved0 & 4)) {
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, host->actions) == cudaSuccess && attr.type == cudaMemoryTypeHost &&
            attr.devicePointer && (reinterpret_cast<uintptr_t>(attr.devicePointer) & 7u) == 0) {
            p.act_host0 = static_cast<const float*>(attr.devicePointer);
            zero_copy_act = true;
        } else {
            cudaGetLastError();
        }
    }
    bool zero_copy_arena = false;
    if (zero_copy && (host->reserved0 & 2) && host->arena_host && host->arena_dev && host->arena_bytes > 0) {
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, host->arena_host) == cudaSuccess && attr.type == cudaMemoryTypeHost &&
            attr.devicePointer) {
            p.arena_lo = static_cast<const char*>(host->arena_dev);
            p.arena_hi = p.arena_lo + host->arena_bytes;
            p.arena_delta = (long long)(static_cast<const char*>(attr.devicePointer) - p.arena_lo);
            zero_copy_arena = p.arena_delta != 0;
        } else {
            cudaGetLastError();
        }
    }
    const int n = host->n_chunks < 1 ? 1 : (host->n_chunks > 8 ? 8 : host->n_chunks);
    cudaSetDevice(h->cfg.device);
    for (int c = 0; c < n; ++c)
        if (!h->hstream[c]) H_CUDA(h, cudaStreamCreateWithFlags(&h->hstream[c], cudaStreamNonBlocking));
    for (int c = 0; c <= 8; ++c)
        if (!h->hevent[c]) H_CUDA(h, cudaEventCreateWithFlags(&h->hevent[c], cudaEventDisableTiming));
    // order the internal streams after whatever the caller has queued
    H_CUDA(h, cudaEventRecord(h->hevent[8], (cudaStream_t)caller_stream));
    float* dev_act = const_cast<float*>(io->actions);
    // (a failure below leaves work queued on the internal streams that still reads / writes the caller's host buffers:
    //  the wrapper drains them before the error is returned)
    auto enqueue = [&]() -> int {
    for (int c = 0; c < n; ++c) {
        cudaStream_t st = h->hstream[c];
        const int lo = (int)((int64_t)E * c / n), hi = (int)((int64_t)E * (c + 1) / n), m = hi - lo;
        if (m <= 0) continue;
        H_CUDA(h, cudaStreamWaitEvent(st, h->hevent[8], 0));
        if (!zero_copy_act)
            H_CUDA(h, cudaMemcpyAsync(dev_act + 2 * (size_t)lo, host->actions + 2 * (size_t)lo, (size_t)m * 2 * sizeof(float),
                                      cudaMemcpyHostToDevice, st));
        if (host->selfplay) {  // car 1: obs block [E + lo, E + hi), action block likewise (agent-major)
            if (launch_policy_act(host->opponent_params, D, io->obs + ((size_t)E + lo) * D, D, m, host->seed,
                                  host->counter * 64 + (uint64_t)c, dev_act + 2 * ((size_t)E + lo), 2, nullptr, nullptr,
                                  nullptr, st)) {
                snprintf(h->err, sizeof(h->err), "rk_step_host: opponent inference launch failed: %s",
                         cudaGetErrorString(cudaGetLastError()));
                return 1;
            }
        }
        p.env_begin = lo;
        p.env_end = hi;
        if (launch_step(p, h->cfg.query_mode, h->cfg.env_kind, st)) {
            snprintf(h->err, sizeof(h->err), "rk_step_host: launch failed: %s", cudaGetErrorString(cudaGetLastError()));
            return 1;
        }
        if (!zero_copy)
            H_CUDA(h, cudaMemcpyAsync(host->obs + (size_t)lo * D, io->obs + (size_t)lo * D, (size_t)m * D * sizeof(float),
                                      cudaMemcpyDeviceToHost, st));
        H_CUDA(h, cudaEventRecord(h->hevent[c], st));
    }
    cudaStream_t last = h->hstream[n - 1];
    for (int c = 0; c < n - 1; ++c) H_CUDA(h, cudaStreamWaitEvent(last, h->hevent[c], 0));
    if (!zero_copy_arena && host->arena_host && host->arena_dev && host->arena_bytes > 0)
        H_CUDA(h, cudaMemcpyAsync(host->arena_host, host->arena_dev, (size_t)host->arena_bytes, cudaMemcpyDeviceToHost, last));
    H_CUDA(h, cudaStreamSynchronize(last));
    return 0;
    };
    const int rc = enqueue();
    if (rc)
        for (int c = 0; c < n; ++c)
            if (h->hstream[c]) cudaStreamSynchronize(h->hstream[c]);
    return rc;
}

int rk_rollout(rk_handle h, const rk_step_io* base, const rk_rollout_io* r, void* stream_) {
    if (!h) return 1;
    if (!base || base->struct_size != (int32_t)sizeof(rk_step_io) || !r || r->struct_size != (int32_t)sizeof(rk_rollout_io)) {
        snprintf(h->err, sizeof(h->err), "rk_rollout: bad io structs (struct_size mismatch)");
        return 1;
    }
    const int E = h->cfg.num_envs, A = h->cfg.num_agents, D = h->D;
    if (r->T <= 0 || !r->learner_params || !r->obs || !r->actions || !r->logprobs || !r->values || !r->rewards ||
        !r->dones || !base->terminated || !base->truncated || base->layout != RK_LAYOUT_AGENT_MAJOR ||
        (r->selfplay && A != 2) || (r->block_policy && !r->opponent_params)) {
        snprintf(h->err, sizeof(h->err),
                 "rk_rollout: needs T > 0, learner parameters, the six rollout buffers, terminated/truncated, the "
                 "agent-major layout (and 2 cars for self-play)");
        return 1;
    }
    StepParams p;
    if (fill_params(h, p, "rk_rollout")) return 1;
    cudaStream_t stream = (cudaStream_t)stream_;
    p.mode = 0;
    p.io = *base;
    p.io.env_begin = 0;
    p.io.env_count = 0;
    const size_t obs_slot = (size_t)A * E * D, act_slot = (size_t)A * E * 2, rew_slot = (size_t)A * E;
    for (int t = 0; t < r->T; ++t) {
        float* obs_t = r->obs + (size_t)t * obs_slot;
        float* act_t = r->actions + (size_t)t * act_slot;
        PolicyJobs jobs{};
        // learner: car 0's block of the agent-major slot (agent/ppo.py:109-112)
        jobs.job[0] = PolicyJob{r->learner_params, obs_t, D, E, r->learner_seed, r->learner_counter0 + (uint64_t)t, act_t, 2,
                                r->logprobs + (size_t)t * E, r->values + (size_t)t * E, nullptr, nullptr, 0, 0};
        int n_jobs = 1;
        if (r->selfplay) {  // opponent: car 1's block (wrappers.py:30-39)
            const float* oobs = (t == 0 && r->opponent_obs0) ? r->opponent_obs0 : obs_t + (size_t)E * D;
            jobs.job[1] = PolicyJob{r->opponent_params, oobs, D, E, r->opponent_seed, r->opponent_counter0 + (uint64_t)t,
                                    act_t + (size_t)E * 2, 2, nullptr, nullptr, nullptr, r->block_policy, r->block_len,
                                    r->pool_stride};
            n_jobs = 2;
        }
        if (launch_policy_jobs(jobs, n_jobs, D, stream)) {
            snprintf(h->err, sizeof(h->err), "rk_rollout: inference launch failed at step %d: %s (parameter blocks must be "
                     "16-byte aligned, block_len a multiple of %d)", t, cudaGetErrorString(cudaGetLastError()), RK_POLICY_BLOCK);
            return 1;
        }
        p.io.actions = act_t;
        p.io.obs = obs_t + obs_slot;
        p.io.reward_f32 = r->rewards + (size_t)t * rew_slot;
        p.io.done_f32 = r->dones + (size_t)(t + 1) * E;
        if (launch_step(p, h->cfg.query_mode, h->cfg.env_kind, stream)) {
            snprintf(h->err, sizeof(h->err), "rk_rollout: step launch failed at step %d: %s", t, cudaGetErrorString(cudaGetLastError()));
            return 1;
        }
    }
    return 0;
}

int rk_observe(rk_handle h, float* dev_obs, int32_t layout, void* stream) {
    if (!h) return 1;
    if (!dev_obs) {
        snprintf(h->err, sizeof(h->err), "rk_observe: null obs");
        return 1;
    }
    StepParams p;
    if (fill_params(h, p, "rk_observe")) return 1;
    p.mode = 2;
    p.io.obs = dev_obs;
    p.io.layout = layout;
    if (launch_step(p, h->cfg.query_mode, h->cfg.env_kind, (cudaStream_t)stream)) {
        snprintf(h->err, sizeof(h->err), "rk_observe: launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        return 1;
    }
    return 0;
}

int rk_set_speed_weight(rk_handle h, double w) {
    if (!h) return 1;
    h->cfg.speed_weight = w;
    return 0;
}

int rk_set_seed(rk_handle h, uint64_t seed) {
    if (!h) return 1;
    cudaSetDevice(h->cfg.device);
    h->cfg.seed = seed;
    // the stream restarts: the shuffle of an environment's k-th reset is Philox(seed; env, k)
    H_CUDA(h, cudaDeviceSynchronize());
    H_CUDA(h, cudaMemset(h->st.reset_count, 0, (size_t)h->cfg.num_envs * sizeof(uint32_t)));
    return 0;
}

int rk_get_state(rk_handle h, double* car_f64, int32_t* car_i32, int32_t* env_i32, double* env_f64) {
    if (!h) return 1;
    cudaSetDevice(h->cfg.device);
    H_CUDA(h, cudaDeviceSynchronize());
    const size_t E = h->cfg.num_envs, C = E * h->cfg.num_agents;
    if (car_f64) {
        std::vector<double> t(C);
        std::vector<float> f(C);
        const double* src[5] = {h->st.x, h->st.y, h->st.ang, h->st.vx, h->st.vy};
        for (int k = 0; k < 5; ++k) {
            H_CUDA(h, cudaMemcpy(t.data(), src[k], C * sizeof(double), cudaMemcpyDeviceToHost));
            for (size_t i = 0; i < C; ++i) car_f64[6 * i + k] = t[i];
        }
        H_CUDA(h, cudaMemcpy(f.data(), h->st.last_steer, C * sizeof(float), cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < C; ++i) car_f64[6 * i + 5] = f[i];
    }
    if (car_i32) {
        std::vector<int32_t> t(C);
        const int32_t* src[4] = {h->st.pidx, h->st.lpidx, h->st.flags, h->st.fstep};
        for (int k = 0; k < 4; ++k) {
            H_CUDA(h, cudaMemcpy(t.data(), src[k], C * sizeof(int32_t), cudaMemcpyDeviceToHost));
            for (size_t i = 0; i < C; ++i) car_i32[4 * i + k] = (k == 2) ? (t[i] & ~(int32_t)F_RAYCACHE) : t[i];   // internal bit
        }
    }
    if (env_i32) {
        std::vector<int32_t> t(E);
        std::vector<uint8_t> u(E);
        H_CUDA(h, cudaMemcpy(t.data(), h->st.steps, E * sizeof(int32_t), cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < E; ++i) env_i32[3 * i] = t[i];
        H_CUDA(h, cudaMemcpy(u.data(), h->st.needs_reset, E, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < E; ++i) env_i32[3 * i + 1] = u[i];
        H_CUDA(h, cudaMemcpy(t.data(), h->st.ep_length, E * sizeof(int32_t), cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < E; ++i) env_i32[3 * i + 2] = t[i];
    }
    if (env_f64) H_CUDA(h, cudaMemcpy(env_f64, h->st.ep_return, E * sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}

int rk_set_state(rk_handle h, const double* car_f64, const int32_t* car_i32, const int32_t* env_i32,
                 const double* env_f64) {
    if (!h) return 1;
    cudaSetDevice(h->cfg.device);
    H_CUDA(h, cudaDeviceSynchronize());
    const size_t E = h->cfg.num_envs, C = E * h->cfg.num_agents;
    if (car_f64) {
        std::vector<double> t(C);
        std::vector<float> f(C);
        double* dst[5] = {h->st.x, h->st.y, h->st.ang, h->st.vx, h->st.vy};
        for (int k = 0; k < 5; ++k) {
            for (size_t i = 0; i < C; ++i) t[i] = car_f64[6 * i + k];
            H_CUDA(h, cudaMemcpy(dst[k], t.data(), C * sizeof(double), cudaMemcpyHostToDevice));
        }
        for (size_t i = 0; i < C; ++i) f[i] = (float)car_f64[6 * i + 5];
        H_CUDA(h, cudaMemcpy(h->st.last_steer, f.data(), C * sizeof(float), cudaMemcpyHostToDevice));
    }
    if (car_i32) {
        std::vector<int32_t> t(C);
        int32_t* dst[4] = {h->st.pidx, h->st.lpidx, h->st.flags, h->st.fstep};
        for (int k = 0; k < 4; ++k) {
            for (size_t i = 0; i < C; ++i) t[i] = (k == 2) ? (car_i32[4 * i + k] & ~(int32_t)F_RAYCACHE) : car_i32[4 * i + k];
            H_CUDA(h, cudaMemcpy(dst[k], t.data(), C * sizeof(int32_t), cudaMemcpyHostToDevice));
        }
    }
    if (env_i32) {
        std::vector<int32_t> t(E);
        std::vector<uint8_t> u(E);
        for (size_t i = 0; i < E; ++i) t[i] = env_i32[3 * i];
        H_CUDA(h, cudaMemcpy(h->st.steps, t.data(), E * sizeof(int32_t), cudaMemcpyHostToDevice));
        for (size_t i = 0; i < E; ++i) u[i] = (uint8_t)(env_i32[3 * i + 1] != 0);
        H_CUDA(h, cudaMemcpy(h->st.needs_reset, u.data(), E, cudaMemcpyHostToDevice));
        for (size_t i = 0; i < E; ++i) t[i] = env_i32[3 * i + 2];
        H_CUDA(h, cudaMemcpy(h->st.ep_length, t.data(), E * sizeof(int32_t), cudaMemcpyHostToDevice));
    }
    if (env_f64) H_CUDA(h, cudaMemcpy(h->st.ep_return, env_f64, E * sizeof(double), cudaMemcpyHostToDevice));
    return 0;
}

int rk_gae(const float* rewards, const float* values, const float* dones, const float* next_value,
           const float* next_done_f32, float gamma, float lam, int32_t T, int32_t E, float* adv, float* ret,
           void* stream) {
    if (!rewards || !values || !dones || !next_value || !next_done_f32 || !adv || !ret || T <= 0 || E <= 0) {
        snprintf(g_create_err, sizeof(g_create_err), "rk_gae: invalid arguments");
        return 1;
    }
    return launch_gae(rewards, values, dones, next_value, next_done_f32, gamma, lam, T, E, adv, ret,
                      (cudaStream_t)stream);
}

int rk_policy_param_count(int32_t obs_dim) { return policy_param_count(obs_dim); }

int rk_gather_minibatch(const int64_t* idx, int32_t n, int32_t obs_dim, const float* obs, const float* act,
                        const float* logp, const float* adv, const float* ret, const float* val, float* o_obs,
                        float* o_act, float* o_logp, float* o_adv, float* o_ret, float* o_val, void* stream) {
    if (!idx || n <= 0 || !obs || !act || !logp || !adv || !ret || !val || !o_obs || !o_act || !o_logp || !o_adv ||
        !o_ret || !o_val) {
        snprintf(g_create_err, sizeof(g_create_err), "rk_gather_minibatch: invalid arguments");
        return 1;
    }
    return launch_gather_minibatch(idx, n, obs_dim, obs, act, logp, adv, ret, val, o_obs, o_act, o_logp, o_adv, o_ret,
                                   o_val, (cudaStream_t)stream);
}

int rk_ppo_loss_grad(const float* mu, const float* v, const float* act, const float* old_logp, const float* adv,
                     const float* ret, const float* v_old, const float* log_std, const float* adv_mean,
                     const float* adv_std, int32_t n, float clip_coef, float vf_coef, float* dmu, float* dv,
                     double* kl_sum, void* stream) {
    if (!mu || !v || !act || !old_logp || !adv || !ret || !v_old || !log_std || !adv_mean || !adv_std || n <= 0 ||
        !dmu || !dv || !kl_sum) {
        snprintf(g_create_err, sizeof(g_create_err), "rk_ppo_loss_grad: invalid arguments");
        return 1;
    }
    return launch_ppo_loss_grad(mu, v, act, old_logp, adv, ret, v_old, log_std, adv_mean, adv_std, n, clip_coef,
                                vf_coef, dmu, dv, kl_sum, (cudaStream_t)stream);
}

int rk_policy_act_pool(const float* params_pool, int64_t pool_stride, const int32_t* block_policy, int32_t block_len,
                       int32_t obs_dim, const float* obs, int64_t obs_stride, int32_t B, uint64_t seed,
                       uint64_t counter, float* action, int64_t act_stride, float* logprob, float* value, float* mean,
                       void* stream) {
    if (!params_pool || !block_policy || !action || !obs || B < 0 || obs_dim <= 0 || obs_dim > 256 || block_len <= 0 ||
        block_len % RK_POLICY_BLOCK != 0 || pool_stride < rk_policy_param_count(obs_dim) || (pool_stride & 3) != 0) {
        snprintf(g_create_err, sizeof(g_create_err),
                 "rk_policy_act_pool: invalid arguments (block_len must be a multiple of %d, pool_stride a multiple "
                 "of 4 and >= rk_policy_param_count)", RK_POLICY_BLOCK);
        return 1;
    }
    if (launch_policy_act(params_pool, obs_dim, obs, obs_stride, B, seed, counter, action, act_stride, logprob, value,
                          mean, (cudaStream_t)stream, block_policy, block_len, pool_stride)) {
        snprintf(g_create_err, sizeof(g_create_err), "rk_policy_act_pool: launch failed: %s",
                 cudaGetErrorString(cudaGetLastError()));
        return 1;
    }
    return 0;
}

int rk_ppo_adv_stats(const int64_t* idx, const float* adv, int32_t n, double* part, void* stream) {
    if (!adv || !part || n <= 0) {
        snprintf(g_create_err, sizeof(g_create_err), "rk_ppo_adv_stats: invalid arguments");
        return 1;
    }
    return launch_adv_stats(idx, adv, n, part, (cudaStream_t)stream);
}

uint64_t rk_ppo_grad_workspace_bytes(void) { return (uint64_t)ppo_grad_workspace_bytes(); }

int rk_ppo_minibatch_grad(const rk_ppo_grad_io* io, void* stream) {
    if (!io || io->struct_size != (int32_t)sizeof(rk_ppo_grad_io)) {
        snprintf(g_create_err, sizeof(g_create_err), "rk_ppo_minibatch_grad: io->struct_size mismatch");
        return 1;
    }
    bool ok = io->n > 0 && io->obs_dim > 0 && io->obs_dim <= RK_PPO_MAX_OBS_DIM &&
              (io->obs_stride == 0 || io->obs_stride >= io->obs_dim) && io->n_global >= 2.0 && io->log_std &&
              io->obs && io->act && io->old_logp && io->adv && io->ret && io->val && io->adv_part && io->workspace &&
              io->flat_grad && io->kl_sum;
    for (int k = 0; k < 12; ++k) ok = ok && io->params[k] != nullptr;
    if (!ok) {
        snprintf(g_create_err, sizeof(g_create_err),
                 "rk_ppo_minibatch_grad: invalid arguments (obs_dim must be 1..%d, no NULL pointers except idx)",
                 RK_PPO_MAX_OBS_DIM);
        return 1;
    }
    PpoGradIO g;
    g.obs_dim = io->obs_dim; g.n = io->n; g.n_global = io->n_global;
    g.obs_stride = io->obs_stride > 0 ? io->obs_stride : io->obs_dim;
    for (int k = 0; k < 12; ++k) g.params[k] = io->params[k];
    g.log_std = io->log_std;
    g.obs = io->obs; g.act = io->act; g.old_logp = io->old_logp; g.adv = io->adv; g.ret = io->ret; g.val = io->val;
    g.idx = io->idx; g.adv_part = io->adv_part; g.clip = io->clip_coef; g.vf_coef = io->vf_coef;
    g.workspace = io->workspace; g.workspace_bytes = (size_t)io->workspace_bytes;
    g.flat_grad = io->flat_grad; g.kl_sum = io->kl_sum; g.kl_sum_f32 = io->kl_sum_f32; g.tensor_cores = io->tensor_cores;
    const int rc = launch_ppo_minibatch_grad(g, (cudaStream_t)stream);
    if (rc) snprintf(g_create_err, sizeof(g_create_err), "rk_ppo_minibatch_grad: %s",
                     rc == 3 ? "workspace too small" : rc == 2 ? "unsupported shape" : cudaGetErrorString(cudaGetLastError()));
    return rc ? 1 : 0;
}

int rk_ppo_adam_step(const rk_adam_io* io, void* stream) {
    if (!io || io->struct_size != (int32_t)sizeof(rk_adam_io)) {
        snprintf(g_create_err, sizeof(g_create_err), "rk_ppo_adam_step: io->struct_size mismatch");
        return 1;
    }
    bool ok = io->world >= 1 && io->flat_grad && io->lr && io->kl_sum && io->state && io->kl_at_stop && io->n_global >= 1.0;
    for (int k = 0; k < 12; ++k)
        ok = ok && io->params[k] && io->exp_avg[k] && io->exp_avg_sq[k] && io->step[k] && io->numel[k] > 0;
    if (!ok) {
        snprintf(g_create_err, sizeof(g_create_err), "rk_ppo_adam_step: invalid arguments");
        return 1;
    }
    PpoAdamIO a;
    for (int k = 0; k < 12; ++k) {
        a.params[k] = io->params[k]; a.exp_avg[k] = io->exp_avg[k]; a.exp_avg_sq[k] = io->exp_avg_sq[k];
        a.step[k] = io->step[k]; a.numel[k] = io->numel[k];
    }
    a.flat_grad = io->flat_grad; a.lr = io->lr;
    a.beta1 = io->beta1; a.beta2 = io->beta2; a.eps = io->eps; a.max_norm = io->max_grad_norm; a.kl_target = io->kl_target;
    a.world = io->world; a.kl_sum = io->kl_sum; a.kl_sum_f32 = io->kl_sum_f32; a.n_global = io->n_global; a.state = io->state; a.kl_at_stop = io->kl_at_stop;
    if (launch_clip_adam(a, (cudaStream_t)stream)) {
        snprintf(g_create_err, sizeof(g_create_err), "rk_ppo_adam_step: launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        return 1;
    }
    return 0;
}

int rk_random_permutation(uint64_t seed, uint64_t counter, int64_t n, int64_t* out, void* stream) {
    if (n == 0) return 0;
    if (!out || n < 0 || n > ((int64_t)1 << 40)) {
        snprintf(g_create_err, sizeof(g_create_err), "rk_random_permutation: invalid arguments");
        return 1;
    }
    return launch_permutation(seed, counter, n, out, (cudaStream_t)stream);
}

int rk_policy_act(const float* params, int32_t obs_dim, const float* obs, int64_t obs_stride, int32_t B,
                  uint64_t seed, uint64_t counter, float* action, int64_t act_stride, float* logprob, float* value,
                  float* mean, void* stream) {
    if (!action || B < 0 || (params && (!obs || obs_dim <= 0 || obs_dim > 256))) {
        snprintf(g_create_err, sizeof(g_create_err), "rk_policy_act: invalid arguments");
        return 1;
    }
    if (launch_policy_act(params, obs_dim, obs, obs_stride, B, seed, counter, action, act_stride, logprob, value,
                          mean, (cudaStream_t)stream)) {
        snprintf(g_create_err, sizeof(g_create_err), "rk_policy_act: launch failed: %s",
                 cudaGetErrorString(cudaGetLastError()));
        return 1;
    }
    return 0;
}

}  // extern "C"
