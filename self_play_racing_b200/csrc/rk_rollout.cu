// Rollout-side kernels that keep the PPO data path on the device:
//   gae_kernel        -- PPO.compute_advantages (reference agent/ppo.py:134-154)
//   policy_act_kernel -- Agent.get_action_and_value(obs) with action=None
//                        (agent/ppo.py:43-56), also used for the frozen
//                        opponent of SelfPlayWrapper.step
//                        (environment/wrappers.py:29-39)
#include <math.h>

#include "rk_types.cuh"

namespace rk {

namespace {

// One thread per environment column, serial over T (the recurrence), loads
// coalesced across the warp.  20 B/transition: HBM-bound.
__global__ void gae_kernel(const float* __restrict__ rewards, const float* __restrict__ values,
                           const float* __restrict__ dones, const float* __restrict__ next_value,
                           const float* __restrict__ next_done, float gamma, float gl, int T, int E,
                           float* __restrict__ adv, float* __restrict__ ret) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    float running = 0.f;
    float nnt = __fsub_rn(1.f, next_done[e]);  // ppo.py:141-143
    float nv = next_value[e];
    for (int t = T - 1; t >= 0; --t) {
        const size_t i = (size_t)t * E + e;
        const float v = values[i];
        // ppo.py:149: delta = r + gamma * nnt * V' - V        (left to right)
        const float delta = __fsub_rn(__fadd_rn(rewards[i], __fmul_rn(__fmul_rn(gamma, nnt), nv)), v);
        // ppo.py:151: A = delta + (gamma*lambda) * nnt * A'
        running = __fadd_rn(delta, __fmul_rn(__fmul_rn(gl, nnt), running));
        adv[i] = running;
        ret[i] = __fadd_rn(running, v);  // ppo.py:152
        nnt = __fsub_rn(1.f, dones[i]);  // ppo.py:145-146 for the next (earlier) step
        nv = v;
    }
}

constexpr int kHidden = 64;
constexpr int kPolicyThreads = 128;

// y[j] = b[j] + sum_i Wt[i][j] * x[i], weights transposed in shared memory so
// that a float4 broadcast load feeds four FMAs.
// x lives in shared memory, one column per thread (stride kPolicyThreads), so
// the loop over inputs can stay rolled without dynamic register indexing.
template <int NIN>
__device__ __forceinline__ void dense64(const float* __restrict__ wt, const float* __restrict__ b,
                                        const float* xs, float* y) {
#pragma unroll
    for (int j = 0; j < kHidden; ++j) y[j] = b[j];
#pragma unroll 4
    for (int i = 0; i < NIN; ++i) {
        const float xi = xs[i * kPolicyThreads];
        const float4* row = reinterpret_cast<const float4*>(wt + i * kHidden);
#pragma unroll
        for (int j4 = 0; j4 < kHidden / 4; ++j4) {
            const float4 w = row[j4];
            y[4 * j4 + 0] = fmaf(w.x, xi, y[4 * j4 + 0]);
            y[4 * j4 + 1] = fmaf(w.y, xi, y[4 * j4 + 1]);
            y[4 * j4 + 2] = fmaf(w.z, xi, y[4 * j4 + 2]);
            y[4 * j4 + 3] = fmaf(w.w, xi, y[4 * j4 + 3]);
        }
    }
}

// Generic-input first layer (obs_dim is a runtime value, x in registers up to 96 wide)
__device__ __forceinline__ void dense_in(const float* __restrict__ wt, const float* __restrict__ b, int nin,
                                         const float* __restrict__ obs_row, float* y) {
#pragma unroll
    for (int j = 0; j < kHidden; ++j) y[j] = b[j];
    for (int i = 0; i < nin; ++i) {
        const float xi = obs_row[i];
        const float4* row = reinterpret_cast<const float4*>(wt + i * kHidden);
#pragma unroll
        for (int j4 = 0; j4 < kHidden / 4; ++j4) {
            const float4 w = row[j4];
            y[4 * j4 + 0] = fmaf(w.x, xi, y[4 * j4 + 0]);
            y[4 * j4 + 1] = fmaf(w.y, xi, y[4 * j4 + 1]);
            y[4 * j4 + 2] = fmaf(w.z, xi, y[4 * j4 + 2]);
            y[4 * j4 + 3] = fmaf(w.w, xi, y[4 * j4 + 3]);
        }
    }
}

// shared-memory layout (floats): [actor W0t | b0 | W2t | b2 | W4 | b4 | log_std | critic W0t | b0 | W2t | b2 | W4 | b4]
__global__ void __launch_bounds__(kPolicyThreads)
policy_act_kernel(const float* __restrict__ params, int obs_dim, const float* __restrict__ obs, int64_t obs_stride,
                  int B, uint64_t seed, uint64_t counter, float* __restrict__ action, int64_t act_stride,
                  float* __restrict__ logprob, float* __restrict__ value, float* __restrict__ mean) {
    extern __shared__ __align__(16) float sm[];
    const int n0 = obs_dim * kHidden;
    float* aW0 = sm;            float* ab0 = aW0 + n0;
    float* aW2 = ab0 + kHidden; float* ab2 = aW2 + kHidden * kHidden;
    float* aW4 = ab2 + kHidden; float* ab4 = aW4 + 2 * kHidden;
    float* lstd = ab4 + 2;
    float* cW0 = lstd + 2;      float* cb0 = cW0 + n0;
    float* cW2 = cb0 + kHidden; float* cb2 = cW2 + kHidden * kHidden;
    float* cW4 = cb2 + kHidden;
    float* hbuf = cW4 + kHidden + 4 + threadIdx.x;  // [kHidden][kPolicyThreads] activations, column per thread
    {
        // global layout: torch [out, in] row-major, in state-dict order (see racing_b200.h)
        const float* g = params;
        for (int k = threadIdx.x; k < n0; k += blockDim.x) aW0[(k % obs_dim) * kHidden + k / obs_dim] = g[k];
        g += n0;
        for (int k = threadIdx.x; k < kHidden; k += blockDim.x) ab0[k] = g[k];
        g += kHidden;
        for (int k = threadIdx.x; k < kHidden * kHidden; k += blockDim.x) aW2[(k % kHidden) * kHidden + k / kHidden] = g[k];
        g += kHidden * kHidden;
        for (int k = threadIdx.x; k < kHidden; k += blockDim.x) ab2[k] = g[k];
        g += kHidden;
        for (int k = threadIdx.x; k < 2 * kHidden + 2 + 2; k += blockDim.x) aW4[k] = g[k];  // W4, b4, log_std
        g += 2 * kHidden + 4;
        if (value != nullptr) {
            for (int k = threadIdx.x; k < n0; k += blockDim.x) cW0[(k % obs_dim) * kHidden + k / obs_dim] = g[k];
            g += n0;
            for (int k = threadIdx.x; k < kHidden; k += blockDim.x) cb0[k] = g[k];
            g += kHidden;
            for (int k = threadIdx.x; k < kHidden * kHidden; k += blockDim.x) cW2[(k % kHidden) * kHidden + k / kHidden] = g[k];
            g += kHidden * kHidden;
            for (int k = threadIdx.x; k < kHidden; k += blockDim.x) cb2[k] = g[k];
            g += kHidden;
            for (int k = threadIdx.x; k < kHidden + 1; k += blockDim.x) cW4[k] = g[k];
        }
    }
    __syncthreads();
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float* orow = obs + (size_t)b * obs_stride;
    float h1[kHidden], h2[kHidden];
    // actor: Linear-Tanh-Linear-Tanh-Linear-Tanh (ppo.py:19-26)
    dense_in(aW0, ab0, obs_dim, orow, h1);
#pragma unroll
    for (int j = 0; j < kHidden; ++j) hbuf[j * kPolicyThreads] = tanhf(h1[j]);
    dense64<kHidden>(aW2, ab2, hbuf, h2);
    float m0 = ab4[0], m1 = ab4[1];
#pragma unroll
    for (int j = 0; j < kHidden; ++j) {
        const float t = tanhf(h2[j]);
        m0 = fmaf(aW4[j], t, m0);
        m1 = fmaf(aW4[kHidden + j], t, m1);
    }
    m0 = tanhf(m0);
    m1 = tanhf(m1);
    if (mean != nullptr) {
        mean[2 * (size_t)b] = m0;
        mean[2 * (size_t)b + 1] = m1;
    }
    // a ~ N(mu, exp(log_std)) clamped to [-1, 1] (ppo.py:47-54); Box-Muller on Philox
    uint32_t c[4] = {(uint32_t)b, (uint32_t)counter, (uint32_t)(counter >> 32), 0x706f6c79u};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    const float rad = sqrtf(-2.f * logf(u01(c[0])));
    float sn, cs;
    sincosf(6.2831853071795865f * u01(c[1]), &sn, &cs);
    const float s0 = expf(lstd[0]), s1 = expf(lstd[1]);
    const float a0 = fminf(fmaxf(fmaf(s0, rad * cs, m0), -1.f), 1.f);
    const float a1 = fminf(fmaxf(fmaf(s1, rad * sn, m1), -1.f), 1.f);
    action[(size_t)b * act_stride] = a0;
    action[(size_t)b * act_stride + 1] = a1;
    if (logprob != nullptr) {
        // torch Normal.log_prob: -(a-mu)^2 / (2 var) - log_std - log(sqrt(2 pi)), summed (ppo.py:56)
        const float kLogSqrt2Pi = 0.9189385332046727f;
        const float l0 = -((a0 - m0) * (a0 - m0)) / (2.f * s0 * s0) - lstd[0] - kLogSqrt2Pi;
        const float l1 = -((a1 - m1) * (a1 - m1)) / (2.f * s1 * s1) - lstd[1] - kLogSqrt2Pi;
        logprob[b] = l0 + l1;
    }
    if (value != nullptr) {
        // critic: Linear-Tanh-Linear-Tanh-Linear (ppo.py:31-37)
        dense_in(cW0, cb0, obs_dim, orow, h1);
#pragma unroll
        for (int j = 0; j < kHidden; ++j) hbuf[j * kPolicyThreads] = tanhf(h1[j]);
        dense64<kHidden>(cW2, cb2, hbuf, h2);
        float v = cW4[kHidden];
#pragma unroll
        for (int j = 0; j < kHidden; ++j) v = fmaf(cW4[j], tanhf(h2[j]), v);
        value[b] = v;
    }
}

// Pool-empty opponent: Box([-1,0],[1,1]).sample() (wrappers.py:30-32, multi_racing_env.py:28-35)
__global__ void random_act_kernel(int B, uint64_t seed, uint64_t counter, float* __restrict__ action,
                                  int64_t act_stride) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    uint32_t c[4] = {(uint32_t)b, (uint32_t)counter, (uint32_t)(counter >> 32), 0x72616e64u};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    action[(size_t)b * act_stride] = 2.f * u01(c[0]) - 1.f;
    action[(size_t)b * act_stride + 1] = u01(c[1]);
}

}  // namespace

int launch_gae(const float* rewards, const float* values, const float* dones, const float* next_value,
               const float* next_done, float gamma, float lam, int T, int E, float* adv, float* ret,
               cudaStream_t stream) {
    const float gl = (float)((double)gamma * (double)lam);
    gae_kernel<<<(E + 127) / 128, 128, 0, stream>>>(rewards, values, dones, next_value, next_done, gamma, gl, T, E,
                                                    adv, ret);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int policy_param_count(int obs_dim) {
    return 2 * (obs_dim * kHidden + kHidden + kHidden * kHidden + kHidden) + (2 * kHidden + 2) + 2 + (kHidden + 1);
}

int launch_policy_act(const float* params, int obs_dim, const float* obs, int64_t obs_stride, int B, uint64_t seed,
                      uint64_t counter, float* action, int64_t act_stride, float* logprob, float* value,
                      float* mean, cudaStream_t stream) {
    if (B <= 0) return 0;
    if (params == nullptr) {
        random_act_kernel<<<(B + 255) / 256, 256, 0, stream>>>(B, seed, counter, action, act_stride);
        count_launch();
        return cudaGetLastError() == cudaSuccess ? 0 : 1;
    }
    const size_t smem = ((size_t)policy_param_count(obs_dim) + 8 + (size_t)kHidden * kPolicyThreads) * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(policy_act_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        attr_set = true;
    }
    policy_act_kernel<<<(B + kPolicyThreads - 1) / kPolicyThreads, kPolicyThreads, smem, stream>>>(
        params, obs_dim, obs, obs_stride, B, seed, counter, action, act_stride, logprob, value, mean);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace rk
