// Rollout-side kernels that keep the PPO data path on the device:
//   gae_kernel        -- PPO.compute_advantages (reference agent/ppo.py:134-154)
//   policy_act_kernel -- Agent.get_action_and_value(obs) with action=None
//                        (agent/ppo.py:43-56), also used for the frozen
//                        opponent of SelfPlayWrapper.step
//                        (environment/wrappers.py:29-39)
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#include "rk_types.cuh"
#include "rk_ppo_loss.cuh"
#include "rk_umma.cuh"

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
    constexpr int kBatch = 8;  // loads of kBatch time steps are issued together: the recurrence is serial, the loads are not
    int t = T - 1;
    for (; t >= kBatch - 1; t -= kBatch) {
        float r[kBatch], v[kBatch], d[kBatch];
#pragma unroll
        for (int k = 0; k < kBatch; ++k) {
            const size_t i = (size_t)(t - k) * E + e;
            r[k] = rewards[i]; v[k] = values[i]; d[k] = dones[i];
        }
#pragma unroll
        for (int k = 0; k < kBatch; ++k) {
            const size_t i = (size_t)(t - k) * E + e;
            // ppo.py:149: delta = r + gamma * nnt * V' - V (left to right); ppo.py:151: A = delta + (gamma*lambda) * nnt * A'
            const float delta = __fsub_rn(__fadd_rn(r[k], __fmul_rn(__fmul_rn(gamma, nnt), nv)), v[k]);
            running = __fadd_rn(delta, __fmul_rn(__fmul_rn(gl, nnt), running));
            adv[i] = running;
            ret[i] = __fadd_rn(running, v[k]);  // ppo.py:152
            nnt = __fsub_rn(1.f, d[k]);         // ppo.py:145-146 for the next (earlier) step
            nv = v[k];
        }
    }
    for (; t >= 0; --t) {
        const size_t i = (size_t)t * E + e;
        const float v = values[i];
        const float delta = __fsub_rn(__fadd_rn(rewards[i], __fmul_rn(__fmul_rn(gamma, nnt), nv)), v);
        running = __fadd_rn(delta, __fmul_rn(__fmul_rn(gl, nnt), running));
        adv[i] = running;
        ret[i] = __fadd_rn(running, v);
        nnt = __fsub_rn(1.f, dones[i]);
        nv = v;
    }
}

constexpr int kHidden = 64;
constexpr int kPolicyThreads = 128;           // threads per CTA
constexpr int kSamplesPerThread = 2;          // register tile: every weight fetched feeds 2 samples
constexpr int kPolicyCols = kPolicyThreads * kSamplesPerThread;

// Packed parameter block (floats), produced by backend.flatten_agent(): every weight
// matrix of a hidden layer is stored TRANSPOSED ([in][out], so one float4 broadcast
// load feeds four output neurons), padded to a multiple of 4 floats for the bulk copy:
//   actor : W0t[obs][64] b0[64] W2t[64][64] b2[64] W4[2][64] b4[2] log_std[2]
//   critic: W0t[obs][64] b0[64] W2t[64][64] b2[64] W4[64] b4[1]  (+ padding)
__host__ __device__ inline int policy_packed_floats(int obs_dim) {
    const int n = 2 * (obs_dim * kHidden + kHidden + kHidden * kHidden + kHidden) + (2 * kHidden + 2) + 2 + (kHidden + 1);
    return (n + 3) / 4 * 4;
}

// tanh_fast (rk_umma.cuh): ex2.approx + rcp.approx, |error| < 2e-7 absolute

// y[s][j] = b[j] + sum_i Wt[i][j] * x_s[i] for the thread's two samples.  x lives
// in shared memory, one column per sample (stride kPolicyCols), so the loop over
// inputs stays rolled with static register indexing of the 2 x 64 accumulators.
__device__ __forceinline__ void dense2(const float* __restrict__ wt, const float* __restrict__ b, int nin,
                                       const float* xs, float (&y0)[kHidden], float (&y1)[kHidden]) {
    // packed fp32 FMA (SASS FFMA2, activation broadcast): 64 instead of 128 FMA instructions per input,
    // bit-identical to scalar fmaf
    float2 p0[kHidden / 2], p1[kHidden / 2];
#pragma unroll
    for (int j = 0; j < kHidden / 2; ++j) { p0[j] = make_float2(b[2 * j], b[2 * j + 1]); p1[j] = p0[j]; }
#pragma unroll 2
    for (int i = 0; i < nin; ++i) {
        const float xa = xs[i * kPolicyCols], xb = xs[i * kPolicyCols + kPolicyThreads];
        const float2 xa2 = make_float2(xa, xa), xb2 = make_float2(xb, xb);
        const float4* row = reinterpret_cast<const float4*>(wt + i * kHidden);
#pragma unroll
        for (int j4 = 0; j4 < kHidden / 4; ++j4) {
            const float4 w = row[j4];
            const float2 wl = make_float2(w.x, w.y), wh = make_float2(w.z, w.w);
            p0[2 * j4] = __ffma2_rn(wl, xa2, p0[2 * j4]); p1[2 * j4] = __ffma2_rn(wl, xb2, p1[2 * j4]);
            p0[2 * j4 + 1] = __ffma2_rn(wh, xa2, p0[2 * j4 + 1]); p1[2 * j4 + 1] = __ffma2_rn(wh, xb2, p1[2 * j4 + 1]);
        }
    }
#pragma unroll
    for (int j = 0; j < kHidden / 2; ++j) {
        y0[2 * j] = p0[j].x; y0[2 * j + 1] = p0[j].y;
        y1[2 * j] = p1[j].x; y1[2 * j + 1] = p1[j].y;
    }
}

__device__ __forceinline__ void stage_obs(float* xs, const float* __restrict__ obs, int64_t obs_stride, int obs_dim,
                                          int b0, int b1, int B) {
    for (int i = 0; i < obs_dim; ++i) {
        xs[i * kPolicyCols] = (b0 < B) ? obs[(size_t)b0 * obs_stride + i] : 0.f;
        xs[i * kPolicyCols + kPolicyThreads] = (b1 < B) ? obs[(size_t)b1 * obs_stride + i] : 0.f;
    }
}

// Fused Agent.get_action_and_value(obs) (agent/ppo.py:43-56).  The packed
// parameter block arrives in shared memory through ONE bulk asynchronous copy
// (cp.async.bulk -> mbarrier) while the threads stage their observations.
//
// The launch carries up to two independent jobs (blockIdx.y): the rollout runs the learner's forward pass and the
// frozen opponent's (self_play: agent/ppo.py:109 and environment/wrappers.py:35-39 of the same step) as ONE grid, so
// that the 256-sample CTAs of both fill the 148 SMs together.  A job without parameters draws the pool-empty
// opponent's uniform Box actions (wrappers.py:30-32).
__global__ void __launch_bounds__(kPolicyThreads, 2)
policy_act_kernel(const PolicyJobs jobs, int obs_dim) {
    extern __shared__ __align__(128) float sm[];
    const PolicyJob& jb = jobs.job[blockIdx.y];
    const int B = jb.B;
    if (blockIdx.x * kPolicyCols >= B) return;
    const float* __restrict__ params = jb.params;
    const float* __restrict__ obs = jb.obs;
    const int64_t obs_stride = jb.obs_stride, act_stride = jb.act_stride, pool_stride = jb.pool_stride;
    const uint64_t seed = jb.seed, counter = jb.counter;
    float* __restrict__ action = jb.action;
    float* __restrict__ logprob = jb.logprob;
    float* __restrict__ value = jb.value;
    float* __restrict__ mean = jb.mean;
    const int32_t* __restrict__ block_policy = jb.block_policy;
    const int block_len = jb.block_len;
    if (params == nullptr) {   // uniform Box([-1,0],[1,1]) samples, the same stream as random_act_kernel
        for (int b = blockIdx.x * kPolicyCols + threadIdx.x; b < min(B, (int)(blockIdx.x + 1) * kPolicyCols); b += kPolicyThreads) {
            uint32_t c[4] = {(uint32_t)b, (uint32_t)counter, (uint32_t)(counter >> 32), 0x72616e64u};
            philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
            action[(size_t)b * act_stride] = 2.f * u01(c[0]) - 1.f;
            action[(size_t)b * act_stride + 1] = u01(c[1]);
        }
        return;
    }
    // a pool of stacked parameter blocks: this CTA's samples all belong to one block of `block_len` samples,
    // whose policy id selects the block it stages (self-play against several snapshots in one launch)
    if (block_policy != nullptr) params += (int64_t)block_policy[(blockIdx.x * kPolicyCols) / block_len] * pool_stride;
    __shared__ __align__(8) unsigned long long bar;
    const int n_packed = policy_packed_floats(obs_dim);
    const int n0 = obs_dim * kHidden;
    const float* aW0 = sm;             const float* ab0 = aW0 + n0;
    const float* aW2 = ab0 + kHidden;  const float* ab2 = aW2 + kHidden * kHidden;
    const float* aW4 = ab2 + kHidden;  const float* ab4 = aW4 + 2 * kHidden;
    const float* lstd = ab4 + 2;
    const float* cW0 = lstd + 2;       const float* cb0 = cW0 + n0;
    const float* cW2 = cb0 + kHidden;  const float* cb2 = cW2 + kHidden * kHidden;
    const float* cW4 = cb2 + kHidden;
    float* xs = sm + n_packed + threadIdx.x;  // this thread's two activation columns
    const unsigned bar_addr = (unsigned)__cvta_generic_to_shared(&bar);
    if (threadIdx.x == 0) {
        const unsigned bytes = (unsigned)n_packed * 4u;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_addr));
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_addr), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"((unsigned)__cvta_generic_to_shared(sm)), "l"(params), "r"(bytes), "r"(bar_addr) : "memory");
    }
    const int b0 = blockIdx.x * kPolicyCols + threadIdx.x, b1 = b0 + kPolicyThreads;
    stage_obs(xs, obs, obs_stride, obs_dim, b0, b1, B);
    __syncthreads();  // the barrier is initialised before anybody polls it
    {
        unsigned done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(bar_addr) : "memory");
    }
    float y0[kHidden], y1[kHidden];
    // actor: Linear-Tanh-Linear-Tanh-Linear-Tanh (ppo.py:19-26)
    dense2(aW0, ab0, obs_dim, xs, y0, y1);
#pragma unroll
    for (int j = 0; j < kHidden; ++j) {
        xs[j * kPolicyCols] = tanh_fast(y0[j]);
        xs[j * kPolicyCols + kPolicyThreads] = tanh_fast(y1[j]);
    }
    dense2(aW2, ab2, kHidden, xs, y0, y1);
    float m[2][2] = {{ab4[0], ab4[1]}, {ab4[0], ab4[1]}};
#pragma unroll
    for (int j = 0; j < kHidden; ++j) {
        const float t0 = tanh_fast(y0[j]), t1 = tanh_fast(y1[j]);
        m[0][0] = fmaf(aW4[j], t0, m[0][0]); m[0][1] = fmaf(aW4[kHidden + j], t0, m[0][1]);
        m[1][0] = fmaf(aW4[j], t1, m[1][0]); m[1][1] = fmaf(aW4[kHidden + j], t1, m[1][1]);
    }
    const float s0 = __expf(lstd[0]), s1 = __expf(lstd[1]);
#pragma unroll
    for (int s = 0; s < kSamplesPerThread; ++s) {
        const int b = s ? b1 : b0;
        if (b >= B) continue;
        const float m0 = tanh_fast(m[s][0]), m1 = tanh_fast(m[s][1]);
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
    }
    if (value != nullptr) {
        // critic: Linear-Tanh-Linear-Tanh-Linear (ppo.py:31-37); the activation columns were
        // overwritten by the actor, so the observations are staged again (L2 resident)
        stage_obs(xs, obs, obs_stride, obs_dim, b0, b1, B);
        dense2(cW0, cb0, obs_dim, xs, y0, y1);
#pragma unroll
        for (int j = 0; j < kHidden; ++j) {
            xs[j * kPolicyCols] = tanh_fast(y0[j]);
            xs[j * kPolicyCols + kPolicyThreads] = tanh_fast(y1[j]);
        }
        dense2(cW2, cb2, kHidden, xs, y0, y1);
        float v0 = cW4[kHidden], v1 = cW4[kHidden];
#pragma unroll
        for (int j = 0; j < kHidden; ++j) {
            v0 = fmaf(cW4[j], tanh_fast(y0[j]), v0);
            v1 = fmaf(cW4[j], tanh_fast(y1[j]), v1);
        }
        if (b0 < B) value[b0] = v0;
        if (b1 < B) value[b1] = v1;
    }
}


// ---------------------------------------------------------------------------
// policy_act_tc_kernel -- the same entry point on the 5th-generation tensor cores (default; RK_B200_POLICY_TC=0 selects
// the FFMA kernel above).  The two 64-wide layers of each MLP run as tcgen05.mma kind::tf32 products with the 3-term
// hi/lo split (fp32-grade accuracy, rk_umma.cuh) chained through tensor memory exactly like the forward half of the
// gradient kernel (rk_train.cu): a thread pair owns a sample = a TMEM lane, reads its accumulator row with tcgen05.ld,
// applies bias / tanh and writes the row back as the A operand of the next layer; the weights sit in shared memory as
// K-major hi / lo tiles (staged once per CTA, re-staged only when a pool launch crosses into a block that plays another
// snapshot).  A CTA walks a contiguous range of 128-sample tiles; two CTAs per SM (256 TMEM columns each) overlap
// one's tensor-core waits with the other's epilogues.  The critic re-uses the observation operand already in TMEM.
// Sampling (Philox Box-Muller, clamp, log-prob) is the code of the FFMA kernel: same counters, same noise.
// ---------------------------------------------------------------------------
constexpr int kPT = 256;      // threads: warps w and w + 4 share TMEM lane quarter w & 3 and split the 64 columns
constexpr int kPXK = 24;      // layer-1 reduction length padded to a multiple of the MMA K
constexpr int kPcAcc = 0, kPcXh = 64, kPcXl = 96, kPcHh = 128, kPcHl = 192, kPcCols = 256;
// per network: W1h W1l [64][24], W2h W2l [64][64], b1 [64], b2 [64], W3 [2][64], b3 [4]
constexpr int kPoW1l = kHidden * kPXK, kPoW2h = 2 * kHidden * kPXK, kPoW2l = kPoW2h + kHidden * kHidden,
              kPob1 = kPoW2l + kHidden * kHidden, kPob2 = kPob1 + kHidden, kPoW3 = kPob2 + kHidden, kPob3 = kPoW3 + 2 * kHidden,
              kPNetFloats = kPob3 + 4;
constexpr size_t kPolicyTcSmemFloats = 2 * (size_t)kPNetFloats + 4 * kUmmaTS;

__global__ void __launch_bounds__(kPT, 2) policy_act_tc_kernel(const PolicyJobs jobs, int obs_dim) {
    extern __shared__ __align__(1024) float sm[];
    const PolicyJob& jb = jobs.job[blockIdx.y];
    const int B = jb.B;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float* __restrict__ action = jb.action;
    const int64_t act_stride = jb.act_stride;
    const uint64_t seed = jb.seed, counter = jb.counter;
    if (jb.params == nullptr) {   // uniform Box([-1,0],[1,1]) samples, the same stream as random_act_kernel
        for (int b = blockIdx.x * kPT + tid; b < B; b += gridDim.x * kPT) {
            uint32_t c[4] = {(uint32_t)b, (uint32_t)counter, (uint32_t)(counter >> 32), 0x72616e64u};
            philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
            action[(size_t)b * act_stride] = 2.f * u01(c[0]) - 1.f;
            action[(size_t)b * act_stride + 1] = u01(c[1]);
        }
        return;
    }
    const int ntiles = (B + kUmmaTS - 1) / kUmmaTS;
    const int per_cta = (ntiles + (int)gridDim.x - 1) / (int)gridDim.x;
    const int t_begin = blockIdx.x * per_cta, t_end = min(ntiles, t_begin + per_cta);
    if (t_begin >= t_end) return;   // (the whole CTA: nothing has been allocated yet)
    const bool want_v = jb.value != nullptr;
    const int nnet = want_v ? 2 : 1;
    const int half = warp >> 2, srow = (warp & 3) * 32 + lane, cbase = half * 32;
    float* OP = sm + 2 * kPNetFloats;   // [2 halves][2 outputs][128] partial output-layer sums
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint32_t tmem_slot;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(kPcCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    unsigned phase = 0;
    const int n0 = obs_dim * kHidden;
    float s0 = 1.f, s1 = 1.f, ls0 = 0.f, ls1 = 0.f;
    int cur_pid = -2;
    for (int tile = t_begin; tile < t_end; ++tile) {
        // ---- the parameter block of this tile's samples (a pool launch: the block's snapshot) ----
        const int pid = jb.block_policy != nullptr ? (int)jb.block_policy[(tile * kUmmaTS) / jb.block_len] : -1;
        if (pid != cur_pid) {
            cur_pid = pid;
            const float* __restrict__ pp = jb.params + (pid >= 0 ? (int64_t)pid * jb.pool_stride : 0);
            __syncthreads();   // (every thread has finished with the previous block's biases / output weights)
            for (int net = 0; net < nnet; ++net) {
                // packed block: actor W0t[obs][64] b0 W2t[64][64] b2 W4[2][64] b4[2] log_std[2]; critic W0t b0 W2t b2 W4[64] b4[1]
                const float* __restrict__ src = pp + (net == 0 ? 0 : n0 + kHidden + kHidden * kHidden + kHidden + 2 * kHidden + 4);
                float* dst = sm + net * kPNetFloats;
                // (all of a thread's loads are issued before the first split: the block comes from L2)
                {
                    float w[kPXK * kHidden / kPT];
#pragma unroll
                    for (int k = 0; k < kPXK * kHidden / kPT; ++k) {
                        const int q = tid + k * kPT, i = q >> 6;
                        w[k] = i < obs_dim ? src[q] : 0.f;
                    }
#pragma unroll
                    for (int k = 0; k < kPXK * kHidden / kPT; ++k) {
                        const int q = tid + k * kPT, i = q >> 6, j = q & 63;
                        uint32_t hi, lo;
                        split_tf32_fast(w[k], hi, lo);
                        dst[umma_off(j, i, kPXK)] = __uint_as_float(hi);
                        dst[kPoW1l + umma_off(j, i, kPXK)] = __uint_as_float(lo);
                    }
                }
                const float* __restrict__ w2 = src + n0 + kHidden;
                {
                    float w[kHidden * kHidden / kPT];
#pragma unroll
                    for (int k = 0; k < kHidden * kHidden / kPT; ++k) w[k] = w2[tid + k * kPT];
#pragma unroll
                    for (int k = 0; k < kHidden * kHidden / kPT; ++k) {
                        const int q = tid + k * kPT, i = q >> 6, j = q & 63;
                        uint32_t hi, lo;
                        split_tf32_fast(w[k], hi, lo);
                        dst[kPoW2h + umma_off(j, i, kHidden)] = __uint_as_float(hi);
                        dst[kPoW2l + umma_off(j, i, kHidden)] = __uint_as_float(lo);
                    }
                }
                const int nout = net == 0 ? 2 : 1;
                for (int q = tid; q < kHidden; q += kPT) { dst[kPob1 + q] = src[n0 + q]; dst[kPob2 + q] = w2[kHidden * kHidden + q]; }
                const float* __restrict__ w4 = w2 + kHidden * kHidden + kHidden;
                for (int q = tid; q < nout * kHidden + nout; q += kPT) dst[kPoW3 + (q < nout * kHidden ? q : 2 * kHidden + q - nout * kHidden)] = w4[q];
            }
            const float* __restrict__ lstd = pp + n0 + kHidden + kHidden * kHidden + kHidden + 2 * kHidden + 2;
            ls0 = lstd[0]; ls1 = lstd[1];
            s0 = __expf(ls0); s1 = __expf(ls1);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the weight tiles are read by the tensor core
            __syncthreads();
        }
        // ---- this tile's observation rows: two threads per sample (columns 0-15 / 16-23), hi + lo into TMEM ----
        const int b = tile * kUmmaTS + srow;
        const bool valid = b < B;
        {
            const float* __restrict__ src = jb.obs + (size_t)(valid ? b : 0) * jb.obs_stride;
#pragma unroll
            for (int c8 = 0; c8 < 16; c8 += 8) {
                if (c8 == 0 || half == 0) {
                    uint32_t hi[8], lo[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const int i = half * 16 + c8 + c;
                        split_tf32_fast((valid && i < obs_dim) ? src[i] : 0.f, hi[c], lo[c]);
                    }
                    tmem_st8(lane_base + kPcXh + half * 16 + c8, hi);
                    tmem_st8(lane_base + kPcXl + half * 16 + c8, lo);
                }
            }
        }
        tmem_publish_and_sync();
        for (int net = 0; net < nnet; ++net) {
            const float* W = sm + net * kPNetFloats;
            const int nout = net == 0 ? 2 : 1;
            if (warp == 0 && elect_one()) issue_product(tmem, kPcAcc, kPcXh, kPcXl, W, W + kPoW1l, kPXK, &bar);
            wait_product(&bar, phase);
            {
                uint32_t v[32];
                tmem_ld32(lane_base + kPcAcc + cbase, v);
#pragma unroll
                for (int c8 = 0; c8 < 32; c8 += 8) {
                    uint32_t hi[8], lo[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        split_tf32_fast(tanh_fast(__uint_as_float(v[c8 + c]) + W[kPob1 + cbase + c8 + c]), hi[c], lo[c]);
                    tmem_st8(lane_base + kPcHh + cbase + c8, hi);
                    tmem_st8(lane_base + kPcHl + cbase + c8, lo);
                }
            }
            tmem_publish_and_sync();
            if (warp == 0 && elect_one()) issue_product(tmem, kPcAcc, kPcHh, kPcHl, W + kPoW2h, W + kPoW2l, kHidden, &bar);
            wait_product(&bar, phase);
            {
                float out[2] = {0.f, 0.f};
                uint32_t v[32];
                tmem_ld32(lane_base + kPcAcc + cbase, v);
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    const int j = cbase + c;
                    const float h = tanh_fast(__uint_as_float(v[c]) + W[kPob2 + j]);
                    out[0] = fmaf(W[kPoW3 + j], h, out[0]);
                    if (nout == 2) out[1] = fmaf(W[kPoW3 + kHidden + j], h, out[1]);
                }
                OP[(half * 2 + 0) * kUmmaTS + srow] = out[0];
                OP[(half * 2 + 1) * kUmmaTS + srow] = out[1];
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();
            if (half == 0 && valid) {
                const float o0 = W[kPob3] + (OP[srow] + OP[2 * kUmmaTS + srow]);
                if (net == 0) {
                    const float o1 = W[kPob3 + 1] + (OP[kUmmaTS + srow] + OP[3 * kUmmaTS + srow]);
                    const float m0 = tanh_fast(o0), m1 = tanh_fast(o1);
                    if (jb.mean != nullptr) {
                        jb.mean[2 * (size_t)b] = m0;
                        jb.mean[2 * (size_t)b + 1] = m1;
                    }
                    // a ~ N(mu, exp(log_std)) clamped to [-1, 1] (ppo.py:47-54); Box-Muller on Philox
                    uint32_t c[4] = {(uint32_t)b, (uint32_t)counter, (uint32_t)(counter >> 32), 0x706f6c79u};
                    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
                    const float rad = sqrtf(-2.f * logf(u01(c[0])));
                    float sn, cs;
                    sincosf(6.2831853071795865f * u01(c[1]), &sn, &cs);
                    const float a0 = fminf(fmaxf(fmaf(s0, rad * cs, m0), -1.f), 1.f);
                    const float a1 = fminf(fmaxf(fmaf(s1, rad * sn, m1), -1.f), 1.f);
                    action[(size_t)b * act_stride] = a0;
                    action[(size_t)b * act_stride + 1] = a1;
                    if (jb.logprob != nullptr) {
                        // torch Normal.log_prob: -(a-mu)^2 / (2 var) - log_std - log(sqrt(2 pi)), summed (ppo.py:56)
                        const float kLogSqrt2Pi = 0.9189385332046727f;
                        const float l0 = -((a0 - m0) * (a0 - m0)) / (2.f * s0 * s0) - ls0 - kLogSqrt2Pi;
                        const float l1 = -((a1 - m1) * (a1 - m1)) / (2.f * s1 * s1) - ls1 - kLogSqrt2Pi;
                        jb.logprob[b] = l0 + l1;
                    }
                } else {
                    jb.value[b] = o0;
                }
            }
            __syncthreads();   // OP is rewritten by the next network / tile
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kPcCols));
}

// ---------------------------------------------------------------------------
// PPO update helpers (reference agent/ppo.py:156-209): one launch gathers a
// minibatch, one launch turns the network outputs into the loss gradients.
// ---------------------------------------------------------------------------
// dst[k] = src[idx[k]] for the six per-sample arrays of a minibatch (ppo.py:170-176,187-195)
__global__ void gather_minibatch_kernel(const int64_t* __restrict__ idx, int n, int obs_dim,
                                        const float* __restrict__ obs, const float* __restrict__ act,
                                        const float* __restrict__ logp, const float* __restrict__ adv,
                                        const float* __restrict__ ret, const float* __restrict__ val,
                                        float* __restrict__ o_obs, float* __restrict__ o_act, float* __restrict__ o_logp,
                                        float* __restrict__ o_adv, float* __restrict__ o_ret, float* __restrict__ o_val) {
    // a warp handles kRows sample rows at once (their random-row loads are all in flight
    // together): lanes copy the obs_dim observation floats, lanes 0..5 the scalars
    constexpr int kRows = 4;
    const int lane = threadIdx.x & 31;
    const int row0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * kRows;
    if (row0 >= n) return;
    int64_t src[kRows];
    float o[kRows], sc[kRows];
#pragma unroll
    for (int k = 0; k < kRows; ++k) src[k] = idx[min(row0 + k, n - 1)];
    const float* scal = lane == 2 ? logp : lane == 3 ? adv : lane == 4 ? ret : val;
#pragma unroll
    for (int k = 0; k < kRows; ++k) {
        o[k] = lane < obs_dim ? obs[(size_t)src[k] * obs_dim + lane] : 0.f;
        sc[k] = lane < 2 ? act[2 * (size_t)src[k] + lane] : (lane < 6 ? scal[src[k]] : 0.f);
    }
#pragma unroll
    for (int k = 0; k < kRows; ++k) {
        const int row = row0 + k;
        if (row >= n) break;
        if (lane < obs_dim) o_obs[(size_t)row * obs_dim + lane] = o[k];
        for (int j = lane + 32; j < obs_dim; j += 32) o_obs[(size_t)row * obs_dim + j] = obs[(size_t)src[k] * obs_dim + j];
        if (lane < 2) o_act[2 * (size_t)row + lane] = sc[k];
        if (lane == 2) o_logp[row] = sc[k];
        if (lane == 3) o_adv[row] = sc[k];
        if (lane == 4) o_ret[row] = sc[k];
        if (lane == 5) o_val[row] = sc[k];
    }
}

// Gradients of  loss = pg_loss + vf_coef * v_loss  (+ an entropy term that is constant in
// the parameters because log_std is a buffer) with respect to the network outputs mu [n,2]
// and v [n], exactly as autograd derives them from ppo.py:173-204:
//   ratio = exp(logp_new - logp_old); pg = mean(max(-A ratio, -A clamp(ratio, 1-c, 1+c)))
//   v_loss = 0.5 mean(max((v - R)^2, (clamp(v - v_old, -c, c) + v_old - R)^2))
// torch.maximum splits the gradient evenly on ties and clamp passes it on its closed
// interval; both conventions are reproduced.  Also accumulates sum(logp_old - logp_new).
__global__ void ppo_loss_grad_kernel(const float* __restrict__ mu, const float* __restrict__ v,
                                     const float* __restrict__ act, const float* __restrict__ old_logp,
                                     const float* __restrict__ adv_raw, const float* __restrict__ ret,
                                     const float* __restrict__ v_old, const float* __restrict__ log_std,
                                     const float* __restrict__ adv_mean, const float* __restrict__ adv_std, int n,
                                     float clip, float vf_coef, float* __restrict__ dmu, float* __restrict__ dv,
                                     double* __restrict__ kl_sum) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float kl = 0.f;
    if (i < n) {
        float d0, d1;
        ppo_policy_grad(mu[2 * (size_t)i], mu[2 * (size_t)i + 1], act[2 * (size_t)i], act[2 * (size_t)i + 1], old_logp[i],
                        adv_raw[i], adv_mean[0], adv_std[0], log_std[0], log_std[1], clip, n, d0, d1, kl);
        dmu[2 * (size_t)i] = d0;
        dmu[2 * (size_t)i + 1] = d1;
        dv[i] = ppo_value_grad(v[i], ret[i], v_old[i], clip, vf_coef, n);
    }
    // block sum of (logp_old - logp_new) -> one atomic per block
    __shared__ float red[32];
    for (int m = 16; m > 0; m >>= 1) kl += __shfl_xor_sync(0xffffffffu, kl, m);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = kl;
    __syncthreads();
    if (threadIdx.x < 32) {
        float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
        for (int m = 16; m > 0; m >>= 1) t += __shfl_xor_sync(0xffffffffu, t, m);
        if (threadIdx.x == 0) atomicAdd(kl_sum, (double)t);
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

int policy_param_count(int obs_dim) { return policy_packed_floats(obs_dim); }

int launch_gather_minibatch(const int64_t* idx, int n, int obs_dim, const float* obs, const float* act,
                            const float* logp, const float* adv, const float* ret, const float* val, float* o_obs,
                            float* o_act, float* o_logp, float* o_adv, float* o_ret, float* o_val, cudaStream_t stream) {
    const int rows_per_block = 8 * 4;  // 8 warps x 4 rows
    gather_minibatch_kernel<<<(n + rows_per_block - 1) / rows_per_block, 8 * 32, 0, stream>>>(
        idx, n, obs_dim, obs, act, logp, adv, ret, val, o_obs, o_act, o_logp, o_adv, o_ret, o_val);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int launch_ppo_loss_grad(const float* mu, const float* v, const float* act, const float* old_logp, const float* adv,
                         const float* ret, const float* v_old, const float* log_std, const float* adv_mean,
                         const float* adv_std, int n, float clip, float vf_coef, float* dmu, float* dv,
                         double* kl_sum, cudaStream_t stream) {
    ppo_loss_grad_kernel<<<(n + 255) / 256, 256, 0, stream>>>(mu, v, act, old_logp, adv, ret, v_old, log_std, adv_mean,
                                                              adv_std, n, clip, vf_coef, dmu, dv, kl_sum);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int launch_policy_jobs(const PolicyJobs& jobs, int n_jobs, int obs_dim, cudaStream_t stream) {
    int maxB = 0;
    for (int k = 0; k < n_jobs; ++k) {
        const PolicyJob& j = jobs.job[k];
        maxB = j.B > maxB ? j.B : maxB;
        if (j.params != nullptr && (reinterpret_cast<uintptr_t>(j.params) & 15u) != 0) return 2;  // the bulk copy needs a 16-byte aligned source
        if (j.block_policy != nullptr && (j.block_len <= 0 || j.block_len % kPolicyCols != 0 || (j.pool_stride & 3) != 0)) return 2;
    }
    if (maxB <= 0) return 0;
    static const bool use_tc = [] { const char* e = getenv("RK_B200_POLICY_TC"); return !(e && e[0] == '0'); }();
    if (use_tc && obs_dim <= kPXK) {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const size_t smem_tc = kPolicyTcSmemFloats * sizeof(float);
        if (first_use_on_device(0))  // the attribute is per device, not per process
            cudaFuncSetAttribute(policy_act_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tc);
        const int ntiles = (maxB + kUmmaTS - 1) / kUmmaTS, ctas = 2 * sms / n_jobs;   // two CTAs per SM over all jobs
        policy_act_tc_kernel<<<dim3(ntiles < ctas ? ntiles : ctas, n_jobs), kPT, smem_tc, stream>>>(jobs, obs_dim);
        count_launch();
        return cudaGetLastError() == cudaSuccess ? 0 : 1;
    }
    const size_t smem = ((size_t)policy_packed_floats(obs_dim) + (size_t)kHidden * kPolicyCols) * sizeof(float);
    if (first_use_on_device(4))  // the attribute is per device, not per process
        cudaFuncSetAttribute(policy_act_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    policy_act_kernel<<<dim3((maxB + kPolicyCols - 1) / kPolicyCols, n_jobs), kPolicyThreads, smem, stream>>>(jobs, obs_dim);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int launch_policy_act(const float* params, int obs_dim, const float* obs, int64_t obs_stride, int B, uint64_t seed,
                      uint64_t counter, float* action, int64_t act_stride, float* logprob, float* value,
                      float* mean, cudaStream_t stream, const int32_t* block_policy, int block_len,
                      int64_t pool_stride) {
    if (B <= 0) return 0;
    if (params == nullptr) {
        random_act_kernel<<<(B + 255) / 256, 256, 0, stream>>>(B, seed, counter, action, act_stride);
        count_launch();
        return cudaGetLastError() == cudaSuccess ? 0 : 1;
    }
    PolicyJobs jobs{};
    jobs.job[0] = PolicyJob{params, obs, obs_stride, B, seed, counter, action, act_stride, logprob, value, mean,
                            block_policy, block_len, pool_stride};
    return launch_policy_jobs(jobs, 1, obs_dim, stream);
}

}  // namespace rk
