// Fused PPO minibatch gradient: PPO.ppo_update's forward + loss + backward
// (reference agent/ppo.py:170-206) for the Agent's two 64-64 tanh MLPs
// (agent/ppo.py:11-37) as ONE kernel.
//
// One CTA owns one network (blockIdx.y: 0 actor_mu, 1 critic) and walks over
// 128-sample tiles of the minibatch.  Per tile everything stays in shared
// memory: the gathered observation rows X[D][128], the hidden activations
// H1, H2 [64][128] (overwritten in place by their pre-activation gradients),
// the weights (transposed copies for the forward pass).  The three 64-wide
// products per tile (layer 1, layer 2, d h1 = d z2 . W2) run as 8x8 register
// tiles fed by 128-bit shared-memory loads; the weight gradients
// dW2 += H1^T dZ2, dW1 += X^T dZ1, dW3 += H2^T dOut are outer-product
// accumulations whose registers PERSIST across all tiles of the CTA, so each
// CTA writes its partial gradient once; a second small kernel sums the partials
// in a fixed order (deterministic, no atomics) into torch's parameter order.
// Arithmetic is fp32 FFMA throughout (what the reference computes on a GPU).
#include "rk_types.cuh"
#include "rk_ppo_loss.cuh"
#include "rk_umma.cuh"

namespace rk {
namespace {

constexpr int kH = 64;          // hidden width (agent/ppo.py:18-22)
constexpr int kTS = 128;        // samples per tile
constexpr int kNT = 128;        // threads per CTA
constexpr int kLD = kTS + 4;    // activation row stride: rows 4 banks apart, 16-byte aligned
constexpr int kWLD = kH + 4;    // W2^T row stride (column reads in the backward product stay conflict-free)
constexpr int kMaxD = 20;       // observation width supported by the register tile of dW1
constexpr int kNetStride = 5648;  // floats per CTA partial (a multiple of 16)
static_assert(kNetStride >= kH * kMaxD + kH + kH * kH + kH + 2 * kH + 2,
              "one partial = W1 [64][D], b1, W2, b2, W3 [2][64], b3 (5,632 was 2 floats short at D = 20: the actor's db3 of "
              "one CTA overlapped dW1[0][0..1] of the next)");

struct NetPtrs { const float *W1, *b1, *W2, *b2, *W3, *b3; };
struct GradArgs {
    NetPtrs net[2];
    const float* log_std;
    const float *obs, *act, *old_logp, *adv, *ret, *val;
    const int64_t* idx;
    const double* adv_part;  // [kAdvBlocks][2] partial (sum, sum of squares) of the minibatch advantages
    double n_global;
    int n, D, obs_stride;
    float clip, vf_coef;
    float* partial;       // [2][gridDim.x][kNetStride]
    double* kl_partial;   // [gridDim.x]
};

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// Blackwell's packed fp32 FMA (SASS FFMA2): two IEEE fp32 FMAs per issue slot, so the
// 8x8 register tiles below cost 32 instead of 64 instructions per step
typedef float2 u64;  // one packed register pair
__device__ __forceinline__ u64 pack2(float lo, float hi) { return make_float2(lo, hi); }
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) { lo = v.x; hi = v.y; }
__device__ __forceinline__ void fma2(u64& d, u64 a, u64 b) { d = __ffma2_rn(a, b, d); }
__device__ __forceinline__ void pack_tile(const float (&acc)[8][8], u64 (&p)[8][4]) {
#pragma unroll
    for (int m = 0; m < 8; ++m)
#pragma unroll
        for (int q = 0; q < 4; ++q) p[m][q] = pack2(acc[m][2 * q], acc[m][2 * q + 1]);
}
__device__ __forceinline__ void unpack_tile(const u64 (&p)[8][4], float (&acc)[8][8]) {
#pragma unroll
    for (int m = 0; m < 8; ++m)
#pragma unroll
        for (int q = 0; q < 4; ++q) unpack2(p[m][q], acc[m][2 * q], acc[m][2 * q + 1]);
}

// acc[m][n] += sum_k A[k][s_m] * W[k][j_n]; s_m = ty*4 + (m&3) + 64*(m>>2), j_n = tx + 8 n.  The weight
// columns are stored permuted (column of output j = tx + 8n at (n>>2)*32 + tx*4 + (n&3): the 8 threads' 128-bit loads are contiguous) so that the thread's 8 outputs
// are two 128-bit loads, while its activation rows tx + 8n land in distinct banks when stored.
__device__ __forceinline__ void tile_product(const float* __restrict__ A, const float* __restrict__ W, int wld, int K,
                                             int ty, int tx, float (&acc)[8][8]) {
    const float* a = A + ty * 4;
    const float* w = W + tx * 4;
    u64 p[8][4];
    pack_tile(acc, p);
    // fragments of step k+1 are loaded before the FMAs of step k (software pipeline)
    float4 a0 = ld4(a), a1 = ld4(a + 64), w0 = ld4(w), w1 = ld4(w + 32);
#pragma unroll 2
    for (int k = 0; k < K; ++k) {
        const int kn = (k + 1 < K) ? k + 1 : k;
        const float4 na0 = ld4(a + kn * kLD), na1 = ld4(a + kn * kLD + 64);
        const float4 nw0 = ld4(w + kn * wld), nw1 = ld4(w + kn * wld + 32);
        const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const u64 wp[4] = {pack2(w0.x, w0.y), pack2(w0.z, w0.w), pack2(w1.x, w1.y), pack2(w1.z, w1.w)};
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const u64 as = pack2(av[m], av[m]);
#pragma unroll
            for (int q = 0; q < 4; ++q) fma2(p[m][q], as, wp[q]);
        }
        a0 = na0; a1 = na1; w0 = nw0; w1 = nw1;
    }
    unpack_tile(p, acc);
}

template <int OUT>
__device__ __forceinline__ void net_body(const GradArgs& g, const NetPtrs& P, float* smem) {
    constexpr bool kActor = OUT == 2;
    const int tid = threadIdx.x, D = g.D;
    const int tx = tid & 7, ty = tid >> 3;
    // shared-memory carve-up
    float* W1t = smem;                   // [D][64]      W1t[i][j] = W1[j][i]
    float* W2t = W1t + kMaxD * kH;       // [64][kWLD]   W2t[i][j] = W2[j][i]
    float* sb1 = W2t + kH * kWLD;        // [64]
    float* sb2 = sb1 + kH;               // [64]
    float* sW3 = sb2 + kH;               // [OUT][64]
    float* X = sW3 + 2 * kH;             // [D][kLD]
    float* H1 = X + kMaxD * kLD;         // [64][kLD]  h1, then dz1
    float* H2 = H1 + kH * kLD;           // [64][kLD]  h2, then dz2
    float* DO = H2 + kH * kLD;           // [2][kLD]   d loss / d (pre-activation output)
    float* red = DO + 2 * kLD;           // [16]

    for (int q = tid; q < kH * D; q += kNT) { const int j = q / D, i = q - j * D; W1t[i * kH + ((j >> 5) * 32 + (j & 7) * 4 + ((j >> 3) & 3))] = P.W1[q]; }
    for (int q = tid; q < kH * kH; q += kNT) { const int j = q >> 6, i = q & 63; W2t[i * kWLD + ((j >> 5) * 32 + (j & 7) * 4 + ((j >> 3) & 3))] = P.W2[q]; }
    for (int q = tid; q < kH; q += kNT) { sb1[q] = P.b1[q]; sb2[q] = P.b2[q]; }
    for (int q = tid; q < OUT * kH; q += kNT) sW3[q] = P.W3[q];
    float b3[OUT];
#pragma unroll
    for (int o = 0; o < OUT; ++o) b3[o] = P.b3[o];

    // minibatch advantage statistics (ppo.py:187, unbiased std), summed in a fixed order
    float adv_mean = 0.f, adv_std = 1.f;
    if (kActor) {
        double s1 = 0.0, s2 = 0.0;
        for (int b = 0; b < kAdvBlocks; ++b) { s1 += g.adv_part[2 * b]; s2 += g.adv_part[2 * b + 1]; }
        const double mean = s1 / g.n_global;
        const double var = (s2 - g.n_global * mean * mean) / (g.n_global - 1.0);
        adv_mean = (float)mean;
        adv_std = (float)sqrt(var > 0.0 ? var : 0.0);
    }
    const float ls0 = kActor ? g.log_std[0] : 0.f, ls1 = kActor ? g.log_std[1] : 0.f;

    // accumulators that live across all tiles of this CTA
    const int kg = tid >> 6, u = tid & 63, ti = u >> 3, tj = u & 7;
    float gW2[8][8], gW1[kMaxD], gb2[8], gb1 = 0.f, gW3 = 0.f, gb3[OUT], kl = 0.f;
#pragma unroll
    for (int m = 0; m < 8; ++m) {
        gb2[m] = 0.f;
#pragma unroll
        for (int n = 0; n < 8; ++n) gW2[m][n] = 0.f;
    }
#pragma unroll
    for (int i = 0; i < kMaxD; ++i) gW1[i] = 0.f;
#pragma unroll
    for (int o = 0; o < OUT; ++o) gb3[o] = 0.f;
    __syncthreads();

    const int ntiles = (g.n + kTS - 1) / kTS;
    const bool vec_rows = (g.obs_stride & 3) == 0 && g.obs_stride >= ((D + 3) & ~3) &&
                          (reinterpret_cast<uintptr_t>(g.obs) & 15u) == 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        // ---- gather the tile: thread <-> sample ----
        const int gi = tile * kTS + tid;
        const bool valid = gi < g.n;
        const int64_t row = valid ? (g.idx ? g.idx[gi] : (int64_t)gi) : 0;
        {
            const float* src = g.obs + row * g.obs_stride;
            if (vec_rows) {   // rows padded to 16 bytes: ceil(D/4) 128-bit loads (the padding lands in unused X rows)
                for (int i = 0; i < D; i += 4) {
                    const float4 v = valid ? __ldg(reinterpret_cast<const float4*>(src + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    X[i * kLD + tid] = v.x; X[(i + 1) * kLD + tid] = v.y; X[(i + 2) * kLD + tid] = v.z; X[(i + 3) * kLD + tid] = v.w;
                }
            } else {
                for (int i = 0; i < D; ++i) X[i * kLD + tid] = valid ? src[i] : 0.f;
            }
        }
        float d_a0 = 0.f, d_a1 = 0.f, d_lp = 0.f, d_adv = 0.f, d_ret = 0.f, d_val = 0.f;
        if (valid) {
            if (kActor) {
                const float2 a = *reinterpret_cast<const float2*>(g.act + 2 * row);
                d_a0 = a.x; d_a1 = a.y; d_lp = g.old_logp[row]; d_adv = g.adv[row];
            } else {
                d_ret = g.ret[row]; d_val = g.val[row];
            }
        }
        __syncthreads();
        float acc[8][8];
        // ---- layer 1: H1 = tanh(X W1^T + b1) ----
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            const float b = sb1[tx + 8 * n];
#pragma unroll
            for (int m = 0; m < 8; ++m) acc[m][n] = b;
        }
        tile_product(X, W1t, kH, D, ty, tx, acc);
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            float* h = H1 + (tx + 8 * n) * kLD + ty * 4;
            st4(h, make_float4(tanh_fast(acc[0][n]), tanh_fast(acc[1][n]), tanh_fast(acc[2][n]), tanh_fast(acc[3][n])));
            st4(h + 64, make_float4(tanh_fast(acc[4][n]), tanh_fast(acc[5][n]), tanh_fast(acc[6][n]), tanh_fast(acc[7][n])));
        }
        __syncthreads();
        // ---- layer 2: H2 = tanh(H1 W2^T + b2) ----
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            const float b = sb2[tx + 8 * n];
#pragma unroll
            for (int m = 0; m < 8; ++m) acc[m][n] = b;
        }
        tile_product(H1, W2t, kWLD, kH, ty, tx, acc);
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            float* h = H2 + (tx + 8 * n) * kLD + ty * 4;
            st4(h, make_float4(tanh_fast(acc[0][n]), tanh_fast(acc[1][n]), tanh_fast(acc[2][n]), tanh_fast(acc[3][n])));
            st4(h + 64, make_float4(tanh_fast(acc[4][n]), tanh_fast(acc[5][n]), tanh_fast(acc[6][n]), tanh_fast(acc[7][n])));
        }
        __syncthreads();
        // ---- output layer + loss gradient: thread <-> sample ----
        {
            float out[OUT];
#pragma unroll
            for (int o = 0; o < OUT; ++o) out[o] = b3[o];
#pragma unroll 4
            for (int j = 0; j < kH; j += 4) {
                const float h0 = H2[j * kLD + tid], h1 = H2[(j + 1) * kLD + tid], h2 = H2[(j + 2) * kLD + tid],
                            h3 = H2[(j + 3) * kLD + tid];
#pragma unroll
                for (int o = 0; o < OUT; ++o) {
                    const float4 w = ld4(sW3 + o * kH + j);
                    out[o] = fmaf(h0, w.x, out[o]); out[o] = fmaf(h1, w.y, out[o]);
                    out[o] = fmaf(h2, w.z, out[o]); out[o] = fmaf(h3, w.w, out[o]);
                }
            }
            float dpre[OUT];
            if (kActor) {
                const float mu0 = tanh_fast(out[0]), mu1 = tanh_fast(out[1]);  // actor_mu ends in nn.Tanh (ppo.py:19)
                float dmu0 = 0.f, dmu1 = 0.f, klv = 0.f;
                if (valid)
                    ppo_policy_grad(mu0, mu1, d_a0, d_a1, d_lp, d_adv, adv_mean, adv_std, ls0, ls1, g.clip, g.n, dmu0,
                                    dmu1, klv);
                kl += klv;
                dpre[0] = dmu0 * (1.f - mu0 * mu0);
                dpre[OUT - 1] = dmu1 * (1.f - mu1 * mu1);
            } else {
                dpre[0] = valid ? ppo_value_grad(out[0], d_ret, d_val, g.clip, g.vf_coef, g.n) : 0.f;
            }
#pragma unroll
            for (int o = 0; o < OUT; ++o) { DO[o * kLD + tid] = dpre[o]; gb3[o] += dpre[o]; }
        }
        __syncthreads();
        // ---- dW3 += H2^T dOut (thread <-> (o, j) or (sample half, j)) ----
        {
            const int o = kActor ? kg : 0;
            const int s0 = kActor ? 0 : kg * 64, s1 = kActor ? kTS : s0 + 64;
            const float* h = H2 + u * kLD;
            const float* d = DO + o * kLD;
            float a = 0.f;
#pragma unroll 4
            for (int s = s0; s < s1; s += 4) {
                const float4 hv = ld4(h + s), dv = ld4(d + s);
                a = fmaf(hv.x, dv.x, a); a = fmaf(hv.y, dv.y, a); a = fmaf(hv.z, dv.z, a); a = fmaf(hv.w, dv.w, a);
            }
            gW3 += a;
        }
        __syncthreads();
        // ---- dZ2 = (dOut W3) * (1 - H2^2), in place ----
        {
            float4 d0[OUT], d1[OUT];
#pragma unroll
            for (int o = 0; o < OUT; ++o) { d0[o] = ld4(DO + o * kLD + ty * 4); d1[o] = ld4(DO + o * kLD + 64 + ty * 4); }
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                const int j = tx + 8 * n;
                float* h = H2 + j * kLD + ty * 4;
                const float4 ha = ld4(h), hb = ld4(h + 64);
                float4 za = make_float4(0.f, 0.f, 0.f, 0.f), zb = za;
#pragma unroll
                for (int o = 0; o < OUT; ++o) {
                    const float w = sW3[o * kH + j];
                    za.x = fmaf(d0[o].x, w, za.x); za.y = fmaf(d0[o].y, w, za.y); za.z = fmaf(d0[o].z, w, za.z); za.w = fmaf(d0[o].w, w, za.w);
                    zb.x = fmaf(d1[o].x, w, zb.x); zb.y = fmaf(d1[o].y, w, zb.y); zb.z = fmaf(d1[o].z, w, zb.z); zb.w = fmaf(d1[o].w, w, zb.w);
                }
                za.x *= 1.f - ha.x * ha.x; za.y *= 1.f - ha.y * ha.y; za.z *= 1.f - ha.z * ha.z; za.w *= 1.f - ha.w * ha.w;
                zb.x *= 1.f - hb.x * hb.x; zb.y *= 1.f - hb.y * hb.y; zb.z *= 1.f - hb.z * hb.z; zb.w *= 1.f - hb.w * hb.w;
                st4(h, za); st4(h + 64, zb);
            }
        }
        __syncthreads();
        // ---- dW2 += H1^T dZ2, db2 += sum dZ2: thread <-> 8 inputs x 8 outputs over its half of the samples ----
        {
            const float* ha = H1 + ti * kLD + kg * 64;
            const float* zb = H2 + tj * kLD + kg * 64;
#pragma unroll 1
            for (int s = 0; s < 64; s += 4) {
                float4 a[8];
#pragma unroll
                for (int m = 0; m < 8; ++m) a[m] = ld4(ha + 8 * m * kLD + s);
                float4 b = ld4(zb + s);
#pragma unroll
                for (int n = 0; n < 8; ++n) {
                    const float4 nb = ld4(zb + 8 * ((n + 1) & 7) * kLD + s);  // next output row, one step ahead
                    gb2[n] += (b.x + b.y) + (b.z + b.w);
#pragma unroll
                    for (int m = 0; m < 8; ++m) {
                        float t = gW2[m][n];
                        t = fmaf(a[m].x, b.x, t); t = fmaf(a[m].y, b.y, t);
                        t = fmaf(a[m].z, b.z, t); t = fmaf(a[m].w, b.w, t);
                        gW2[m][n] = t;
                    }
                    b = nb;
                }
            }
        }
        // ---- dH1 = dZ2 W2 (outputs i = tx + 8 n), then dZ1 = dH1 * (1 - H1^2) in place ----
#pragma unroll
        for (int m = 0; m < 8; ++m)
#pragma unroll
            for (int n = 0; n < 8; ++n) acc[m][n] = 0.f;
        {
            const float* a = H2 + ty * 4;
            const float* w = W2t + tx * kWLD;
            u64 p[8][4];
            pack_tile(acc, p);
            float4 a0 = ld4(a), a1 = ld4(a + 64);
            float wv[8];
#pragma unroll
            for (int n = 0; n < 8; ++n) wv[n] = w[8 * n * kWLD];
#pragma unroll 2
            for (int k = 0; k < kH; ++k) {
                const int kn = (k + 1 < kH) ? k + 1 : k;
                const float4 na0 = ld4(a + kn * kLD), na1 = ld4(a + kn * kLD + 64);
                float nw[8];
#pragma unroll
                for (int n = 0; n < 8; ++n) nw[n] = w[8 * n * kWLD + ((kn >> 5) * 32 + (kn & 7) * 4 + ((kn >> 3) & 3))];  // column of W2^T row k (permuted)
                const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                const u64 wp[4] = {pack2(wv[0], wv[1]), pack2(wv[2], wv[3]), pack2(wv[4], wv[5]), pack2(wv[6], wv[7])};
#pragma unroll
                for (int m = 0; m < 8; ++m) {
                    const u64 as = pack2(av[m], av[m]);
#pragma unroll
                    for (int q = 0; q < 4; ++q) fma2(p[m][q], as, wp[q]);
                }
                a0 = na0; a1 = na1;
#pragma unroll
                for (int n = 0; n < 8; ++n) wv[n] = nw[n];
            }
            unpack_tile(p, acc);
        }
        __syncthreads();  // every thread has finished reading H1 (dW2) before it is overwritten
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            float* h = H1 + (tx + 8 * n) * kLD + ty * 4;
            const float4 ha = ld4(h), hb = ld4(h + 64);
            st4(h, make_float4(acc[0][n] * (1.f - ha.x * ha.x), acc[1][n] * (1.f - ha.y * ha.y),
                               acc[2][n] * (1.f - ha.z * ha.z), acc[3][n] * (1.f - ha.w * ha.w)));
            st4(h + 64, make_float4(acc[4][n] * (1.f - hb.x * hb.x), acc[5][n] * (1.f - hb.y * hb.y),
                                    acc[6][n] * (1.f - hb.z * hb.z), acc[7][n] * (1.f - hb.w * hb.w)));
        }
        __syncthreads();
        // ---- dW1 += X^T dZ1, db1 += sum dZ1: thread <-> output j over its half of the samples ----
        {
            const float* z = H1 + u * kLD + kg * 64;
            const float* x = X + kg * 64;
#pragma unroll 2
            for (int s = 0; s < 64; s += 4) {
                const float4 zv = ld4(z + s);
                gb1 += (zv.x + zv.y) + (zv.z + zv.w);
#pragma unroll
                for (int i = 0; i < kMaxD; ++i) {
                    if (i < D) {
                        const float4 xv = ld4(x + i * kLD + s);
                        float t = gW1[i];
                        t = fmaf(xv.x, zv.x, t); t = fmaf(xv.y, zv.y, t); t = fmaf(xv.z, zv.z, t); t = fmaf(xv.w, zv.w, t);
                        gW1[i] = t;
                    }
                }
            }
        }
        __syncthreads();  // X, H1, H2 are rewritten by the next tile
    }

    // ---- combine the two sample halves and write this CTA's partial gradient ----
    // layout = torch's parameter order of one Sequential: W1 [64][D], b1, W2 [64][64], b2, W3 [OUT][64], b3
    float* stage = H1;  // >= kNetStride floats (H1 and H2 are contiguous)
    const int oW1 = 0, ob1 = kH * D, oW2 = ob1 + kH, ob2 = oW2 + kH * kH, oW3 = ob2 + kH, ob3 = oW3 + OUT * kH;
    for (int pass = 1; pass >= 0; --pass) {
        if (kg == pass) {
            const bool add = pass == 0;
#pragma unroll
            for (int m = 0; m < 8; ++m)
#pragma unroll
                for (int n = 0; n < 8; ++n) {
                    float* d = stage + oW2 + (tj + 8 * n) * kH + (ti + 8 * m);
                    *d = add ? *d + gW2[m][n] : gW2[m][n];
                }
            if (ti == 0) {
#pragma unroll
                for (int n = 0; n < 8; ++n) {
                    float* d = stage + ob2 + tj + 8 * n;
                    *d = add ? *d + gb2[n] : gb2[n];
                }
            }
#pragma unroll
            for (int i = 0; i < kMaxD; ++i)
                if (i < D) {
                    float* d = stage + oW1 + u * D + i;
                    *d = add ? *d + gW1[i] : gW1[i];
                }
            {
                float* d = stage + ob1 + u;
                *d = add ? *d + gb1 : gb1;
            }
            if (kActor) {
                stage[oW3 + kg * kH + u] = gW3;          // kg is the output index for the actor
            } else {
                float* d = stage + oW3 + u;
                *d = add ? *d + gW3 : gW3;
            }
        }
        __syncthreads();
    }
    // db3 and the KL sum: fixed-order block reduction
    float r[OUT + 1];
#pragma unroll
    for (int o = 0; o < OUT; ++o) r[o] = gb3[o];
    r[OUT] = kl;
#pragma unroll
    for (int o = 0; o <= OUT; ++o) {
        for (int m = 16; m > 0; m >>= 1) r[o] += __shfl_xor_sync(0xffffffffu, r[o], m);
        if ((tid & 31) == 0) red[o * 4 + (tid >> 5)] = r[o];
    }
    __syncthreads();
    if (tid < OUT) stage[ob3 + tid] = (red[tid * 4] + red[tid * 4 + 1]) + (red[tid * 4 + 2] + red[tid * 4 + 3]);
    if (kActor && tid == 0)
        g.kl_partial[blockIdx.x] = (double)((red[OUT * 4] + red[OUT * 4 + 1]) + (red[OUT * 4 + 2] + red[OUT * 4 + 3]));
    __syncthreads();
    float* dst = g.partial + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * kNetStride;
    const int np = ob3 + OUT;
    for (int q = tid; q < np; q += kNT) dst[q] = stage[q];
}

__global__ void __launch_bounds__(kNT, 2) ppo_mlp_grad_kernel(const GradArgs g) {
    extern __shared__ __align__(16) float train_smem[];
    if (blockIdx.y == 0) net_body<2>(g, g.net[0], train_smem);
    else net_body<1>(g, g.net[1], train_smem);
}

// ---------------------------------------------------------------------------
// Tensor-core variant (opt-in): the three per-sample products of a tile -- layer 1, layer 2 and
// dH1 = dZ2 . W2 -- run on the 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM)
// as a chain THROUGH tensor memory: thread t owns sample t = TMEM lane t; it reads its accumulator row
// with tcgen05.ld, applies bias / tanh / the loss gradient, and writes the row back with tcgen05.st as
// the A operand of the next product (A comes from TMEM, B = the weights in shared memory in the K-major
// core-matrix layout).  fp32 accuracy is kept with the 3-pass split x = hi + lo (both exactly
// representable in TF32): A.B ~ Ah.Bh + Ah.Bl + Al.Bh, accumulated in fp32.  The weight-gradient
// accumulations stay on the CUDA cores (their operands are needed feature-major) and read
// the same plain [feature][sample] tiles as the FFMA kernel, now written by the thread-per-sample epilogues.
// Building blocks verified stand-alone in tools/umma_selftest.cu.
// ---------------------------------------------------------------------------
constexpr int kTmemCols = 512;
constexpr int kColAcc0 = 0, kColAcc1 = 64, kColXh = 128, kColXl = 160, kColHh = 192, kColHl = 256, kColZh = 320, kColZl = 384;
constexpr int kXK = 24;  // layer-1 reduction length padded to a multiple of the MMA K (8)

constexpr int kNT2 = 256;   // threads of the tensor-core variant: 8 warps; warps w and w + 4 share TMEM lane quarter w & 3

template <int OUT>
__device__ __forceinline__ void net_body_tc(const GradArgs& g, const NetPtrs& P, float* smem) {
    constexpr bool kActor = OUT == 2;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, D = g.D;
    const int half = warp >> 2;                    // which 32 of the 64 accumulator columns this thread handles
    const int srow = (warp & 3) * 32 + lane;       // sample of the tile = TMEM lane
    const int cbase = half * 32;
    // shared-memory carve-up: operand tiles of the tensor-core products first (16-byte aligned core matrices)
    float* W1h = smem;                      // [64][kXK]  B of layer 1 (K-major, rows = output j)
    float* W1l = W1h + kH * kXK;
    float* W2h = W1l + kH * kXK;            // [64][64]   B of layer 2: rows = output j, k = input i
    float* W2l = W2h + kH * kH;
    float* W2th = W2l + kH * kH;            // [64][64]   B of dH1: rows = input i, k = output j
    float* W2tl = W2th + kH * kH;
    float* sb1 = W2tl + kH * kH;            // [64]
    float* sb2 = sb1 + kH;                  // [64]
    float* sW3 = sb2 + kH;                  // [OUT][64]
    float* X = sW3 + 2 * kH;                // 2 x [kMaxD][kLD]  plain feature-major tiles of the CUDA-core phases (X double-buffered)
    float* H1 = X + 2 * kMaxD * kLD;        // [64][kLD]  h1, then dz1
    float* H2 = H1 + kH * kLD;              // [64][kLD]  h2, then dz2
    float* DO = H2 + kH * kLD;              // [2][kLD]   d loss / d (pre-activation output)
    float* OP = DO + 2 * kLD;               // [2 halves][2][kLD]  partial output-layer sums of the two column halves
    float* red = OP + 4 * kLD;              // [32]
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint32_t tmem_slot;

    for (int q = tid; q < kH * kXK; q += kNT2) {
        const int j = q / kXK, i = q - j * kXK;
        uint32_t hi = 0, lo = 0;
        if (i < D) split_tf32(P.W1[j * D + i], hi, lo);
        W1h[umma_off(j, i, kXK)] = __uint_as_float(hi);
        W1l[umma_off(j, i, kXK)] = __uint_as_float(lo);
    }
    for (int q = tid; q < kH * kH; q += kNT2) {
        const int j = q >> 6, i = q & 63;
        uint32_t hi, lo;
        split_tf32(P.W2[q], hi, lo);
        W2h[umma_off(j, i, kH)] = __uint_as_float(hi); W2l[umma_off(j, i, kH)] = __uint_as_float(lo);
        W2th[umma_off(i, j, kH)] = __uint_as_float(hi); W2tl[umma_off(i, j, kH)] = __uint_as_float(lo);
    }
    for (int q = tid; q < kH; q += kNT2) { sb1[q] = P.b1[q]; sb2[q] = P.b2[q]; }
    for (int q = tid; q < OUT * kH; q += kNT2) sW3[q] = P.W3[q];
    for (int q = tid; q < 2 * kMaxD * kLD; q += kNT2) X[q] = 0.f;
    float b3[OUT];
#pragma unroll
    for (int o = 0; o < OUT; ++o) b3[o] = P.b3[o];
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the weight tiles are read by the tensor core (async proxy)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);   // this warp's quarter of the TMEM lanes
    unsigned phase = 0;

    float adv_mean = 0.f, adv_std = 1.f;
    if (kActor) {
        double s1 = 0.0, s2 = 0.0;
        for (int b = 0; b < kAdvBlocks; ++b) { s1 += g.adv_part[2 * b]; s2 += g.adv_part[2 * b + 1]; }
        const double mean = s1 / g.n_global;
        const double var = (s2 - g.n_global * mean * mean) / (g.n_global - 1.0);
        adv_mean = (float)mean;
        adv_std = (float)sqrt(var > 0.0 ? var : 0.0);
    }
    const float ls0 = kActor ? g.log_std[0] : 0.f, ls1 = kActor ? g.log_std[1] : 0.f;

    // roles in the CUDA-core weight-gradient phases
    const int kg = tid >> 7;                              // dW2: half of the tile's samples
    const int ti = (tid & 127) >> 3, tj = tid & 7;        // dW2: inputs i = ti + 16 m (m < 4), outputs j = tj + 8 n (n < 8)
    const int u = tid & 63, grp = tid >> 6;               // dW1 / dW3: output j = u, sample quarter (or output x sample half)
    float2 gW2[4][8];   // FFMA2 lanes = partial sums over even / odd samples (added at the end)
    float2 gW1[kMaxD];
    float gb2[8], gb1 = 0.f, gW3 = 0.f, gb3[OUT], kl = 0.f;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
        gb2[n] = 0.f;
#pragma unroll
        for (int m = 0; m < 4; ++m) gW2[m][n] = make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < kMaxD; ++i) gW1[i] = make_float2(0.f, 0.f);
#pragma unroll
    for (int o = 0; o < OUT; ++o) gb3[o] = 0.f;

    const int ntiles = (g.n + kTS - 1) / kTS;
    // One tile's observation rows: two threads per sample (columns 0-15 / 16-23 of the padded row); the row goes
    // to a plain X tile (dW1) and, split into hi + lo, to the TMEM columns that are the A operand of layer 1.
    auto gather = [&](int tile, float* Xbuf, bool& valid, float& a0, float& a1, float& lp, float& adv, float& ret,
                      float& val) {
        const int gi = tile * kTS + srow;
        valid = gi < g.n;
        const int64_t row = valid ? (g.idx ? g.idx[gi] : (int64_t)gi) : 0;
        const float* src = g.obs + row * g.obs_stride;
        const int c_lo = half ? 16 : 0, c_hi = half ? kXK : 16;
        for (int c0 = c_lo; c0 < c_hi; c0 += 8) {
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int i = c0 + c;
                const float v = (valid && i < D) ? src[i] : 0.f;
                if (i < D) Xbuf[i * kLD + srow] = v;
                split_tf32(v, hi[c], lo[c]);
            }
            tmem_st8(lane_base + kColXh + c0, hi);
            tmem_st8(lane_base + kColXl + c0, lo);
        }
        a0 = a1 = lp = adv = ret = val = 0.f;
        if (valid && half == 0) {
            if (kActor) {
                const float2 a = *reinterpret_cast<const float2*>(g.act + 2 * row);
                a0 = a.x; a1 = a.y; lp = g.old_logp[row]; adv = g.adv[row];
            } else {
                ret = g.ret[row]; val = g.val[row];
            }
        }
    };
    // dW1 += X^T dZ1, db1 += sum dZ1 of one tile: thread <-> (sample quarter, output j)
    auto dw1 = [&](const float* Xbuf) {
        const float* z = H1 + u * kLD + grp * 32;
        const float* x = Xbuf + grp * 32;
#pragma unroll 2
        for (int s = 0; s < 32; s += 4) {
            const float4 zv = ld4(z + s);
            gb1 += (zv.x + zv.y) + (zv.z + zv.w);
#pragma unroll
            for (int i = 0; i < kMaxD; ++i) {
                if (i < D) {
                    const float4 xv = ld4(x + i * kLD + s);
                    float2 t = gW1[i];
                    t = __ffma2_rn(make_float2(xv.x, xv.y), make_float2(zv.x, zv.y), t);
                    t = __ffma2_rn(make_float2(xv.z, xv.w), make_float2(zv.z, zv.w), t);
                    gW1[i] = t;
                }
            }
        }
    };
    // Software pipeline across tiles: the tensor core never waits for the CUDA cores to have nothing to do --
    // dW1 of tile t-1 runs while layer 1 of tile t is in flight, the gather of tile t+1 while layer 2 is,
    // dW2 while dH1 is.
    bool valid = false, n_valid = false;
    float d_a0 = 0.f, d_a1 = 0.f, d_lp = 0.f, d_adv = 0.f, d_ret = 0.f, d_val = 0.f;
    float n_a0 = 0.f, n_a1 = 0.f, n_lp = 0.f, n_adv = 0.f, n_ret = 0.f, n_val = 0.f;
    int buf = 0;
    bool have_prev = false;
    if ((int)blockIdx.x < ntiles) gather(blockIdx.x, X, valid, d_a0, d_a1, d_lp, d_adv, d_ret, d_val);
    tmem_publish_and_sync();
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        float* Xcur = X + buf * kMaxD * kLD;
        float* Xoth = X + (buf ^ 1) * kMaxD * kLD;
        // ---- layer 1 on the tensor core: ACC0 = X W1^T; meanwhile dW1 of the previous tile ----
        if (warp == 0 && elect_one()) issue_product(tmem, kColAcc0, kColXh, kColXl, W1h, W1l, kXK, &bar);
        if (have_prev) dw1(Xoth);
        __syncthreads();   // dW1 has finished reading H1 (dZ1 of the previous tile) before the epilogue overwrites it
        wait_product(&bar, phase);
        {
            uint32_t v[32];
            tmem_ld32(lane_base + kColAcc0 + cbase, v);
#pragma unroll
            for (int c8 = 0; c8 < 32; c8 += 8) {
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const int j = cbase + c8 + c;
                    const float h = tanh_fast(__uint_as_float(v[c8 + c]) + sb1[j]);
                    H1[j * kLD + srow] = h;
                    split_tf32(h, hi[c], lo[c]);
                }
                tmem_st8(lane_base + kColHh + cbase + c8, hi);
                tmem_st8(lane_base + kColHl + cbase + c8, lo);
            }
        }
        tmem_publish_and_sync();
        // ---- layer 2: ACC1 = H1 W2^T; meanwhile the next tile's rows are gathered (layer 1 has released the X columns) ----
        if (warp == 0 && elect_one()) issue_product(tmem, kColAcc1, kColHh, kColHl, W2h, W2l, kH, &bar);
        if (tile + (int)gridDim.x < ntiles) gather(tile + gridDim.x, Xoth, n_valid, n_a0, n_a1, n_lp, n_adv, n_ret, n_val);
        wait_product(&bar, phase);
        {
            float out[OUT];
#pragma unroll
            for (int o = 0; o < OUT; ++o) out[o] = 0.f;
            uint32_t v[32];
            tmem_ld32(lane_base + kColAcc1 + cbase, v);
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                const int j = cbase + c;
                const float h = tanh_fast(__uint_as_float(v[c]) + sb2[j]);
                H2[j * kLD + srow] = h;
#pragma unroll
                for (int o = 0; o < OUT; ++o) out[o] = fmaf(h, sW3[o * kH + j], out[o]);
            }
#pragma unroll
            for (int o = 0; o < OUT; ++o) OP[(half * 2 + o) * kLD + srow] = out[o];
        }
        __syncthreads();
        // ---- output layer + loss gradient: one thread per sample (column half 0) ----
        if (half == 0) {
            float out[OUT], dpre[OUT];
#pragma unroll
            for (int o = 0; o < OUT; ++o) out[o] = b3[o] + (OP[o * kLD + srow] + OP[(2 + o) * kLD + srow]);
            if (kActor) {
                const float mu0 = tanh_fast(out[0]), mu1 = tanh_fast(out[OUT - 1]);  // actor_mu ends in nn.Tanh (ppo.py:19)
                float dmu0 = 0.f, dmu1 = 0.f, klv = 0.f;
                if (valid)
                    ppo_policy_grad(mu0, mu1, d_a0, d_a1, d_lp, d_adv, adv_mean, adv_std, ls0, ls1, g.clip, g.n, dmu0,
                                    dmu1, klv);
                kl += klv;
                dpre[0] = dmu0 * (1.f - mu0 * mu0);
                dpre[OUT - 1] = dmu1 * (1.f - mu1 * mu1);
            } else {
                dpre[0] = valid ? ppo_value_grad(out[0], d_ret, d_val, g.clip, g.vf_coef, g.n) : 0.f;
            }
#pragma unroll
            for (int o = 0; o < OUT; ++o) { DO[o * kLD + srow] = dpre[o]; gb3[o] += dpre[o]; }
        }
        __syncthreads();
        // ---- dW3 += H2^T dOut: thread <-> (output o, sample half, j) [actor] or (sample quarter, j) [critic] ----
        {
            const int o = kActor ? (grp & 1) : 0;
            const int s0 = kActor ? (grp >> 1) * 64 : grp * 32, s1 = s0 + (kActor ? 64 : 32);
            const float* h = H2 + u * kLD;
            const float* d = DO + o * kLD;
            float a = 0.f;
#pragma unroll 4
            for (int s = s0; s < s1; s += 4) {
                const float4 hv = ld4(h + s), dv = ld4(d + s);
                a = fmaf(hv.x, dv.x, a); a = fmaf(hv.y, dv.y, a); a = fmaf(hv.z, dv.z, a); a = fmaf(hv.w, dv.w, a);
            }
            gW3 += a;
        }
        __syncthreads();
        // ---- dZ2 = (dOut W3) * (1 - H2^2): plain tile in place + A operand of the dH1 product ----
        {
            float dpre[OUT];
#pragma unroll
            for (int o = 0; o < OUT; ++o) dpre[o] = DO[o * kLD + srow];
#pragma unroll
            for (int c8 = 0; c8 < 32; c8 += 8) {
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const int j = cbase + c8 + c;
                    const float h = H2[j * kLD + srow];
                    float z = 0.f;
#pragma unroll
                    for (int o = 0; o < OUT; ++o) z = fmaf(dpre[o], sW3[o * kH + j], z);
                    z *= 1.f - h * h;
                    H2[j * kLD + srow] = z;
                    split_tf32(z, hi[c], lo[c]);
                }
                tmem_st8(lane_base + kColZh + cbase + c8, hi);
                tmem_st8(lane_base + kColZl + cbase + c8, lo);
            }
        }
        tmem_publish_and_sync();
        // ---- dH1 = dZ2 W2 on the tensor core (ACC0) while the CUDA cores accumulate dW2 += H1^T dZ2 ----
        if (warp == 0 && elect_one()) issue_product(tmem, kColAcc0, kColZh, kColZl, W2th, W2tl, kH, &bar);
        {
            const float* ha = H1 + ti * kLD + kg * 64;
            const float* zb = H2 + tj * kLD + kg * 64;
#pragma unroll 1
            for (int s = 0; s < 64; s += 4) {
                float4 a[4];
#pragma unroll
                for (int m = 0; m < 4; ++m) a[m] = ld4(ha + 16 * m * kLD + s);
                float4 b = ld4(zb + s);
#pragma unroll
                for (int n = 0; n < 8; ++n) {
                    const float4 nb = ld4(zb + 8 * ((n + 1) & 7) * kLD + s);
                    gb2[n] += (b.x + b.y) + (b.z + b.w);
#pragma unroll
                    for (int m = 0; m < 4; ++m) {
                        float2 t = gW2[m][n];
                        t = __ffma2_rn(make_float2(a[m].x, a[m].y), make_float2(b.x, b.y), t);
                        t = __ffma2_rn(make_float2(a[m].z, a[m].w), make_float2(b.z, b.w), t);
                        gW2[m][n] = t;
                    }
                    b = nb;
                }
            }
        }
        wait_product(&bar, phase);
        __syncthreads();  // every thread has finished reading H1 (dW2) before it is overwritten
        // ---- dZ1 = dH1 * (1 - H1^2), in place ----
        {
            uint32_t v[32];
            tmem_ld32(lane_base + kColAcc0 + cbase, v);
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                float* h = H1 + (cbase + c) * kLD + srow;
                const float hv = *h;
                *h = __uint_as_float(v[c]) * (1.f - hv * hv);
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        // dW1 of this tile is deferred to the next iteration (or to the epilogue below)
        have_prev = true;
        buf ^= 1;
        valid = n_valid; d_a0 = n_a0; d_a1 = n_a1; d_lp = n_lp; d_adv = n_adv; d_ret = n_ret; d_val = n_val;
    }
    if (have_prev) dw1(X + (buf ^ 1) * kMaxD * kLD);
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols));

    // ---- combine the thread groups in a fixed order and write this CTA's partial gradient ----
    // layout = torch's parameter order of one Sequential: W1 [64][D], b1, W2 [64][64], b2, W3 [OUT][64], b3
    float* stage = H1;
    const int oW1 = 0, ob1 = kH * D, oW2 = ob1 + kH, ob2 = oW2 + kH * kH, oW3 = ob2 + kH, ob3 = oW3 + OUT * kH;
    for (int q = tid; q < kNetStride; q += kNT2) stage[q] = 0.f;
    __syncthreads();
    for (int pass = 0; pass < 4; ++pass) {
        if (kg == pass) {   // dW2, db2: two sample halves
#pragma unroll
            for (int m = 0; m < 4; ++m)
#pragma unroll
                for (int n = 0; n < 8; ++n) stage[oW2 + (tj + 8 * n) * kH + (ti + 16 * m)] += gW2[m][n].x + gW2[m][n].y;
            if (ti == 0) {
#pragma unroll
                for (int n = 0; n < 8; ++n) stage[ob2 + tj + 8 * n] += gb2[n];
            }
        }
        if (grp == pass) {  // dW1, db1, dW3: four groups
#pragma unroll
            for (int i = 0; i < kMaxD; ++i)
                if (i < D) stage[oW1 + u * D + i] += gW1[i].x + gW1[i].y;
            stage[ob1 + u] += gb1;
            stage[oW3 + (kActor ? (grp & 1) : 0) * kH + u] += gW3;
        }
        __syncthreads();
    }
    // db3 and the KL sum: fixed-order block reduction (only column-half-0 threads hold contributions)
    float r[OUT + 1];
#pragma unroll
    for (int o = 0; o < OUT; ++o) r[o] = gb3[o];
    r[OUT] = kl;
#pragma unroll
    for (int o = 0; o <= OUT; ++o) {
        for (int m = 16; m > 0; m >>= 1) r[o] += __shfl_xor_sync(0xffffffffu, r[o], m);
        if (lane == 0) red[o * 8 + warp] = r[o];
    }
    __syncthreads();
    if (tid < OUT) stage[ob3 + tid] = (red[tid * 8] + red[tid * 8 + 1]) + (red[tid * 8 + 2] + red[tid * 8 + 3]);
    if (kActor && tid == 0)
        g.kl_partial[blockIdx.x] = (double)((red[OUT * 8] + red[OUT * 8 + 1]) + (red[OUT * 8 + 2] + red[OUT * 8 + 3]));
    __syncthreads();
    float* dst = g.partial + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * kNetStride;
    const int np = ob3 + OUT;
    for (int q = tid; q < np; q += kNT2) dst[q] = stage[q];
}

__global__ void __launch_bounds__(kNT2, 1) ppo_mlp_grad_tc_kernel(const GradArgs g) {
    extern __shared__ __align__(1024) float train_smem[];
    if (blockIdx.y == 0) net_body_tc<2>(g, g.net[0], train_smem);
    else net_body_tc<1>(g, g.net[1], train_smem);
}

constexpr size_t kTrainTcSmemFloats = 2 * (size_t)kH * kXK + 4 * (size_t)kH * kH + 2 * kH + 2 * kH + 2 * (size_t)kMaxD * kLD +
                                      2 * (size_t)kH * kLD + 2 * kLD + 4 * kLD + 32;
static_assert(kTrainTcSmemFloats * sizeof(float) <= 226 * 1024, "one CTA per SM");
static_assert((kH * kXK) % 4 == 0, "16-byte aligned carve-up");

// ---------------------------------------------------------------------------
// Tensor-core variant 2 (rk_ppo_grad_io.tensor_cores = 2): the WEIGHT GRADIENTS dW2 = dZ2^T H1 (+ db2) and
// dW1 = dZ1^T X (+ db1) run on tcgen05 as well -- every 64-wide product of the update is then a tensor-core product
// and the CUDA cores keep the epilogues (bias, tanh, loss gradient, the TF32 hi/lo splits) and dW3.
//   * Both operands of a weight-gradient product are needed FEATURE-major (K = the tile's 128 samples), so they can
//     not come from tensor memory (a TMEM lane is a sample).  The thread-per-sample epilogues write them straight
//     into shared-memory operand tiles in the K-major no-swizzle core-matrix layout, with the k-cores of a row
//     group 144 bytes apart instead of 128: the 32 threads of a warp (32 consecutive samples) then store one feature
//     row into 32 distinct banks (tools/umma_dw_selftest.cu verifies descriptors with LBO = 144, SBO = 4608).
//     The plain [feature][sample] tiles of variant 1 are gone.
//   * The hi and lo parts of an operand are STACKED along M or N so that one series of 16 instructions (K = 8 samples
//     each) computes several terms of the split at once; the accumulators persist in TMEM across all tiles of the CTA
//     and are read (and their parts added) once at the end.  The bias gradients ride along as one more column: H1's
//     operand tile is preceded by a constant row of ones, X's tile holds a one in column D.
//   * Three 36 KB operand buffers, adjacent in the order [ones | P | R | Q], are time-multiplexed within a tile (227 KB
//     of shared memory hold the 76 KB of weight tiles, these three and X^T hi/lo, nothing more):
//       P: h1.hi (epilogue 1)                         -> dz1.hi (after the products that read h1.hi)
//       R: dz2.lo -> h1.lo (fetched back from the TMEM columns that fed layer 2) -> dz1.lo
//       Q: h2 (epilogue 2) -> dz2.hi (in place)
//     dW2|db2:  [R; Q] . [ones | P]^T   M = 128, N = 72  (dz2.lo.h1.hi and dz2.hi.h1.hi in lanes 0-63 / 64-127; issued
//               together with dH1 and completing under the dZ1 arithmetic), then Q . R'^T  M = 64, N = 64 (dz2.hi.h1.lo)
//               once h1.lo is in R;
//     dW1|db1:  [P'; R''] . [Xh; Xl]^T  M = 128, N = 48  (all four terms of the split), issued at the end of the tile and
//               completing under the next tile's layer 1.
//   * The next tile's observation rows are loaded into registers while layer 2 runs and stored (TMEM A operand of
//     layer 1 + X^T tile of dW1) when the previous users of those buffers have completed.
// ---------------------------------------------------------------------------
constexpr int kLboF = 36, kSboF = 32 * kLboF;   // floats: 144 B between the k-cores of a row group, 4608 B between 8-row groups
__device__ __forceinline__ int poff(int r, int k) { return (r >> 3) * kSboF + (k >> 2) * kLboF + (r & 7) * 4 + (k & 3); }
// TMEM columns: the X operand of layer 1 (written at the end of the previous tile) shares the columns of dZ2 (written
// after layer 1 has completed, consumed by dH1 before the next X arrives)
constexpr int kC2Acc = 0, kC2dW2 = 64, kC2dW2c = 136, kC2dW1 = 200, kC2Hh = 248, kC2Hl = 312, kC2Zh = 376, kC2Zl = 440,
              kC2Xh = kC2Zh, kC2Xl = kC2Zl;
static_assert(kC2Zl + kH <= kTmemCols, "TMEM columns");

// D[M x N] (+)= A[M x 128] . B[N x 128]^T, both operands padded K-major tiles in shared memory (M = 64: row r in TMEM
// lane 32 * (r / 16) + r % 16; M = 128: row r in lane r)
__device__ __forceinline__ void issue_ss(uint32_t tmem_d, const float* A, const float* B, int M, int N, uint32_t& acc) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    uint64_t da = umma_desc(smem_u32(A), 4 * kLboF, 4 * kSboF), db = umma_desc(smem_u32(B), 4 * kLboF, 4 * kSboF);
#pragma unroll
    for (int kb = 0; kb < kTS / 8; ++kb) {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                     ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
        acc = 1;
        da += (2 * 4 * kLboF) >> 4;   // next 8 samples: two k-cores further (the start-address field counts 16-byte units)
        db += (2 * 4 * kLboF) >> 4;
    }
}
// smem operand tiles written by this thread (generic proxy) become visible to the tensor core (async proxy)
__device__ __forceinline__ void publish_operands_and_sync() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tmem_publish_and_sync();
}

template <int OUT>
__device__ __forceinline__ void net_body_tc2(const GradArgs& g, const NetPtrs& P, float* smem) {
    constexpr bool kActor = OUT == 2;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, D = g.D;
    const int half = warp >> 2;                    // which 32 of the 64 accumulator columns this thread handles
    const int srow = (warp & 3) * 32 + lane;       // sample of the tile = TMEM lane
    const int cbase = half * 32;
    float* W1h = smem;                      // [64][kXK]  B of layer 1 (K-major, rows = output j)
    float* W1l = W1h + kH * kXK;
    float* W2h = W1l + kH * kXK;            // [64][64]   B of layer 2: rows = output j, k = input i
    float* W2l = W2h + kH * kH;
    float* W2th = W2l + kH * kH;            // [64][64]   B of dH1: rows = input i, k = output j
    float* W2tl = W2th + kH * kH;
    float* sb1 = W2tl + kH * kH;            // [64]
    float* sb2 = sb1 + kH;                  // [64]
    float* sW3 = sb2 + kH;                  // [OUT][64]
    float* Ob = sW3 + 2 * kH;               // [8 rows][128 samples]   row 0 = ones: with P the B operand of dW2|db2
    float* Pb = Ob + kSboF;                 // [64]  h1.hi, later dz1.hi
    float* Rb = Pb + 8 * kSboF;             // [64]  dz2.lo, h1.lo, dz1.lo    ([P; R] = A of dW1, [R; Q] = A of dW2)
    float* Qb = Rb + 8 * kSboF;             // [64]  h2, later dz2.hi
    float* XTh = Qb + 8 * kSboF;            // [24]  X^T hi (row D = ones)   ([Xh; Xl] = B of dW1)
    float* XTl = XTh + 3 * kSboF;
    float* DO = XTl + 3 * kSboF;            // [2][kLD]   d loss / d (pre-activation output)
    float* OP = DO + 2 * kLD;               // [2 halves][2][kLD]  partial output-layer sums of the two column halves
    float* red = OP + 4 * kLD;              // [32]
    __shared__ __align__(8) unsigned long long bar, bar2;
    __shared__ uint32_t tmem_slot;

    for (int q = tid; q < kH * kXK; q += kNT2) {
        const int j = q / kXK, i = q - j * kXK;
        uint32_t hi = 0, lo = 0;
        if (i < D) split_tf32(P.W1[j * D + i], hi, lo);
        W1h[umma_off(j, i, kXK)] = __uint_as_float(hi);
        W1l[umma_off(j, i, kXK)] = __uint_as_float(lo);
    }
    for (int q = tid; q < kH * kH; q += kNT2) {
        const int j = q >> 6, i = q & 63;
        uint32_t hi, lo;
        split_tf32(P.W2[q], hi, lo);
        W2h[umma_off(j, i, kH)] = __uint_as_float(hi); W2l[umma_off(j, i, kH)] = __uint_as_float(lo);
        W2th[umma_off(i, j, kH)] = __uint_as_float(hi); W2tl[umma_off(i, j, kH)] = __uint_as_float(lo);
    }
    for (int q = tid; q < kH; q += kNT2) { sb1[q] = P.b1[q]; sb2[q] = P.b2[q]; }
    for (int q = tid; q < OUT * kH; q += kNT2) sW3[q] = P.W3[q];
    for (int q = tid; q < 8 * kTS; q += kNT2) Ob[poff(q >> 7, q & 127)] = (q >> 7) == 0 ? 1.f : 0.f;
    float b3[OUT];
#pragma unroll
    for (int o = 0; o < OUT; ++o) b3[o] = P.b3[o];
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the weight tiles are read by the tensor core (async proxy)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);   // this warp's quarter of the TMEM lanes
    unsigned phase = 0, phase2 = 0;

    float adv_mean = 0.f, adv_std = 1.f;
    if (kActor) {
        double s1 = 0.0, s2 = 0.0;
        for (int b = 0; b < kAdvBlocks; ++b) { s1 += g.adv_part[2 * b]; s2 += g.adv_part[2 * b + 1]; }
        const double mean = s1 / g.n_global;
        const double var = (s2 - g.n_global * mean * mean) / (g.n_global - 1.0);
        adv_mean = (float)mean;
        adv_std = (float)sqrt(var > 0.0 ? var : 0.0);
    }
    const float ls0 = kActor ? g.log_std[0] : 0.f, ls1 = kActor ? g.log_std[1] : 0.f;

    const int u = tid & 63, grp = tid >> 6;               // dW3: output j = u, (output o x sample half) or sample quarter
    float gW3 = 0.f, gb3[OUT], kl = 0.f;
#pragma unroll
    for (int o = 0; o < OUT; ++o) gb3[o] = 0.f;

    const int ntiles = (g.n + kTS - 1) / kTS;
    // One tile's observation rows: two threads per sample (columns 0-15 / 16-23 of the padded row), loaded into registers
    // (the row index of a tile is fetched one tile earlier than its row: two dependent global loads would otherwise
    //  sit in the critical path of every tile)
    //  (the load is unconditional from a clamped address and nothing depends on it until the row is gathered)
    auto row_of = [&](int tile) -> int64_t {
        const int gi = tile * kTS + srow;
        const int gc = (tile < ntiles && gi < g.n) ? gi : 0;
        return g.idx ? g.idx[gc] : (int64_t)gc;
    };
    auto gather_load = [&](int tile, int64_t row, float (&x)[16], bool& valid, float& a0, float& a1, float& lp, float& adv,
                           float& ret, float& val) {
        valid = tile * kTS + srow < g.n;
        const float* src = g.obs + row * g.obs_stride;
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            const int i = half * 16 + c;
            x[c] = (valid && i < D) ? src[i] : 0.f;
        }
        a0 = a1 = lp = adv = ret = val = 0.f;
        if (valid && half == 0) {
            if (kActor) {
                const float2 a = *reinterpret_cast<const float2*>(g.act + 2 * row);
                a0 = a.x; a1 = a.y; lp = g.old_logp[row]; adv = g.adv[row];
            } else {
                ret = g.ret[row]; val = g.val[row];
            }
        }
    };
    // ... split into hi + lo: the A operand of layer 1 (TMEM columns) ...
    auto store_x_tmem = [&](const float (&x)[16]) {
#pragma unroll
        for (int c8 = 0; c8 < 16; c8 += 8) {
            if (c8 == 0 || half == 0) {
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) split_tf32_fast(x[c8 + c], hi[c], lo[c]);
                tmem_st8(lane_base + kC2Xh + half * 16 + c8, hi);
                tmem_st8(lane_base + kC2Xl + half * 16 + c8, lo);
            }
        }
    };
    // ... and the B operand of dW1 (X^T tile; column D is the constant one that yields db1)
    auto store_x_smem = [&](const float (&x)[16]) {
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            if (c < 8 || half == 0) {
                const int i = half * 16 + c;
                uint32_t hi, lo;
                split_tf32_fast(x[c], hi, lo);
                if (i == D) hi = 0x3f800000u;
                XTh[poff(i, srow)] = __uint_as_float(hi);
                XTl[poff(i, srow)] = __uint_as_float(lo);
            }
        }
    };

    bool valid = false, n_valid = false;
    float d_a0 = 0.f, d_a1 = 0.f, d_lp = 0.f, d_adv = 0.f, d_ret = 0.f, d_val = 0.f;
    float n_a0 = 0.f, n_a1 = 0.f, n_lp = 0.f, n_adv = 0.f, n_ret = 0.f, n_val = 0.f;
    float xr[16], xn[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) { xr[c] = 0.f; xn[c] = 0.f; }
    const bool any_tile = (int)blockIdx.x < ntiles;
    int64_t next_row = row_of(blockIdx.x + gridDim.x);
    if (any_tile) {
        gather_load(blockIdx.x, row_of(blockIdx.x), xr, valid, d_a0, d_a1, d_lp, d_adv, d_ret, d_val);
        store_x_tmem(xr);
    }
    tmem_publish_and_sync();
    uint32_t acc2 = 0, acc2c = 0, acc1 = 0;   // (thread 0) the gradient accumulators are overwritten by their first instruction only
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const bool more = tile + (int)gridDim.x < ntiles;
        // ---- layer 1 on the tensor core: ACC = X W1^T (its completion also covers dW1 of the previous tile) ----
        if (warp == 0 && elect_one()) issue_product(tmem, kC2Acc, kC2Xh, kC2Xl, W1h, W1l, kXK, &bar);
        wait_product(&bar, phase);
        {
            uint32_t v[32];
            tmem_ld32(lane_base + kC2Acc + cbase, v);
#pragma unroll
            for (int c8 = 0; c8 < 32; c8 += 8) {
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const int j = cbase + c8 + c;
                    const float h = tanh_fast(__uint_as_float(v[c8 + c]) + sb1[j]);
                    split_tf32_fast(h, hi[c], lo[c]);
                    Pb[poff(j, srow)] = __uint_as_float(hi[c]);
                }
                tmem_st8(lane_base + kC2Hh + cbase + c8, hi);
                tmem_st8(lane_base + kC2Hl + cbase + c8, lo);
            }
        }
        tmem_publish_and_sync();
        // ---- layer 2: ACC = H1 W2^T; meanwhile the next tile's rows are on their way into registers ----
        if (warp == 0 && elect_one()) issue_product(tmem, kC2Acc, kC2Hh, kC2Hl, W2h, W2l, kH, &bar);
        if (more) gather_load(tile + gridDim.x, next_row, xn, n_valid, n_a0, n_a1, n_lp, n_adv, n_ret, n_val);
        next_row = row_of(tile + 2 * gridDim.x);
        wait_product(&bar, phase);
        {
            float out[OUT];
#pragma unroll
            for (int o = 0; o < OUT; ++o) out[o] = 0.f;
            uint32_t v[32];
            tmem_ld32(lane_base + kC2Acc + cbase, v);
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                const int j = cbase + c;
                const float h = tanh_fast(__uint_as_float(v[c]) + sb2[j]);
                Qb[poff(j, srow)] = h;
#pragma unroll
                for (int o = 0; o < OUT; ++o) out[o] = fmaf(h, sW3[o * kH + j], out[o]);
            }
#pragma unroll
            for (int o = 0; o < OUT; ++o) OP[(half * 2 + o) * kLD + srow] = out[o];
        }
        __syncthreads();
        // ---- output layer + loss gradient: one thread per sample (column half 0) ----
        if (half == 0) {
            float out[OUT], dpre[OUT];
#pragma unroll
            for (int o = 0; o < OUT; ++o) out[o] = b3[o] + (OP[o * kLD + srow] + OP[(2 + o) * kLD + srow]);
            if (kActor) {
                const float mu0 = tanh_fast(out[0]), mu1 = tanh_fast(out[OUT - 1]);  // actor_mu ends in nn.Tanh (ppo.py:19)
                float dmu0 = 0.f, dmu1 = 0.f, klv = 0.f;
                if (valid)
                    ppo_policy_grad(mu0, mu1, d_a0, d_a1, d_lp, d_adv, adv_mean, adv_std, ls0, ls1, g.clip, g.n, dmu0,
                                    dmu1, klv);
                kl += klv;
                dpre[0] = dmu0 * (1.f - mu0 * mu0);
                dpre[OUT - 1] = dmu1 * (1.f - mu1 * mu1);
            } else {
                dpre[0] = valid ? ppo_value_grad(out[0], d_ret, d_val, g.clip, g.vf_coef, g.n) : 0.f;
            }
#pragma unroll
            for (int o = 0; o < OUT; ++o) { DO[o * kLD + srow] = dpre[o]; gb3[o] += dpre[o]; }
        }
        __syncthreads();
        // ---- dW3 += H2^T dOut on the CUDA cores (2 x 64 outputs): rows of the h2 tile are read along the samples ----
        {
            const int o = kActor ? (grp & 1) : 0;
            const int s0 = kActor ? (grp >> 1) * 64 : grp * 32, s1 = s0 + (kActor ? 64 : 32);
            const float* h = Qb + (u >> 3) * kSboF + (u & 7) * 4;
            const float* d = DO + o * kLD;
            float a = 0.f;
#pragma unroll 4
            for (int s = s0; s < s1; s += 4) {
                const float4 hv = ld4(h + (s >> 2) * kLboF), dv = ld4(d + s);
                a = fmaf(hv.x, dv.x, a); a = fmaf(hv.y, dv.y, a); a = fmaf(hv.z, dv.z, a); a = fmaf(hv.w, dv.w, a);
            }
            gW3 += a;
        }
        __syncthreads();
        // ---- dZ2 = (dOut W3) * (1 - H2^2): hi in place (A of dW2), lo next to it, both as the A operand of dH1 ----
        {
            float dpre[OUT];
#pragma unroll
            for (int o = 0; o < OUT; ++o) dpre[o] = DO[o * kLD + srow];
#pragma unroll
            for (int c8 = 0; c8 < 32; c8 += 8) {
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const int j = cbase + c8 + c;
                    const int q = poff(j, srow);
                    const float h = Qb[q];
                    float z = 0.f;
#pragma unroll
                    for (int o = 0; o < OUT; ++o) z = fmaf(dpre[o], sW3[o * kH + j], z);
                    z *= 1.f - h * h;
                    split_tf32_fast(z, hi[c], lo[c]);
                    Qb[q] = __uint_as_float(hi[c]);
                    Rb[q] = __uint_as_float(lo[c]);
                }
                tmem_st8(lane_base + kC2Zh + cbase + c8, hi);
                tmem_st8(lane_base + kC2Zl + cbase + c8, lo);
            }
        }
        publish_operands_and_sync();
        // ---- dH1 = dZ2 W2 (ACC, own barrier) and dW2|db2 += [dz2.lo; dz2.hi]^T [1 | h1.hi] ----
        if (warp == 0 && elect_one()) {
            issue_product(tmem, kC2Acc, kC2Zh, kC2Zl, W2th, W2tl, kH, &bar);
            issue_ss(tmem + kC2dW2, Rb, Ob, 128, kH + 8, acc2);
            mma_commit(&bar2);
        }
        store_x_smem(xr);   // (X^T tile: free since layer 1's completion covered dW1 of the previous tile; needed by dW1 below)
        wait_product(&bar, phase);
        // ---- dZ1 = dH1 * (1 - H1^2) while the weight-gradient product runs; h1 = hi + lo comes back from the TMEM
        //      columns that fed layer 2 ----
        float zhi[32], zlo[32], h1lo[32];
        {
            uint32_t d[32], hh[32], hl[32];
            tmem_ld32(lane_base + kC2Acc + cbase, d);
            tmem_ld32(lane_base + kC2Hh + cbase, hh);
            tmem_ld32(lane_base + kC2Hl + cbase, hl);
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                const float h = __uint_as_float(hh[c]) + __uint_as_float(hl[c]);
                const float z = __uint_as_float(d[c]) * (1.f - h * h);
                uint32_t zh, zl;
                split_tf32_fast(z, zh, zl);
                zhi[c] = __uint_as_float(zh); zlo[c] = __uint_as_float(zl); h1lo[c] = __uint_as_float(hl[c]);
            }
        }
        wait_product(&bar2, phase2);   // h1.hi (P) and dz2.lo (R) have been consumed
#pragma unroll
        for (int c = 0; c < 32; ++c) {
            const int q = poff(cbase + c, srow);
            Rb[q] = h1lo[c];   // B of the last term of dW2
            Pb[q] = zhi[c];    // A of dW1
        }
        publish_operands_and_sync();
        if (warp == 0 && elect_one()) {
            issue_ss(tmem + kC2dW2c, Qb, Rb, 64, kH, acc2c);   // dW2 += dz2.hi^T h1.lo
            mma_commit(&bar);
        }
        if (more) store_x_tmem(xn);   // dH1 has released the dZ2 columns the X operand shares
        wait_product(&bar, phase);
#pragma unroll
        for (int c = 0; c < 32; ++c) Rb[poff(cbase + c, srow)] = zlo[c];
        publish_operands_and_sync();
        // ---- dW1|db1 += [dz1.hi; dz1.lo]^T [Xh | 1 | Xl]: completes under the next tile's layer 1 ----
        if (warp == 0 && elect_one()) issue_ss(tmem + kC2dW1, Pb, XTh, 128, 2 * kXK, acc1);
        valid = n_valid; d_a0 = n_a0; d_a1 = n_a1; d_lp = n_lp; d_adv = n_adv; d_ret = n_ret; d_val = n_val;
#pragma unroll
        for (int c = 0; c < 16; ++c) xr[c] = xn[c];
    }
    if (any_tile) {
        if (warp == 0 && elect_one()) mma_commit(&bar);
        wait_product(&bar, phase);
    }

    // ---- this CTA's partial gradient, in torch's parameter order of one Sequential: W1 [64][D], b1, W2 [64][64], b2, W3, b3
    float* stage = Qb;
    const int oW1 = 0, ob1 = kH * D, oW2 = ob1 + kH, ob2 = oW2 + kH * kH, oW3 = ob2 + kH, ob3 = oW3 + OUT * kH;
    for (int q = tid; q < kNetStride; q += kNT2) stage[q] = 0.f;
    __syncthreads();
    // The M = 128 accumulators hold the lo-part terms in lanes 0-63 and the hi-part terms in lanes 64-127 (gradient row
    // j = lane & 63); the M = 64 accumulator holds row j in lane 32 * (j / 16) + j % 16.  Three rounds of += through
    // shared memory in a fixed order; the two warps of a lane quarter split the columns.
    for (int round = 0; round < 3; ++round) {
        if (any_tile && round < 2 && ((warp & 3) >> 1) == round) {
            const int j = (warp & 1) * 32 + lane;
            {
                uint32_t v[32];
                tmem_ld32(lane_base + kC2dW2 + 8 + cbase, v);
#pragma unroll
                for (int c = 0; c < 32; ++c) stage[oW2 + j * kH + cbase + c] += __uint_as_float(v[c]);
            }
            if (half == 0) {
                uint32_t b[8];
                tmem_ld8(lane_base + kC2dW2, b);
                stage[ob2 + j] += __uint_as_float(b[0]);
            } else {
#pragma unroll
                for (int c8 = 0; c8 < 2 * kXK; c8 += 8) {
                    uint32_t w[8];
                    tmem_ld8(lane_base + kC2dW1 + c8, w);
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const int i = (c8 + c) % kXK;   // columns 24.. are the X.lo terms of the same element
                        if (i < D) stage[oW1 + j * D + i] += __uint_as_float(w[c]);
                        else if (i == D) stage[ob1 + j] += __uint_as_float(w[c]);
                    }
                }
            }
        }
        if (any_tile && round == 2) {
            const int j = (warp & 3) * 16 + lane;
            uint32_t v[32];
            tmem_ld32(lane_base + kC2dW2c + cbase, v);
            if (lane < 16) {
#pragma unroll
                for (int c = 0; c < 32; ++c) stage[oW2 + j * kH + cbase + c] += __uint_as_float(v[c]);
            }
        }
        __syncthreads();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols));
    for (int pass = 0; pass < 4; ++pass) {
        if (grp == pass) stage[oW3 + (kActor ? (grp & 1) : 0) * kH + u] += gW3;   // dW3: four thread groups, fixed order
        __syncthreads();
    }
    // db3 and the KL sum: fixed-order block reduction (only column-half-0 threads hold contributions)
    float r[OUT + 1];
#pragma unroll
    for (int o = 0; o < OUT; ++o) r[o] = gb3[o];
    r[OUT] = kl;
#pragma unroll
    for (int o = 0; o <= OUT; ++o) {
        for (int m = 16; m > 0; m >>= 1) r[o] += __shfl_xor_sync(0xffffffffu, r[o], m);
        if (lane == 0) red[o * 8 + warp] = r[o];
    }
    __syncthreads();
    if (tid < OUT) stage[ob3 + tid] = (red[tid * 8] + red[tid * 8 + 1]) + (red[tid * 8 + 2] + red[tid * 8 + 3]);
    if (kActor && tid == 0)
        g.kl_partial[blockIdx.x] = (double)((red[OUT * 8] + red[OUT * 8 + 1]) + (red[OUT * 8 + 2] + red[OUT * 8 + 3]));
    __syncthreads();
    float* dst = g.partial + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * kNetStride;
    const int np = ob3 + OUT;
    for (int q = tid; q < np; q += kNT2) dst[q] = stage[q];
}

__global__ void __launch_bounds__(kNT2, 1) ppo_mlp_grad_tc2_kernel(const GradArgs g) {
    extern __shared__ __align__(1024) float train_smem[];
    if (blockIdx.y == 0) net_body_tc2<2>(g, g.net[0], train_smem);
    else net_body_tc2<1>(g, g.net[1], train_smem);
}

constexpr size_t kTrainTc2SmemFloats = 2 * (size_t)kH * kXK + 4 * (size_t)kH * kH + 2 * kH + 2 * kH + (9 + 8 + 8 + 3 + 3) * (size_t)kSboF +
                                       2 * kLD + 4 * kLD + 32;
static_assert(kTrainTc2SmemFloats * sizeof(float) <= 227 * 1024, "one CTA per SM");
static_assert(8 * (size_t)kSboF >= (size_t)kNetStride, "the staging area must hold one partial gradient");

// flat_grad[q] = sum over CTAs of partial[net][cta][q'] in a fixed order; *kl_sum = sum kl_partial.
// Block = 64 gradient elements x 4 groups of CTAs (each thread sums its quarter with 4 independent
// chains, the quarters are combined through shared memory in a fixed order).
__global__ void __launch_bounds__(256) ppo_grad_reduce_kernel(const float* __restrict__ partial,
                                                              const double* __restrict__ kl_partial, int ncta, int n_actor,
                                                              int n_total, float* __restrict__ flat_grad,
                                                              double* __restrict__ kl_sum, float* __restrict__ kl_sum_f32) {
    __shared__ float part[4][64];
    const int lane64 = threadIdx.x & 63, grp = threadIdx.x >> 6;
    const int q = blockIdx.x * 64 + lane64;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    if (q < n_total) {
        const int net = q >= n_actor, local = net ? q - n_actor : q;
        const float* src = partial + (size_t)net * ncta * kNetStride + local;
        const int per = (ncta + 3) >> 2, c0 = grp * per, c1 = min(ncta, c0 + per);
        int c = c0;
        for (; c + 4 <= c1; c += 4) {
            s0 += src[(size_t)c * kNetStride]; s1 += src[(size_t)(c + 1) * kNetStride];
            s2 += src[(size_t)(c + 2) * kNetStride]; s3 += src[(size_t)(c + 3) * kNetStride];
        }
        for (; c < c1; ++c) s0 += src[(size_t)c * kNetStride];
    }
    part[grp][lane64] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (grp == 0 && q < n_total) flat_grad[q] = (part[0][lane64] + part[1][lane64]) + (part[2][lane64] + part[3][lane64]);
    if (blockIdx.x == 0 && threadIdx.x < 32) {
        double sk = 0.0;
        for (int c = threadIdx.x; c < ncta; c += 32) sk += kl_partial[c];
        for (int m = 16; m > 0; m >>= 1) sk += __shfl_xor_sync(0xffffffffu, sk, m);
        if (threadIdx.x == 0) {
            *kl_sum = sk;
            if (kl_sum_f32) *kl_sum_f32 = (float)sk;
        }
    }
}

// partial (sum, sum of squares) of adv[idx[k]] in float64: kAdvBlocks blocks, fixed order inside each
__global__ void __launch_bounds__(512) adv_stats_kernel(const int64_t* __restrict__ idx, const float* __restrict__ adv,
                                                        int n, double* __restrict__ part) {
    double s1 = 0.0, s2 = 0.0;
    const int stride = gridDim.x * blockDim.x;
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    for (; k + 3 * stride < n; k += 4 * stride) {   // four independent gathers in flight
        float a[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) a[r] = adv[idx ? idx[k + r * stride] : (int64_t)(k + r * stride)];
#pragma unroll
        for (int r = 0; r < 4; ++r) { s1 += (double)a[r]; s2 += (double)a[r] * (double)a[r]; }
    }
    for (; k < n; k += stride) {
        const double a = (double)adv[idx ? idx[k] : (int64_t)k];
        s1 += a; s2 += a * a;
    }
    __shared__ double r1[16], r2[16];
    for (int m = 16; m > 0; m >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, m); s2 += __shfl_xor_sync(0xffffffffu, s2, m); }
    if ((threadIdx.x & 31) == 0) { r1[threadIdx.x >> 5] = s1; r2[threadIdx.x >> 5] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += r1[w]; b += r2[w]; }
        part[2 * blockIdx.x] = a; part[2 * blockIdx.x + 1] = b;
    }
}

// A seeded pseudo-random permutation of 0..n-1 without sorting: a 6-round balanced Feistel network over
// the next even power of two with cycle walking (out-of-range images are mapped again until they fall
// below n).  Replaces np.random.shuffle(b_inds) (agent/ppo.py:168) / torch.randperm, which sorts.
__device__ __forceinline__ uint32_t feistel_round(uint32_t x, uint32_t key) {
    x = (x ^ key) * 0x9E3779B1u;
    x ^= x >> 15; x *= 0x85EBCA77u; x ^= x >> 13; x *= 0xC2B2AE3Du; x ^= x >> 16;
    return x;
}
__global__ void permutation_kernel(uint64_t seed, uint64_t counter, int64_t n, int half_bits, int64_t* __restrict__ out) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    uint32_t keys[4] = {(uint32_t)counter, (uint32_t)(counter >> 32), 0x7065726du, 0u};
    philox4x32_10(keys, (uint32_t)seed, (uint32_t)(seed >> 32));
    const uint32_t mask = (1u << half_bits) - 1u;
    uint64_t x = (uint64_t)k;
    do {
        uint32_t l = (uint32_t)(x >> half_bits) & mask, r = (uint32_t)x & mask;
#pragma unroll
        for (int round = 0; round < 6; ++round) {
            const uint32_t t = l ^ (feistel_round(r, keys[round & 3] + 0x9E3779B9u * (uint32_t)round) & mask);
            l = r; r = t;
        }
        x = ((uint64_t)l << half_bits) | r;
    } while ((int64_t)x >= n);
    out[k] = (int64_t)x;
}

// Gradient clipping + Adam + the KL early stop of one minibatch step (agent/ppo.py:178-182,205-207) as one
// single-CTA kernel over the ~11k parameters, operating in place on the torch optimizer's own state
// tensors (exp_avg, exp_avg_sq, step), so that optimizer.state_dict() stays the reference's.
//   approx_kl = kl_sum / n_global > kl_target  ->  the step is NOT applied and `state[0]` latches: every
//   later call is a no-op until the host clears it (the reference breaks out of all epochs).
// Otherwise: g /= world; g *= min(1, max_norm / (||g|| + 1e-6)) (nn.utils.clip_grad_norm_), then torch's
// Adam update (no weight decay, no amsgrad) with the learning rate read from its device tensor.
struct AdamArgs {
    float* p[12];
    float* m[12];
    float* v[12];
    float* step[12];
    int off[13];
    const float* grad;
    const float* lr;
    float beta1, beta2, eps, max_norm, kl_target;
    int world;
    const double* kl_sum;
    const float* kl_sum_f32;  // if set, used instead (it travelled through the gradient's all-reduce)
    double n_global;
    int* state;        // [0] stopped, [1] optimizer steps applied since the host cleared it, [2] scratch counter
    float* kl_at_stop;
};

__global__ void __launch_bounds__(256) clip_adam_kernel(const AdamArgs a) {
    __shared__ float red[8];
    __shared__ float s_clip;
    const int tid = threadIdx.x;
    if (a.state[0]) return;
    const double approx_kl = (a.kl_sum_f32 ? (double)*a.kl_sum_f32 : *a.kl_sum) / a.n_global;
    if (approx_kl > (double)a.kl_target) {
        if (blockIdx.x == 0 && tid == 0) { *a.kl_at_stop = (float)approx_kl; }
        // state[0] is latched by the LAST block to pass here, so that no block of this launch reads the flag set
        if (tid == 0) {
            __threadfence();
            if (atomicAdd(&a.state[2], 1) == (int)gridDim.x - 1) { a.state[2] = 0; a.state[0] = 1; }
        }
        return;
    }
    // every block computes the global gradient norm itself, in the same fixed order (deterministic)
    const float inv_world = 1.f / (float)a.world;
    const int n = a.off[12];
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int e = tid;
    for (; e + 768 < n; e += 1024) {
        const float g0 = a.grad[e] * inv_world, g1 = a.grad[e + 256] * inv_world, g2 = a.grad[e + 512] * inv_world,
                    g3 = a.grad[e + 768] * inv_world;
        s0 = fmaf(g0, g0, s0); s1 = fmaf(g1, g1, s1); s2 = fmaf(g2, g2, s2); s3 = fmaf(g3, g3, s3);
    }
    for (; e < n; e += 256) { const float gv = a.grad[e] * inv_world; s0 = fmaf(gv, gv, s0); }
    float ss = (s0 + s1) + (s2 + s3);
    for (int m = 16; m > 0; m >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, m);
    if ((tid & 31) == 0) red[tid >> 5] = ss;
    __syncthreads();
    if (tid == 0) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += red[w];
        s_clip = fminf(a.max_norm / (sqrtf(t) + 1e-6f), 1.f);
    }
    __syncthreads();
    const float scale = s_clip * inv_world;
    const float t_new = a.step[0][0] + 1.f;   // advanced only by the last block to finish (below)
    const float bc1 = (float)(1.0 - pow((double)a.beta1, (double)t_new));
    const float bc2_sqrt = (float)sqrt(1.0 - pow((double)a.beta2, (double)t_new));
    const float step_size = *a.lr / bc1;
    const int q = blockIdx.x * 256 + tid;
    if (q < n) {
        int k = 0;
#pragma unroll
        for (int j = 1; j < 12; ++j) k += (q >= a.off[j]);
        const int el = q - a.off[k];
        float *pp = a.p[k] + el, *pm = a.m[k] + el, *pv = a.v[k] + el;
        const float gv = a.grad[q] * scale;
        const float mm = *pm + (gv - *pm) * (1.f - a.beta1);           // torch: exp_avg.lerp_(grad, 1 - beta1)
        const float vv = a.beta2 * *pv + (1.f - a.beta2) * gv * gv;
        *pm = mm; *pv = vv;
        *pp -= step_size * mm / (sqrtf(vv) / bc2_sqrt + a.eps);
    }
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(&a.state[2], 1) == (int)gridDim.x - 1) {   // every other block has read step[] and finished
            a.state[2] = 0;
            for (int k = 0; k < 12; ++k) a.step[k][0] = t_new;
            a.state[1] += 1;
        }
    }
}

constexpr size_t kTrainSmemFloats = (size_t)kMaxD * kH + (size_t)kH * kWLD + 2 * kH + 2 * kH + (size_t)kMaxD * kLD +
                                    2 * (size_t)kH * kLD + 2 * kLD + 16;
static_assert(2 * (size_t)kH * kLD >= (size_t)kNetStride, "the staging area must hold one partial gradient");
static_assert((kMaxD * kH) % 4 == 0 && (kH * kWLD) % 4 == 0 && (kMaxD * kLD) % 4 == 0, "16-byte aligned carve-up");

int train_grid(int n) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int ntiles = (n + kTS - 1) / kTS;
    return ntiles < sms ? (ntiles > 0 ? ntiles : 1) : sms;
}

}  // namespace

size_t ppo_grad_workspace_bytes() {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return 2 * (size_t)sms * kNetStride * sizeof(float) + (size_t)sms * sizeof(double);
}

int launch_adv_stats(const int64_t* idx, const float* adv, int n, double* part, cudaStream_t stream) {
    adv_stats_kernel<<<kAdvBlocks, 512, 0, stream>>>(idx, adv, n, part);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int launch_ppo_minibatch_grad(const PpoGradIO& io, cudaStream_t stream) {
    if (io.obs_dim < 1 || io.obs_dim > kMaxD) return 2;
    if (io.n <= 0) return 2;
    GradArgs g;
    for (int k = 0; k < 2; ++k)
        g.net[k] = NetPtrs{io.params[6 * k], io.params[6 * k + 1], io.params[6 * k + 2], io.params[6 * k + 3],
                           io.params[6 * k + 4], io.params[6 * k + 5]};
    g.log_std = io.log_std;
    g.obs = io.obs; g.act = io.act; g.old_logp = io.old_logp; g.adv = io.adv; g.ret = io.ret; g.val = io.val;
    g.idx = io.idx; g.adv_part = io.adv_part; g.n_global = io.n_global; g.n = io.n; g.D = io.obs_dim;
    g.obs_stride = io.obs_stride;
    g.clip = io.clip; g.vf_coef = io.vf_coef;
    int ncta = train_grid(io.n);
    if (io.tensor_cores) ncta = (ncta + 1) / 2;   // one CTA per SM: half of the SMs per network
    if (io.workspace_bytes < 2 * (size_t)ncta * kNetStride * sizeof(float) + (size_t)ncta * sizeof(double)) return 3;
    g.kl_partial = reinterpret_cast<double*>(io.workspace);
    g.partial = reinterpret_cast<float*>(g.kl_partial + ((ncta + 1) & ~1));
    if ((size_t)((ncta + 1) & ~1) * sizeof(double) + 2 * (size_t)ncta * kNetStride * sizeof(float) > io.workspace_bytes) return 3;
    const size_t smem = kTrainSmemFloats * sizeof(float);
    if (first_use_on_device(1))  // the attribute is per device, not per process
        cudaFuncSetAttribute(ppo_mlp_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (io.tensor_cores == 2) {
        const size_t smem_tc = kTrainTc2SmemFloats * sizeof(float);
        if (first_use_on_device(3))
            cudaFuncSetAttribute(ppo_mlp_grad_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tc);
        ppo_mlp_grad_tc2_kernel<<<dim3(ncta, 2), kNT2, smem_tc, stream>>>(g);
    } else if (io.tensor_cores) {
        const size_t smem_tc = kTrainTcSmemFloats * sizeof(float);
        if (first_use_on_device(2))
            cudaFuncSetAttribute(ppo_mlp_grad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tc);
        ppo_mlp_grad_tc_kernel<<<dim3(ncta, 2), kNT2, smem_tc, stream>>>(g);
    } else {
        ppo_mlp_grad_kernel<<<dim3(ncta, 2), kNT, smem, stream>>>(g);
    }
    count_launch();
    const int n_actor = kH * io.obs_dim + kH + kH * kH + kH + 2 * kH + 2;
    const int n_total = n_actor + kH * io.obs_dim + kH + kH * kH + kH + kH + 1;
    ppo_grad_reduce_kernel<<<(n_total + 63) / 64, 256, 0, stream>>>(g.partial, g.kl_partial, ncta, n_actor, n_total,
                                                                       io.flat_grad, io.kl_sum, io.kl_sum_f32);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int launch_clip_adam(const PpoAdamIO& io, cudaStream_t stream) {
    AdamArgs a;
    int off = 0;
    for (int k = 0; k < 12; ++k) {
        a.p[k] = io.params[k]; a.m[k] = io.exp_avg[k]; a.v[k] = io.exp_avg_sq[k]; a.step[k] = io.step[k];
        a.off[k] = off;
        off += io.numel[k];
    }
    a.off[12] = off;
    a.grad = io.flat_grad; a.lr = io.lr;
    a.beta1 = io.beta1; a.beta2 = io.beta2; a.eps = io.eps; a.max_norm = io.max_norm; a.kl_target = io.kl_target;
    a.world = io.world; a.kl_sum = io.kl_sum; a.kl_sum_f32 = io.kl_sum_f32; a.n_global = io.n_global; a.state = io.state; a.kl_at_stop = io.kl_at_stop;
    clip_adam_kernel<<<(off + 255) / 256, 256, 0, stream>>>(a);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int launch_permutation(uint64_t seed, uint64_t counter, int64_t n, int64_t* out, cudaStream_t stream) {
    if (n <= 0) return 0;
    int bits = 2;
    while (((int64_t)1 << bits) < n) bits += 2;   // even number of bits: two equal Feistel halves
    permutation_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(seed, counter, n, bits / 2, out);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace rk
