// FMA-pipe peak micro-benchmark: the FP32 / FP64 roofline denominators of the
// step kernel are not in MEASURED_PEAKS.json, so bench.py
// measures them live with this kernel: 8 independent dependent-FMA chains per
// thread, every SM saturated, timed with CUDA events.
#include "rk_types.cuh"

namespace rk {
namespace {

template <typename T>
__global__ void __launch_bounds__(256) fma_chain_kernel(T* out, int iters, T a, T b) {
    T x0 = (T)threadIdx.x, x1 = x0 + (T)1, x2 = x0 + (T)2, x3 = x0 + (T)3;
    T x4 = x0 + (T)4, x5 = x0 + (T)5, x6 = x0 + (T)6, x7 = x0 + (T)7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = x0 * a + b; x1 = x1 * a + b; x2 = x2 * a + b; x3 = x3 * a + b;
            x4 = x4 * a + b; x5 = x5 * a + b; x6 = x6 * a + b; x7 = x7 * a + b;
        }
    }
    const T s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == (T)123456789) out[0] = s;  // never true; keeps the chains alive
}

// the same with Blackwell's packed fp32 FMA (fma.rn.f32x2 -> SASS FFMA2): 8 independent
// 64-bit chains per thread, two FMAs per instruction
__global__ void __launch_bounds__(256) fma2_chain_kernel(float* out, int iters, float a, float b) {
    unsigned long long x[8], aa, bb;
    asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
#pragma unroll
    for (int u = 0; u < 8; ++u) asm("mov.b64 %0, {%1, %2};" : "=l"(x[u]) : "f"((float)threadIdx.x + u), "f"((float)u));
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int u = 0; u < 8; ++u) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[u]) : "l"(aa), "l"(bb));
    }
    float s = 0.f;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[u]));
        s += lo + hi;
    }
    if (s == 123456789.f) out[0] = s;
}

// legacy tensor-core path: mma.sync m16n8k8 TF32 -> FP32, 8 independent accumulator fragments per warp
// (what a 3xTF32 version of the update's 64-wide products could draw on; measured, not assumed)
__global__ void __launch_bounds__(256) mma_tf32_chain_kernel(float* out, int iters) {
    float c[8][4];
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int r = 0; r < 4; ++r) c[u][r] = (float)(threadIdx.x + u + r);
    const unsigned a0 = 0x3f800000u, a1 = 0x3f000000u, a2 = 0x3e800000u, a3 = 0x3f400000u, b0 = 0x3f800000u, b1 = 0x3f000000u;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[u][0]), "+f"(c[u][1]), "+f"(c[u][2]), "+f"(c[u][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    float s = 0.f;
#pragma unroll
    for (int u = 0; u < 8; ++u) s += c[u][0] + c[u][1] + c[u][2] + c[u][3];
    if (s == 123456789.f) out[0] = s;
}

template <typename T>
double measure(int iters, bool packed = false, bool tensor = false) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    T* out = nullptr;
    if (cudaMalloc(&out, sizeof(T)) != cudaSuccess) return -1.0;
    const int grid = sms * 8, block = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        if (tensor) mma_tf32_chain_kernel<<<grid, block>>>((float*)out, iters);
        else if (packed) fma2_chain_kernel<<<grid, block>>>((float*)out, iters, 1.0000001f, 1e-7f);
        else fma_chain_kernel<T><<<grid, block>>>(out, iters, (T)1.0000001, (T)1e-7);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        count_launch();
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        // per thread-iteration: 64 FMAs (scalar / packed chains) or 8 warp-wide m16n8k8 MMAs = 8 * 1024 / 32 FMAs
        const double flops = 2.0 * (tensor ? 256.0 : 64.0) * (double)iters * (double)grid * block;
        if (rep > 0 && ms > 0.f) best = fmax(best, flops / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    return cudaGetLastError() == cudaSuccess ? best : -1.0;
}

}  // namespace
}  // namespace rk

extern "C" RK_API double rk_fma_peak(int32_t use_fp64, int32_t iters) {
    if (iters <= 0) iters = 1024;
    if (use_fp64 == 3) return rk::measure<float>(iters, false, true);  // mma.sync m16n8k8 TF32 (legacy tensor path)
    if (use_fp64 == 2) return rk::measure<float>(iters, true);  // packed fp32 (FFMA2), same 64 FMAs per thread-iteration
    return use_fp64 ? rk::measure<double>(iters) : rk::measure<float>(iters);
}
