// probe: tcgen05.mma with the A operand in TMEM (written by tcgen05.st, thread <-> row), B K-major in smem, 3-pass split
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
constexpr int M = 128, N = 64, K = 64;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ int off(int r, int f, int F) { return ((r >> 3) * (F >> 2) + (f >> 2)) * 32 + (r & 7) * 4 + (f & 3); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) | ((uint64_t)((sbo >> 4) & 0x3fff) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }
__global__ void __launch_bounds__(128) k(const float* A, const float* B, float* D) {
    extern __shared__ __align__(1024) float sm[];
    float* sBh = sm; float* sBl = sm + N * K;
    __shared__ __align__(8) unsigned long long bar; __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int e = tid; e < N * K; e += 128) { int n = e / K, kk = e % K; float v = B[e], h = tf32_hi(v); sBh[off(n, kk, K)] = h; sBl[off(n, kk, K)] = tf32_hi(v - h); }
    if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(256)); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    const uint32_t colD = 0, colAh = 64, colAl = 128;
    // thread t writes row t of A (hi and lo parts) into its TMEM lane, 8 columns per instruction
    {
        const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
        for (int c0 = 0; c0 < K; c0 += 8) {
            uint32_t h[8], l[8];
            for (int c = 0; c < 8; ++c) { float v = A[tid * K + c0 + c], hv = tf32_hi(v); h[c] = __float_as_uint(hv); l[c] = __float_as_uint(tf32_hi(v - hv)); }
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                         ::"r"(lane_base + colAh + c0), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]), "r"(h[4]), "r"(h[5]), "r"(h[6]), "r"(h[7]) : "memory");
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                         ::"r"(lane_base + colAl + c0), "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]), "r"(l[4]), "r"(l[5]), "r"(l[6]), "r"(l[7]) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        uint32_t acc = 0;
        for (int pass = 0; pass < 3; ++pass) {
            const uint32_t a_col = (pass == 2) ? colAl : colAh;
            const float* b = (pass == 1) ? sBl : sBh;
            for (int kb = 0; kb < K / 8; ++kb) {
                const uint64_t db = make_desc(smem_u32(b) + kb * 256, 128, (K / 4) * 128);
                const uint32_t ta = tmem + a_col + kb * 8;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
                             ::"r"(tmem + colD), "r"(ta), "l"(db), "r"(idesc), "r"(acc) : "memory");
                acc = 1;
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    unsigned done = 0; int spins = 0;
    while (!done && ++spins < (1 << 22))
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c = 0; c < N; ++c) {
        uint32_t v; const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + colD + c;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        D[tid * N + c] = done ? __uint_as_float(v) : -12345.f;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}
int main() {
    std::vector<float> A(M * K), B(N * K), D(M * N);
    srand(4);
    for (auto& x : A) x = (float)rand() / RAND_MAX * 2 - 1;
    for (auto& x : B) x = (float)rand() / RAND_MAX * 2 - 1;
    float *dA, *dB, *dD; cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * N * K * 4);
    k<<<1, 128, 2 * N * K * 4>>>(dA, dB, dD);
    cudaError_t e = cudaDeviceSynchronize(); if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double me = 0;
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double r = 0; for (int s = 0; s < K; ++s) r += (double)A[m * K + s] * B[n * K + s]; me = fmax(me, fabs(r - D[m * N + n])); }
    printf("A from TMEM, 3-pass: max |D - A B^T| = %.3e  %s\n", me, me < 5e-5 ? "OK" : "MISMATCH");
    return 0;
}
