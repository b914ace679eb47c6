// Building block of the tensor-core weight gradients (rk_train.cu, ppo_mlp_grad_tc2_kernel), stand-alone:
//   D[64 x N] += A[64 x K] . B[N x K]^T,  K = 128 samples per tile, accumulated over several tiles in tensor memory,
// with BOTH operands in shared memory, K-major, no swizzle, core matrices (8 rows x 16 bytes) PADDED along K
// (144 bytes between the k-cores of a row group, so that 32 threads that own 32 consecutive samples store one
// feature row without bank conflicts), 3-pass TF32 split (hi.hi + hi.lo + lo.hi), M = 64 accumulator
// (row r -> TMEM lane 32*(r/16) + r%16), N = 72 and N = 24, at a non-zero TMEM column.
// nvcc -gencode arch=compute_100a,code=sm_100a -o umma_dw_selftest umma_dw_selftest.cu && ./umma_dw_selftest
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
constexpr int kLbo = 144, kSbo = 32 * kLbo;   // bytes
// float offset of (row r, sample k) in a padded K-major tile
__device__ __forceinline__ int poff(int r, int k) { return ((r >> 3) * kSbo + (k >> 2) * kLbo) / 4 + (r & 7) * 4 + (k & 3); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) | ((uint64_t)((sbo >> 4) & 0x3fff) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ float tf32_rn(float x) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return __uint_as_float(r); }
__global__ void __launch_bounds__(128) k(const float* A, const float* B, float* Dall, int N, int tiles, int col0) {
    extern __shared__ __align__(1024) float sm[];
    const int K = 128;
    const int rgA = 8, rgB = N / 8;
    float* sAh = sm; float* sAl = sAh + rgA * kSbo / 4; float* sBh = sAl + rgA * kSbo / 4; float* sBl = sBh + rgB * kSbo / 4;
    __shared__ __align__(8) unsigned long long bar; __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(256)); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::); }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    unsigned phase = 0;
    for (int t = 0; t < tiles; ++t) {
        // thread = sample k: stores one feature row at a time (the access pattern of the real epilogues)
        for (int r = 0; r < 64; ++r) { const float v = A[((size_t)t * 64 + r) * K + tid]; const float h = tf32_rn(v); sAh[poff(r, tid)] = h; sAl[poff(r, tid)] = tf32_rn(v - h); }
        for (int r = 0; r < N; ++r) { const float v = B[((size_t)t * N + r) * K + tid]; const float h = tf32_rn(v); sBh[poff(r, tid)] = h; sBl[poff(r, tid)] = tf32_rn(v - h); }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (tid == 0) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(64 >> 4) << 24);
            for (int pass = 0; pass < 3; ++pass) {
                const uint32_t a = smem_u32(pass == 2 ? sAl : sAh), b = smem_u32(pass == 1 ? sBl : sBh);
                for (int kb = 0; kb < K / 8; ++kb) {
                    const uint64_t da = make_desc(a + kb * 2 * kLbo, kLbo, kSbo), db = make_desc(b + kb * 2 * kLbo, kLbo, kSbo);
                    const uint32_t acc = (t > 0 || pass > 0 || kb > 0);
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                                 ::"r"(tmem + col0), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        }
        unsigned done = 0; int spins = 0;
        while (!done && ++spins < (1 << 22))
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(phase) : "memory");
        phase ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        __syncthreads();
    }
    for (int c = 0; c < N; ++c) {
        uint32_t v; const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + col0 + c;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (lane < 16) Dall[(warp * 16 + lane) * N + c] = __uint_as_float(v);   // row = 16 * lane quarter + lane
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}
int main() {
    const int K = 128, tiles = 3;
    int bad = 0;
    for (int N : {72, 24, 64}) {
        std::vector<float> A((size_t)tiles * 64 * K), B((size_t)tiles * N * K), D(64 * N);
        srand(5 + N);
        for (auto& x : A) x = (float)rand() / RAND_MAX * 2 - 1;
        for (auto& x : B) x = (float)rand() / RAND_MAX * 2 - 1;
        float *dA, *dB, *dD; cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
        cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
        const size_t smem = (size_t)(2 * 8 + 2 * (N / 8)) * kSbo;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k<<<1, 128, smem>>>(dA, dB, dD, N, tiles, N == 24 ? 136 : 64);
        cudaError_t e = cudaDeviceSynchronize(); if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        double err = 0, ref = 0;
        for (int i = 0; i < 64; ++i) for (int j = 0; j < N; ++j) {
            double a = 0;
            for (int t = 0; t < tiles; ++t) for (int s = 0; s < K; ++s) a += (double)A[((size_t)t * 64 + i) * K + s] * B[((size_t)t * N + j) * K + s];
            err = fmax(err, fabs(a - D[i * N + j])); ref = fmax(ref, fabs(a));
        }
        printf("N=%2d tiles=%d smem=%zu : max |D - A B^T| = %.3e (max |ref| %.2f) %s\n", N, tiles, smem, err, ref, err < 2e-4 ? "OK" : "FAIL");
        bad += !(err < 2e-4);
        cudaFree(dA); cudaFree(dB); cudaFree(dD);
    }
    printf(bad ? "FAILED\n" : "OK\n");
    return bad;
}
