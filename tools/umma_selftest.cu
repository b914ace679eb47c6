// Stand-alone check of the tcgen05 building blocks used by the tensor-core gradient kernel:
// D[128 x 64] (fp32, TMEM) = A[128 x K] * B[64 x K]^T with kind::tf32, operands in shared memory in the
// canonical K-major no-swizzle layout (8-row x 16-byte core matrices), one CTA, one issuing thread.
//   nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/_umma_selftest.bin tools/umma_selftest.cu
// Modes: 1 pass (plain TF32) and 3 passes (hi*hi + hi*lo + lo*hi: fp32-grade products).
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

constexpr int M = 128, N = 64, K = 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, no swizzle: element (r, k) of a [rows x K] tile at ((r/8) * (K/4) + k/4) * 128 + (r%8) * 16 + (k%4) * 4 bytes
__device__ __forceinline__ int kmajor_off(int r, int k, int Kt) { return ((r >> 3) * (Kt >> 2) + (k >> 2)) * 32 + (r & 7) * 4 + (k & 3); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
    return d;                 // base_offset 0, lbo_mode 0, layout_type 0 = SWIZZLE_NONE
}

__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

__global__ void __launch_bounds__(128) umma_test_kernel(const float* A, const float* B, float* D, int passes) {
    extern __shared__ __align__(1024) float sm[];
    float* sAh = sm;                 // [M x K] hi
    float* sAl = sAh + M * K;        // lo
    float* sBh = sAl + M * K;        // [N x K]
    float* sBl = sBh + N * K;
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int e = tid; e < M * K; e += 128) {
        const int r = e / K, k = e % K;
        const float v = A[e], h = tf32_hi(v);
        sAh[kmajor_off(r, k, K)] = h;
        sAl[kmajor_off(r, k, K)] = tf32_hi(v - h);
    }
    for (int e = tid; e < N * K; e += 128) {
        const int r = e / K, k = e % K;
        const float v = B[e], h = tf32_hi(v);
        sBh[kmajor_off(r, k, K)] = h;
        sBl[kmajor_off(r, k, K)] = tf32_hi(v - h);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(64));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;

    if (tid == 0) {
        // instruction descriptor: D fp32, A/B tf32, both K-major, N = 64, M = 128
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint32_t lbo = 128, sbo = (K / 4) * 128;
        uint32_t acc = 0;
        for (int pass = 0; pass < passes; ++pass) {
            const float* a = (pass == 2) ? sAl : sAh;
            const float* b = (pass == 1) ? sBl : sBh;
            for (int k = 0; k < K / 8; ++k) {
                const uint64_t da = make_desc(smem_u32(a) + k * 256, lbo, sbo);
                const uint64_t db = make_desc(smem_u32(b) + k * 256, lbo, sbo);
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                    ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
                acc = 1;
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    {
        unsigned done = 0;
        int spins = 0;
        while (!done && ++spins < (1 << 22))   // bounded: a wrong descriptor must not hang the GPU
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
        if (!done && tid == 0) D[0] = __int_as_float(0x7fc00000);   // NaN marker: the MMA never signalled completion
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // warp w reads TMEM lanes 32w..32w+31 (its quarter), 32 columns at a time: thread <-> accumulator row
    uint32_t v[32];
    for (int c0 = 0; c0 < N; c0 += 32) {
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                     "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                     "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                       "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                       "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                       "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int c = 0; c < 32; ++c) D[(warp * 32 + lane) * N + c0 + c] = __uint_as_float(v[c]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64));
}

int main() {
    std::vector<float> hA(M * K), hB(N * K), hD(M * N);
    srand(1);
    for (auto& x : hA) x = (float)rand() / RAND_MAX * 2.f - 1.f;
    for (auto& x : hB) x = (float)rand() / RAND_MAX * 2.f - 1.f;
    float *dA, *dB, *dD;
    cudaMalloc(&dA, hA.size() * 4); cudaMalloc(&dB, hB.size() * 4); cudaMalloc(&dD, hD.size() * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice);
    const size_t smem = (size_t)(2 * M * K + 2 * N * K) * 4;
    cudaFuncSetAttribute(umma_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int rc = 0;
    for (int passes : {1, 3}) {
        cudaMemset(dD, 0, hD.size() * 4);
        umma_test_kernel<<<1, 128, smem>>>(dA, dB, dD, passes);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("passes %d: CUDA error %s\n", passes, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
        double max_err = 0, max_ref = 0;
        for (int m = 0; m < M; ++m)
            for (int n = 0; n < N; ++n) {
                double ref = 0;
                for (int k = 0; k < K; ++k) ref += (double)hA[m * K + k] * (double)hB[n * K + k];
                max_err = fmax(max_err, fabs(ref - (double)hD[m * N + n]));
                max_ref = fmax(max_ref, fabs(ref));
            }
        printf("passes %d: max |D - A B^T| = %.3e (max |ref| %.2f)\n", passes, max_err, max_ref);
        if (passes == 1 && max_err > 2e-2) rc = 1;
        if (passes == 3 && max_err > 2e-5) rc = 1;
    }
    printf(rc ? "FAILED\n" : "OK\n");
    return rc;
}
