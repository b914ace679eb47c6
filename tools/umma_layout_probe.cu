// D[M x N] = A[M x K] * B[N x K]^T, K = 64, N = 64; each operand either stored K-major (tile rows = mn index, 64 k-features)
// or stored "transposed" (tile rows = k index, features = mn index) and described MN-major.  Reports the lane of every row.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// tile [R rows x F features] : core (r/8, f/4) at ((r/8)*(F/4) + f/4)*128 B, within (r%8)*16 + (f%4)*4
__device__ __forceinline__ int off(int r, int f, int F) { return ((r >> 3) * (F >> 2) + (f >> 2)) * 32 + (r & 7) * 4 + (f & 3); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) | ((uint64_t)((sbo >> 4) & 0x3fff) << 32) | ((uint64_t)1 << 46);
}
__global__ void __launch_bounds__(128) k(const float* A, const float* B, float* Dall, int M, int a_mn, int b_mn, int swap) {
    extern __shared__ __align__(1024) float sm[];
    const int K = 64, N = 64;
    float* sA = sm; float* sB = sm + 128 * 64;
    __shared__ __align__(8) unsigned long long bar; __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int e = tid; e < M * K; e += 128) { int m = e / K, kk = e % K; if (a_mn) sA[off(kk, m, M)] = A[e]; else sA[off(m, kk, K)] = A[e]; }
    for (int e = tid; e < N * K; e += 128) { int n = e / K, kk = e % K; if (b_mn) sB[off(kk, n, N)] = B[e]; else sB[off(n, kk, K)] = B[e]; }
    if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(64)); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    {
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
        for (int c = 0; c < 64; ++c) { uint32_t z = 0x7f800000u; asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr + c), "r"(z) : "memory"); }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        for (int kb = 0; kb < K / 8; ++kb) {
            uint64_t da, db;
            // K-major tile [mn rows x 64 k]: k-cores 128 B apart (LBO), 8-row groups (K/4)*128 apart (SBO); next MMA: +2 cores
            // transposed tile [k rows x MN feats]: 8-k-row groups (MN/4)*128 apart, mn-cores 128 B apart; next MMA: +1 row group
            if (!a_mn) da = make_desc(smem_u32(sA) + kb * 256, 128, (K / 4) * 128);
            else { uint32_t kdir = (M / 4) * 128, mndir = 128; da = make_desc(smem_u32(sA) + kb * kdir, swap ? mndir : kdir, swap ? kdir : mndir); }
            if (!b_mn) db = make_desc(smem_u32(sB) + kb * 256, 128, (K / 4) * 128);
            else { uint32_t kdir = (N / 4) * 128, mndir = 128; db = make_desc(smem_u32(sB) + kb * kdir, swap ? mndir : kdir, swap ? kdir : mndir); }
            uint32_t acc = kb > 0;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                         ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    unsigned done = 0; int spins = 0;
    while (!done && ++spins < (1 << 22))
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c = 0; c < 64; ++c) {
        uint32_t v; const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        Dall[(warp * 32 + lane) * 64 + c] = done ? __uint_as_float(v) : -12345.f;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64));
}
int main() {
    const int K = 64, N = 64;
    auto tf = [](float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xffffe000u; memcpy(&x, &u, 4); return x; };
    float *dA, *dB, *dD; cudaMalloc(&dA, 128 * 64 * 4); cudaMalloc(&dB, 64 * 64 * 4); cudaMalloc(&dD, 128 * 64 * 4);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 128 * 64 * 4);
    int cases[][4] = {{64, 0, 0, 0}, {128, 1, 0, 0}, {128, 1, 0, 1}, {128, 0, 1, 0}, {128, 0, 1, 1}, {64, 1, 1, 0}, {64, 1, 1, 1}};
    for (auto& cs : cases) {
        const int M = cs[0];
        std::vector<float> A(M * K), B(N * K), D(128 * 64);
        srand(3);
        for (auto& x : A) x = tf((float)rand() / RAND_MAX * 2 - 1);
        for (auto& x : B) x = tf((float)rand() / RAND_MAX * 2 - 1);
        cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
        k<<<1, 128, 2 * 128 * 64 * 4>>>(dA, dB, dD, M, cs[1], cs[2], cs[3]);
        cudaError_t e = cudaDeviceSynchronize(); if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        std::vector<double> R(M * N, 0.0);
        for (int i = 0; i < M; ++i) for (int j = 0; j < N; ++j) { double a = 0; for (int s = 0; s < K; ++s) a += (double)A[i * K + s] * B[j * K + s]; R[i * N + j] = a; }
        int good = 0, untouched = 0; char map[600] = ""; int pos = 0;
        for (int lane = 0; lane < 128; ++lane) {
            if (std::isinf(D[lane * 64])) { untouched++; continue; }
            int best = -1; double be = 1e30;
            for (int i = 0; i < M; ++i) { double err = 0; for (int j = 0; j < N; ++j) err = fmax(err, fabs(R[i * N + j] - D[lane * 64 + j])); if (err < be) { be = err; best = i; } }
            if (be < 1e-3) { good++; if (lane % 16 == 0 && pos < 560) pos += sprintf(map + pos, " L%d=r%d", lane, best); }
        }
        printf("M=%3d a_mn=%d b_mn=%d swap=%d : %3d lanes match a row, %3d untouched;%s\n", M, cs[1], cs[2], cs[3], good, untouched, map);
    }
    return 0;
}
