// tcgen05 / TMEM building blocks shared by the tensor-core kernels (rk_train.cu: the PPO minibatch gradient; rk_rollout.cu: policy
// inference): K-major no-swizzle operand tiles and their descriptors, the TF32 hi/lo splits, TMEM loads / stores, the 3-term
// product with A in tensor memory, mbarrier waits.  Stand-alone checks of the building blocks: tools/umma_selftest.cu,
// tools/umma_dw_selftest.cu.
#pragma once
#include <cstdint>

namespace rk {
namespace {

constexpr int kUmmaH = 64;     // hidden width = N of the per-sample products
constexpr int kUmmaTS = 128;   // samples per tile = M

// tanh(x) = 1 - 2 / (exp(2x) + 1); ex2.approx.ftz directly (one MUFU): __expf adds a denormal-range fix-up (compare +
// two multiplies per call) that only matters where the result is -1 anyway
__device__ __forceinline__ float tanh_fast(float x) {
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.8853900817779268f));
    return 1.f - __fdividef(2.f, e + 1.f);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// K-major no-swizzle operand tile [rows][F features]: 8-row x 16-byte core matrices, feature cores contiguous
__device__ __forceinline__ int umma_off(int r, int f, int F) { return ((r >> 3) * (F >> 2) + (f >> 2)) * 32 + (r & 7) * 4 + (f & 3); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ float tf32_rn(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    const float h = tf32_rn(x);
    hi = __float_as_uint(h);
    lo = __float_as_uint(tf32_rn(x - h));
}
// The same split in three instructions (variant 2): hi = x rounded to TF32's 11 significant bits (add half an ulp to the
// bit pattern, clear the low 13 bits: round-to-nearest, ties away), lo = x - hi exactly; lo is handed to the tensor core
// as it is -- kind::tf32 reads the upper 19 bits of each 32-bit operand, so hi + lo carries >= 21 bits of x.
__device__ __forceinline__ void split_tf32_fast(float x, uint32_t& hi, uint32_t& lo) {
    hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// D[128 x 64] (+)= A[128 x K] . B[64 x K]^T, 3-pass split; A hi/lo in TMEM columns, B hi/lo K-major tiles in smem
__device__ __forceinline__ void issue_product(uint32_t tmem, int colD, int colAh, int colAl, const float* Bh, const float* Bl,
                                              int K, unsigned long long* bar) {
    // instruction descriptor: D fp32, A/B tf32, both K-major, N = 64, M = 128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kUmmaH >> 3) << 17) | ((uint32_t)(kUmmaTS >> 4) << 24);
    const uint32_t sbo = (uint32_t)(K >> 2) * 128u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t acc = 0;
#pragma unroll 1
    for (int pass = 0; pass < 3; ++pass) {
        const int colA = (pass == 2) ? colAl : colAh;
        const uint32_t b = smem_u32((pass == 1) ? Bl : Bh);
        for (int kb = 0; kb < (K >> 3); ++kb) {
            const uint64_t db = umma_desc(b + kb * 256, 128, sbo);
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
                         ::"r"(tmem + colD), "r"(tmem + colA + kb * 8), "l"(db), "r"(idesc), "r"(acc) : "memory");
            acc = 1;
        }
    }
    if (bar != nullptr)   // (nullptr: the caller issues more products and commits them together)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Bounded: if the tensor core never signalled completion (a wrong descriptor) the kernel finishes with wrong
// results, which the parity tests catch, instead of hanging the GPU.
__device__ __forceinline__ bool wait_product(unsigned long long* bar, unsigned& phase) {
    unsigned done = 0;
    const uint32_t a = smem_u32(bar);
    // each try_wait may suspend the thread in hardware for up to the hinted time, so the (bounded) loop turns rarely
    for (int spins = 0; !done && spins < (1 << 16); ++spins)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(a), "r"(phase), "r"(20000u) : "memory");
    phase ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    return done != 0;
}
// publish this thread's tcgen05.st writes to the thread that issues the next MMA
__device__ __forceinline__ void tmem_publish_and_sync() {
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void mma_commit(unsigned long long* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// one lane of a converged warp (the tensor-core instructions are issued by a single thread)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

}  // namespace
}  // namespace rk
