// Inline-PTX wrappers for the Blackwell (sm_100a) tensor-core path: mbarrier, 1-D bulk TMA copies
// (cp.async.bulk -> UBLKCP), tcgen05.mma (-> UTCHMMA) with shared-memory matrix descriptors, TMEM
// allocation and tcgen05.ld (-> LDTM).  No CUTLASS: the descriptor bit layouts are restated from the PTX ISA
// (matrix-descriptor / instruction-descriptor tables for tcgen05.mma kind::f16).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cb {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug traps (launch failure reported to the host) instead of hanging the GPU.  The bound is
// wall-clock (%globaltimer, 20 s), not a spin count: a legitimate long stall (time slicing between green contexts or
// processes, a profiler replaying the kernel) must not kill the context.  The timer is only read every 2^16 failed polls.
__device__ __forceinline__ uint64_t globaltimer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    uint64_t t0 = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 0xffffu) == 0) {
            const uint64_t now = globaltimer_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 20000000000ull) __trap();
        }
    }
}

// ---------------------------------------------------------------- TMA: 1-D bulk copy global -> shared
// size must be a multiple of 16 bytes, both addresses 16-byte aligned; completion is signalled on `bar` (complete_tx).
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---------------------------------------------------------------- TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 32 lanes x 32-bit, 16 consecutive columns: thread i of the warp receives lane (base_lane + i), columns c..c+15
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ---------------------------------------------------------------- descriptors
// Shared-memory matrix descriptor, SWIZZLE_NONE ("interleaved" 8x16-byte core matrices), Blackwell version bit set.
//   K-major : addr(r, k)  = start + (r % 8) * 16 + (r / 8) * SBO + (k % 8) * 2 + (k / 8) * LBO      (r = M or N index)
//   MN-major: addr(mn, k) = start + (mn % 8) * 2 + (mn / 8) * SBO + (k % 8) * 16 + (k / 8) * LBO
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;   // descriptor version (sm_100)
    return d;                 // base_offset = 0, lbo_mode = 0, layout_type = SWIZZLE_NONE (0)
}

// Instruction descriptor for kind::f16 with bf16 operands and fp32 accumulation.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4)                       // D format  = F32
           | (1u << 7)                     // A format  = BF16
           | (1u << 10)                    // B format  = BF16
           | ((uint32_t)a_mn_major << 15)  // A major   (0 = K, 1 = MN)
           | ((uint32_t)b_mn_major << 16)  // B major
           | ((uint32_t)(N >> 3) << 17)    // N / 8
           | ((uint32_t)(M >> 4) << 24);   // M / 16
}

// Instruction descriptor for kind::f16 with fp16 operands (A / B format fields 0) and fp32 accumulation: the trunk carriers.
// (tcgen05 rejects mixed fp16 x bf16 operands -- tools/mma_mixed_format_probe.cu -- so every trunk tensor is fp16.)
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4)                       // D format  = F32
           | ((uint32_t)a_mn_major << 15)  // A major   (0 = K, 1 = MN)
           | ((uint32_t)b_mn_major << 16)  // B major
           | ((uint32_t)(N >> 3) << 17)    // N / 8
           | ((uint32_t)(M >> 4) << 24);   // M / 16
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread.
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Split form for hot issue loops: the 64-bit descriptors are assembled from a precomputed high word (SBO + version) and a
// low word (start address + LBO) so that stepping through taps / K blocks is ONE integer add per operand.  Measured on
// B200 (tools/mma_microbench.cu): an M128 x N<=96 x K16 SS MMA occupies the tensor pipe for ~56 cycles, so the single
// issuing thread must spend well under that per MMA or it, not the tensor pipe, becomes the bottleneck.
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr, uint32_t lbo_bytes) {
    return ((smem_addr >> 4) & 0x3FFF) | (((lbo_bytes >> 4) & 0x3FFF) << 16);
}
__device__ __forceinline__ uint32_t desc_hi(uint32_t sbo_bytes) { return ((sbo_bytes >> 4) & 0x3FFF) | (1u << 14); }
__device__ __forceinline__ void mma_bf16_parts(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                               uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
        ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Warp-uniform issue.  tcgen05.mma / tcgen05.commit take their descriptors from UNIFORM registers.  Issued from inside an
// `if (lane == 0)` branch the operands live in per-thread registers the compiler cannot prove uniform, and every MMA becomes a
// "waterfall" (R2UR + ELECT + BRA.U.ANY per operand): ~170 cycles per MMA on one thread -- measured as THE limiter of the
// 32-channel conv kernels (ncu source page, profiles/r02_v2_ncu_full_conv_4_32_raw.csv: the issuing warp never waits).  Instead
// the whole warp runs the issue loop converged (warp index made provably uniform with a shuffle, all descriptor arithmetic on
// warp-uniform values -> uniform datapath) and only the instruction itself is predicated on an elected lane.
__device__ __forceinline__ int uniform_warp_idx() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }
__device__ __forceinline__ uint32_t elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred;
}
__device__ __forceinline__ void mma_f16_elect(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                              uint32_t accumulate, uint32_t leader) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "setp.ne.b32 q, %7, 0;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
        ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate), "r"(leader)
        : "memory");
}
__device__ __forceinline__ void mma_commit_elect(uint64_t* bar, uint32_t leader) {
    asm volatile(
        "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar)), "r"(leader)
        : "memory");
}
// mbarrier arrives when all tcgen05.mma previously issued by this thread have completed (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

}  // namespace umma
}  // namespace cb
