// Microbenchmark 3: does the tensor pipe's ~50-cycle floor for small-N tcgen05.mma (M128 x N<=96 x K16, bf16) depend on the
// shared-memory layout of the A operand?  Variants: SWIZZLE_NONE (8x16-byte core matrices, what the conv kernels use),
// SWIZZLE_32B / 64B / 128B K-major atoms, and A taken from TMEM (no shared-memory A fetch at all).  Data content is irrelevant.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I cleanba_b200/csrc tools/mma_microbench3.cu -o tools/bin/mma_microbench3
#include <cstdio>
#include <cuda_runtime.h>
#include "umma.cuh"
using namespace cb::umma;

__device__ __forceinline__ void mma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc)
        : "memory");
}

// SW: 0 none, 6 = 32B, 4 = 64B, 2 = 128B (sm_100 descriptor layout_type, bits 61-63); 99 = A in TMEM
template <int M, int N, int SW>
__global__ void __launch_bounds__(128) k_time(int iters, int issuers, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar[2];
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;
    if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_barrier_init(); }
    fence_proxy_async();
    if (warp == 0) tmem_alloc(&slot, 512);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = slot;
    if (warp >= 1 && warp <= issuers && lane == 0) {
        const uint32_t a = smem_u32(smem) + (warp - 1) * 16384, b = smem_u32(smem) + 48 * 1024;
        constexpr uint32_t idesc = make_idesc_bf16(M, N, 0, 0);
        constexpr uint32_t sbo = SW == 0 ? 128 : (SW == 6 ? 256 : (SW == 4 ? 512 : 1024));
        const uint32_t a_hi = desc_hi(sbo) | ((uint32_t)(SW == 99 ? 0 : SW) << 29), b_hi = desc_hi(128);
        const uint32_t a_lo = desc_lo(a, SW == 0 ? 2816 : 16), b_lo = desc_lo(b, N * 16);
        const uint32_t d = tm + (warp - 1) * 256;
        const uint32_t a_t = tm + 480 + (warp - 1) * 16;     // A in TMEM: 128 lanes x 8 columns (16 bf16)
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 32; ++u) {
                if (SW == 99) mma_bf16_ts(d, a_t, b_lo, b_hi, idesc, 1);
                else mma_bf16_parts(d, a_lo + (SW == 0 ? (u & 7) : 2 * (u & 1)), a_hi, b_lo, b_hi, idesc, 1);
            }
        }
        mma_commit(&bar[warp - 1]);
        mbar_wait(&bar[warp - 1], 0);
        if (blockIdx.x == 0 && warp == 1) out[0] = clock64() - t0;
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 512);
}

template <int M, int N, int SW> void run(long long* c, const char* name) {
    cudaFuncSetAttribute(k_time<M, N, SW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
    for (int issuers = 1; issuers <= 2; ++issuers) {
        k_time<M, N, SW><<<148, 128, 128 * 1024>>>(100, issuers, c);
        cudaError_t e = cudaDeviceSynchronize();
        long long cyc = 0; cudaMemcpy(&cyc, c, 8, cudaMemcpyDeviceToHost);
        printf("%-12s M=%3d N=%3d issuers=%d: %6.1f cycles per MMA per issuer (%6.1f per MMA overall; ideal math %5.1f) %s\n", name, M, N,
               issuers, cyc / 3200.0, cyc / 3200.0 / issuers, M * N * 16 / 4096.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
}

int main() {
    long long* c; cudaMalloc(&c, 8);
    run<128, 32, 0>(c, "SWIZZLE_NONE"); run<128, 64, 0>(c, "SWIZZLE_NONE"); run<128, 96, 0>(c, "SWIZZLE_NONE");
    run<128, 32, 6>(c, "SWIZZLE_32B"); run<128, 64, 6>(c, "SWIZZLE_32B"); run<128, 96, 6>(c, "SWIZZLE_32B");
    run<128, 32, 4>(c, "SWIZZLE_64B"); run<128, 64, 4>(c, "SWIZZLE_64B"); run<128, 96, 4>(c, "SWIZZLE_64B");
    run<128, 32, 2>(c, "SWIZZLE_128B"); run<128, 64, 2>(c, "SWIZZLE_128B"); run<128, 96, 2>(c, "SWIZZLE_128B");
    run<128, 32, 99>(c, "A_in_TMEM"); run<128, 64, 99>(c, "A_in_TMEM"); run<128, 96, 99>(c, "A_in_TMEM");
    return 0;
}
