// Microbenchmark 2: true tensor-pipe time of tcgen05.mma (bf16, K=16, SWIZZLE_NONE smem operands) with an issue loop that is
// NOT instruction-bound: 32 MMAs fully unrolled with loop-invariant descriptors; optionally two issuing warps.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I cleanba_b200/csrc tools/mma_microbench2.cu -o tools/bin/mma_microbench2
#include <cstdio>
#include <cuda_runtime.h>
#include "umma.cuh"
using namespace cb::umma;

template <int M, int N>
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
        const uint32_t a_hi = desc_hi(128), b_hi = desc_hi(128);
        const uint32_t a_lo = desc_lo(a, 2816), b_lo = desc_lo(b, N * 16);
        const uint32_t d = tm + (warp - 1) * 256;
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 32; ++u) mma_bf16_parts(d, a_lo + (u & 7), a_hi, b_lo, b_hi, idesc, 1);
        }
        mma_commit(&bar[warp - 1]);
        mbar_wait(&bar[warp - 1], 0);
        if (blockIdx.x == 0 && warp == 1) out[0] = clock64() - t0;
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 512);
}

template <int M, int N> void run(long long* c) {
    cudaFuncSetAttribute(k_time<M, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
    for (int issuers = 1; issuers <= 2; ++issuers) {
        k_time<M, N><<<148, 128, 128 * 1024>>>(100, issuers, c);
        cudaError_t e = cudaDeviceSynchronize();
        long long cyc = 0; cudaMemcpy(&cyc, c, 8, cudaMemcpyDeviceToHost);
        printf("M=%3d N=%3d issuers=%d: %6.1f cycles per MMA per issuer (%6.1f per MMA overall; ideal math %5.1f) %s\n", M, N, issuers,
               cyc / 3200.0, cyc / 3200.0 / issuers, (M < 128 ? 128.0 : M) * N * 16 / 4096.0 * 0 + M * N * 16 / 4096.0,
               e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
}

int main() {
    long long* c; cudaMalloc(&c, 8);
    run<128, 16>(c); run<128, 32>(c); run<128, 48>(c); run<128, 64>(c); run<128, 96>(c); run<128, 128>(c); run<128, 256>(c);
    run<64, 16>(c); run<64, 32>(c); run<64, 64>(c); run<64, 96>(c);
    return 0;
}
