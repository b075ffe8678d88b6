// Microbenchmark: cycles per tcgen05.mma (kind::f16, bf16, M=128, K=16, SWIZZLE_NONE operands in shared memory) as a
// function of N, operand major-ness and whether consecutive MMAs target the same accumulator.  One CTA per SM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I cleanba_b200/csrc tools/mma_microbench.cu -o tools/bin/mma_microbench
#include <cstdio>
#include <cuda_runtime.h>
#include "umma.cuh"
using namespace cb::umma;

__global__ void __launch_bounds__(128) k_bench(int N, int mn_major, int iters, int rotate_acc, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;  // small bf16 values
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    fence_proxy_async();
    if (warp == 0) tmem_alloc(&slot, 512);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = slot;
    if (warp == 1 && lane == 0) {
        const uint32_t a = smem_u32(smem), b = a + 64 * 1024;
        const uint32_t idesc = make_idesc_bf16(128, N, mn_major, mn_major);
        // K-major: rows 16 B apart (SBO 128), k-chunks 2816 B apart (like a conv window plane); MN-major: LBO 128, SBO 1056/1024
        const uint64_t da = mn_major ? make_desc(a, 128, 1056) : make_desc(a, 2816, 128);
        const uint64_t db = mn_major ? make_desc(b, 128, 1024) : make_desc(b, (uint32_t)N * 16, 128);
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            uint32_t d = tm + (rotate_acc ? ((i & 1) * 256) : 0);
            mma_bf16(d, da + (uint64_t)((i & 7) * 2), db, idesc, i > 1);
        }
        mma_commit(&bar);
        mbar_wait(&bar, 0);
        long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 512);
}

int main() {
    long long* d; cudaMalloc(&d, 8);
    cudaFuncSetAttribute(k_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    int Ns[] = {16, 32, 48, 64, 96, 128, 192, 256};
    for (int mn = 0; mn < 2; ++mn)
        for (int rot = 0; rot < 2; ++rot)
            for (int N : Ns) {
                const int iters = 2000;
                k_bench<<<148, 128, 200 * 1024>>>(N, mn, iters, rot, d);
                cudaError_t e = cudaDeviceSynchronize();
                long long c = 0; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
                printf("%s rotate_acc=%d N=%3d : %7.1f cycles/MMA  (ideal math %5.1f)  %s\n", mn ? "MN-major" : "K-major ", rot, N,
                       (double)c / iters, 128.0 * N * 16 / 4096.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
            }
    return 0;
}
