// Probe: tcgen05.mma cta_group::1 with M = 64: (1) which TMEM lanes receive rows 0..63, (2) cycles per MMA vs N.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I cleanba_b200/csrc tools/mma_m64_probe.cu -o tools/bin/mma_m64_probe
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include "umma.cuh"
using namespace cb::umma;

__global__ void __launch_bounds__(128) k_layout(int M, float* out /*[128]*/) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __nv_bfloat16* A = reinterpret_cast<__nv_bfloat16*>(smem);            // K-major: (m%8)*8 + (m/8)*64 + (k%8) + (k/8)*1024 elements
    __nv_bfloat16* B = reinterpret_cast<__nv_bfloat16*>(smem + 8192);
    for (int i = threadIdx.x; i < 4096; i += 128) { A[i] = __float2bfloat16(0.f); B[i] = __float2bfloat16(0.f); }
    __syncthreads();
    if (threadIdx.x < M) A[(threadIdx.x % 8) * 8 + (threadIdx.x / 8) * 64] = __float2bfloat16((float)(threadIdx.x + 1));  // A[m][0] = m+1
    if (threadIdx.x == 0) { B[0] = __float2bfloat16(1.f); mbar_init(&bar, 1); fence_barrier_init(); }                        // B[0][0] = 1
    fence_proxy_async();
    if (warp == 0) tmem_alloc(&slot, 32);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = slot;
    if (warp == 1 && lane == 0) {
        mma_bf16(tm, make_desc(smem_u32(A), 2048, 128), make_desc(smem_u32(B), 2048, 128), make_idesc_bf16(M, 16, 0, 0), 0);
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    float v[16];
    tmem_ld16(tm + ((uint32_t)(warp * 32) << 16), v);
    out[threadIdx.x] = v[0];
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 32);
}

__global__ void __launch_bounds__(128) k_time(int M, int N, int iters, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    fence_proxy_async();
    if (warp == 0) tmem_alloc(&slot, 512);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = slot;
    if (warp == 1 && lane == 0) {
        const uint32_t a = smem_u32(smem), b = a + 48 * 1024;
        const uint32_t idesc = make_idesc_bf16(M, N, 1, 1);      // MN-major like wgrad
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) mma_bf16(tm, make_desc(a + (i & 7) * 32, 128, 1056), make_desc(b, 128, 1024), idesc, i > 0);
        mma_commit(&bar);
        mbar_wait(&bar, 0);
        if (blockIdx.x == 0) out[0] = clock64() - t0;
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 512);
}

int main() {
    float* d; cudaMalloc(&d, 512); long long* c; cudaMalloc(&c, 8);
    cudaFuncSetAttribute(k_layout, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(k_time, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
    for (int M : {128, 64}) {
        k_layout<<<1, 128, 64 * 1024>>>(M, d);
        cudaError_t e = cudaDeviceSynchronize();
        float h[128]; cudaMemcpy(h, d, 512, cudaMemcpyDeviceToHost);
        printf("M=%d (%s): TMEM lane -> row+1 :", M, cudaGetErrorString(e));
        for (int i = 0; i < 128; ++i) printf(" %g", h[i]);
        printf("\n");
    }
    for (int M : {128, 64})
        for (int N : {16, 32, 48, 64, 96}) {
            k_time<<<148, 128, 128 * 1024>>>(M, N, 2000, c);
            cudaError_t e = cudaDeviceSynchronize();
            long long cyc = 0; cudaMemcpy(&cyc, c, 8, cudaMemcpyDeviceToHost);
            printf("M=%3d N=%3d: %.1f cycles/MMA %s\n", M, N, cyc / 2000.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    return 0;
}
