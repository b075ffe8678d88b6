// Probe: does tcgen05.mma kind::f16 accept A = fp16 with B = bf16 in ONE instruction (instruction-descriptor a_format = F16,
// b_format = BF16)?  A[128x16] = fp16(1.5), B[16x16] = bf16(3.0): D must be 16 * 1.5 * 3.0 = 72 everywhere.  (If B were read as
// fp16 the bit pattern 0x4040 would be 2.125 -> 51; if A were read as bf16, 0x3E00 would be 0.125 -> 6.)
// Also checks MN-major operands (the wgrad form) and fp16 x fp16.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I cleanba_b200/csrc tools/mma_mixed_format_probe.cu -o tools/bin/mma_mixed_format_probe
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "umma.cuh"
using namespace cb::umma;

__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int afmt, int bfmt, int a_mn, int b_mn) {
    return (1u << 4) | ((uint32_t)afmt << 7) | ((uint32_t)bfmt << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__global__ void __launch_bounds__(128) k_probe(uint32_t abits, uint32_t bbits, int afmt, int bfmt, int mn, float* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint16_t* A = reinterpret_cast<uint16_t*>(smem);            // 128 x 16 elements, 4 KB
    uint16_t* B = reinterpret_cast<uint16_t*>(smem + 8192);     // 16 x 16 elements
    for (int i = threadIdx.x; i < 2048; i += 128) A[i] = (uint16_t)abits;
    for (int i = threadIdx.x; i < 256; i += 128) B[i] = (uint16_t)bbits;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    fence_proxy_async();
    if (warp == 0) tmem_alloc(&slot, 32);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc(128, 16, afmt, bfmt, mn, mn);
        // K-major: core matrix = 8 rows x 16 B; SBO = 128 (next 8 rows), LBO = bytes between the two K halves
        // MN-major: core matrix = 8 k x 16 B (8 mn); SBO = next 8 mn, LBO = next 8 k.  With constant data any consistent tiling works.
        const uint64_t ad = make_desc(smem_u32(A), mn ? 128 : 2048, mn ? 256 : 128);
        const uint64_t bd = make_desc(smem_u32(B), mn ? 128 : 256, mn ? 256 : 128);
        mma_bf16(tm, ad, bd, idesc, 0);
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    float v[16];
    tmem_ld16(tm + ((uint32_t)(warp * 32) << 16), v);
    if (threadIdx.x == 0) { out[0] = v[0]; out[1] = v[15]; }
    if (threadIdx.x == 127) { out[2] = v[7]; }
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 32);
}

int main(int argc, char** argv) {
    const int only = argc > 1 ? atoi(argv[1]) : -1;   // a failing case poisons the context: run cases in separate processes
    int idx = 0;
    float* d; cudaMalloc(&d, 16);
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024);
    struct { const char* name; uint32_t a, b; int af, bf, mn; float want; } cases[] = {
        {"A bf16(1.5)  x B bf16(3.0), K-major ", 0x3FC0, 0x4040, 1, 1, 0, 72.f},
        {"A fp16(1.5)  x B fp16(3.0), K-major ", 0x3E00, 0x4200, 0, 0, 0, 72.f},
        {"A fp16(1.5)  x B fp16(3.0), MN-major", 0x3E00, 0x4200, 0, 0, 1, 72.f},
        {"A fp16(1.5)  x B bf16(3.0), K-major ", 0x3E00, 0x4040, 0, 1, 0, 72.f},
        {"A bf16(1.5)  x B fp16(3.0), K-major ", 0x3FC0, 0x4200, 1, 0, 0, 72.f},
        {"A fp16(1.5)  x B bf16(3.0), MN-major", 0x3E00, 0x4040, 0, 1, 1, 72.f},
    };
    for (auto& c : cases) {
        if (only >= 0 && idx++ != only) continue;
        cudaMemset(d, 0, 16);
        k_probe<<<1, 128, 32 * 1024>>>(c.a, c.b, c.af, c.bf, c.mn, d);
        cudaError_t e = cudaDeviceSynchronize();
        float h[4] = {0, 0, 0, 0};
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("%s: D[0][0]=%g D[0][15]=%g D[127][7]=%g (want %g) %s %s\n", c.name, h[0], h[1], h[2], c.want,
               (h[0] == c.want && h[1] == c.want && h[2] == c.want) ? "OK" : "MISMATCH", e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
    return 0;
}
