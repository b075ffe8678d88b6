// tcgen05 implementation of the 3872 -> 256 dense layer (cleanba/cleanba_ppo.py:185-188): forward, dX and dW.
//
// The dense layer is a sum over the 121 pixels of the last feature map of [samples x 32] x [32 x 256] products.  To feed
// it to tcgen05.mma with the same SWIZZLE_NONE 8x16-byte core matrices as the convolutions, the last conv also writes
// its (relu'd) output in a *sample-minor* copy  featT[plane][chunk][pixel][sample][8]  (bf16 hi/mid/lo), and the loss
// heads write the gradient of the pre-activation as  dpreT[plane][j / 8][sample][8]  (bf16 hi/mid).  With the sample
// index adjacent to the 8-channel vector, a block of 128 samples is a K-major A operand (forward, dX) and a block of
// 64 samples is an MN-major operand (dW) -- both plain contiguous byte ranges, i.e. 1-D bulk TMA copies again.
//   forward : D[128 samples, 64 outputs]  += featT(p) * W(p)          over 121 pixels (K = 32 per pixel)
//   dX      : D[128 samples, 32 channels]  = dpreT * W(p)^T            per pixel (K = 256)
//   dW      : D[4 pixels x 32 ch, 64 outputs] += featT^T * dpreT       over all samples (K = samples)
// Precision: the same split-bf16 scheme as the convs (forward 3x3 planes with N-stacked weight planes, gradients 2 planes).
#include "common.cuh"
#include "kernels.h"
#include "umma.cuh"

namespace cb {
using namespace umma;

constexpr int DN_PIX = 121, DN_PIXPAD = 124, DN_C = 32, DN_CH = 4;   // pixels, padded pixel slots, channels, chunks
constexpr int DN_THREADS = 192;                                      // warps 0-3 epilogue, 4 TMA, 5 MMA

// ------------------------------------------------------------------------------------------------ weight packing
// fwd image: [nsplit 4][pixel 121][kstep 2][kc 2][plane 3][64 outputs][8]   (B tile of one K=16 step: K-major, N = 192 stacked)
// dx  image: [pixel 121][kstep 16][kc 2][plane 2][32 channels][8]           (B = W(p)^T: N = 64 stacked, K = 256 outputs)
__global__ void k_pack_dense(const float* __restrict__ w, bf16* __restrict__ fwd, bf16* __restrict__ dx) {
    const long long nf = 4LL * DN_PIX * 2 * 2 * 3 * 64 * 8, nd = (long long)DN_PIX * 16 * 2 * 2 * 32 * 8;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nf + nd; e += (long long)gridDim.x * blockDim.x) {
        float v; int plane;
        if (e < nf) {
            int k8 = e % 8, n = (e / 8) % 64; plane = (e / 512) % 3;
            int kc = (e / 1536) % 2, ks = (e / 3072) % 2, p = (e / 6144) % DN_PIX, ns = (int)(e / (6144LL * DN_PIX));
            int c = ks * 16 + kc * 8 + k8;
            v = w[((long long)p * DN_C + c) * HIDDEN + ns * 64 + n];
        } else {
            long long d = e - nf;
            int k8 = d % 8, c = (d / 8) % 32; plane = (d / 256) % 2;
            int kc = (d / 512) % 2, ks = (d / 1024) % 16, p = (int)(d / 16384);
            int j = ks * 16 + kc * 8 + k8;
            v = w[((long long)p * DN_C + c) * HIDDEN + j];
        }
        bf16 h, m, l;
        split_bf16(v, h, m, l);
        bf16 o = plane == 0 ? h : (plane == 1 ? m : l);
        if (e < nf) fwd[e] = o; else dx[e - nf] = o;
    }
}
long long dense_pack_fwd_elems() { return 4LL * DN_PIX * 2 * 2 * 3 * 64 * 8; }
long long dense_pack_dx_elems() { return (long long)DN_PIX * 16 * 2 * 2 * 32 * 8; }
long long dense_featT_elems(int npad) { return (long long)DN_CH * DN_PIXPAD * npad * 8; }   // per plane
long long dense_dpreT_elems(int npad) { return (long long)(HIDDEN / 8) * npad * 8; }          // per plane
int launch_pack_dense(const float* w, bf16* fwd, bf16* dx, cudaStream_t st) {
    k_pack_dense<<<296, 256, 0, st>>>(w, fwd, dx);
    CB_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------------ forward
constexpr int DF_STAGES = 4;
constexpr int DF_A_BYTES = 3 * DN_CH * 128 * 16;     // 24 KB: [plane][chunk][128 samples][8]
constexpr int DF_B_BYTES = 2 * 2 * 3 * 64 * 16;      // 12 KB: [kstep][kc][plane][64][8]
constexpr int DF_STAGE = DF_A_BYTES + DF_B_BYTES;
constexpr int DF_SMEM = 1024 + DF_STAGES * DF_STAGE;

__global__ void __launch_bounds__(DN_THREADS) k_dense_fwd_umma(DenseUmmaArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    griddep_launch();
    griddep_wait();
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + DF_STAGES;
    uint64_t* done = empty + DF_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
    uint8_t* stages = smem + 1024;
    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * 128, ns = blockIdx.y;
    // small batches (the actor's N = 60) split the 121 pixels over gridDim.z CTAs; partial sums are added by k_dense_finish
    const int ppc = (DN_PIX + gridDim.z - 1) / gridDim.z;
    const int p0 = blockIdx.z * ppc, p1 = min(DN_PIX, p0 + ppc);
    if (threadIdx.x == 0) {
        for (int s = 0; s < DF_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(done, 1);
        fence_barrier_init();
    }
    if (warp == 5) tmem_alloc(tmem_slot, 256);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (warp == 4) {
        int s = 0; uint32_t ph = 0;
        for (int p = p0; p < p1; ++p) {
            mbar_wait(&empty[s], ph ^ 1);
            if (lane == 0) mbar_arrive_expect_tx(&full[s], DF_STAGE);
            uint8_t* dst = stages + s * DF_STAGE;
            if (lane < 12) {
                const int pl = lane / DN_CH, j = lane % DN_CH;
                const bf16* src = (pl == 0 ? a.ft_hi : (pl == 1 ? a.ft_mid : a.ft_lo)) + (((long long)j * DN_PIXPAD + p) * a.npad + m0) * 8;
                bulk_g2s(dst + lane * 2048, src, 2048, &full[s]);
            } else if (lane == 12) {
                bulk_g2s(dst + DF_A_BYTES, a.w_fwd + ((long long)ns * DN_PIX + p) * (DF_B_BYTES / 2), DF_B_BYTES, &full[s]);
            }
            __syncwarp();
            if (++s == DF_STAGES) { s = 0; ph ^= 1; }
        }
    } else if (warp == 5) {
        constexpr uint32_t ID3 = make_idesc_bf16(128, 192, 0, 0), ID2 = make_idesc_bf16(128, 128, 0, 0), ID1 = make_idesc_bf16(128, 64, 0, 0);
        const uint32_t a_hi_w = desc_hi(128), b_hi_w = desc_hi(128);
        const uint32_t leader = elect_one();
        int s = 0; uint32_t ph = 0;
        for (int p = p0; p < p1; ++p) {
            mbar_wait(&full[s], ph);
            tc_fence_after();
            {   // warp-uniform issue (umma.cuh)
                const uint32_t sa = smem_u32(stages + s * DF_STAGE);
                const uint32_t a_lo0 = desc_lo(sa, 2048), b_lo0 = desc_lo(sa + DF_A_BYTES, 192 * 16);
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                    const uint32_t al = a_lo0 + ks * (2 * 2048 / 16), bl = b_lo0 + ks * (2 * 192 * 16 / 16);
                    mma_f16_elect(tmem_base, al, a_hi_w, bl, b_hi_w, ID3, (p != p0) || (ks != 0), leader);
                    mma_f16_elect(tmem_base, al + (DN_CH * 2048 / 16), a_hi_w, bl, b_hi_w, ID2, 1, leader);
                    mma_f16_elect(tmem_base, al + (2 * DN_CH * 2048 / 16), a_hi_w, bl, b_hi_w, ID1, 1, leader);
                }
                mma_commit_elect(&empty[s], leader);
            }
            __syncwarp();
            if (++s == DF_STAGES) { s = 0; ph ^= 1; }
        }
        mma_commit_elect(done, leader);
        __syncwarp();
    } else {
        mbar_wait(done, 0);
        tc_fence_after();
        const int b = m0 + warp * 32 + lane;
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            float v[16], t[16];
            tmem_ld16(taddr + 128 + h * 16, v);
            tmem_ld16(taddr + 64 + h * 16, t);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += t[i];
            tmem_ld16(taddr + h * 16, t);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += t[i];
            if (b < a.n && gridDim.z > 1) {
                float* o = a.part + ((long long)blockIdx.z * a.n + b) * HIDDEN + ns * 64 + h * 16;
#pragma unroll
                for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(o + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            } else if (b < a.n) {
                const int j0 = ns * 64 + h * 16;
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                    float4 o;
                    o.x = fmaxf(v[i] + a.bias[j0 + i], 0.f); o.y = fmaxf(v[i + 1] + a.bias[j0 + i + 1], 0.f);
                    o.z = fmaxf(v[i + 2] + a.bias[j0 + i + 2], 0.f); o.w = fmaxf(v[i + 3] + a.bias[j0 + i + 3], 0.f);
                    *reinterpret_cast<float4*>(a.hidden + (long long)b * HIDDEN + j0 + i) = o;
                }
            }
        }
    }
    tc_fence_before(); __syncthreads();
    if (warp == 5) tmem_dealloc(tmem_base, 256);
}

// ------------------------------------------------------------------------------------------------ dX
// One CTA = 128 samples x a range of pixels.  A = dpreT (hi, mid) stays resident in shared memory (128 KB); B = W(p)^T is
// streamed per pixel (2 stages x 32 KB); two TMEM accumulators of 64 columns ([W_hi | W_mid] stacked).
constexpr int DX_A_BYTES = 2 * 32 * 128 * 16;        // 128 KB
constexpr int DX_B_BYTES = 16 * 2 * 2 * 32 * 16;     // 32 KB per pixel
constexpr int DX_SMEM = 1024 + DX_A_BYTES + 2 * DX_B_BYTES;

__global__ void __launch_bounds__(DN_THREADS) k_dense_dx_umma(DenseUmmaArgs a, int pix_per_cta) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);      // [2]
    uint64_t* empty = full + 2;                               // [2]
    uint64_t* tfull = empty + 2;                              // [2]
    uint64_t* tempty = tfull + 2;                             // [2]
    uint64_t* abar = tempty + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(abar + 1);
    uint8_t* sa = smem + 1024;
    uint8_t* sb = sa + DX_A_BYTES;
    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * 128;
    const int p0 = blockIdx.y * pix_per_cta, p1 = min(DN_PIX, p0 + pix_per_cta);
    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 4); }
        mbar_init(abar, 1);
        fence_barrier_init();
    }
    if (warp == 5) tmem_alloc(tmem_slot, 128);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (warp == 4) {
        if (lane == 0) mbar_arrive_expect_tx(abar, DX_A_BYTES);
        for (int i = lane; i < 64; i += 32) {
            const int pl = i / 32, jc = i % 32;
            bulk_g2s(sa + i * 2048, (pl == 0 ? a.dp_hi : a.dp_mid) + ((long long)jc * a.npad + m0) * 8, 2048, abar);
        }
        int s = 0; uint32_t ph = 0;
        for (int p = p0; p < p1; ++p) {
            mbar_wait(&empty[s], ph ^ 1);
            if (lane == 0) {
                mbar_arrive_expect_tx(&full[s], DX_B_BYTES);
                bulk_g2s(sb + s * DX_B_BYTES, a.w_dx + (long long)p * (DX_B_BYTES / 2), DX_B_BYTES, &full[s]);
            }
            __syncwarp();
            if (++s == 2) { s = 0; ph ^= 1; }
        }
    } else if (warp == 5) {
        constexpr uint32_t ID2 = make_idesc_bf16(128, 64, 0, 0), ID1 = make_idesc_bf16(128, 32, 0, 0);
        const uint32_t hw = desc_hi(128);
        const uint32_t leader = elect_one();
        mbar_wait(abar, 0);
        const uint32_t a_lo0 = desc_lo(smem_u32(sa), 2048);
        int s = 0; uint32_t ph = 0; int acc = 0; uint32_t aph = 0;
        for (int p = p0; p < p1; ++p) {
            mbar_wait(&tempty[acc], aph ^ 1);
            mbar_wait(&full[s], ph);
            tc_fence_after();
            {   // warp-uniform issue (umma.cuh)
                const uint32_t b_lo0 = desc_lo(smem_u32(sb + s * DX_B_BYTES), 64 * 16);
                const uint32_t d = tmem_base + acc * 64;
#pragma unroll
                for (int ks = 0; ks < 16; ++ks) {
                    const uint32_t al = a_lo0 + ks * (2 * 2048 / 16), bl = b_lo0 + ks * (2 * 64 * 16 / 16);
                    mma_f16_elect(d, al, hw, bl, hw, ID2, ks != 0, leader);
                    mma_f16_elect(d, al + (32 * 2048 / 16), hw, bl, hw, ID1, 1, leader);
                }
                mma_commit_elect(&empty[s], leader);
                mma_commit_elect(&tfull[acc], leader);
            }
            __syncwarp();
            if (++s == 2) { s = 0; ph ^= 1; }
            if (++acc == 2) { acc = 0; aph ^= 1; }
        }
    } else {
        int acc = 0; uint32_t aph = 0;
        const int b = m0 + warp * 32 + lane;
        const float gS = *a.gscale;                 // loss scale carried by the trunk's gradient tensors
        for (int p = p0; p < p1; ++p) {
            mbar_wait(&tfull[acc], aph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + acc * 64;
            float v[32], t[16];
            tmem_ld16(taddr + 32, v); tmem_ld16(taddr + 48, v + 16);
            tmem_ld16(taddr, t);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += t[i];
            tmem_ld16(taddr + 16, t);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[16 + i] += t[i];
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
            if (b < a.n) {
                const int h = p / 11, w = p - h * 11;
                const long long q = (long long)b * 144 + (h + 1) * 12 + (w + 1);      // 11x11 map, shared borders (common.cuh)
#pragma unroll
                for (int jc = 0; jc < DN_CH; ++jc) {
                    // relu gate of the forward feature (hi plane of the sample-minor copy), loss scale, carrier split
                    uint4 m = *reinterpret_cast<const uint4*>(a.ft_hi + (((long long)jc * DN_PIXPAD + p) * a.npad + b) * 8);
                    float mf[8], o[8];
                    unpack8(m, mf);
#pragma unroll
                    for (int e = 0; e < 8; ++e) o[e] = mf[e] > 0.f ? v[jc * 8 + e] * gS : 0.f;
                    store_planes8(a.out, ((long long)jc * a.out.plane_px + q) * 8, o);
                }
            }
            if (++acc == 2) { acc = 0; aph ^= 1; }
        }
    }
    tc_fence_before(); __syncthreads();
    if (warp == 5) tmem_dealloc(tmem_base, 128);
}

// ------------------------------------------------------------------------------------------------ dW
// One CTA = 4 pixels (M = 4 x 32 channels) x 64 outputs; reduction over all samples in blocks of 64.
constexpr int DW_BLOCK = 64, DW_STAGES = 3;
constexpr int DW_A_BYTES = 2 * 16 * DW_BLOCK * 16;   // X hi|mid: [plane][pixel 4][chunk 4][64 samples][8] = 32 KB
constexpr int DW_B_BYTES = 2 * 8 * DW_BLOCK * 16;    // dpre hi|mid: [plane][jchunk 8][64 samples][8] = 16 KB
constexpr int DW_STAGE = DW_A_BYTES + DW_B_BYTES;
constexpr int DW_SMEM = 1024 + DW_STAGES * DW_STAGE;

__global__ void __launch_bounds__(DN_THREADS) k_dense_dw_umma(DenseUmmaArgs a, float* __restrict__ dw) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + DW_STAGES;
    uint64_t* done = empty + DW_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
    uint8_t* stages = smem + 1024;
    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    const int pg = blockIdx.x * 4, j0 = blockIdx.y * 64;
    const int nblk = a.npad / DW_BLOCK;
    if (threadIdx.x == 0) {
        for (int s = 0; s < DW_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(done, 1);
        fence_barrier_init();
    }
    if (warp == 5) tmem_alloc(tmem_slot, 128);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (warp == 4) {
        int s = 0; uint32_t ph = 0;
        for (int blk = 0; blk < nblk; ++blk) {
            mbar_wait(&empty[s], ph ^ 1);
            if (lane == 0) mbar_arrive_expect_tx(&full[s], DW_STAGE);
            uint8_t* dst = stages + s * DW_STAGE;
            const long long b0 = (long long)blk * DW_BLOCK;
            for (int i = lane; i < 48; i += 32) {
                if (i < 32) {          // X: plane pl, pixel pi, chunk jc  (pixel slots up to DN_PIXPAD exist and are zero beyond 120)
                    const int pl = i / 16, pi = (i % 16) / 4, jc = i % 4;
                    const bf16* src = (pl == 0 ? a.ft_hi : a.ft_mid) + (((long long)jc * DN_PIXPAD + pg + pi) * a.npad + b0) * 8;
                    bulk_g2s(dst + i * 1024, src, 1024, &full[s]);
                } else {               // dpre: plane pl, output chunk jc
                    const int k = i - 32, pl = k / 8, jc = k % 8;
                    const bf16* src = (pl == 0 ? a.dp_hi : a.dp_mid) + ((long long)(j0 / 8 + jc) * a.npad + b0) * 8;
                    bulk_g2s(dst + DW_A_BYTES + k * 1024, src, 1024, &full[s]);
                }
            }
            __syncwarp();
            if (++s == DW_STAGES) { s = 0; ph ^= 1; }
        }
    } else if (warp == 5) {
        constexpr uint32_t ID2 = make_idesc_bf16(128, 128, 1, 1), ID1 = make_idesc_bf16(128, 64, 1, 1);
        const uint32_t hw = desc_hi(1024);         // M / N groups (8 channels / 8 outputs) are 1 KB apart
        const uint32_t leader = elect_one();
        int s = 0; uint32_t ph = 0;
        for (int blk = 0; blk < nblk; ++blk) {
            mbar_wait(&full[s], ph);
            tc_fence_after();
            {   // warp-uniform issue (umma.cuh)
                const uint32_t base = smem_u32(stages + s * DW_STAGE);
                const uint32_t a_lo0 = desc_lo(base, 128), b_lo0 = desc_lo(base + DW_A_BYTES, 128);
#pragma unroll
                for (int ks = 0; ks < DW_BLOCK / 16; ++ks) {
                    const uint32_t al = a_lo0 + ks * 16, bl = b_lo0 + ks * 16;
                    mma_f16_elect(tmem_base, al, hw, bl, hw, ID2, (blk | ks) != 0, leader);
                    mma_f16_elect(tmem_base, al + (16 * 1024 / 16), hw, bl, hw, ID1, 1, leader);
                }
                mma_commit_elect(&empty[s], leader);
            }
            __syncwarp();
            if (++s == DW_STAGES) { s = 0; ph ^= 1; }
        }
        mma_commit_elect(done, leader);
        __syncwarp();
    } else {
        mbar_wait(done, 0);
        tc_fence_after();
        const int m = warp * 32 + lane;            // row = pixel-in-group * 32 + channel
        const int p = pg + m / 32, c = m % 32;
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            float v[16], t[16];
            tmem_ld16(taddr + 64 + h * 16, v);
            tmem_ld16(taddr + h * 16, t);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += t[i];
            if (p < DN_PIX) {
                float* o = dw + ((long long)p * DN_C + c) * HIDDEN + j0 + h * 16;
#pragma unroll
                for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(o + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            }
        }
    }
    tc_fence_before(); __syncthreads();
    if (warp == 5) tmem_dealloc(tmem_base, 128);
}

// column sums of dpre (bias gradient), two deterministic stages
__global__ void k_colsum_partial(const float* __restrict__ d, int n, float* __restrict__ part) {
    const int j = threadIdx.x;        // 256 columns
    float s = 0.f;
    for (int b = blockIdx.x; b < n; b += gridDim.x) s += d[(long long)b * HIDDEN + j];
    part[blockIdx.x * HIDDEN + j] = s;
}
__global__ void k_colsum_final(const float* __restrict__ part, int nparts, float* __restrict__ out) {
    const int j = threadIdx.x;
    float s = 0.f;
    for (int i = 0; i < nparts; ++i) s += part[i * HIDDEN + j];
    out[j] = s;
}

// dpre [n][256] fp32 -> dpreT[plane][j / 8][sample][8] bf16 (hi, mid); rows n .. npad-1 are written as zeros (dW reduces over them)
__global__ void k_dpre_transpose(const float* __restrict__ dpre, int n, int npad, bf16* __restrict__ hi, bf16* __restrict__ mid) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= npad * (HIDDEN / 8)) return;
    const int jc = t % (HIDDEN / 8), b = t / (HIDDEN / 8);
    float v[8];
    if (b < n) {
        const float4* p = reinterpret_cast<const float4*>(dpre + (long long)b * HIDDEN + jc * 8);
        float4 x = p[0], y = p[1];
        v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w; v[4] = y.x; v[5] = y.y; v[6] = y.z; v[7] = y.w;
    } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = 0.f;
    }
    store_bf16_split8(hi, mid, nullptr, ((long long)jc * npad + b) * 8, v);
}
int launch_dpre_transpose(const float* dpre, int n, int npad, bf16* dp_hi, bf16* dp_mid, cudaStream_t st) {
    int total = npad * (HIDDEN / 8);
    k_dpre_transpose<<<(total + 255) / 256, 256, 0, st>>>(dpre, n, npad, dp_hi, dp_mid);
    CB_LAUNCH_CHECK();
    return 0;
}

template <typename K>
static int set_smem(K kernel, int bytes) {
    CB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    return 0;
}

int dense_umma_init() {   // once per device, outside any graph capture
    if (set_smem(k_dense_fwd_umma, DF_SMEM)) return -1;
    if (set_smem(k_dense_dx_umma, DX_SMEM)) return -1;
    if (set_smem(k_dense_dw_umma, DW_SMEM)) return -1;
    return 0;
}

__global__ void k_dense_finish_umma(const float* __restrict__ part, int nsplit, int n, const float* __restrict__ bias,
                                    float* __restrict__ hidden) {
    griddep_launch();
    griddep_wait();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)n * HIDDEN) return;
    float s = 0.f;
    for (int z = 0; z < nsplit; ++z) s += part[(long long)z * n * HIDDEN + i];
    hidden[i] = fmaxf(s + bias[i % HIDDEN], 0.f);
}

int launch_dense_fwd_umma(const DenseUmmaArgs& a, cudaStream_t st) {
    const int tiles = a.npad / 128;
    const int psplit = tiles >= 16 ? 1 : (tiles >= 4 ? 4 : 11);
    dim3 grid(tiles, 4, psplit);
    launch_pdl(k_dense_fwd_umma, grid, dim3(DN_THREADS), (size_t)DF_SMEM, st, a);
    CB_LAUNCH_CHECK();
    if (psplit > 1) {
        long long tot = (long long)a.n * HIDDEN;
        launch_pdl(k_dense_finish_umma, dim3((unsigned)((tot + 255) / 256)), dim3(256), 0, st, (const float*)a.part, psplit, a.n, a.bias, a.hidden);
        CB_LAUNCH_CHECK();
    }
    return 0;
}

int launch_dense_bwd_umma(const DenseUmmaArgs& a, const float* dpre, float* dw, float* db, float* scratch, cudaStream_t st) {
    // dW, db
    dim3 gw((DN_PIX + 3) / 4, 4);
    k_dense_dw_umma<<<gw, DN_THREADS, DW_SMEM, st>>>(a, dw);
    CB_LAUNCH_CHECK();
    k_colsum_partial<<<64, HIDDEN, 0, st>>>(dpre, a.n, scratch);
    CB_LAUNCH_CHECK();
    k_colsum_final<<<1, HIDDEN, 0, st>>>(scratch, 64, db);
    CB_LAUNCH_CHECK();
    // dX (gradient tensors keep exact zeros on the padding ring and up to the 128-pixel tile boundary)
    const size_t npr = (size_t)((a.NP + 127) / 128 * 128);
    for (int c = 0; c < DN_CH; ++c) {
        CB_CUDA(cudaMemsetAsync(a.out.hi + (long long)c * a.out.plane_px * 8, 0, npr * 8 * sizeof(f16), st));
        CB_CUDA(cudaMemsetAsync(a.out.mid + (long long)c * a.out.plane_px * 8, 0, npr * 8 * sizeof(f16), st));
    }
    const int tiles = a.npad / 128;
    int psplit = tiles >= 74 ? 2 : (tiles >= 30 ? 4 : 11);
    int ppc = (DN_PIX + psplit - 1) / psplit;
    dim3 gx(tiles, (DN_PIX + ppc - 1) / ppc);
    k_dense_dx_umma<<<gx, DN_THREADS, DX_SMEM, st>>>(a, ppc);
    CB_LAUNCH_CHECK();
    return 0;
}

}  // namespace cb
