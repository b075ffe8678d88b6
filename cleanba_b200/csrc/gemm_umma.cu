// Generic tcgen05 GEMMs on fp16x2 carrier "row planes", used by the Nature-CNN trunk (nature.cu;
// cleanba/legacy_scripts/cleanba_ppo_envpool_impala_atari_wrapper_naturecnn.py:143-178): the strided VALID convolutions
// (8x8 s4, 4x4 s2, 3x3 s1) and the 3136 -> 512 dense layer are GEMMs over im2col matrices.
//
// Row planes: a matrix X[R, K] is stored as plane[K / 8][Rpad][8] (Rpad = R rounded up to 128, rows >= R are zero), once per
// carrier plane (hi, mid; the frame im2col has hi only: exact).  A 128-row x 16-K tile is two contiguous 2 KB blocks, which is
// exactly a K-major SWIZZLE_NONE tcgen05 operand (rows = M) and, read the other way, an MN-major operand (rows = K of the MMA):
// every operand of the three GEMMs below is fetched with 1-D bulk TMA copies and never re-laid out.
//
//   k_gemm_umma<NB>        C[R, N]  = A[R, K] * W[K, N]      forward (W = layer weights) and dgrad (W = layer weights transposed);
//                          persistent CTAs over (row tile, N block of NB columns), K streamed in blocks of 64 together with the
//                          packed weight tiles, two MMAs per K step (A_hi*[W_hi|W_mid], A_mid*W_hi), two TMEM accumulators so the
//                          epilogue of a tile overlaps the MMAs of the next one.
//   k_gemm_wgrad_umma<NB>  dW[K, N] = A[R, K]^T * G[R, N]    reduction over rows; M = 16 K-chunks of A (MN-major), N = [G_hi|G_mid];
//                          accumulators live in TMEM across all row blocks of a CTA, per-CTA partials reduced in a fixed order.
#include "common.cuh"
#include "kernels.h"
#include "umma.cuh"

namespace cb {
using namespace umma;

constexpr int GM_TILE = 128;
constexpr int GM_KB = 64;                 // K per pipeline stage (4 K steps)
constexpr int GM_SLOT = GM_TILE * 16;     // bytes of one (chunk, 128 rows) block
constexpr int GM_THREADS = 320;           // warps 0-7: two epilogue groups, warp 8: TMA, warp 9: MMA
constexpr int GM_MAXST = 6;

struct GemmSmem { int a_bytes, b_bytes, stage_bytes, stages, total; };
__host__ __device__ inline GemmSmem gemm_smem(int apl, int NB) {
    GemmSmem L;
    L.a_bytes = apl * (GM_KB / 8) * GM_SLOT;
    L.b_bytes = (GM_KB / 16) * 2 * 2 * NB * 16;
    L.stage_bytes = L.a_bytes + L.b_bytes;
    int st = (226 * 1024 - 1024) / L.stage_bytes;
    L.stages = st > GM_MAXST ? GM_MAXST : st;
    L.total = 1024 + L.stages * L.stage_bytes;
    return L;
}

template <int NB, int APL>
__global__ void __launch_bounds__(GM_THREADS) k_gemm_umma(GemmArgs a, int ntiles, int nblocks) {
    extern __shared__ __align__(1024) uint8_t smem[];
    griddep_launch();
    const GemmSmem L = gemm_smem(APL, NB);
    const int NSTAGES = L.stages;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);          // [GM_MAXST]
    uint64_t* empty = full + GM_MAXST;                           // [GM_MAXST]
    uint64_t* tfull = empty + GM_MAXST;                          // [2]
    uint64_t* tempty = tfull + 2;                                // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    uint8_t* stages = smem + 1024;
    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    constexpr int ACC_COLS = 2 * NB;                             // block 0: hi*hi, block 1: (hi*mid + mid*hi) * 2^11
    constexpr uint32_t TMEM_COLS = 2 * ACC_COLS <= 128 ? 128 : (2 * ACC_COLS <= 256 ? 256 : 512);
    const int nkb = a.K / GM_KB;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 4); }
        fence_barrier_init();
    }
    if (warp == 9) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    griddep_wait();

    if (warp == 8) {
        // ===================== TMA producer =====================
        int s = 0; uint32_t ph = 0;
        const int ncopy = APL * (GM_KB / 8);
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int m = tile / nblocks, nb = tile - m * nblocks;
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(&empty[s], ph ^ 1);
                if (lane == 0) mbar_arrive_expect_tx(&full[s], (uint32_t)L.stage_bytes);
                __syncwarp();
                uint8_t* dst = stages + s * L.stage_bytes;
                if (lane < ncopy) {
                    const int pl = lane / (GM_KB / 8), c = lane % (GM_KB / 8);
                    const f16* src = (pl == 0 ? a.a.hi : a.a.mid) + ((long long)(kb * (GM_KB / 8) + c) * a.a.rpad + (long long)m * GM_TILE) * 8;
                    bulk_g2s(dst + lane * GM_SLOT, src, GM_SLOT, &full[s]);
                } else if (lane == 31) {
                    bulk_g2s(dst + L.a_bytes, a.wp + ((long long)nb * nkb + kb) * (L.b_bytes / 2), L.b_bytes, &full[s]);
                }
                __syncwarp();
                if (++s == NSTAGES) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 9) {
        // ===================== MMA issuer =====================
        constexpr uint32_t IDESC2 = make_idesc_f16(GM_TILE, 2 * NB, 0, 0);
        constexpr uint32_t IDESC1 = make_idesc_f16(GM_TILE, NB, 0, 0);
        const uint32_t hw = desc_hi(128);
        const uint32_t leader = elect_one();
        int s = 0; uint32_t ph = 0;
        int acc = 0; uint32_t aph = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            mbar_wait(&tempty[acc], aph ^ 1);
            const uint32_t d_tmem = tmem_base + acc * ACC_COLS;
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(&full[s], ph);
                tc_fence_after();
                {   // warp-uniform issue (umma.cuh)
                    const uint32_t sa = smem_u32(stages + s * L.stage_bytes);
                    const uint32_t a_lo0 = desc_lo(sa, GM_SLOT);                      // K chunks GM_SLOT apart (LBO), 8-row groups 128 B (SBO)
                    const uint32_t b_lo0 = desc_lo(sa + L.a_bytes, 2 * NB * 16);
#pragma unroll
                    for (int ks = 0; ks < GM_KB / 16; ++ks) {
                        const uint32_t al = a_lo0 + ks * (2 * GM_SLOT / 16), bl = b_lo0 + ks * (2 * 2 * NB);
                        mma_f16_elect(d_tmem, al, hw, bl, hw, IDESC2, (kb | ks) != 0, leader);
                        if (APL == 2) mma_f16_elect(d_tmem + NB, al + ((GM_KB / 8) * GM_SLOT / 16), hw, bl, hw, IDESC1, 1, leader);
                    }
                    mma_commit_elect(&empty[s], leader);
                    if (kb == nkb - 1) mma_commit_elect(&tfull[acc], leader);
                }
                __syncwarp();
                if (++s == NSTAGES) { s = 0; ph ^= 1; }
            }
            if (++acc == 2) { acc = 0; aph ^= 1; }
        }
    } else {
        // ===================== epilogue: group g = warp / 4 owns accumulator g (every second tile of this CTA) ==========
        const int grp = warp >> 2, quad = warp & 3;
        uint32_t aph = 0;
        for (int tile = blockIdx.x + grp * gridDim.x; tile < ntiles; tile += 2 * gridDim.x) {
            const int m = tile / nblocks, nb = tile - m * nblocks;
            const long long row = (long long)m * GM_TILE + quad * 32 + lane;
            const bool valid = row < a.R;
            mbar_wait(&tfull[grp], aph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + grp * ACC_COLS;
#pragma unroll 1
            for (int h = 0; h < NB / 16; ++h) {
                float v[16], t[16];
                tmem_ld16(taddr + NB + h * 16, t);
                tmem_ld16(taddr + h * 16, v);
                const int col0 = nb * NB + h * 16;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    float x = fmaf(t[i], MID_INV, v[i]) * a.acc_scale;
                    if (a.bias) x += a.bias[col0 + i];
                    if (a.relu) x = fmaxf(x, 0.f);
                    v[i] = valid ? x : 0.f;
                }
                if (a.out_hi) {
                    Planes o;
                    o.hi = a.out_hi; o.mid = a.out_mid; o.plane_px = a.out_rpad;
                    if (col0 < a.N) store_planes8(o, ((long long)(col0 / 8) * a.out_rpad + row) * 8, v);
                    if (col0 + 8 < a.N) store_planes8(o, ((long long)(col0 / 8 + 1) * a.out_rpad + row) * 8, v + 8);
                }
                if (a.out_f32 && valid) {
#pragma unroll
                    for (int i = 0; i < 16; i += 4)
                        if (col0 + i < a.N)
                            *reinterpret_cast<float4*>(a.out_f32 + row * a.out_ld + col0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[grp]);
            aph ^= 1;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 9) tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int NB, int APL>
static int launch_gemm_t(const GemmArgs& a, int num_sms, cudaStream_t st) {
    const GemmSmem L = gemm_smem(APL, NB);
    CB_CHECK(L.stages >= 2, "gemm_umma<%d>: stage of %d bytes does not fit twice", NB, L.stage_bytes);
    static std::atomic<unsigned> attr_done{0};
    int dev = 0;
    CB_CUDA(cudaGetDevice(&dev));
    if (!(attr_done.load() & (1u << dev))) {
        CB_CUDA(cudaFuncSetAttribute(k_gemm_umma<NB, APL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_done.fetch_or(1u << dev);
    }
    const int nblocks = (a.N + NB - 1) / NB;
    const long long ntiles = (a.Rpad / GM_TILE) * nblocks;
    const int grid = (int)(ntiles < num_sms ? ntiles : num_sms);
    launch_pdl(k_gemm_umma<NB, APL>, dim3(grid), dim3(GM_THREADS), (size_t)L.total, st, a, (int)ntiles, nblocks);
    CB_LAUNCH_CHECK();
    return 0;
}

int launch_gemm_umma(const GemmArgs& a, int NB, int num_sms, cudaStream_t st) {
    CB_CHECK(a.K % GM_KB == 0 && a.Rpad % GM_TILE == 0 && a.a.rpad >= a.Rpad, "gemm_umma: K=%d must be a multiple of %d, rows padded to %d", a.K, GM_KB, GM_TILE);
    CB_CHECK(a.N % 16 == 0 || a.N % 8 == 0, "gemm_umma: N=%d must be a multiple of 8", a.N);
    if (NB == 32 && !a.a.mid) return launch_gemm_t<32, 1>(a, num_sms, st);
    if (NB == 32) return launch_gemm_t<32, 2>(a, num_sms, st);
    if (NB == 64 && a.a.mid) return launch_gemm_t<64, 2>(a, num_sms, st);
    if (NB == 128 && a.a.mid) return launch_gemm_t<128, 2>(a, num_sms, st);
    CB_CHECK(false, "gemm_umma: unsupported N block %d", NB);
}

// ------------------------------------------------------------------------------------------------ weight images
// image[nb][ks][kc(2)][2*NB][8] fp16: B tile of K step ks for the NB output columns of block nb, rows [0,NB) = hi, [NB,2NB) = mid.
//   transpose = 0: B[k][n] = w[k * Nl + n]   (forward:  K = Kl, N = Nl)
//   transpose = 1: B[k][n] = w[n * Nl + k]   (dgrad:    K = Nl, N = Kl)
__global__ void k_pack_gemm(const float* __restrict__ w, int Kl, int Nl, int transpose, int NB, f16* __restrict__ out, long long total) {
    const int Kg = transpose ? Nl : Kl, Ng = transpose ? Kl : Nl;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int k8 = (int)(e % 8);
        const int n2 = (int)((e / 8) % (2 * NB));
        const int kc = (int)((e / (16LL * NB)) % 2);
        const int ks = (int)((e / (32LL * NB)) % (Kg / 16));
        const int nb = (int)(e / (32LL * NB * (Kg / 16)));
        const int plane = n2 / NB, n = nb * NB + n2 % NB, k = ks * 16 + kc * 8 + k8;
        float v = 0.f;
        if (n < Ng) v = transpose ? w[(long long)n * Nl + k] : w[(long long)k * Nl + n];
        f16 h, m;
        split_f16(v, h, m);
        out[e] = plane == 0 ? h : m;
    }
}
long long gemm_pack_elems(int Kl, int Nl, int transpose, int NB) {
    const int Kg = transpose ? Nl : Kl, Ng = transpose ? Kl : Nl;
    return (long long)((Ng + NB - 1) / NB) * (Kg / 16) * 2 * 2 * NB * 8;
}
int launch_pack_gemm(const float* w, int Kl, int Nl, int transpose, int NB, f16* out, cudaStream_t st) {
    const long long total = gemm_pack_elems(Kl, Nl, transpose, NB);
    k_pack_gemm<<<296, 256, 0, st>>>(w, Kl, Nl, transpose, NB, out, total);
    CB_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------------ wgrad
constexpr int GW_THREADS = 192;           // warps 0-3 epilogue, 4 TMA, 5 MMA
constexpr int GW_MCH = 16;                // K chunks (of the layer) per M tile = 128 rows of dW

struct GwSmem { int a_bytes, b_bytes, stage_bytes, stages, total; };
__host__ __device__ inline GwSmem gw_smem(int apl, int NB) {
    GwSmem L;
    L.a_bytes = apl * GW_MCH * GM_SLOT;
    L.b_bytes = 2 * (NB / 8) * GM_SLOT;
    L.stage_bytes = L.a_bytes + L.b_bytes;
    int st = (226 * 1024 - 1024) / L.stage_bytes;
    L.stages = st > 4 ? 4 : st;
    L.total = 1024 + L.stages * L.stage_bytes;
    return L;
}

template <int NB, int APL>
__global__ void __launch_bounds__(GW_THREADS) k_gemm_wgrad_umma(GemmWgradArgs a, int nmt, int nnb, int nsplit, float* __restrict__ partial) {
    extern __shared__ __align__(1024) uint8_t smem[];
    griddep_launch();
    const GwSmem L = gw_smem(APL, NB);
    const int NSTAGES = L.stages;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + 4;
    uint64_t* done = empty + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
    uint8_t* stages = smem + 1024;
    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    constexpr int COLS = 3 * NB;                                 // D1 = A_hi * [G_hi | G_mid] (2 NB), D2 = A_mid * G_hi (NB)
    constexpr uint32_t TMEM_COLS = COLS <= 128 ? 128 : 256;
    const int item = blockIdx.x;                                 // (mt, nb, sp), sp fastest
    const int sp = item % nsplit, nb = (item / nsplit) % nnb, mt = item / (nsplit * nnb);
    const int nrb = (int)(a.Rpad / GM_TILE);
    const int kch = a.K / 8;
    const int mch = min(GW_MCH, kch - mt * GW_MCH);              // valid K chunks of this M tile

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(done, 1);
        fence_barrier_init();
    }
    // M groups past the layer's K (last M tile) are never loaded: keep them finite
    for (int s = 0; s < NSTAGES; ++s)
        for (int pl = 0; pl < APL; ++pl)
            for (int c = mch; c < GW_MCH; ++c) {
                uint4* z = reinterpret_cast<uint4*>(stages + s * L.stage_bytes + (pl * GW_MCH + c) * GM_SLOT);
                for (int t = threadIdx.x; t < GM_SLOT / 16; t += GW_THREADS) z[t] = make_uint4(0, 0, 0, 0);
            }
    fence_proxy_async();
    if (warp == 5) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    griddep_wait();

    if (warp == 4) {
        int s = 0; uint32_t ph = 0;
        const int ncopy_a = APL * mch, ncopy_b = 2 * (NB / 8);
        const uint32_t tx = (uint32_t)((ncopy_a + ncopy_b) * GM_SLOT);
        for (int rb = sp; rb < nrb; rb += nsplit) {
            mbar_wait(&empty[s], ph ^ 1);
            if (lane == 0) mbar_arrive_expect_tx(&full[s], tx);
            __syncwarp();
            uint8_t* dst = stages + s * L.stage_bytes;
            const long long r0 = (long long)rb * GM_TILE;
            for (int i = lane; i < ncopy_a + ncopy_b; i += 32) {
                if (i < ncopy_a) {
                    const int pl = i / mch, c = i % mch;
                    const f16* src = (pl == 0 ? a.a.hi : a.a.mid) + ((long long)(mt * GW_MCH + c) * a.a.rpad + r0) * 8;
                    bulk_g2s(dst + (pl * GW_MCH + c) * GM_SLOT, src, GM_SLOT, &full[s]);
                } else {
                    const int k = i - ncopy_a, pl = k / (NB / 8), c = k % (NB / 8);
                    const f16* src = (pl == 0 ? a.g.hi : a.g.mid) + ((long long)(nb * (NB / 8) + c) * a.g.rpad + r0) * 8;
                    bulk_g2s(dst + L.a_bytes + k * GM_SLOT, src, GM_SLOT, &full[s]);
                }
            }
            __syncwarp();
            if (++s == NSTAGES) { s = 0; ph ^= 1; }
        }
    } else if (warp == 5) {
        constexpr uint32_t ID2 = make_idesc_f16(GM_TILE, 2 * NB, 1, 1), ID1 = make_idesc_f16(GM_TILE, NB, 1, 1);
        const uint32_t hw = desc_hi(GM_SLOT);                    // M / N groups (8 channels) are one slot apart
        const uint32_t leader = elect_one();
        int s = 0; uint32_t ph = 0;
        uint32_t accum = 0;
        for (int rb = sp; rb < nrb; rb += nsplit) {
            mbar_wait(&full[s], ph);
            tc_fence_after();
            {   // warp-uniform issue (umma.cuh)
                const uint32_t base = smem_u32(stages + s * L.stage_bytes);
                const uint32_t a_lo0 = desc_lo(base, 128), b_lo0 = desc_lo(base + L.a_bytes, 128);   // K step of 8 rows = 128 B (LBO)
#pragma unroll
                for (int ks = 0; ks < GM_TILE / 16; ++ks) {
                    mma_f16_elect(tmem_base, a_lo0 + ks * 16, hw, b_lo0 + ks * 16, hw, ID2, ks == 0 ? accum : 1u, leader);
                    if (APL == 2)
                        mma_f16_elect(tmem_base + 2 * NB, a_lo0 + (GW_MCH * GM_SLOT / 16) + ks * 16, hw, b_lo0 + ks * 16, hw, ID1, ks == 0 ? accum : 1u, leader);
                }
                accum = 1;
                mma_commit_elect(&empty[s], leader);
            }
            __syncwarp();
            if (++s == NSTAGES) { s = 0; ph ^= 1; }
        }
        mma_commit_elect(done, leader);
        __syncwarp();
    } else if (warp < 4) {
        mbar_wait(done, 0);
        tc_fence_after();
        const int m = warp * 32 + lane;
        const int k = mt * GM_TILE + m;
        const bool any = sp < nrb;                               // a split without row blocks contributes zeros
        const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
        float* out = partial + ((long long)sp * a.K + k) * a.N + nb * NB;
#pragma unroll 1
        for (int h = 0; h < NB / 16; ++h) {
            float v[16], t[16];
            tmem_ld16(lane_addr + h * 16, v);                    // A_hi * G_hi
            tmem_ld16(lane_addr + NB + h * 16, t);               // A_hi * G_mid (carries 2^11)
            if (APL == 2) {
                float u[16];
                tmem_ld16(lane_addr + 2 * NB + h * 16, u);       // A_mid * G_hi (carries 2^11)
#pragma unroll
                for (int i = 0; i < 16; ++i) t[i] += u[i];
            }
            if (k < a.K) {
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                    float4 o = make_float4(fmaf(t[i], MID_INV, v[i]), fmaf(t[i + 1], MID_INV, v[i + 1]), fmaf(t[i + 2], MID_INV, v[i + 2]),
                                           fmaf(t[i + 3], MID_INV, v[i + 3]));
                    if (!any) o = make_float4(0.f, 0.f, 0.f, 0.f);
                    *reinterpret_cast<float4*>(out + h * 16 + i) = o;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem_base, TMEM_COLS);
}

// dW[i] = scale * inv * sum_sp partial[sp][i]  (fixed order)
__global__ void k_gemm_wgrad_reduce(const float* __restrict__ partial, int nsplit, long long count, float scale,
                                    const float* __restrict__ inv_scale, float* __restrict__ dw) {
    griddep_launch();
    griddep_wait();
    const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i >= count) return;
    float4 s = *reinterpret_cast<const float4*>(partial + i);
    for (int sp = 1; sp < nsplit; ++sp) {
        const float4 t = *reinterpret_cast<const float4*>(partial + (long long)sp * count + i);
        s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
    }
    const float f = scale * (inv_scale ? *inv_scale : 1.f);
    *reinterpret_cast<float4*>(dw + i) = make_float4(s.x * f, s.y * f, s.z * f, s.w * f);
}

template <int NB, int APL>
static int launch_gemm_wgrad_t(const GemmWgradArgs& a, float* partial, long long partial_cap, int num_sms, cudaStream_t st) {
    const GwSmem L = gw_smem(APL, NB);
    CB_CHECK(L.stages >= 2, "gemm_wgrad_umma<%d>: stage of %d bytes does not fit twice", NB, L.stage_bytes);
    static std::atomic<unsigned> attr_done{0};
    int dev = 0;
    CB_CUDA(cudaGetDevice(&dev));
    if (!(attr_done.load() & (1u << dev))) {
        CB_CUDA(cudaFuncSetAttribute(k_gemm_wgrad_umma<NB, APL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_done.fetch_or(1u << dev);
    }
    const int nmt = (a.K / 8 + GW_MCH - 1) / GW_MCH, nnb = a.N / NB;
    const int nrb = (int)(a.Rpad / GM_TILE);
    int nsplit = (2 * num_sms + nmt * nnb - 1) / (nmt * nnb);
    if (nsplit > nrb) nsplit = nrb;
    if (nsplit < 1) nsplit = 1;
    const long long count = (long long)a.K * a.N;
    while (nsplit > 1 && (long long)nsplit * count > partial_cap) --nsplit;
    CB_CHECK((long long)nsplit * count <= partial_cap, "gemm_wgrad: partial buffer too small (%lld floats needed)", (long long)nsplit * count);
    launch_pdl(k_gemm_wgrad_umma<NB, APL>, dim3(nmt * nnb * nsplit), dim3(GW_THREADS), (size_t)L.total, st, a, nmt, nnb, nsplit, partial);
    CB_LAUNCH_CHECK();
    launch_pdl(k_gemm_wgrad_reduce, dim3((unsigned)((count / 4 + 255) / 256)), dim3(256), 0, st, (const float*)partial, nsplit, count, a.scale,
               a.inv_scale, a.dw);
    CB_LAUNCH_CHECK();
    return 0;
}

int launch_gemm_wgrad_umma(const GemmWgradArgs& a, float* partial, long long partial_cap, int num_sms, cudaStream_t st) {
    CB_CHECK(a.K % 8 == 0 && a.Rpad % GM_TILE == 0 && a.g.mid, "gemm_wgrad_umma: bad shapes (K=%d, Rpad=%lld)", a.K, a.Rpad);
    if (a.N % 64 == 0 && a.a.mid) return launch_gemm_wgrad_t<64, 2>(a, partial, partial_cap, num_sms, st);
    if (a.N % 32 == 0 && !a.a.mid) return launch_gemm_wgrad_t<32, 1>(a, partial, partial_cap, num_sms, st);
    if (a.N % 32 == 0) return launch_gemm_wgrad_t<32, 2>(a, partial, partial_cap, num_sms, st);
    CB_CHECK(false, "gemm_wgrad_umma: N=%d must be a multiple of 32", a.N);
}

}  // namespace cb
