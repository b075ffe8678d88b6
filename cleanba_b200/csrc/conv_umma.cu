// tcgen05 / TMEM / TMA implementation of the 3x3 SAME convolutions of the IMPALA-ResNet trunk
// (cleanba/cleanba_ppo.py:149-171: nn.Conv(channels, (3,3)) in ConvSequence / ResidualBlock), used for the
// forward conv, for dgrad (same kernel, flipped + transposed weights) and, below, for wgrad.
//
// Implicit GEMM by "flat shifted windows".  Activations live on a zero-padded grid flattened over
// (image, y', x') and split in planes of 8 channels (common.cuh), so
//   * a tile of 128 consecutive flat pixels is the M dimension of a tcgen05.mma (TMEM lane = pixel),
//   * the A operand of filter tap (ky,kx) is the same 128 rows shifted by (ky-1)*Wp + (kx-1) pixels, i.e. it is
//     just a different START ADDRESS inside one contiguous window of 128 + 2*Wp + 2 pixels,
//   * that window is ONE contiguous byte range per channel plane -> one 1-D bulk TMA copy per plane per tile,
//   * a plane [pixel][8 ch] of bf16 is exactly a column of SWIZZLE_NONE 8x16-byte core matrices, both K-major
//     (conv / dgrad: K = channels) and MN-major (wgrad: K = pixels), so no data is ever re-laid out.
//
// Precision (parity bar: 1e-4 against an fp32 reference, with discontinuous relu / max-pool gates downstream):
// every fp32 value is carried as an exact 3-way bf16 split x = hi + mid + lo (24 significant bits).  The six
// significant cross products are obtained with THREE tcgen05.mma per K step by stacking the weight planes along N:
//     D[:, 0:3C] += A_hi  * [W_hi | W_mid | W_lo]        (N = 3C)
//     D[:, 0:2C] += A_mid * [W_hi | W_mid]               (N = 2C)
//     D[:, 0:C ] += A_lo  * [W_hi]                       (N = C)
// and the epilogue adds the three C-column blocks.  The MMAs are bound by the A-operand shared-memory fetch (ncu:
// sm__pipe_tc_cycles_active ~80%), which does not depend on N, so the extra precision costs no tensor-pipe time.
//
// Warp roles (320 threads): warps 0-3 / 4-7 two epilogue groups, one per TMEM accumulator (TMEM lane quadrant = warp % 4),
// which prefetch the residual / gate operands of their next tile before waiting for its accumulator; warp 8 TMA producer;
// warp 9 MMA issuer.
#include "common.cuh"
#include "kernels.h"
#include "umma.cuh"

namespace cb {
using namespace umma;

constexpr int TILE_M = 128;
constexpr int CONV_STAGES = 3;
constexpr int CONV_THREADS = 320;   // warps 0-7: two epilogue groups (one per TMEM accumulator), warp 8: TMA, warp 9: MMA

__host__ __device__ constexpr int conv_steps(int cin_chunks) { return cin_chunks == 1 ? 5 : 9 * (cin_chunks / 2); }
// bytes of the packed weight image: per step [kc(2)][3*cout][8] bf16
__host__ __device__ constexpr int conv_wbytes(int cin_chunks, int cout) { return conv_steps(cin_chunks) * 2 * 3 * cout * 16; }

long long packed_conv_elems(int cin_chunks, int cout) { return (long long)conv_wbytes(cin_chunks, cout) / 2; }

struct ConvSmemLayout {
    int win, plane_bytes, nplanes, stage_bytes, w_bytes, total;
};
__host__ __device__ inline ConvSmemLayout conv_smem_layout(int cin_chunks, int cout, int Wp, int apl = 3) {
    ConvSmemLayout L;
    // frames path: the zero-weight half of the last K step reads one pixel past the 3x3 window -> load it too
    L.win = TILE_M + 2 * Wp + 2 + (cin_chunks == 1 ? 1 : 0);
    L.plane_bytes = L.win * 16;
    L.nplanes = cin_chunks == 1 ? 1 : apl * cin_chunks;      // hi planes, mid planes[, lo planes] (frames: hi only, exact)
    L.stage_bytes = L.nplanes * L.plane_bytes;
    L.w_bytes = conv_wbytes(cin_chunks, cout);
    L.total = 1024 + L.w_bytes + CONV_STAGES * L.stage_bytes;
    return L;
}
int umma_conv_smem_bytes(int cin_chunks, int cout, int Wp) { return conv_smem_layout(cin_chunks, cout, Wp, 3).total; }

// APL = bf16 planes of the A operand: 1 (frames, exact), 2 (gradient tensors, 16 bits: A_hi*[W_hi|W_mid] + A_mid*[W_hi])
// or 3 (forward activations, 24 bits, the three MMAs above).
template <int CIN_CHUNKS, int COUT, int APL>
__global__ void __launch_bounds__(CONV_THREADS) k_conv_umma(ConvArgs a, int ntiles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const ConvSmemLayout L = conv_smem_layout(CIN_CHUNKS, COUT, a.g.Wp, APL);
    // [0,1024): barriers + tmem pointer; then the weight image; then the activation stages
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);          // [CONV_STAGES]
    uint64_t* empty = full + CONV_STAGES;                        // [CONV_STAGES]
    uint64_t* tfull = empty + CONV_STAGES;                       // [2]
    uint64_t* tempty = tfull + 2;                                // [2]
    uint64_t* wbar = tempty + 2;                                 // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wbar + 1);
    uint8_t* wsm = smem + 1024;
    uint8_t* stages = wsm + L.w_bytes;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int STEPS = conv_steps(CIN_CHUNKS);
    constexpr int NBLK = APL == 2 ? 2 : 3;                       // C-column blocks per accumulator
    constexpr int ACC_COLS = NBLK * COUT;
    constexpr uint32_t TMEM_COLS = (2 * ACC_COLS <= 64) ? 64 : ((2 * ACC_COLS <= 128) ? 128 : 256);
    constexpr int NPLANES = CIN_CHUNKS == 1 ? 1 : APL * CIN_CHUNKS;

    if (threadIdx.x == 0) {
        for (int s = 0; s < CONV_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 4); }
        mbar_init(wbar, 1);
        fence_barrier_init();
    }
    if (warp == 9) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 8) {
        // ===================== TMA producer (lanes share the bulk copies of a stage) =====================
        if (lane == 0) {
            mbar_arrive_expect_tx(wbar, (uint32_t)L.w_bytes);
            bulk_g2s(wsm, a.wp, L.w_bytes, wbar);
        }
        int s = 0; uint32_t ph = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            mbar_wait(&empty[s], ph ^ 1);
            if (lane == 0) mbar_arrive_expect_tx(&full[s], (uint32_t)L.stage_bytes);
            const long long q_lo = (long long)tile * TILE_M - a.g.Wp - 1;   // inside the front guard for tile 0
            uint8_t* dst = stages + s * L.stage_bytes;
            if (lane < NPLANES) {
                const int pl = lane / CIN_CHUNKS, j = lane % CIN_CHUNKS;      // pl: 0 hi, 1 mid, 2 lo
                const bf16* src = pl == 0 ? a.in.hi : (pl == 1 ? a.in.mid : a.in.lo);
                bulk_g2s(dst + lane * L.plane_bytes, src + ((long long)j * a.in.plane_px + q_lo) * 8, L.plane_bytes, &full[s]);
            }
            __syncwarp();
            if (++s == CONV_STAGES) { s = 0; ph ^= 1; }
        }
    } else if (warp == 9) {
        // ===================== MMA issuer =====================
        constexpr uint32_t IDESC3 = make_idesc_bf16(TILE_M, 3 * COUT, 0, 0);
        constexpr uint32_t IDESC2 = make_idesc_bf16(TILE_M, 2 * COUT, 0, 0);
        constexpr uint32_t IDESC1 = make_idesc_bf16(TILE_M, COUT, 0, 0);
        mbar_wait(wbar, 0);
        int s = 0; uint32_t ph = 0;
        int acc = 0; uint32_t aph = 0;
        // Per-step descriptor low words relative to the stage base, in 16-byte units (hoisted out of the tile loop; the
        // issue loop below is fully unrolled and costs one integer add per operand per MMA).
        const uint32_t win16 = (uint32_t)L.win;                  // plane stride in 16-byte units
        const uint32_t b_hi = desc_hi(128), a_hi = desc_hi(128);
        const uint32_t b_lo0 = desc_lo(smem_u32(wsm), 3 * COUT * 16);
        uint32_t a_rel[STEPS];
#pragma unroll
        for (int step = 0; step < STEPS; ++step) {
            if (CIN_CHUNKS == 1) {
                // frame stack: 8-channel pixels, two taps per K=16 step (see pack.cu for the matching weights)
                if (step < 3) a_rel[step] = (uint32_t)(step * a.g.Wp) | (1u << 16);                 // LBO = 16 bytes
                else if (step == 3) a_rel[step] = 2u | ((uint32_t)a.g.Wp << 16);                    // LBO = one image row
                else a_rel[step] = (uint32_t)(2 * a.g.Wp + 2) | (1u << 16);
            } else {
                constexpr int HALF = CIN_CHUNKS > 1 ? CIN_CHUNKS / 2 : 1;
                const int tap = step / HALF, pair = step % HALF;
                a_rel[step] = ((uint32_t)(pair * 2) * win16 + (uint32_t)((tap / 3) * a.g.Wp + (tap % 3))) | (win16 << 16);
            }
        }
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            mbar_wait(&tempty[acc], aph ^ 1);
            mbar_wait(&full[s], ph);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t d_tmem = tmem_base + acc * ACC_COLS;
                const uint32_t st16 = (smem_u32(stages + s * L.stage_bytes) >> 4);   // stage base (hi planes), 16-byte units
                const uint32_t mid16 = CIN_CHUNKS * win16, lo16 = 2 * CIN_CHUNKS * win16;
#pragma unroll
                for (int step = 0; step < STEPS; ++step) {
                    const uint32_t a_lo = st16 + a_rel[step];
                    const uint32_t b_lo = b_lo0 + step * (2 * 3 * COUT);
                    if (APL == 2) {
                        mma_bf16_parts(d_tmem, a_lo, a_hi, b_lo, b_hi, IDESC2, step > 0);
                        mma_bf16_parts(d_tmem, a_lo + mid16, a_hi, b_lo, b_hi, IDESC1, 1);
                    } else {
                        mma_bf16_parts(d_tmem, a_lo, a_hi, b_lo, b_hi, IDESC3, step > 0);
                        if (APL == 3) {
                            mma_bf16_parts(d_tmem, a_lo + mid16, a_hi, b_lo, b_hi, IDESC2, 1);
                            mma_bf16_parts(d_tmem, a_lo + lo16, a_hi, b_lo, b_hi, IDESC1, 1);
                        }
                    }
                }
                mma_commit(&empty[s]);      // smem stage reusable once these MMAs have read it
                mma_commit(&tfull[acc]);    // accumulator complete
            }
            __syncwarp();
            if (++s == CONV_STAGES) { s = 0; ph ^= 1; }
            if (++acc == 2) { acc = 0; aph ^= 1; }
        }
    } else {
        // ===================== epilogue: group g = warp / 4 owns accumulator g (every second tile of this CTA) ==========
        const int grp = warp >> 2, quad = warp & 3;
        uint32_t aph = 0;
        for (int tile = blockIdx.x + grp * gridDim.x; tile < ntiles; tile += 2 * gridDim.x) {
            const long long q = (long long)tile * TILE_M + quad * 32 + lane;
            EpiPrefetch<COUT> pre;
            epi_prefetch<COUT>(a.ep, a.g, q, pre);          // global loads in flight while the MMAs of this tile run
            mbar_wait(&tfull[grp], aph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + grp * ACC_COLS;
            float v[COUT];
            {   // sum the column blocks, smallest contributions first
                float t[16];
#pragma unroll
                for (int h = 0; h < COUT / 16; ++h) {
                    tmem_ld16(taddr + (NBLK - 1) * COUT + h * 16, v + h * 16);
#pragma unroll
                    for (int blk = NBLK - 2; blk >= 0; --blk) {
                        tmem_ld16(taddr + blk * COUT + h * 16, t);
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[h * 16 + i] += t[i];
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[grp]);
            epi_finish<COUT>(a.ep, a.g, q, v, pre);
            aph ^= 1;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 9) tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int CIN_CHUNKS, int COUT, int APL>
static int launch_conv_umma_t(const ConvArgs& a, int num_sms, cudaStream_t st) {
    ConvSmemLayout L = conv_smem_layout(CIN_CHUNKS, COUT, a.g.Wp, APL);
    CB_CHECK(L.total <= 227 * 1024, "conv_umma<%d,%d>: %d bytes of shared memory needed", CIN_CHUNKS, COUT, L.total);
    // Opt in to the device maximum once per device: the attribute is per function (contexts on other host threads launch
    // the same instantiation with other window sizes concurrently), and nothing but launches may happen while a
    // CUDA graph is being captured.
    static std::atomic<unsigned> attr_done{0};
    int dev = 0;
    CB_CUDA(cudaGetDevice(&dev));
    if (!(attr_done.load() & (1u << dev))) {
        CB_CUDA(cudaFuncSetAttribute(k_conv_umma<CIN_CHUNKS, COUT, APL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_done.fetch_or(1u << dev);
    }
    int ntiles = (int)((a.g.NP + TILE_M - 1) / TILE_M);
    int ctas_per_sm = L.total <= 110 * 1024 ? 2 : 1;
    int grid = ntiles < num_sms * ctas_per_sm ? ntiles : num_sms * ctas_per_sm;
    k_conv_umma<CIN_CHUNKS, COUT, APL><<<grid, CONV_THREADS, L.total, st>>>(a, ntiles);
    CB_LAUNCH_CHECK();
    return 0;
}

int launch_conv_umma(const ConvArgs& a, int num_sms, cudaStream_t st) {
    CB_CHECK(a.g.Wp + 1 <= GUARD && TILE_M + a.g.Wp + 2 <= GUARD, "conv_umma: guard too small for Wp=%d", a.g.Wp);
    const int apl = a.in.lo ? 3 : (a.in.mid ? 2 : 1);
    if (a.cin_chunks == 1 && a.cout == 16 && apl == 1) return launch_conv_umma_t<1, 16, 1>(a, num_sms, st);
    if (a.cin_chunks == 2 && a.cout == 16 && apl == 3) return launch_conv_umma_t<2, 16, 3>(a, num_sms, st);
    if (a.cin_chunks == 2 && a.cout == 16 && apl == 2) return launch_conv_umma_t<2, 16, 2>(a, num_sms, st);
    if (a.cin_chunks == 2 && a.cout == 32 && apl == 3) return launch_conv_umma_t<2, 32, 3>(a, num_sms, st);
    if (a.cin_chunks == 4 && a.cout == 16 && apl == 2) return launch_conv_umma_t<4, 16, 2>(a, num_sms, st);
    if (a.cin_chunks == 4 && a.cout == 16 && apl == 3) return launch_conv_umma_t<4, 16, 3>(a, num_sms, st);
    if (a.cin_chunks == 4 && a.cout == 32 && apl == 3) return launch_conv_umma_t<4, 32, 3>(a, num_sms, st);
    if (a.cin_chunks == 4 && a.cout == 32 && apl == 2) return launch_conv_umma_t<4, 32, 2>(a, num_sms, st);
    CB_CHECK(false, "conv_umma: unsupported shape cin_chunks=%d cout=%d planes=%d", a.cin_chunks, a.cout, apl);
}

// =================================================================================================
// wgrad on tcgen05:  dW[(ky,kx), ci, co] = sum_q X[q + d(ky,kx)][ci] * G[q][co],  db[co] = sum_q G[q][co].
// GEMM view per 64-pixel block:  D_kx[m, co] += A_kx[m, q] * B[q, co]  with the reduction over pixels (K), where
//   m = ky * C + ci  stacks the three filter ROWS (three bulk copies of the same planes shifted by one image row) and
//   kx is a 16-byte shift of the start address.  A is MN-major (M = channels contiguous), B (= G) is MN-major too,
//   so both operands are again the untouched chunk planes.  One extra M group of bf16 ones yields the bias gradient.
// Precision: same 3-way split as above, the three G planes are stacked along N (they are adjacent in shared memory):
//   D += X_hi * [G_hi|G_mid|G_lo],  D[:, :2C] += X_mid * [G_hi|G_mid],  D[:, :C] += X_lo * [G_hi].
// Three accumulators (kx) x 3*Cout columns stay in TMEM across ALL pixel blocks of a CTA; one partial per CTA is
// written at the end and reduced in a fixed order by k_wgrad_umma_reduce (deterministic).
constexpr int WG_BLOCK = 64;              // pixels per pipeline stage (K block)
constexpr int WG_STAGES = 3;
constexpr int WG_WINX = WG_BLOCK + 2;     // pixels per shifted copy
constexpr int WG_PLANE = WG_WINX * 16;    // bytes
constexpr int WG_THREADS = 192;

struct WgSmemLayout {
    int groups;        // real M groups = 3 * cin_chunks
    int xplanes;       // X planes used: 1 (frames) .. 3
    int gplanes;       // G planes (2 or 3), stacked along N
    int a_bytes;       // one A region (one split plane): (groups + 1) copies (last = ones / zeros)
    int b_chunk;       // bytes of one (plane, chunk) of G
    int b_bytes;       // all of B: 3 planes x cout/8 chunks
    int stage_bytes, stages, total;
};
// Measured (tools/mma_microbench2.cu, profiles/): one thread can issue a tcgen05.mma every ~50 cycles; the tensor pipe
// itself needs ~39 cycles for M=128 and <= 25 for M=64 (N <= 32).  wgrad therefore uses M = 64 whenever the stacked
// (ky, ci) rows fit (<= 64) and runs two CTAs (two issuing threads) per SM whenever shared memory allows.
__host__ __device__ constexpr int wg_mrows(int cin_chunks) { return (3 * cin_chunks + 1) * 8 <= 64 ? 64 : 128; }
__host__ __device__ inline WgSmemLayout wg_smem_layout(int cin_chunks, int cout, int xpl, int gpl) {
    WgSmemLayout L;
    L.groups = 3 * cin_chunks;
    L.xplanes = xpl;
    L.gplanes = gpl;
    L.a_bytes = (L.groups + 1) * WG_PLANE;
    L.b_chunk = WG_BLOCK * 16;
    L.b_bytes = gpl * (cout / 8) * L.b_chunk;
    L.stage_bytes = L.xplanes * L.a_bytes + L.b_bytes;
    L.stages = (1024 + WG_STAGES * L.stage_bytes + (wg_mrows(cin_chunks) / 8) * WG_PLANE <= 112 * 1024) ? WG_STAGES : 2;
    // the MMA reads M/8 groups from the A base: pad so the unused groups stay inside the allocation
    L.total = 1024 + L.stages * L.stage_bytes + (wg_mrows(cin_chunks) / 8) * WG_PLANE;
    return L;
}

// XPL = planes of X used (1 frames, 2 or 3), GPL = planes of G (2 or 3); plane p of X multiplies the first GPL - p planes of G.
template <int CIN_CHUNKS, int COUT, int XPL, int GPL>
__global__ void __launch_bounds__(WG_THREADS) k_wgrad_umma(WgradArgs a, int nblocks, float* __restrict__ partial) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const WgSmemLayout L = wg_smem_layout(CIN_CHUNKS, COUT, XPL, GPL);
    const int NSTAGES = L.stages;
    constexpr int WG_MROWS = wg_mrows(CIN_CHUNKS);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + WG_STAGES;
    uint64_t* done = empty + WG_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
    uint8_t* stages = smem + 1024;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int ACC_COLS = GPL * COUT;
    constexpr uint32_t TMEM_COLS = (3 * ACC_COLS <= 128) ? 128 : ((3 * ACC_COLS <= 256) ? 256 : 512);
    constexpr int GROUPS = 3 * CIN_CHUNKS;
    constexpr int NCOPY_A = XPL * GROUPS, NCOPY_B = GPL * (COUT / 8);

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(done, 1);
        fence_barrier_init();
    }
    // constant "ones" group (hi region) and "zeros" groups (mid / lo regions) of every stage
    for (int s = 0; s < NSTAGES; ++s)
        for (int pl = 0; pl < XPL; ++pl) {
            uint32_t* g = reinterpret_cast<uint32_t*>(stages + s * L.stage_bytes + pl * L.a_bytes + GROUPS * WG_PLANE);
            const uint32_t val = pl == 0 ? 0x3F803F80u : 0u;   // bf16 1.0 x2
            for (int t = threadIdx.x; t < WG_PLANE / 4; t += WG_THREADS) g[t] = val;
        }
    fence_proxy_async();
    if (warp == 5) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int Wp = a.g.Wp;

    if (warp == 4) {
        int s = 0; uint32_t ph = 0;
        const uint32_t tx = (uint32_t)(NCOPY_A * WG_PLANE + NCOPY_B * L.b_chunk);
        for (int blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
            mbar_wait(&empty[s], ph ^ 1);
            if (lane == 0) mbar_arrive_expect_tx(&full[s], tx);
            uint8_t* dst = stages + s * L.stage_bytes;
            const long long q0 = (long long)blk * WG_BLOCK;
            for (int i = lane; i < NCOPY_A + NCOPY_B; i += 32) {
                if (i < NCOPY_A) {
                    const int pl = i / GROUPS, g = i % GROUPS, ky = g / CIN_CHUNKS, j = g % CIN_CHUNKS;
                    const bf16* src = pl == 0 ? a.x.hi : (pl == 1 ? a.x.mid : a.x.lo);
                    const long long q = q0 + (ky - 1) * Wp - 1;
                    bulk_g2s(dst + pl * L.a_bytes + g * WG_PLANE, src + ((long long)j * a.x.plane_px + q) * 8, WG_PLANE, &full[s]);
                } else {
                    const int k = i - NCOPY_A, pl = k / (COUT / 8), j = k % (COUT / 8);
                    const bf16* src = pl == 0 ? a.gy.hi : (pl == 1 ? a.gy.mid : a.gy.lo);
                    bulk_g2s(dst + XPL * L.a_bytes + k * L.b_chunk, src + ((long long)j * a.gy.plane_px + q0) * 8, L.b_chunk, &full[s]);
                }
            }
            __syncwarp();
            if (++s == NSTAGES) { s = 0; ph ^= 1; }
        }
    } else if (warp == 5) {
        constexpr uint32_t IDESC_A = make_idesc_bf16(WG_MROWS, GPL * COUT, 1, 1);                         // X plane 0
        constexpr uint32_t IDESC_B = make_idesc_bf16(WG_MROWS, (GPL > 1 ? GPL - 1 : 1) * COUT, 1, 1);     // X plane 1
        constexpr uint32_t IDESC_C = make_idesc_bf16(WG_MROWS, (GPL > 2 ? GPL - 2 : 1) * COUT, 1, 1);     // X plane 2
        int s = 0; uint32_t ph = 0;
        uint32_t accum = 0;
        for (int blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
            mbar_wait(&full[s], ph);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t a_base = smem_u32(stages + s * L.stage_bytes);
                // K step = 16 pixels = 2 core-matrix groups of 8 pixels (128 bytes each, LBO); M groups WG_PLANE apart,
                // N groups (8 channels of one G plane chunk) b_chunk apart (SBO)
                const uint32_t a_lo0 = desc_lo(a_base, 128), a_hi_w = desc_hi(WG_PLANE);
                const uint32_t b_lo0 = desc_lo(a_base + XPL * L.a_bytes, 128), b_hi_w = desc_hi(L.b_chunk);
                const uint32_t apl16 = (uint32_t)L.a_bytes >> 4;
#pragma unroll
                for (int ks = 0; ks < WG_BLOCK / 16; ++ks) {
                    const uint32_t b_lo = b_lo0 + ks * 16;
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        const uint32_t d_tmem = tmem_base + kx * ACC_COLS;
                        const uint32_t a_lo = a_lo0 + ks * 16 + kx;
                        mma_bf16_parts(d_tmem, a_lo, a_hi_w, b_lo, b_hi_w, IDESC_A, ks == 0 ? accum : 1u);
                        if (XPL > 1 && GPL > 1) mma_bf16_parts(d_tmem, a_lo + apl16, a_hi_w, b_lo, b_hi_w, IDESC_B, 1);
                        if (XPL > 2 && GPL > 2) mma_bf16_parts(d_tmem, a_lo + 2 * apl16, a_hi_w, b_lo, b_hi_w, IDESC_C, 1);
                    }
                }
                accum = 1;
                mma_commit(&empty[s]);
            }
            __syncwarp();
            if (++s == NSTAGES) { s = 0; ph ^= 1; }
        }
        if (lane == 0) mma_commit(done);
        __syncwarp();
    } else {
        mbar_wait(done, 0);
        tc_fence_after();
        // D rows: m = ky*C + ci (and m = 3C .. 3C+7: bias rows).  M = 128: TMEM lane = m.  M = 64: row m lives in lane
        // (m / 16) * 32 + m % 16, i.e. the first 16 lanes of every warp's quadrant (probed, profiles/r01_mma_m64_layout_probe.txt).
        const int m = WG_MROWS == 128 ? warp * 32 + lane : (lane < 16 ? warp * 16 + lane : (1 << 30));
        constexpr int MROWS_USED = GROUPS * 8 + 8;
        static_assert(MROWS_USED <= WG_MROWS, "stacked rows exceed the MMA M");
        float* out = partial + (long long)blockIdx.x * (3 * MROWS_USED * COUT);
        for (int kx = 0; kx < 3; ++kx) {
            float v[COUT];
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + kx * ACC_COLS;
            float t[16];
#pragma unroll
            for (int h = 0; h < COUT / 16; ++h) {
                tmem_ld16(taddr + (GPL - 1) * COUT + h * 16, v + h * 16);
#pragma unroll
                for (int blk = GPL - 2; blk >= 0; --blk) {
                    tmem_ld16(taddr + blk * COUT + h * 16, t);
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[h * 16 + i] += t[i];
                }
            }
            if (m < MROWS_USED) {
#pragma unroll
                for (int c = 0; c < COUT; c += 4)
                    *reinterpret_cast<float4*>(out + ((long long)kx * MROWS_USED + m) * COUT + c) =
                        make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem_base, TMEM_COLS);
}

// dW[(ky*3+kx)][ci][co] = scale * sum_cta partial[cta][kx][ky*Cpad + ci][co];  db[co] = sum_cta partial[cta][1][3*Cpad][co]
__global__ void k_wgrad_umma_reduce(const float* __restrict__ partial, int nctas, int cpad, int cin_real, int cout,
                                    float scale, float* __restrict__ dw, float* __restrict__ db) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int nw = 9 * cin_real * cout;
    if (i >= nw + cout) return;
    const int mrows = 3 * cpad + 8;
    long long src;
    if (i < nw) {
        int co = i % cout, ci = (i / cout) % cin_real, tap = i / (cout * cin_real);
        int ky = tap / 3, kx = tap % 3;
        src = ((long long)kx * mrows + ky * cpad + ci) * cout + co;
    } else {
        src = ((long long)1 * mrows + 3 * cpad) * cout + (i - nw);
    }
    float s = 0.f;
    for (int b = 0; b < nctas; ++b) s += partial[(long long)b * (3 * mrows * cout) + src];
    if (i < nw) dw[i] = s * scale; else db[i - nw] = s;
}

template <int CIN_CHUNKS, int COUT, int XPL, int GPL>
static int launch_wgrad_umma_t(const WgradArgs& a, float* partial, int num_sms, cudaStream_t st) {
    WgSmemLayout L = wg_smem_layout(CIN_CHUNKS, COUT, XPL, GPL);
    CB_CHECK(L.total <= 227 * 1024, "wgrad_umma<%d,%d>: %d bytes of shared memory needed", CIN_CHUNKS, COUT, L.total);
    static std::atomic<unsigned> attr_done{0};
    int dev = 0;
    CB_CUDA(cudaGetDevice(&dev));
    if (!(attr_done.load() & (1u << dev))) {
        CB_CUDA(cudaFuncSetAttribute(k_wgrad_umma<CIN_CHUNKS, COUT, XPL, GPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_done.fetch_or(1u << dev);
    }
    int nblocks = (int)((a.g.NP + WG_BLOCK - 1) / WG_BLOCK);
    int ctas_per_sm = (L.total <= 112 * 1024 && 3 * GPL * COUT <= 256) ? 2 : 1;   // two CTAs must also share the 512 TMEM columns
    int grid = nblocks < num_sms * ctas_per_sm ? nblocks : num_sms * ctas_per_sm;
    k_wgrad_umma<CIN_CHUNKS, COUT, XPL, GPL><<<grid, WG_THREADS, L.total, st>>>(a, nblocks, partial);
    CB_LAUNCH_CHECK();
    int nw = 9 * a.cin_real * a.cout;
    k_wgrad_umma_reduce<<<(nw + a.cout + 255) / 256, 256, 0, st>>>(partial, grid, CIN_CHUNKS * 8, a.cin_real, a.cout, a.scale, a.dw, a.db);
    CB_LAUNCH_CHECK();
    return 0;
}

int launch_wgrad_umma(const WgradArgs& a, float* partial, int num_sms, cudaStream_t st) {
    CB_CHECK(TILE_M + a.g.Wp + 2 <= GUARD, "wgrad_umma: guard too small for Wp=%d", a.g.Wp);
    // gradients are carried with 2 planes (16 bits): X_hi*[G_hi|G_mid] + X_mid*[G_hi]; with 3-plane G the full 6 products
    const int gpl = a.gy.lo ? 3 : 2;
    const int xpl = a.x.mid ? (gpl == 3 && a.x.lo ? 3 : 2) : 1;
    if (gpl == 2) {
        if (a.cin_chunks == 1 && a.cout == 16) return launch_wgrad_umma_t<1, 16, 1, 2>(a, partial, num_sms, st);
        if (a.cin_chunks == 2 && a.cout == 16 && xpl == 2) return launch_wgrad_umma_t<2, 16, 2, 2>(a, partial, num_sms, st);
        if (a.cin_chunks == 2 && a.cout == 32 && xpl == 2) return launch_wgrad_umma_t<2, 32, 2, 2>(a, partial, num_sms, st);
        if (a.cin_chunks == 4 && a.cout == 32 && xpl == 2) return launch_wgrad_umma_t<4, 32, 2, 2>(a, partial, num_sms, st);
    } else {
        if (a.cin_chunks == 1 && a.cout == 16) return launch_wgrad_umma_t<1, 16, 1, 3>(a, partial, num_sms, st);
        if (a.cin_chunks == 2 && a.cout == 16 && xpl == 3) return launch_wgrad_umma_t<2, 16, 3, 3>(a, partial, num_sms, st);
        if (a.cin_chunks == 2 && a.cout == 32 && xpl == 3) return launch_wgrad_umma_t<2, 32, 3, 3>(a, partial, num_sms, st);
        if (a.cin_chunks == 4 && a.cout == 32 && xpl == 3) return launch_wgrad_umma_t<4, 32, 3, 3>(a, partial, num_sms, st);
    }
    CB_CHECK(false, "wgrad_umma: unsupported shape cin_chunks=%d cout=%d", a.cin_chunks, a.cout);
}

}  // namespace cb
