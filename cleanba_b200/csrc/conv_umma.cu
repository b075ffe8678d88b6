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
//   * a plane [pixel][8 ch] of fp16 is exactly a column of SWIZZLE_NONE 8x16-byte core matrices, both K-major
//     (conv / dgrad: K = channels) and MN-major (wgrad: K = pixels), so no data is ever re-laid out.
//
// Precision (parity bar: 1e-4 against an fp32 reference, with discontinuous relu / max-pool gates downstream):
// every fp32 value is carried as the fp16x2 pair x = hi + mid * 2^-11 (22 significant bits, common.cuh); fp16 x fp16 products
// are exact in the fp32 accumulator.  The three significant cross products cost TWO tcgen05.mma per K step by stacking the
// weight planes along N:
//     D[:, 0:2C] += A_hi  * [W_hi | W_mid]       (N = 2C)     column block 0: hi*hi,  block 1: hi*mid
//     D[:, C:2C] += A_mid * [W_hi]               (N = C)      block 1 += mid*hi
// and the epilogue evaluates  block0 + 2^-11 * block1  (the mid*mid term, 2^-22 relative, is dropped).  Round 1 carried three
// bf16 planes and issued three MMAs per K step; the MMAs are bound by the A-operand shared-memory fetch (32 cycles per MMA
// regardless of N), so two planes are 2/3 of the tensor-pipe time as well as 2/3 of the bytes.
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
constexpr int CONV_MAXST = 8;        // stages of the input-window ring: as many as fit (memory latency, not bandwidth,
                                     // bounds these kernels: profiles/r01_v5_ncu_sweep_learner_mb3840.txt)
constexpr int CONV_THREADS = 320;   // warps 0-7: two epilogue groups (one per TMEM accumulator), warp 8: TMA, warp 9: MMA

__host__ __device__ constexpr int conv_steps(int cin_chunks) { return cin_chunks == 1 ? 5 : 9 * (cin_chunks / 2); }
// bytes of the packed weight image: per step [kc(2)][2*cout][8] fp16 (hi | mid rows, pack.cu)
__host__ __device__ constexpr int conv_wbytes(int cin_chunks, int cout) { return conv_steps(cin_chunks) * 2 * 2 * cout * 16; }

long long packed_conv_elems(int cin_chunks, int cout) { return (long long)conv_wbytes(cin_chunks, cout) / 2; }

struct ConvSmemLayout {
    int win, plane_bytes, nplanes, stage_bytes, w_bytes, stages, ctas_per_sm, total;
};
__host__ __device__ inline ConvSmemLayout conv_smem_layout(int cin_chunks, int cout, int Wp, int max_ctas = 2) {
    const int apl = 2;
    ConvSmemLayout L;
    // frames path: the zero-weight half of the last K step reads one pixel past the 3x3 window -> load it too
    L.win = TILE_M + 2 * Wp + 2 + (cin_chunks == 1 ? 1 : 0);
    L.plane_bytes = L.win * 16;
    L.nplanes = cin_chunks == 1 ? 1 : apl * cin_chunks;      // hi planes, mid planes (frames: hi only, exact)
    L.stage_bytes = L.nplanes * L.plane_bytes;
    L.w_bytes = conv_wbytes(cin_chunks, cout);
    // two CTAs per SM (two MMA-issuing threads) when three stages fit in half an SM, else one CTA with a deeper ring
    L.ctas_per_sm = (max_ctas >= 2 && 1024 + L.w_bytes + 3 * L.stage_bytes <= 112 * 1024) ? 2 : 1;
    const int budget = (L.ctas_per_sm == 2 ? 112 : 226) * 1024 - 1024 - L.w_bytes;
    L.stages = budget / L.stage_bytes;
    if (L.stages > CONV_MAXST) L.stages = CONV_MAXST;
    L.total = 1024 + L.w_bytes + L.stages * L.stage_bytes;
    return L;
}
int umma_conv_smem_bytes(int cin_chunks, int cout, int Wp) { return conv_smem_layout(cin_chunks, cout, Wp).total; }

// One kernel for the forward conv and for dgrad (activations and gradients use the same carrier); CIN_CHUNKS == 1 is the
// frame stack (one exact plane, ONE MMA per step: A_hi * [W_hi | W_mid]).
// Registers: the 32-channel instantiations need ~130 registers per thread, so only ONE of the two CTAs per SM that the shared-
// memory layout provides for is resident at a time (2 x 320 x 130 > 65,536; the grid of 2 x SMs runs as two waves).  Capping them
// at 96 (__launch_bounds__(320, 2), 88 bytes of spills) makes both resident and was measured SLOWER at 21x21 (forward 0.160 vs
// 0.150 ms, dgrad 0.144 vs 0.140 per 3840 frames) and mixed at 11x11: profiles/r02_v9_conv_epilogue_ab.txt.  Not kept.
// Three CTAs per SM for the 16-channel instantiations (registers capped at 64, 3-4 stage rings, 24 epilogue warps) were measured
// slower as well (forward 0.925 vs 0.843 ms per four launches): these kernels already move 5.9-6.0 TB/s.
template <int CIN_CHUNKS, int COUT, int MAXCTAS = 2>
__global__ void __launch_bounds__(CONV_THREADS, 1) k_conv_umma(ConvArgs a, int ntiles) {
    constexpr int APL = CIN_CHUNKS == 1 ? 1 : 2;
    extern __shared__ __align__(1024) uint8_t smem[];
    griddep_launch();
    const ConvSmemLayout L = conv_smem_layout(CIN_CHUNKS, COUT, a.g.Wp, MAXCTAS);
    // [0,1024): barriers + tmem pointer; then the weight image; then the activation stages
    const int NSTAGES = L.stages;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);          // [CONV_MAXST]
    uint64_t* empty = full + CONV_MAXST;                         // [CONV_MAXST]
    uint64_t* tfull = empty + CONV_MAXST;                        // [2]
    uint64_t* tempty = tfull + 2;                                // [2]
    uint64_t* wbar = tempty + 2;                                 // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wbar + 1);
    uint8_t* wsm = smem + 1024;
    uint8_t* stages = wsm + L.w_bytes;

    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    constexpr int STEPS = conv_steps(CIN_CHUNKS);
    constexpr int ACC_COLS = 2 * COUT;                           // column block 0: hi*hi, block 1: (hi*mid + mid*hi) * 2^11
    constexpr uint32_t TMEM_COLS = (2 * ACC_COLS <= 64) ? 64 : ((2 * ACC_COLS <= 128) ? 128 : 256);
    constexpr int NPLANES = CIN_CHUNKS == 1 ? 1 : APL * CIN_CHUNKS;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 4); }
        mbar_init(wbar, 1);
        fence_barrier_init();
    }
    if (warp == 9) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // the packed weights do not depend on the preceding kernels: their load is issued before the dependency wait
    if (warp == 8 && lane == 0) {
        mbar_arrive_expect_tx(wbar, (uint32_t)L.w_bytes);
        bulk_g2s(wsm, a.wp, L.w_bytes, wbar);
    }
    griddep_wait();

    if (warp == 8) {
        // ===================== TMA producer (lanes share the bulk copies of a stage) =====================
        int s = 0; uint32_t ph = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            mbar_wait(&empty[s], ph ^ 1);
            if (lane == 0) mbar_arrive_expect_tx(&full[s], (uint32_t)L.stage_bytes);
            const long long q_lo = (long long)tile * TILE_M - a.g.Wp - 1;   // inside the front guard for tile 0
            uint8_t* dst = stages + s * L.stage_bytes;
            if (lane < NPLANES) {
                const int pl = lane / CIN_CHUNKS, j = lane % CIN_CHUNKS;      // pl: 0 hi, 1 mid
                const f16* src = pl == 0 ? a.in.hi : a.in.mid;
                bulk_g2s(dst + lane * L.plane_bytes, src + ((long long)j * a.in.plane_px + q_lo) * 8, L.plane_bytes, &full[s]);
            }
            __syncwarp();
            if (++s == NSTAGES) { s = 0; ph ^= 1; }
        }
    } else if (warp == 9) {
        // ===================== MMA issuer =====================
        constexpr uint32_t IDESC2 = make_idesc_f16(TILE_M, 2 * COUT, 0, 0);
        constexpr uint32_t IDESC1 = make_idesc_f16(TILE_M, COUT, 0, 0);
        const uint32_t leader = elect_one();
        mbar_wait(wbar, 0);
        int s = 0; uint32_t ph = 0;
        int acc = 0; uint32_t aph = 0;
        // Per-step descriptor low words relative to the stage base, in 16-byte units (hoisted out of the tile loop; the
        // issue loop below is fully unrolled and costs one integer add per operand per MMA).
        const uint32_t win16 = (uint32_t)L.win;                  // plane stride in 16-byte units
        const uint32_t b_hi = desc_hi(128), a_hi = desc_hi(128);
        constexpr int WPL = 2;
        const uint32_t b_lo0 = desc_lo(smem_u32(wsm), WPL * COUT * 16);
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
            {   // the whole warp runs the issue code converged (uniform datapath); only the elected lane issues (umma.cuh)
                const uint32_t d_tmem = tmem_base + acc * ACC_COLS;
                const uint32_t st16 = (smem_u32(stages + s * L.stage_bytes) >> 4);   // stage base (hi planes), 16-byte units
                const uint32_t mid16 = CIN_CHUNKS * win16;
#pragma unroll
                for (int step = 0; step < STEPS; ++step) {
                    const uint32_t a_lo = st16 + a_rel[step];
                    const uint32_t b_lo = b_lo0 + step * (2 * WPL * COUT);
                    mma_f16_elect(d_tmem, a_lo, a_hi, b_lo, b_hi, IDESC2, step > 0, leader);                 // A_hi * [W_hi | W_mid]
                    if (APL == 2) mma_f16_elect(d_tmem + COUT, a_lo + mid16, a_hi, b_lo, b_hi, IDESC1, 1, leader);   // block 1 += A_mid * W_hi
                }
                mma_commit_elect(&empty[s], leader);      // smem stage reusable once these MMAs have read it
                mma_commit_elect(&tfull[acc], leader);    // accumulator complete
            }
            __syncwarp();
            if (++s == NSTAGES) { s = 0; ph ^= 1; }
            if (++acc == 2) { acc = 0; aph ^= 1; }
        }
    } else {
        // ===================== epilogue: group g = warp / 4 owns accumulator g (every second tile of this CTA) ==========
        const int grp = warp >> 2, quad = warp & 3;
        // the layer's bias once per CTA in shared memory (the 1 KB header has room) instead of COUT global loads per tile
        float* bias_sm = reinterpret_cast<float*>(smem + 512);
        if (threadIdx.x < COUT) bias_sm[threadIdx.x] = a.ep.bias ? a.ep.bias[threadIdx.x] : 0.f;
        asm volatile("bar.sync 1, 256;" ::: "memory");      // the eight epilogue warps only
        // Position of this thread's pixel inside its image, carried from tile to tile: r = q mod P advances by a constant, and
        // the row is one float multiply (exact: (r + 0.5) / Wp is at least 0.5 / Wp away from an integer and r < 2^20).
        const uint32_t P = (uint32_t)a.g.P, Wp = (uint32_t)a.g.Wp;
        const float inv_wp = 1.f / (float)a.g.Wp;
        uint32_t r = (uint32_t)(((long long)(blockIdx.x + grp * gridDim.x) * TILE_M + quad * 32 + lane) % a.g.P);
        const uint32_t dr = (uint32_t)(((long long)2 * gridDim.x * TILE_M) % a.g.P);
        uint32_t aph = 0;
        for (int tile = blockIdx.x + grp * gridDim.x; tile < ntiles; tile += 2 * gridDim.x) {
            const long long q = (long long)tile * TILE_M + quad * 32 + lane;
            const uint32_t y = __float2uint_rz(((float)r + 0.5f) * inv_wp), x = r - y * Wp;
            const bool tail = q >= a.g.NP;
            const bool in = !tail & (y >= 1u) & (y <= (uint32_t)a.g.H) & (x >= 1u) & (x <= (uint32_t)a.g.W);
            r += dr;
            if (r >= P) r -= P;
            EpiPrefetch<COUT> pre;
            epi_prefetch_known<COUT>(a.ep, q, in, tail, pre);   // global loads in flight while the MMAs of this tile run
            mbar_wait(&tfull[grp], aph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + grp * ACC_COLS;
            float v[COUT];
            {   // block0 + 2^-11 * block1
                float t[16];
#pragma unroll
                for (int h = 0; h < COUT / 16; ++h) {
                    tmem_ld16(taddr + COUT + h * 16, t);
                    tmem_ld16(taddr + h * 16, v + h * 16);
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[h * 16 + i] = fmaf(t[i], MID_INV, v[h * 16 + i]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[grp]);
            epi_finish<COUT, true>(a.ep, a.g, q, v, pre, bias_sm);
            aph ^= 1;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 9) tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int CIN_CHUNKS, int COUT, int MAXCTAS = 2>
static int launch_conv_umma_t(const ConvArgs& a, int num_sms, cudaStream_t st) {
    ConvSmemLayout L = conv_smem_layout(CIN_CHUNKS, COUT, a.g.Wp, MAXCTAS);
    CB_CHECK(L.stages >= 2 && L.total <= 227 * 1024, "conv_umma<%d,%d>: %d bytes of shared memory needed", CIN_CHUNKS, COUT, L.total);
    CB_CHECK(a.g.P < (1 << 20), "conv_umma: image of %d padded pixels (the epilogue's row arithmetic is exact below 2^20)", a.g.P);
    // Opt in to the device maximum once per device: the attribute is per function (contexts on other host threads launch
    // the same instantiation with other window sizes concurrently), and nothing but launches may happen while a
    // CUDA graph is being captured.
    static std::atomic<unsigned> attr_done{0};
    int dev = 0;
    CB_CUDA(cudaGetDevice(&dev));
    if (!(attr_done.load() & (1u << dev))) {
        CB_CUDA(cudaFuncSetAttribute(k_conv_umma<CIN_CHUNKS, COUT, MAXCTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_done.fetch_or(1u << dev);
    }
    int ntiles = (int)((a.g.NP + TILE_M - 1) / TILE_M);
    int grid = ntiles < num_sms * L.ctas_per_sm ? ntiles : num_sms * L.ctas_per_sm;
    // (halving the grid of the 16-channel kernels for small batches was measured too: actor step 0.126 -> 0.136 ms, not kept)
    launch_pdl(k_conv_umma<CIN_CHUNKS, COUT, MAXCTAS>, dim3(grid), dim3(CONV_THREADS), (size_t)L.total, st, a, ntiles);
    CB_LAUNCH_CHECK();
    return 0;
}

int launch_conv_umma(const ConvArgs& a, int num_sms, cudaStream_t st) {
    CB_CHECK(a.g.Wp + 1 <= GUARD && TILE_M + a.g.Wp + 2 <= GUARD, "conv_umma: guard too small for Wp=%d", a.g.Wp);
    CB_CHECK((a.cin_chunks == 1) == (a.in.mid == nullptr), "conv_umma: single-plane inputs are the frame stack only");
    if (a.cin_chunks == 1 && a.cout == 16) return launch_conv_umma_t<1, 16>(a, num_sms, st);
    if (a.cin_chunks == 2 && a.cout == 16) return launch_conv_umma_t<2, 16>(a, num_sms, st);
    if (a.cin_chunks == 2 && a.cout == 32) return launch_conv_umma_t<2, 32>(a, num_sms, st);
    if (a.cin_chunks == 4 && a.cout == 16) return launch_conv_umma_t<4, 16>(a, num_sms, st);
    // Cin = Cout = 32: ~130 registers per thread allow only one resident CTA per SM, so the layout is asked for ONE CTA per SM
    // with the whole shared memory as an 8-stage ring (grid = SMs) instead of two half-size CTAs that ran as two waves: forward /
    // dgrad at 21x21 -4..5 %, at 11x11 -8 %, and the n = 60 actor step 0.141 -> 0.123 ms (half as many CTAs load the 37 KB weight
    // image); CLEANBA_CONV32_CTAS=2 restores the old shape (profiles/r02_v9_conv_epilogue_ab.txt)
    static const int c32 = [] { const char* e = getenv("CLEANBA_CONV32_CTAS"); return e ? atoi(e) : 1; }();
    if (a.cin_chunks == 4 && a.cout == 32) return c32 == 1 ? launch_conv_umma_t<4, 32, 1>(a, num_sms, st) : launch_conv_umma_t<4, 32>(a, num_sms, st);
    CB_CHECK(false, "conv_umma: unsupported shape cin_chunks=%d cout=%d", a.cin_chunks, a.cout);
}

// =================================================================================================
// First ConvSequence, forward: the frame conv FUSED with its max-pool (cleanba_ppo.py:167-168).
// The 84x84x16 fp32 conv output (473 KB per frame: the largest tensor of the network) is never written to HBM.  A CTA
// processes BANDS of 3 pooled rows = 7 consecutive conv rows = 602 consecutive flat pixels = 5 MMA tiles of 128 (the
// flat-shifted-window trick needs consecutive pixels, and whole rows ARE consecutive).  The epilogue warps move each
// accumulator (x 1/255 + bias) into a shared-memory band buffer instead of HBM; after the 5th tile they pool the band
// (3x3 stride 2, SAME: windows (2i..2i+2, 2j..2j+2) clipped at 84, first maximum in row-major order like k_pool_fwd) and
// write what the separate pool kernel wrote: the fp32 stream, the relu'd 3-way split planes and the arg-max bytes of the
// pooled 44x44 padded grid (borders zero).  The TMA producer and the MMA issuer run ahead of the pooling by the two
// TMEM accumulators.  HBM traffic per frame: 118 KB in + 340 KB out instead of 1.4 MB through the two kernels.
constexpr int C0_HP = 85, C0_P = C0_HP * C0_HP, C0_HO = 42, C0_WPO = 43, C0_PO = C0_WPO * C0_WPO;   // shared borders (common.cuh)
constexpr int C0_BAND_K = 3;                                   // pooled rows per band
constexpr int C0_BANDS = C0_HO / C0_BAND_K;                    // 14 bands per frame
constexpr int C0_BAND_TILES = 5;                               // ceil(7 * 85 / 128)
constexpr int C0_BAND_PX = C0_BAND_TILES * TILE_M;             // 640 pixels in the band buffer
constexpr int C0_COUT = 16;
constexpr int C0_STAGES = 8;
constexpr int C0_WIN = TILE_M + 2 * C0_HP + 3;                 // 301 pixels per input window (see conv_smem_layout)
constexpr int C0_STAGE_BYTES = C0_WIN * 16;
constexpr int C0_W_BYTES = conv_wbytes(1, C0_COUT);
constexpr int C0_BAND_BYTES = C0_BAND_PX * C0_COUT * 4;
constexpr int C0_SMEM = 1024 + C0_W_BYTES + C0_STAGES * C0_STAGE_BYTES + C0_BAND_BYTES;
static_assert(C0_HO % C0_BAND_K == 0, "bands must tile the pooled rows");
static_assert(2 * C0_SMEM <= 226 * 1024, "two CTAs per SM");

struct Conv0PoolArgs {
    int n;                    // frames
    const f16* x_hi;          // unpacked frames: chunk plane [n * 86 * 86][8] (4 real channels), flat pixel 0
    const f16* wp;            // packed forward weight image of the frame conv
    const float* bias;        // [16]
    float acc_scale;          // 1/255
    Planes out;               // pooled planes (raw: the residual input of the first block)
    Planes out_r;             // rectified pooled planes (the first block's conv operand)
    uint8_t* amax;            // arg-max bytes [2][n * 44 * 44][8] or null (actor contexts)
    uint8_t* bits;            // relu gate bits of out_r [n * 44 * 44][2] or null (common.cuh ConvEpilogue::bits_out)
};

// 16-byte chunk c (0..3) of band pixel pb.  A pixel PAIR is one 128-byte row (all 32 banks); the 3-bit index (pixel parity, c)
// is XOR-swizzled with the pair index, a Latin square: the epilogue's stores (8 lanes = 8 consecutive pixels, same c) and the
// pool's loads (8 lanes = 8 consecutive pooled columns = pixel stride 2, same parity and c) both touch 8 distinct 16-byte
// bank groups, i.e. both are conflict free.
__device__ __forceinline__ float4* c0_band_ptr(uint8_t* band, int pb, int c) {
    const int pair = pb >> 1;
    return reinterpret_cast<float4*>(band + (pair << 7) + (((((pb & 1) << 2) | c) ^ (pair & 7)) << 4));
}

__global__ void __launch_bounds__(CONV_THREADS) k_conv0_pool_umma(Conv0PoolArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    griddep_launch();
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);          // [C0_STAGES]
    uint64_t* empty = full + C0_STAGES;                          // [C0_STAGES]
    uint64_t* tfull = empty + C0_STAGES;                         // [C0_BAND_TILES]: one TMEM accumulator per tile of a band
    uint64_t* tempty = tfull + C0_BAND_TILES;                    // [C0_BAND_TILES]
    uint64_t* wbar = tempty + C0_BAND_TILES;                     // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wbar + 1);
    uint8_t* wsm = smem + 1024;
    uint8_t* stages = wsm + C0_W_BYTES;
    uint8_t* band = stages + C0_STAGES * C0_STAGE_BYTES;

    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    constexpr int STEPS = conv_steps(1);
    constexpr int ACC_COLS = 2 * C0_COUT;
    constexpr uint32_t TMEM_COLS = 256;                          // 5 accumulators x 32 columns: the MMAs of the NEXT band run
    const int nbands = a.n * C0_BANDS;                           // while this band is being pooled

    if (threadIdx.x == 0) {
        for (int s = 0; s < C0_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int s = 0; s < C0_BAND_TILES; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 4); }
        mbar_init(wbar, 1);
        fence_barrier_init();
    }
    if (warp == 9) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (warp == 8 && lane == 0) {
        mbar_arrive_expect_tx(wbar, (uint32_t)C0_W_BYTES);
        bulk_g2s(wsm, a.wp, C0_W_BYTES, wbar);
    }
    griddep_wait();

    if (warp == 8) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (int bnd = blockIdx.x; bnd < nbands; bnd += gridDim.x) {
                const int img = bnd / C0_BANDS, b = bnd - img * C0_BANDS;
                const long long qb = (long long)img * C0_P + (2 * C0_BAND_K * b + 1) * C0_HP;     // first conv row of the band
                for (int t = 0; t < C0_BAND_TILES; ++t) {
                    mbar_wait(&empty[s], ph ^ 1);
                    mbar_arrive_expect_tx(&full[s], (uint32_t)C0_STAGE_BYTES);
                    const long long q_lo = qb + t * TILE_M - C0_HP - 1;
                    bulk_g2s(stages + s * C0_STAGE_BYTES, a.x_hi + q_lo * 8, C0_STAGE_BYTES, &full[s]);
                    if (++s == C0_STAGES) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 9) {
        // ===================== MMA issuer (frames: one exact fp16 plane, two taps per K = 16 step) =====================
        constexpr uint32_t IDESC2 = make_idesc_f16(TILE_M, 2 * C0_COUT, 0, 0);
        const uint32_t leader = elect_one();
        mbar_wait(wbar, 0);
        int s = 0; uint32_t ph = 0;
        uint32_t aph = 0;                                        // accumulator phase: flips once per band
        const uint32_t b_hi = desc_hi(128), a_hi = desc_hi(128);
        const uint32_t b_lo0 = desc_lo(smem_u32(wsm), 2 * C0_COUT * 16);
        uint32_t a_rel[STEPS];
#pragma unroll
        for (int step = 0; step < STEPS; ++step) {
            if (step < 3) a_rel[step] = (uint32_t)(step * C0_HP) | (1u << 16);
            else if (step == 3) a_rel[step] = 2u | ((uint32_t)C0_HP << 16);
            else a_rel[step] = (uint32_t)(2 * C0_HP + 2) | (1u << 16);
        }
        for (int bnd = blockIdx.x; bnd < nbands; bnd += gridDim.x) {
            for (int t = 0; t < C0_BAND_TILES; ++t) {
                mbar_wait(&tempty[t], aph ^ 1);
                mbar_wait(&full[s], ph);
                tc_fence_after();
                {   // warp-uniform issue (umma.cuh)
                    const uint32_t d_tmem = tmem_base + t * ACC_COLS;
                    const uint32_t st16 = (smem_u32(stages + s * C0_STAGE_BYTES) >> 4);
#pragma unroll
                    for (int step = 0; step < STEPS; ++step)
                        mma_f16_elect(d_tmem, st16 + a_rel[step], a_hi, b_lo0 + step * (2 * 2 * C0_COUT), b_hi, IDESC2, step > 0, leader);
                    mma_commit_elect(&empty[s], leader);
                    mma_commit_elect(&tfull[t], leader);
                }
                __syncwarp();
                if (++s == C0_STAGES) { s = 0; ph ^= 1; }
            }
            aph ^= 1;
        }
    } else {
        // ===================== epilogue warps 0-7: accumulators -> band buffer, then pool the band =====================
        const int grp = warp >> 2, quad = warp & 3;
        const int etid = threadIdx.x;                            // 0..255
        uint32_t aph = 0;                                        // accumulator phase: flips once per band
        float bias[C0_COUT];
#pragma unroll
        for (int e = 0; e < C0_COUT; ++e) bias[e] = a.bias[e];
        // stream + relu'd planes + arg-max bytes of one (pooled pixel, chunk)
        auto emit = [&](int img, int ypo, int xp, int jc, const float* v, const int* am) {
            const long long qo = (long long)img * C0_PO + ypo * C0_WPO + xp;
            const long long so = ((long long)jc * a.n * C0_PO + qo) * 8;
            store_planes8(a.out, ((long long)jc * a.out.plane_px + qo) * 8, v);
            float rl[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) rl[e] = fmaxf(v[e], 0.f);
            store_planes8(a.out_r, ((long long)jc * a.out_r.plane_px + qo) * 8, rl);
            if (a.amax) {
                uint2 pk;
                pk.x = am[0] | (am[1] << 8) | (am[2] << 16) | (am[3] << 24);
                pk.y = am[4] | (am[5] << 8) | (am[6] << 16) | (am[7] << 24);
                *reinterpret_cast<uint2*>(a.amax + so) = pk;
            }
            if (a.bits) {
                uint32_t bb = 0;
#pragma unroll
                for (int e = 0; e < 8; ++e) bb |= (v[e] > 0.f ? 1u : 0u) << e;
                a.bits[qo * 2 + jc] = (uint8_t)bb;
            }
        };
        auto emit_zero = [&](int img, int ypo, int xp, int jc) {
            float v[8];
            int am[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) { v[e] = 0.f; am[e] = 15; }
            emit(img, ypo, xp, jc, v, am);
        };
        for (int bnd = blockIdx.x; bnd < nbands; bnd += gridDim.x) {
            const int img = bnd / C0_BANDS, b = bnd - img * C0_BANDS;
            for (int t = grp; t < C0_BAND_TILES; t += 2) {        // group 0: tiles 0, 2, 4; group 1: tiles 1, 3
                mbar_wait(&tfull[t], aph);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + t * ACC_COLS;
                float v[C0_COUT], u[16];
                tmem_ld16(taddr + C0_COUT, u);
                tmem_ld16(taddr, v);
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = fmaf(u[i], MID_INV, v[i]);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[t]);
                const int pb = t * TILE_M + quad * 32 + lane;
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    *c0_band_ptr(band, pb, c) = make_float4(v[c * 4 + 0] * a.acc_scale + bias[c * 4 + 0], v[c * 4 + 1] * a.acc_scale + bias[c * 4 + 1],
                                                            v[c * 4 + 2] * a.acc_scale + bias[c * 4 + 2], v[c * 4 + 3] * a.acc_scale + bias[c * 4 + 3]);
            }
            aph ^= 1;
            asm volatile("bar.sync 1, 256;" ::: "memory");       // the band is complete in shared memory
            // 3 pooled rows x 42 interior columns x 2 chunks = 252 items, one per thread; threads 252..255 write the 12 border
            // pixels of these rows, and the first / last band of a frame also writes the border row 0 / 43.
            if (etid < C0_BAND_K * C0_HO * 2) {
                const int j = etid % C0_HO, jc = (etid / C0_HO) & 1, r = etid / (2 * C0_HO);
                float v[8];
                int am[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) { v[e] = -INFINITY; am[e] = 15; }
                const int y0 = 2 * (C0_BAND_K * b + r);
#pragma unroll
                for (int dy = 0; dy < 3; ++dy) {
                    if (y0 + dy >= 84) continue;
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) {
                        if (2 * j + dx >= 84) continue;
                        const int pb = (2 * r + dy) * C0_HP + 2 * j + dx + 1;
                        const float4 f0 = *c0_band_ptr(band, pb, 2 * jc), f1 = *c0_band_ptr(band, pb, 2 * jc + 1);
                        const float o[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
#pragma unroll
                        for (int e = 0; e < 8; ++e)
                            if (o[e] > v[e]) { v[e] = o[e]; am[e] = dy * 3 + dx; }
                    }
                }
                emit(img, C0_BAND_K * b + r + 1, j + 1, jc, v, am);
            } else {
                // shared borders: the zero pixel that starts each of these rows (= the right border of the row above) ...
                for (int k = etid - C0_BAND_K * C0_HO * 2; k < C0_BAND_K * 2; k += 4)      // (row, chunk)
                    emit_zero(img, C0_BAND_K * b + (k >> 1) + 1, 0, k & 1);
            }
            if (b == 0 && etid < 2 * C0_WPO)                                               // ... and the zero row above the image
                emit_zero(img, 0, etid % C0_WPO, etid / C0_WPO);
            asm volatile("bar.sync 1, 256;" ::: "memory");       // band buffer free for the next band
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 9) tmem_dealloc(tmem_base, TMEM_COLS);
}

int launch_conv0_pool_umma(const ConvArgs& a, Planes out, Planes out_r, uint8_t* amax, uint8_t* bits, int num_sms, cudaStream_t st) {
    CB_CHECK(a.g.H == 84 && a.g.W == 84 && a.cin_chunks == 1 && a.cout == C0_COUT && !a.transpose,
             "conv0_pool_umma: frame conv (84x84, 4 -> 16 channels) only");
    CB_CHECK(C0_HP + 1 <= GUARD && TILE_M + C0_HP + 2 + TILE_M <= GUARD + 128, "conv0_pool_umma: guard too small");
    static std::atomic<unsigned> attr_done{0};
    int dev = 0;
    CB_CUDA(cudaGetDevice(&dev));
    if (!(attr_done.load() & (1u << dev))) {
        CB_CUDA(cudaFuncSetAttribute(k_conv0_pool_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, C0_SMEM));
        attr_done.fetch_or(1u << dev);
    }
    Conv0PoolArgs p;
    p.n = a.g.n; p.x_hi = a.in.hi; p.wp = a.wp; p.bias = a.ep.bias; p.acc_scale = a.ep.acc_scale;
    p.out = out; p.out_r = out_r; p.amax = amax; p.bits = bits;
    const int nbands = a.g.n * C0_BANDS;
    const int grid = nbands < 2 * num_sms ? nbands : 2 * num_sms;
    launch_pdl(k_conv0_pool_umma, dim3(grid), dim3(CONV_THREADS), (size_t)C0_SMEM, st, p);
    CB_LAUNCH_CHECK();
    return 0;
}

// =================================================================================================
// Second / third ConvSequence, forward: the sequence conv (16 -> 32 at 42x42, 32 -> 32 at 21x21; 3-plane activations, three MMAs
// per K step as in k_conv_umma) FUSED with its max-pool, by the band scheme of k_conv0_pool_umma: a band = CP_K pooled rows =
// 2*CP_K + 1 conv rows = consecutive flat pixels = up to CP_MAXT tiles, each with its own TMEM accumulator (96 columns), so
// the MMAs of the next band overlap the pooling of this one.  The last band of an image may be shorter (21 = 4*5 + 1,
// 11 = 2*5 + 1).  pad_lo = 0 (42 -> 21: windows 2i .. 2i+2) or 1 (21 -> 11: windows 2i-1 .. 2i+1), clipped to the image.
// Band shapes (template parameters K = pooled rows per band, MAXT = tiles per band):
//   42x42 (Wp = 43): K = 3 -> 7 conv rows = 301 px = 3 tiles; 3 accumulators = 256 TMEM columns and ~95 KB of shared memory, so
//                    TWO CTAs (16 epilogue warps, two MMA issuers) share an SM; 7 bands x 7 rows = 49 conv rows per image for 43
//                    needed (K = 2: 11 x 5 = 55; measured 0.516 -> 0.428 ms per 3840 frames).  With K = 5 (4 tiles, 512 columns,
//                    one CTA per SM) the 8 epilogue warps were the bottleneck;
//   21x21 (Wp = 22): K = 5 -> 11 conv rows = 242 px = 2 tiles, one CTA per SM (the 37 KB weight image + 8-plane stages leave no room for two).
constexpr int CP_COUT = 32;
constexpr int CP_ACC_COLS = 2 * CP_COUT;

struct ConvPoolArgs {
    ConvGeom gi, go;          // conv grid (input of the pool) and pooled grid
    int pad_lo;
    int bands_per_img;        // ceil(go.H / K)
    Planes in;                // input activations of the conv (raw output of the previous stage)
    const f16* wp;            // packed forward weight image
    const float* bias;        // [32]
    Planes out;               // pooled planes (raw)
    Planes out_r;             // rectified pooled planes
    uint8_t* amax;            // arg-max bytes or null
    uint8_t* bits;            // relu gate bits of out_r [go.NP][4] or null
};

struct ConvPoolSmem { int win, plane_bytes, stage_bytes, w_bytes, stages, band_bytes, total, ctas_per_sm; };
__host__ __device__ inline ConvPoolSmem conv_pool_smem(int cin_chunks, int Wp, int K, int MAXT) {
    ConvPoolSmem L;
    L.win = TILE_M + 2 * Wp + 2;
    L.plane_bytes = L.win * 16;
    L.stage_bytes = 2 * cin_chunks * L.plane_bytes;
    L.w_bytes = conv_wbytes(cin_chunks, CP_COUT);
    const int band_px = ((2 * K + 1) * Wp + TILE_M - 1) / TILE_M * TILE_M;
    L.band_bytes = band_px * CP_COUT * 4;
    // two CTAs per SM when the accumulators fit 256 TMEM columns AND two stages fit half an SM's shared memory
    const int fixed = 1024 + L.w_bytes + L.band_bytes;
    const int st2 = (113 * 1024 - fixed) / L.stage_bytes;
    L.ctas_per_sm = (MAXT * CP_ACC_COLS <= 256 && st2 >= 2) ? 2 : 1;
    int st = L.ctas_per_sm == 2 ? st2 : (226 * 1024 - fixed) / L.stage_bytes;
    L.stages = st > 4 ? 4 : st;
    L.total = 1024 + L.w_bytes + L.stages * L.stage_bytes + L.band_bytes;
    return L;
}

// 16-byte chunk c (0..7) of band pixel pb (one pixel = 32 channels = one 128-byte row): XOR with the pixel index keeps the
// epilogue's per-lane pixel stores conflict free; the pool's stride-2 pixel loads see a 2-way conflict.
__device__ __forceinline__ float4* cp_band_ptr(uint8_t* band, int pb, int c) {
    return reinterpret_cast<float4*>(band + (pb << 7) + ((c ^ (pb & 7)) << 4));
}

template <int CIN_CHUNKS, int CP_K, int CP_MAXT>
__global__ void __launch_bounds__(CONV_THREADS) k_conv_pool_umma(ConvPoolArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    griddep_launch();
    const ConvPoolSmem L = conv_pool_smem(CIN_CHUNKS, a.gi.Wp, CP_K, CP_MAXT);
    const int NSTAGES = L.stages;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);          // [4]
    uint64_t* empty = full + 4;                                  // [4]
    uint64_t* tfull = empty + 4;                                 // [CP_MAXT]
    uint64_t* tempty = tfull + CP_MAXT;                          // [CP_MAXT]
    uint64_t* wbar = tempty + CP_MAXT;                           // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wbar + 1);
    uint8_t* wsm = smem + 1024;
    uint8_t* stages = wsm + L.w_bytes;
    uint8_t* band = stages + NSTAGES * L.stage_bytes;

    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    constexpr int STEPS = conv_steps(CIN_CHUNKS);
    constexpr uint32_t TMEM_COLS = CP_MAXT * CP_ACC_COLS <= 128 ? 128 : (CP_MAXT * CP_ACC_COLS <= 256 ? 256 : 512);   // CP_MAXT accumulators x 64 columns
    constexpr int NPLANES = 2 * CIN_CHUNKS;
    const int Wp = a.gi.Wp, Ho = a.go.H;
    const int nbands = a.gi.n * a.bands_per_img;
    // band b of an image: pooled rows [b*CP_K, b*CP_K + kb), conv rows from padded row 2*b*CP_K - pad_lo + 1, (2*kb + 1) of them
    auto band_rows = [&](int b) { const int r = Ho - b * CP_K; return r < CP_K ? r : CP_K; };
    auto band_tiles = [&](int kb) { return ((2 * kb + 1) * Wp + TILE_M - 1) / TILE_M; };

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int s = 0; s < CP_MAXT; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 4); }
        mbar_init(wbar, 1);
        fence_barrier_init();
    }
    if (warp == 9) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (warp == 8 && lane == 0) {
        mbar_arrive_expect_tx(wbar, (uint32_t)L.w_bytes);
        bulk_g2s(wsm, a.wp, L.w_bytes, wbar);
    }
    griddep_wait();

    if (warp == 8) {
        // ===================== TMA producer (lanes share the bulk copies of a stage) =====================
        int s = 0; uint32_t ph = 0;
        for (int bnd = blockIdx.x; bnd < nbands; bnd += gridDim.x) {
            const int img = bnd / a.bands_per_img, b = bnd - img * a.bands_per_img;
            const long long qb = (long long)img * a.gi.P + (long long)(2 * b * CP_K - a.pad_lo + 1) * Wp;
            const int nt = band_tiles(band_rows(b));
            for (int t = 0; t < nt; ++t) {
                mbar_wait(&empty[s], ph ^ 1);
                if (lane == 0) mbar_arrive_expect_tx(&full[s], (uint32_t)L.stage_bytes);
                const long long q_lo = qb + t * TILE_M - Wp - 1;
                uint8_t* dst = stages + s * L.stage_bytes;
                if (lane < NPLANES) {
                    const int pl = lane / CIN_CHUNKS, j = lane % CIN_CHUNKS;      // pl: 0 hi, 1 mid
                    const f16* src = pl == 0 ? a.in.hi : a.in.mid;
                    bulk_g2s(dst + lane * L.plane_bytes, src + ((long long)j * a.in.plane_px + q_lo) * 8, L.plane_bytes, &full[s]);
                }
                __syncwarp();
                if (++s == NSTAGES) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 9) {
        // ===================== MMA issuer: two MMAs per K step (see k_conv_umma) =====================
        constexpr uint32_t IDESC2 = make_idesc_f16(TILE_M, 2 * CP_COUT, 0, 0);
        constexpr uint32_t IDESC1 = make_idesc_f16(TILE_M, CP_COUT, 0, 0);
        const uint32_t leader = elect_one();
        mbar_wait(wbar, 0);
        int s = 0; uint32_t ph = 0;
        uint32_t accph = 0;                                      // bit t: phase of accumulator t (flips each time it is used)
        const uint32_t win16 = (uint32_t)L.win;
        const uint32_t b_hi = desc_hi(128), a_hi = desc_hi(128);
        const uint32_t b_lo0 = desc_lo(smem_u32(wsm), 2 * CP_COUT * 16);
        uint32_t a_rel[STEPS];
#pragma unroll
        for (int step = 0; step < STEPS; ++step) {
            constexpr int HALF = CIN_CHUNKS / 2;
            const int tap = step / HALF, pair = step % HALF;
            a_rel[step] = ((uint32_t)(pair * 2) * win16 + (uint32_t)((tap / 3) * Wp + (tap % 3))) | (win16 << 16);
        }
        const uint32_t mid16 = CIN_CHUNKS * win16;
        for (int bnd = blockIdx.x; bnd < nbands; bnd += gridDim.x) {
            const int b = bnd % a.bands_per_img;
            const int nt = band_tiles(band_rows(b));
            for (int t = 0; t < nt; ++t) {
                mbar_wait(&tempty[t], ((accph >> t) & 1) ^ 1);
                mbar_wait(&full[s], ph);
                tc_fence_after();
                {   // warp-uniform issue (umma.cuh)
                    const uint32_t d_tmem = tmem_base + t * CP_ACC_COLS;
                    const uint32_t st16 = (smem_u32(stages + s * L.stage_bytes) >> 4);
#pragma unroll
                    for (int step = 0; step < STEPS; ++step) {
                        const uint32_t a_lo = st16 + a_rel[step];
                        const uint32_t b_lo = b_lo0 + step * (2 * 2 * CP_COUT);
                        mma_f16_elect(d_tmem, a_lo, a_hi, b_lo, b_hi, IDESC2, step > 0, leader);
                        mma_f16_elect(d_tmem + CP_COUT, a_lo + mid16, a_hi, b_lo, b_hi, IDESC1, 1, leader);
                    }
                    mma_commit_elect(&empty[s], leader);
                    mma_commit_elect(&tfull[t], leader);
                }
                __syncwarp();
                accph ^= 1u << t;
                if (++s == NSTAGES) { s = 0; ph ^= 1; }
            }
        }
    } else {
        // ===================== epilogue warps 0-7: accumulators -> band buffer, then pool the band =====================
        const int grp = warp >> 2, quad = warp & 3;
        const int etid = threadIdx.x;                            // 0..255
        uint32_t accph = 0;
        const int Wo = a.go.W, Wpo = a.go.Wp;
        constexpr int NCH = CP_COUT / 8;
        float bias[CP_COUT];
#pragma unroll
        for (int e = 0; e < CP_COUT; ++e) bias[e] = a.bias[e];
        auto emit = [&](int img, int ypo, int xp, int jc, const float* v, const int* am) {
            const long long qo = (long long)img * a.go.P + ypo * Wpo + xp;
            const long long so = ((long long)jc * a.go.NP + qo) * 8;
            store_planes8(a.out, ((long long)jc * a.out.plane_px + qo) * 8, v);
            float rl[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) rl[e] = fmaxf(v[e], 0.f);
            store_planes8(a.out_r, ((long long)jc * a.out_r.plane_px + qo) * 8, rl);
            if (a.amax) {
                uint2 pk;
                pk.x = am[0] | (am[1] << 8) | (am[2] << 16) | (am[3] << 24);
                pk.y = am[4] | (am[5] << 8) | (am[6] << 16) | (am[7] << 24);
                *reinterpret_cast<uint2*>(a.amax + so) = pk;
            }
            if (a.bits) {
                uint32_t bb = 0;
#pragma unroll
                for (int e = 0; e < 8; ++e) bb |= (v[e] > 0.f ? 1u : 0u) << e;
                a.bits[qo * NCH + jc] = (uint8_t)bb;
            }
        };
        auto emit_zero = [&](int img, int ypo, int xp, int jc) {
            float v[8];
            int am[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) { v[e] = 0.f; am[e] = 15; }
            emit(img, ypo, xp, jc, v, am);
        };
        // pooling item -> (pooled row r of the band, chunk jc, column j).  Every band shape in use has <= 256 items, so a thread's
        // item is `etid` in every band: decoded once here instead of three integer divisions per item (~90 of ~420 instructions)
        const int j_0 = etid % Wo, jc_0 = (etid / Wo) % NCH, r_0 = etid / (Wo * NCH);
        for (int bnd = blockIdx.x; bnd < nbands; bnd += gridDim.x) {
            const int img = bnd / a.bands_per_img, b = bnd - img * a.bands_per_img;
            const int kb = band_rows(b), nt = band_tiles(kb);
            for (int t = grp; t < nt; t += 2) {
                mbar_wait(&tfull[t], (accph >> t) & 1);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + t * CP_ACC_COLS;
                const int pb = t * TILE_M + quad * 32 + lane;
#pragma unroll
                for (int h = 0; h < CP_COUT / 16; ++h) {
                    float v[16], u[16];
                    tmem_ld16(taddr + CP_COUT + h * 16, u);
                    tmem_ld16(taddr + h * 16, v);
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = fmaf(u[i], MID_INV, v[i]);
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        *cp_band_ptr(band, pb, h * 4 + c) = make_float4(v[c * 4 + 0] + bias[h * 16 + c * 4 + 0], v[c * 4 + 1] + bias[h * 16 + c * 4 + 1],
                                                                        v[c * 4 + 2] + bias[h * 16 + c * 4 + 2], v[c * 4 + 3] + bias[h * 16 + c * 4 + 3]);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[t]);
                accph ^= 1u << t;
            }
            // accumulators this group did not touch in this band keep their phase; the other group's bits are tracked by that group
            asm volatile("bar.sync 1, 256;" ::: "memory");       // the band is complete in shared memory
            const int nitems = kb * Wo * NCH;
            for (int it = etid; it < nitems; it += 256) {
                int j = j_0, jc = jc_0, r = r_0;
                if (it != etid) { j = it % Wo; jc = (it / Wo) % NCH; r = it / (Wo * NCH); }
                float v[8];
                int am[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) { v[e] = -INFINITY; am[e] = 15; }
                const int y0 = 2 * (b * CP_K + r) - a.pad_lo, x0 = 2 * j - a.pad_lo;
#pragma unroll
                for (int dy = 0; dy < 3; ++dy) {
                    if (y0 + dy < 0 || y0 + dy >= a.gi.H) continue;
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) {
                        if (x0 + dx < 0 || x0 + dx >= a.gi.W) continue;
                        const int pb = (2 * r + dy) * Wp + x0 + dx + 1;
                        const float4 f0 = *cp_band_ptr(band, pb, 2 * jc), f1 = *cp_band_ptr(band, pb, 2 * jc + 1);
                        const float o[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
#pragma unroll
                        for (int e = 0; e < 8; ++e)
                            if (o[e] > v[e]) { v[e] = o[e]; am[e] = dy * 3 + dx; }
                    }
                }
                emit(img, b * CP_K + r + 1, j + 1, jc, v, am);
            }
            // zero border of the pooled grid (shared borders, common.cuh): the pixel that starts each of this band's rows, plus
            // the row above the image
            const int nside = kb * NCH;
            const int ntop = (b == 0) ? Wpo * NCH : 0;
            for (int it = etid; it < nside + ntop; it += 256) {
                if (it < nside) {
                    emit_zero(img, b * CP_K + it / NCH + 1, 0, it % NCH);
                } else {
                    const int k = it - nside;
                    emit_zero(img, 0, k % Wpo, k / Wpo);
                }
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");       // band buffer free for the next band
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 9) tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int CIN_CHUNKS, int CP_K, int CP_MAXT>
static int launch_conv_pool_umma_t(ConvPoolArgs p, int num_sms, cudaStream_t st) {
    const ConvPoolSmem L = conv_pool_smem(CIN_CHUNKS, p.gi.Wp, CP_K, CP_MAXT);
    CB_CHECK(L.stages >= 2 && L.total <= 227 * 1024, "conv_pool_umma<%d>: %d bytes of shared memory needed", CIN_CHUNKS, L.total);
    CB_CHECK(((2 * CP_K + 1) * p.gi.Wp + TILE_M - 1) / TILE_M <= CP_MAXT, "conv_pool_umma: band of %d rows x %d does not fit %d tiles",
             2 * CP_K + 1, p.gi.Wp, CP_MAXT);
    p.bands_per_img = (p.go.H + CP_K - 1) / CP_K;
    {   // the last band's windows may run past the last image: they must stay inside the back guard
        const int bl = p.bands_per_img - 1, kbl = p.go.H - bl * CP_K;
        const int ntl = ((2 * kbl + 1) * p.gi.Wp + TILE_M - 1) / TILE_M;
        const int over = (2 * bl * CP_K - p.pad_lo + 1) * p.gi.Wp + ntl * TILE_M + p.gi.Wp + 1 - p.gi.P;
        CB_CHECK(over <= GUARD, "conv_pool_umma: the last band reads %d pixels past the image (guard %d)", over, GUARD);
    }
    static std::atomic<unsigned> attr_done{0};
    int dev = 0;
    CB_CUDA(cudaGetDevice(&dev));
    if (!(attr_done.load() & (1u << dev))) {
        CB_CUDA(cudaFuncSetAttribute(k_conv_pool_umma<CIN_CHUNKS, CP_K, CP_MAXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_done.fetch_or(1u << dev);
    }
    const int nbands = p.gi.n * p.bands_per_img;
    const int per_sm = L.ctas_per_sm;
    const int grid = nbands < per_sm * num_sms ? nbands : per_sm * num_sms;
    launch_pdl(k_conv_pool_umma<CIN_CHUNKS, CP_K, CP_MAXT>, dim3(grid), dim3(CONV_THREADS), (size_t)L.total, st, p);
    CB_LAUNCH_CHECK();
    return 0;
}

int launch_conv_pool_umma(const ConvArgs& a, ConvGeom go, int pad_lo, Planes out, Planes out_r, uint8_t* amax, uint8_t* bits, int num_sms,
                          cudaStream_t st) {
    CB_CHECK(a.cout == CP_COUT && !a.transpose && a.in.mid && (a.cin_chunks == 2 || a.cin_chunks == 4),
             "conv_pool_umma: 16|32 -> 32 channel sequence convs only");
    CB_CHECK(a.g.Wp + 1 <= GUARD, "conv_pool_umma: guard too small for Wp=%d", a.g.Wp);
    ConvPoolArgs p;
    p.gi = a.g; p.go = go; p.pad_lo = pad_lo; p.bands_per_img = 0;
    p.in = a.in; p.wp = a.wp; p.bias = a.ep.bias; p.out = out; p.out_r = out_r; p.amax = amax; p.bits = bits;
    static const int k42 = [] { const char* e = getenv("CLEANBA_CP42_K"); return e ? atoi(e) : 3; }();      // A/B knob (2 | 3)
    if (a.cin_chunks == 2 && k42 == 2) return launch_conv_pool_umma_t<2, 2, 2>(p, num_sms, st);
    if (a.cin_chunks == 2) return launch_conv_pool_umma_t<2, 3, 3>(p, num_sms, st);
    return launch_conv_pool_umma_t<4, 5, 2>(p, num_sms, st);
}

// =================================================================================================
// wgrad on tcgen05:  dW[(ky,kx), ci, co] = sum_q X[q + d(ky,kx)][ci] * G[q][co],  db[co] = sum_q G[q][co].
// GEMM view per 128-pixel block:  D_kx[m, n] += A_kx[m, q] * B[q, n]  with the reduction over pixels (K), where
//   m = (ky, ci) stacks the three filter ROWS (three bulk copies of the same planes shifted by one image row) and
//   kx is a 16-byte shift of the A start address.  A is MN-major (M = channels contiguous), B (= G) is MN-major too,
//   so both operands are again the untouched chunk planes.  One extra M group of fp16 ones yields the bias gradient.
// Precision: X and G are fp16x2 carriers (22 bits); n = (G plane, co) stacks the two G planes along N; the blocks that involve
//   one mid plane carry 2^11 and are folded in by the epilogue / the reduce kernel (G also carries the loss scale S).
// Measured limits that shape the kernel (profiles/r01_v5_*): the TMA unit retires one bulk copy per ~50 cycles per SM, so
//   copies are >= 2 KB (128-pixel blocks); one thread issues at most one MMA per ~50 cycles and the tensor pipe needs
//   (A bytes + B bytes) / 128 cycles per MMA, so the MMA count per block is minimised per shape:
//   * Cin = 4 (frames, exact in fp16):  M = 64,  one MMA per (K step, kx), two CTAs per SM.
//   * Cin = 16: the X planes are stacked along M as well ([X_hi groups | ones | X_mid groups | zeros] = 14 groups, M = 128):
//     ONE MMA per (K step, kx) yields X_hi*G_hi, X_hi*G_mid, X_mid*G_hi (and X_mid*G_mid for free); two CTAs per SM.
//   * Cin = 32: 13 groups per plane do not stack; two issuing threads instead, each with its own accumulators
//     (X_hi * [G_hi|G_mid] and X_mid * [G_hi]), one CTA per SM with a 3-stage ring.
// The accumulators stay in TMEM across ALL pixel blocks of a CTA; one partial per CTA is written at the end and
// reduced in a fixed order by k_wgrad_umma_reduce (deterministic).
constexpr int WG_BLOCK = 128;             // pixels per pipeline stage (K block)
constexpr int WG_WINX = WG_BLOCK + 2;     // pixels per shifted copy
constexpr int WG_PLANE = WG_WINX * 16;    // bytes of one copy (one M group)
constexpr int WG_BCHUNK = WG_BLOCK * 16;  // bytes of one (plane, chunk) of G (one N group)
constexpr int WG_THREADS = 224;           // warps 0-3 epilogue, 4 TMA producer, 5 / 6 MMA issuers
constexpr int WG_MAXST = 4;

template <int CIN_CHUNKS>
struct WgShape {
    static constexpr int XPL = CIN_CHUNKS == 1 ? 1 : 2;         // X planes used
    static constexpr bool MSTACK = CIN_CHUNKS == 2;             // both X planes in one MMA (stacked along M)
    static constexpr bool DUAL = CIN_CHUNKS == 4;               // two issuers, separate accumulators per X plane
    static constexpr int GROUPS = 3 * CIN_CHUNKS;               // real M groups per plane: (ky, chunk)
    static constexpr int HROWS = (GROUPS + 1) * 8;              // rows per plane incl. the ones / zeros group
    static constexpr int ROWS = MSTACK ? 2 * HROWS : HROWS;     // rows of D that are written to the partial
    static constexpr int MMA_M = ROWS <= 64 ? 64 : 128;
};

struct WgSmemLayout { int a_bytes, b_bytes, stage_bytes, stages, total, ctas_per_sm; };
template <int CIN_CHUNKS, int COUT>
__host__ __device__ inline WgSmemLayout wg_smem_layout() {
    using S = WgShape<CIN_CHUNKS>;
    WgSmemLayout L;
    L.a_bytes = (S::GROUPS + 1) * WG_PLANE;                 // one X plane: real groups + ones (hi) / zeros (mid)
    L.b_bytes = 2 * (COUT / 8) * WG_BCHUNK;
    L.stage_bytes = S::XPL * L.a_bytes + L.b_bytes;
    // the MMA reads MMA_M / 8 groups from its A base; the unused ones fall into the B region of the same stage (+ a pad
    // after the last stage where B is smaller than the overrun)
    const int over = (S::MMA_M / 8) * WG_PLANE - ((S::MSTACK ? 2 : 1) * L.a_bytes + L.b_bytes);
    const int pad = over > 0 ? (over + 1023) / 1024 * 1024 : 0;
    // (three CTAs = three MMA issuers per SM with 2-stage rings: no change, 0.882 vs 0.872 ms per four Cin = 16 launches -- the
    // tensor pipe's per-instruction time, not the issuing thread, is the limit; profiles/r02_v9_conv_epilogue_ab.txt)
    L.ctas_per_sm = S::DUAL ? 1 : 2;
    const int budget = (S::DUAL ? 226 : 113) * 1024 - 1024 - pad;
    int st = budget / L.stage_bytes;
    L.stages = st > WG_MAXST ? WG_MAXST : st;
    L.total = 1024 + L.stages * L.stage_bytes + pad;
    return L;
}

template <int CIN_CHUNKS, int COUT>
__global__ void __launch_bounds__(WG_THREADS) k_wgrad_umma(WgradArgs a, int nblocks, float* __restrict__ partial) {
    using S = WgShape<CIN_CHUNKS>;
    extern __shared__ __align__(1024) uint8_t smem[];
    griddep_launch();
    const WgSmemLayout L = wg_smem_layout<CIN_CHUNKS, COUT>();
    const int NSTAGES = L.stages;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + WG_MAXST;
    uint64_t* done = empty + WG_MAXST;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
    uint8_t* stages = smem + 1024;
    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    constexpr int XPL = S::XPL, GROUPS = S::GROUPS, NCH = COUT / 8;
    constexpr int ACC0 = 2 * COUT;                             // columns of one kx accumulator: [G_hi | G_mid]
    constexpr int ACC1 = S::DUAL ? COUT : 0;                   // second issuer: X_mid * G_hi
    constexpr int COLS = 3 * (ACC0 + ACC1);
    constexpr uint32_t TMEM_COLS = COLS <= 128 ? 128 : (COLS <= 256 ? 256 : 512);
    constexpr int NCOPY_A = XPL * GROUPS, NCOPY_B = 2 * NCH;
    constexpr int NISSUE = S::DUAL ? 2 : 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NISSUE); }
        mbar_init(done, NISSUE);
        fence_barrier_init();
    }
    // constant "ones" group (hi region) and "zeros" group (mid region) of every stage
    for (int s = 0; s < NSTAGES; ++s)
        for (int pl = 0; pl < XPL; ++pl) {
            uint32_t* g = reinterpret_cast<uint32_t*>(stages + s * L.stage_bytes + pl * L.a_bytes + GROUPS * WG_PLANE);
            const uint32_t val = pl == 0 ? 0x3C003C00u : 0u;   // fp16 1.0 x2
            for (int t = threadIdx.x; t < WG_PLANE / 4; t += WG_THREADS) g[t] = val;
        }
    fence_proxy_async();
    if (warp == 5) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int Wp = a.g.Wp;
    griddep_wait();

    if (warp == 4) {
        // ===================== TMA producer: one bulk copy per lane =====================
        int s = 0; uint32_t ph = 0;
        const uint32_t tx = (uint32_t)(NCOPY_A * WG_PLANE + NCOPY_B * WG_BCHUNK);
        for (int blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
            mbar_wait(&empty[s], ph ^ 1);
            if (lane == 0) mbar_arrive_expect_tx(&full[s], tx);
            __syncwarp();
            uint8_t* dst = stages + s * L.stage_bytes;
            const long long q0 = (long long)blk * WG_BLOCK;
            for (int i = lane; i < NCOPY_A + NCOPY_B; i += 32) {
                if (i < NCOPY_A) {
                    const int pl = i / GROUPS, g = i % GROUPS, ky = g / CIN_CHUNKS, j = g % CIN_CHUNKS;
                    const f16* src = pl == 0 ? a.x.hi : a.x.mid;
                    const long long q = q0 + (ky - 1) * Wp - 1;
                    bulk_g2s(dst + pl * L.a_bytes + g * WG_PLANE, src + ((long long)j * a.x.plane_px + q) * 8, WG_PLANE, &full[s]);
                } else {
                    const int k = i - NCOPY_A, pl = k / NCH, j = k % NCH;
                    const f16* src = pl == 0 ? a.gy.hi : a.gy.mid;
                    bulk_g2s(dst + XPL * L.a_bytes + k * WG_BCHUNK, src + ((long long)j * a.gy.plane_px + q0) * 8, WG_BCHUNK, &full[s]);
                }
            }
            __syncwarp();
            if (++s == NSTAGES) { s = 0; ph ^= 1; }
        }
    } else if (warp == 5 || (S::DUAL && warp == 6)) {
        // ===================== MMA issuer(s) =====================
        const int who = warp - 5;                               // 0: X_hi (or both planes when stacked), 1: X_mid (DUAL)
        const uint32_t idesc = make_idesc_f16(S::MMA_M, who == 0 ? 2 * COUT : COUT, 1, 1);
        const uint32_t d0 = tmem_base + (who == 0 ? 0 : 3 * ACC0);
        const uint32_t dstep = who == 0 ? ACC0 : ACC1;
        const uint32_t leader = elect_one();
        int s = 0; uint32_t ph = 0;
        uint32_t accum = 0;
        for (int blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
            mbar_wait(&full[s], ph);
            tc_fence_after();
            {   // warp-uniform issue (umma.cuh)
                const uint32_t st_base = smem_u32(stages + s * L.stage_bytes);
                // K step = 16 pixels = 2 core-matrix groups of 8 pixels (128 bytes each, LBO); M groups WG_PLANE apart,
                // N groups (8 channels of one G plane chunk) WG_BCHUNK apart (SBO)
                const uint32_t a_lo0 = desc_lo(st_base + who * L.a_bytes, 128), a_hi_w = desc_hi(WG_PLANE);
                const uint32_t b_lo0 = desc_lo(st_base + XPL * L.a_bytes, 128), b_hi_w = desc_hi(WG_BCHUNK);
#pragma unroll
                for (int ks = 0; ks < WG_BLOCK / 16; ++ks) {
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx)
                        mma_f16_elect(d0 + kx * dstep, a_lo0 + ks * 16 + kx, a_hi_w, b_lo0 + ks * 16, b_hi_w, idesc,
                                      ks == 0 ? accum : 1u, leader);
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
        // D rows: M = 128: TMEM lane = row.  M = 64: row m lives in lane (m / 16) * 32 + m % 16, i.e. the first 16 lanes of
        // every warp's quadrant (probed, profiles/r01_mma_m64_layout_probe.txt).
        const int m = S::MMA_M == 128 ? warp * 32 + lane : (lane < 16 ? warp * 16 + lane : (1 << 30));
        float* out = partial + (long long)blockIdx.x * (3 * S::ROWS * COUT);
        const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
        for (int kx = 0; kx < 3; ++kx) {
            float v[COUT], t[16];
#pragma unroll
            for (int h = 0; h < COUT / 16; ++h) {
                tmem_ld16(lane_addr + kx * ACC0 + COUT + h * 16, t);                   // * G_mid (carries 2^11)
                tmem_ld16(lane_addr + kx * ACC0 + h * 16, v + h * 16);                 // * G_hi
#pragma unroll
                for (int i = 0; i < 16; ++i) v[h * 16 + i] = fmaf(t[i], MID_INV, v[h * 16 + i]);
                if (S::DUAL) {
                    tmem_ld16(lane_addr + 3 * ACC0 + kx * ACC1 + h * 16, t);           // X_mid * G_hi (same rows, carries 2^11)
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[h * 16 + i] = fmaf(t[i], MID_INV, v[h * 16 + i]);
                }
            }
            if (m < S::ROWS) {
#pragma unroll
                for (int c = 0; c < COUT; c += 4)
                    *reinterpret_cast<float4*>(out + ((long long)kx * S::ROWS + m) * COUT + c) =
                        make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem_base, TMEM_COLS);
}

// dW[(ky*3+kx)][ci][co] = scale / S * sum_cta (P[cta][kx][ky*Cpad + ci][co] + 2^-11 * P[cta][kx][hrows + ky*Cpad + ci][co] if stacked)
// db[co] = 1 / S * sum_cta P[cta][1][3*Cpad][co]              (S = the minibatch's loss scale, *inv_scale = 1 / S)
// One block per 32 consecutive outputs; warp w of 32 sums the partials of CTAs w, w+32, ... (coalesced 128-byte rows, loads of
// four CTAs in flight), the 32 warp sums are added in a fixed order (deterministic).
constexpr int WG_RED_WARPS = 32;
__global__ void __launch_bounds__(WG_RED_WARPS * 32) k_wgrad_umma_reduce(const float* __restrict__ partial, int nctas, int cpad, int rows,
                                                                          int hrows, int stacked, int cin_real, int cout, float scale,
                                                                          const float* __restrict__ inv_scale, float* __restrict__ dw,
                                                                          float* __restrict__ db) {
    __shared__ float sm[WG_RED_WARPS][32];
    griddep_launch();
    griddep_wait();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + lane;
    const int nw = 9 * cin_real * cout;
    float s = 0.f;
    if (i < nw + cout) {
        long long src;
        bool two = false;
        if (i < nw) {
            int co = i % cout, ci = (i / cout) % cin_real, tap = i / (cout * cin_real);
            int ky = tap / 3, kx = tap % 3;
            src = ((long long)kx * rows + ky * cpad + ci) * cout + co;
            two = stacked != 0;
        } else {
            src = ((long long)1 * rows + 3 * cpad) * cout + (i - nw);
        }
        const long long stride = (long long)3 * rows * cout;
        const long long off2 = two ? (long long)hrows * cout : 0;
#pragma unroll 4
        for (int b = w; b < nctas; b += WG_RED_WARPS) {
            const float* p = partial + (long long)b * stride + src;
            float v = p[0];
            if (two) v = fmaf(p[off2], MID_INV, v);      // rows of the X_mid groups
            s += v;
        }
    }
    sm[w][lane] = s;
    __syncthreads();
    if (w == 0 && i < nw + cout) {
        float t = sm[0][lane];
#pragma unroll
        for (int k = 1; k < WG_RED_WARPS; ++k) t += sm[k][lane];
        const float inv = inv_scale ? *inv_scale : 1.f;
        if (i < nw) dw[i] = t * scale * inv; else db[i - nw] = t * inv;
    }
}

template <int CIN_CHUNKS, int COUT>
static int launch_wgrad_umma_t(const WgradArgs& a, float* partial, int num_sms, cudaStream_t st) {
    using S = WgShape<CIN_CHUNKS>;
    WgSmemLayout L = wg_smem_layout<CIN_CHUNKS, COUT>();
    CB_CHECK(L.stages >= 2 && L.total <= 227 * 1024, "wgrad_umma<%d,%d>: %d bytes of shared memory needed", CIN_CHUNKS, COUT, L.total);
    static std::atomic<unsigned> attr_done{0};
    int dev = 0;
    CB_CUDA(cudaGetDevice(&dev));
    if (!(attr_done.load() & (1u << dev))) {
        CB_CUDA(cudaFuncSetAttribute(k_wgrad_umma<CIN_CHUNKS, COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_done.fetch_or(1u << dev);
    }
    int nblocks = (int)((a.g.NP + WG_BLOCK - 1) / WG_BLOCK);
    int grid = nblocks < num_sms * L.ctas_per_sm ? nblocks : num_sms * L.ctas_per_sm;
    launch_pdl(k_wgrad_umma<CIN_CHUNKS, COUT>, dim3(grid), dim3(WG_THREADS), (size_t)L.total, st, a, nblocks, partial);
    CB_LAUNCH_CHECK();
    int nw = 9 * a.cin_real * a.cout;
    launch_pdl(k_wgrad_umma_reduce, dim3((nw + a.cout + 31) / 32), dim3(WG_RED_WARPS * 32), 0, st, (const float*)partial, grid, CIN_CHUNKS * 8,
               S::ROWS, S::HROWS, S::MSTACK ? 1 : 0, a.cin_real, a.cout, a.scale, a.inv_scale, a.dw, a.db);
    CB_LAUNCH_CHECK();
    return 0;
}

int launch_wgrad_umma(const WgradArgs& a, float* partial, int num_sms, cudaStream_t st) {
    CB_CHECK(WG_BLOCK + a.g.Wp + 2 <= GUARD, "wgrad_umma: guard too small for Wp=%d", a.g.Wp);
    CB_CHECK(a.gy.mid, "wgrad_umma: gradient tensors carry two fp16 planes");
    CB_CHECK(a.cin_chunks == 1 || a.x.mid, "wgrad_umma: activation planes hi and mid needed");
    if (a.cin_chunks == 1 && a.cout == 16) return launch_wgrad_umma_t<1, 16>(a, partial, num_sms, st);
    if (a.cin_chunks == 2 && a.cout == 16) return launch_wgrad_umma_t<2, 16>(a, partial, num_sms, st);
    if (a.cin_chunks == 2 && a.cout == 32) return launch_wgrad_umma_t<2, 32>(a, partial, num_sms, st);
    if (a.cin_chunks == 4 && a.cout == 32) return launch_wgrad_umma_t<4, 32>(a, partial, num_sms, st);
    CB_CHECK(false, "wgrad_umma: unsupported shape cin_chunks=%d cout=%d", a.cin_chunks, a.cout);
}

}  // namespace cb
