// Persistent tail of the ACTOR's trunk (cleanba/cleanba_ppo.py:245-261 get_action_and_value -> Network.__call__ :178-189):
// ConvSequence 1 and ConvSequence 2 -- ten 3x3 convolutions and two max-pools -- of a rollout step in ONE kernel.
//
// Why: at the actor's batch (local_num_envs = 60 frames) these ten layers are 5 % of a B200's tensor throughput but ten
// dependent launches; each costs ~7 us of launch / prologue / drain latency around ~1 us of MMAs
// (profiles/r02_v5_ncu_sweep_actor_n60.txt), i.e. more than half of the whole step.  A conv layer only couples the pixels of
// ONE frame, so a thread-block cluster that owns a frame can run all ten layers back to back with nothing but a cluster
// barrier between them -- no grid-wide dependency exists.
//
// Shape: one cluster of `csize` CTAs (1 or 2) per frame; the CTAs of a cluster split the 128-pixel tiles (convs) or the row
// bands (conv + pool) of their frame.  Every layer is the same flat-shifted-window implicit GEMM as conv_umma.cu (same warp
// roles, same packed weight images, same fp16x2 carrier epilogue, so results are bit-identical to the per-layer kernels); the
// activations go through the L2-resident planes the per-layer path uses, so the rest of the step (dense layer, heads) is
// unchanged.  Between layers: the epilogue threads' plane stores are made visible to the next layer's bulk-TMA loads
// (generic -> async proxy fence + release / acquire cluster barrier).  The pipeline (mbarrier ring, accumulator barriers, their
// phases) is set up ONCE and keeps running across layers; the packed weights are double-buffered, layer l + 1's image streams
// in while layer l computes.  TMEM (128 columns) is allocated once.
//
// Tiles are FRAME-aligned here (tile t of frame i covers flat pixels i*P + 128 t ...), so the last tile of a frame laps into
// the next frame's pixels: those accumulators are discarded (another cluster owns them).
#include "common.cuh"
#include "kernels.h"
#include "umma.cuh"

namespace cb {
using namespace umma;

constexpr int AT_THREADS = 320;      // warps 0-7: two epilogue groups, warp 8: TMA producer, warp 9: MMA issuer (conv_umma.cu)
constexpr int AT_TILE_M = 128;
constexpr int AT_COUT = 32;
constexpr int AT_ACC_COLS = 2 * AT_COUT;
constexpr int AT_NST = 3;                        // activation-window ring
constexpr uint32_t AT_TMEM_COLS = 128;           // two accumulators of 64 columns

__host__ __device__ constexpr int at_steps(int cin_chunks) { return 9 * (cin_chunks / 2); }
__host__ __device__ constexpr int at_wbytes(int cin_chunks) { return at_steps(cin_chunks) * 2 * 2 * AT_COUT * 16; }
constexpr int AT_WSLOT = at_wbytes(4);                                   // 36,864 bytes: one packed 32 -> 32 weight image
constexpr int AT_STAGE_SLOT = 8 * (AT_TILE_M + 2 * 22 + 2) * 16;         // 22,272 bytes: 8 planes of a 21x21 window (the largest)
constexpr int AT_BAND = 2 * AT_TILE_M * AT_COUT * 4;                     // 32,768 bytes: two tiles of fp32 conv outputs
constexpr int AT_SMEM = 1024 + 2 * AT_WSLOT + AT_NST * AT_STAGE_SLOT + AT_BAND;
static_assert(4 * (AT_TILE_M + 2 * 43 + 2) * 16 <= AT_STAGE_SLOT, "the 42x42 window (4 planes) fits a stage slot");
static_assert(AT_SMEM <= 227 * 1024, "shared memory");

struct AtBars {                      // first 1 KB of shared memory
    uint64_t full[AT_NST], empty[AT_NST], tfull[2], tempty[2], wb[2];
    uint32_t tmem_slot;
};

// Pipeline state every thread carries across layers (each role advances the fields it uses; all roles see the same sequence).
struct AtState {
    int s;               // ring slot (producer / issuer)
    uint32_t ph;         // ring phase
    uint32_t accph;      // bit t: phase of accumulator t
    uint32_t layer;      // layers executed so far by this CTA: weight buffer layer & 1, its barrier phase (layer >> 1) & 1
};

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_nctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_id_x() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_count_x() { uint32_t r; asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r)); return r; }

// End of a layer: this cluster's plane stores (generic proxy) become visible to the next layer's bulk-TMA loads (async proxy)
// and to the residual loads of ANY thread of the cluster (release / acquire at cluster scope).
__device__ __forceinline__ void layer_barrier(uint32_t csize) {
    tc_fence_before();
    asm volatile("fence.proxy.async;" ::: "memory");
    if (csize > 1) {
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    } else {
        __threadfence_block();
        __syncthreads();
    }
    tc_fence_after();
}

__device__ __forceinline__ uint8_t* at_wsm(uint8_t* smem, uint32_t layer) { return smem + 1024 + (layer & 1) * AT_WSLOT; }
__device__ __forceinline__ uint8_t* at_stage(uint8_t* smem, int s) { return smem + 1024 + 2 * AT_WSLOT + s * AT_STAGE_SLOT; }
__device__ __forceinline__ uint8_t* at_band(uint8_t* smem) { return smem + 1024 + 2 * AT_WSLOT + AT_NST * AT_STAGE_SLOT; }
// the descriptor address field is relative to the CTA's own shared window (a CTA of cluster rank > 0 sees a non-zero window offset)
__device__ __forceinline__ uint32_t at_desc16(const void* p) { return (smem_u32(p) >> 4) & 0x3FFFu; }

// Issue the bulk load of a layer's packed weights into the buffer that layer will use (one thread).
__device__ __forceinline__ void at_load_weights(uint8_t* smem, AtBars* B, uint32_t layer, const f16* wp, int w_bytes) {
    mbar_arrive_expect_tx(&B->wb[layer & 1], (uint32_t)w_bytes);
    bulk_g2s(at_wsm(smem, layer), wp, w_bytes, &B->wb[layer & 1]);
}

// One 128-pixel tile: 9 taps x CIN_CHUNKS / 2 K-steps, two MMAs each (conv_umma.cu).  Warp-uniform issue.
template <int CIN_CHUNKS>
__device__ __forceinline__ void at_issue_tile(uint32_t d_tmem, uint32_t st16, uint32_t b_lo0, int Wp, uint32_t win16, uint32_t leader) {
    constexpr int STEPS = at_steps(CIN_CHUNKS), COUT = AT_COUT, HALF = CIN_CHUNKS / 2;
    constexpr uint32_t IDESC2 = make_idesc_f16(AT_TILE_M, 2 * COUT, 0, 0);
    constexpr uint32_t IDESC1 = make_idesc_f16(AT_TILE_M, COUT, 0, 0);
    const uint32_t b_hi = desc_hi(128), a_hi = desc_hi(128);
    const uint32_t mid16 = CIN_CHUNKS * win16;
#pragma unroll
    for (int step = 0; step < STEPS; ++step) {
        const int tap = step / HALF, pair = step % HALF;
        const uint32_t a_lo = st16 + (((uint32_t)(pair * 2) * win16 + (uint32_t)((tap / 3) * Wp + (tap % 3))) | (win16 << 16));
        const uint32_t b_lo = b_lo0 + step * (2 * 2 * COUT);
        mma_f16_elect(d_tmem, a_lo, a_hi, b_lo, b_hi, IDESC2, step > 0, leader);                 // A_hi * [W_hi | W_mid]
        mma_f16_elect(d_tmem + COUT, a_lo + mid16, a_hi, b_lo, b_hi, IDESC1, 1, leader);         // block 1 += A_mid * W_hi
    }
}

// ------------------------------------------------------------------------------------------------ 3x3 conv, 32 -> 32
// Tiles i = 0, 1, ... of this CTA are frame tiles t = crank + i * csize; accumulator i & 1 = epilogue group i & 1.
__device__ __forceinline__ void tail_conv32(const ConvArgs& a, int img, uint32_t crank, uint32_t csize, uint8_t* smem, uint32_t tmem_base,
                                            AtState& S, int warp, int lane) {
    constexpr int CIN_CHUNKS = 4, COUT = AT_COUT, NPLANES = 2 * CIN_CHUNKS;
    AtBars* B = reinterpret_cast<AtBars*>(smem);
    const int Wp = a.g.Wp;
    const int win = AT_TILE_M + 2 * Wp + 2, plane_bytes = win * 16, stage_bytes = NPLANES * plane_bytes;
    const long long q_base = (long long)img * a.g.P, q_limit = q_base + a.g.P;
    const int ntl = (a.g.P + AT_TILE_M - 1) / AT_TILE_M;                 // tiles of one frame
    const int mine = ntl > (int)crank ? (ntl - (int)crank + (int)csize - 1) / (int)csize : 0;

    if (warp == 8) {
        for (int i = 0; i < mine; ++i) {
            const int t = (int)crank + i * (int)csize;
            mbar_wait(&B->empty[S.s], S.ph ^ 1);
            if (lane == 0) mbar_arrive_expect_tx(&B->full[S.s], (uint32_t)stage_bytes);
            const long long q_lo = q_base + (long long)t * AT_TILE_M - Wp - 1;
            uint8_t* dst = at_stage(smem, S.s);
            if (lane < NPLANES) {
                const int pl = lane / CIN_CHUNKS, j = lane % CIN_CHUNKS;
                const f16* src = pl == 0 ? a.in.hi : a.in.mid;
                bulk_g2s(dst + lane * plane_bytes, src + ((long long)j * a.in.plane_px + q_lo) * 8, plane_bytes, &B->full[S.s]);
            }
            __syncwarp();
            if (++S.s == AT_NST) { S.s = 0; S.ph ^= 1; }
        }
    } else if (warp == 9) {
        const uint32_t leader = elect_one();
        mbar_wait(&B->wb[S.layer & 1], (S.layer >> 1) & 1);
        const uint32_t b_lo0 = desc_lo(smem_u32(at_wsm(smem, S.layer)), 2 * COUT * 16);
        for (int i = 0; i < mine; ++i) {
            const int acc = i & 1;
            mbar_wait(&B->tempty[acc], ((S.accph >> acc) & 1) ^ 1);
            mbar_wait(&B->full[S.s], S.ph);
            tc_fence_after();
            at_issue_tile<CIN_CHUNKS>(tmem_base + acc * AT_ACC_COLS, at_desc16(at_stage(smem, S.s)), b_lo0, Wp, (uint32_t)win, leader);
            mma_commit_elect(&B->empty[S.s], leader);
            mma_commit_elect(&B->tfull[acc], leader);
            __syncwarp();
            S.accph ^= 1u << acc;
            if (++S.s == AT_NST) { S.s = 0; S.ph ^= 1; }
        }
    } else {
        const int grp = warp >> 2, quad = warp & 3;
        for (int i = grp; i < mine; i += 2) {
            const int t = (int)crank + i * (int)csize;
            const long long q = q_base + (long long)t * AT_TILE_M + quad * 32 + lane;
            const bool own = q < q_limit;                        // the rest of the frame's last tile belongs to the next frame
            EpiPrefetch<COUT> pre;
            pre.in = false; pre.tail = !own; pre.mbits = 0;
            if (own) epi_prefetch<COUT>(a.ep, a.g, q, pre);
            mbar_wait(&B->tfull[grp], (S.accph >> grp) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + grp * AT_ACC_COLS;
            float v[COUT];
            {
                float u[16];
#pragma unroll
                for (int h = 0; h < COUT / 16; ++h) {
                    tmem_ld16(taddr + COUT + h * 16, u);
                    tmem_ld16(taddr + h * 16, v + h * 16);
#pragma unroll
                    for (int k = 0; k < 16; ++k) v[h * 16 + k] = fmaf(u[k], MID_INV, v[h * 16 + k]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&B->tempty[grp]);
            if (own) epi_finish<COUT>(a.ep, a.g, q, v, pre);
            S.accph ^= 1u << grp;
        }
    }
}

// ------------------------------------------------------------------------------------------------ conv (-> 32) + max-pool
// The band scheme of k_conv_pool_umma (conv_umma.cu): band b = K pooled rows = 2K + 1 conv rows = up to two 128-pixel tiles,
// tile t in accumulator t (epilogue group t); this CTA takes bands b = crank, crank + csize, ...
struct TailPool {
    ConvGeom gi, go;
    int pad_lo, bands_per_img;
    Planes in;
    const f16* wp;
    const float* bias;
    Planes out, out_r;
};

__device__ __forceinline__ float4* at_band_ptr(uint8_t* band, int pb, int c) {
    return reinterpret_cast<float4*>(band + (pb << 7) + ((c ^ (pb & 7)) << 4));
}

template <int CIN_CHUNKS, int CP_K>
__device__ __forceinline__ void tail_conv_pool(const TailPool& a, int img, uint32_t crank, uint32_t csize, uint8_t* smem, uint32_t tmem_base,
                                               AtState& S, int warp, int lane) {
    constexpr int COUT = AT_COUT, NPLANES = 2 * CIN_CHUNKS, NCH = COUT / 8;
    AtBars* B = reinterpret_cast<AtBars*>(smem);
    const int Wp = a.gi.Wp, Ho = a.go.H, Wo = a.go.W, Wpo = a.go.Wp;
    const int win = AT_TILE_M + 2 * Wp + 2, plane_bytes = win * 16, stage_bytes = NPLANES * plane_bytes;
    uint8_t* band = at_band(smem);
    auto band_rows = [&](int b) { const int r = Ho - b * CP_K; return r < CP_K ? r : CP_K; };
    auto band_tiles = [&](int kb) { return ((2 * kb + 1) * Wp + AT_TILE_M - 1) / AT_TILE_M; };

    if (warp == 8) {
        for (int b = (int)crank; b < a.bands_per_img; b += (int)csize) {
            const long long qb = (long long)img * a.gi.P + (long long)(2 * b * CP_K - a.pad_lo + 1) * Wp;
            const int nt = band_tiles(band_rows(b));
            for (int t = 0; t < nt; ++t) {
                mbar_wait(&B->empty[S.s], S.ph ^ 1);
                if (lane == 0) mbar_arrive_expect_tx(&B->full[S.s], (uint32_t)stage_bytes);
                const long long q_lo = qb + t * AT_TILE_M - Wp - 1;
                uint8_t* dst = at_stage(smem, S.s);
                if (lane < NPLANES) {
                    const int pl = lane / CIN_CHUNKS, j = lane % CIN_CHUNKS;
                    const f16* src = pl == 0 ? a.in.hi : a.in.mid;
                    bulk_g2s(dst + lane * plane_bytes, src + ((long long)j * a.in.plane_px + q_lo) * 8, plane_bytes, &B->full[S.s]);
                }
                __syncwarp();
                if (++S.s == AT_NST) { S.s = 0; S.ph ^= 1; }
            }
        }
    } else if (warp == 9) {
        const uint32_t leader = elect_one();
        mbar_wait(&B->wb[S.layer & 1], (S.layer >> 1) & 1);
        const uint32_t b_lo0 = desc_lo(smem_u32(at_wsm(smem, S.layer)), 2 * COUT * 16);
        for (int b = (int)crank; b < a.bands_per_img; b += (int)csize) {
            const int nt = band_tiles(band_rows(b));
            for (int t = 0; t < nt; ++t) {
                mbar_wait(&B->tempty[t], ((S.accph >> t) & 1) ^ 1);
                mbar_wait(&B->full[S.s], S.ph);
                tc_fence_after();
                at_issue_tile<CIN_CHUNKS>(tmem_base + t * AT_ACC_COLS, at_desc16(at_stage(smem, S.s)), b_lo0, Wp, (uint32_t)win, leader);
                mma_commit_elect(&B->empty[S.s], leader);
                mma_commit_elect(&B->tfull[t], leader);
                __syncwarp();
                S.accph ^= 1u << t;
                if (++S.s == AT_NST) { S.s = 0; S.ph ^= 1; }
            }
        }
    } else {
        const int grp = warp >> 2, quad = warp & 3;
        const int etid = threadIdx.x;                            // 0..255
        float bias[COUT];
#pragma unroll
        for (int e = 0; e < COUT; ++e) bias[e] = a.bias[e];
        auto emit = [&](int ypo, int xp, int jc, const float* v) {
            const long long qo = (long long)img * a.go.P + ypo * Wpo + xp;
            store_planes8(a.out, ((long long)jc * a.out.plane_px + qo) * 8, v);
            float rl[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) rl[e] = fmaxf(v[e], 0.f);
            store_planes8(a.out_r, ((long long)jc * a.out_r.plane_px + qo) * 8, rl);
        };
        auto emit_zero = [&](int ypo, int xp, int jc) {
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = 0.f;
            emit(ypo, xp, jc, v);
        };
        for (int b = (int)crank; b < a.bands_per_img; b += (int)csize) {
            const int kb = band_rows(b), nt = band_tiles(kb);
            if (grp < nt) {                                      // tile t = grp of this band (a band has at most two tiles)
                const int t = grp;
                mbar_wait(&B->tfull[t], (S.accph >> t) & 1);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + t * AT_ACC_COLS;
                const int pb = t * AT_TILE_M + quad * 32 + lane;
#pragma unroll
                for (int h = 0; h < COUT / 16; ++h) {
                    float v[16], u[16];
                    tmem_ld16(taddr + COUT + h * 16, u);
                    tmem_ld16(taddr + h * 16, v);
#pragma unroll
                    for (int k = 0; k < 16; ++k) v[k] = fmaf(u[k], MID_INV, v[k]);
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        *at_band_ptr(band, pb, h * 4 + c) = make_float4(v[c * 4 + 0] + bias[h * 16 + c * 4 + 0], v[c * 4 + 1] + bias[h * 16 + c * 4 + 1],
                                                                        v[c * 4 + 2] + bias[h * 16 + c * 4 + 2], v[c * 4 + 3] + bias[h * 16 + c * 4 + 3]);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&B->tempty[t]);
                S.accph ^= 1u << t;
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");       // the band is complete in shared memory
            const int nitems = kb * Wo * NCH;
            for (int it = etid; it < nitems; it += 256) {
                const int j = it % Wo, jc = (it / Wo) % NCH, r = it / (Wo * NCH);
                float v[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = -INFINITY;
                const int y0 = 2 * (b * CP_K + r) - a.pad_lo, x0 = 2 * j - a.pad_lo;
#pragma unroll
                for (int dy = 0; dy < 3; ++dy) {
                    if (y0 + dy < 0 || y0 + dy >= a.gi.H) continue;
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) {
                        if (x0 + dx < 0 || x0 + dx >= a.gi.W) continue;
                        const int pb = (2 * r + dy) * Wp + x0 + dx + 1;
                        const float4 f0 = *at_band_ptr(band, pb, 2 * jc), f1 = *at_band_ptr(band, pb, 2 * jc + 1);
                        const float o[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
#pragma unroll
                        for (int e = 0; e < 8; ++e) v[e] = fmaxf(v[e], o[e]);
                    }
                }
                emit(b * CP_K + r + 1, j + 1, jc, v);
            }
            // shared borders (common.cuh): the zero pixel that starts each of this band's rows, plus the row above the image
            const int nside = kb * NCH;
            const int ntop = (b == 0) ? Wpo * NCH : 0;
            for (int it = etid; it < nside + ntop; it += 256) {
                if (it < nside) {
                    emit_zero(b * CP_K + it / NCH + 1, 0, it % NCH);
                } else {
                    const int k = it - nside;
                    emit_zero(0, k % Wpo, k / Wpo);
                }
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");       // band buffer free for the next band
        }
    }
}

struct ActorTailArgs {
    int n;
    TailPool pool[2];        // sequence conv + pool of ConvSequence 1 (16 -> 32 at 42x42) and 2 (32 -> 32 at 21x21)
    ConvArgs conv[8];        // the four residual-block convs of each sequence
};

__global__ void __launch_bounds__(AT_THREADS, 1) k_actor_tail(const __grid_constant__ ActorTailArgs A) {
    extern __shared__ __align__(1024) uint8_t smem[];
    griddep_launch();
    AtBars* B = reinterpret_cast<AtBars*>(smem);
    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    const uint32_t crank = cluster_ctarank(), csize = cluster_nctarank();
    if (threadIdx.x == 0) {
        for (int s = 0; s < AT_NST; ++s) { mbar_init(&B->full[s], 1); mbar_init(&B->empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&B->tfull[s], 1); mbar_init(&B->tempty[s], 4); mbar_init(&B->wb[s], 1); }
        fence_barrier_init();
    }
    if (warp == 9) tmem_alloc(&B->tmem_slot, AT_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = B->tmem_slot;
    const int nclusters = (int)cluster_count_x();
    // layer k of a frame: 0 = sequence-1 conv + pool, 1..4 its residual convs, 5 = sequence-2 conv + pool, 6..9 its residual convs
    auto wp_of = [&](int k) -> const f16* { return k == 0 ? A.pool[0].wp : (k == 5 ? A.pool[1].wp : A.conv[k < 5 ? k - 1 : k - 2].wp); };
    auto wbytes_of = [&](int k) { return k == 0 ? at_wbytes(2) : at_wbytes(4); };
    AtState S = {0, 0u, 0u, 0u};
    const bool loader = warp == 8 && lane == 0;
    if (loader && (int)cluster_id_x() < A.n) at_load_weights(smem, B, 0, wp_of(0), wbytes_of(0));   // weights do not depend on the
    griddep_wait();                                                                                // preceding kernels
    for (int img = (int)cluster_id_x(); img < A.n; img += nclusters) {
        const bool more = img + nclusters < A.n;
        // the image of the NEXT layer streams into the other weight buffer while this layer computes
        auto prefetch = [&](int knext, bool exists) {
            if (loader && exists) at_load_weights(smem, B, S.layer + 1, wp_of(knext), wbytes_of(knext));
        };
        prefetch(1, true);
        tail_conv_pool<2, 2>(A.pool[0], img, crank, csize, smem, tmem_base, S, warp, lane);
        layer_barrier(csize); ++S.layer;
#pragma unroll 1
        for (int k = 0; k < 4; ++k) {
            prefetch(k + 2, true);
            tail_conv32(A.conv[k], img, crank, csize, smem, tmem_base, S, warp, lane);
            layer_barrier(csize); ++S.layer;
        }
        prefetch(6, true);
        tail_conv_pool<4, 5>(A.pool[1], img, crank, csize, smem, tmem_base, S, warp, lane);
        layer_barrier(csize); ++S.layer;
#pragma unroll 1
        for (int k = 4; k < 8; ++k) {
            prefetch(k == 7 ? 0 : k + 3, k < 7 || more);
            tail_conv32(A.conv[k], img, crank, csize, smem, tmem_base, S, warp, lane);
            layer_barrier(csize); ++S.layer;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 9) tmem_dealloc(tmem_base, AT_TMEM_COLS);
}

int launch_actor_tail(const ActorTailHost& h, int csize, int num_sms, cudaStream_t st) {
    CB_CHECK(csize == 1 || csize == 2, "actor_tail: cluster size %d not built (1 | 2)", csize);
    ActorTailArgs A;
    memset(&A, 0, sizeof(A));
    A.n = h.n;
    for (int s = 0; s < 2; ++s) {
        const ConvArgs& c = h.pool[s];
        CB_CHECK(c.cout == AT_COUT && c.cin_chunks == (s == 0 ? 2 : 4), "actor_tail: sequence conv %d has an unexpected shape", s);
        TailPool& p = A.pool[s];
        p.gi = c.g; p.go = h.pool_go[s]; p.pad_lo = h.pad_lo[s];
        const int K = s == 0 ? 2 : 5;
        p.bands_per_img = (p.go.H + K - 1) / K;
        CB_CHECK(((2 * K + 1) * p.gi.Wp + AT_TILE_M - 1) / AT_TILE_M <= 2, "actor_tail: a band of sequence %d needs more than two tiles", s);
        p.in = c.in; p.wp = c.wp; p.bias = c.ep.bias; p.out = h.pool_out[s]; p.out_r = h.pool_out_r[s];
    }
    for (int k = 0; k < 8; ++k) {
        CB_CHECK(h.conv[k].cout == AT_COUT && h.conv[k].cin_chunks == 4, "actor_tail: residual conv %d has an unexpected shape", k);
        A.conv[k] = h.conv[k];
    }
    static std::atomic<unsigned> attr_done{0};
    int dev = 0;
    CB_CUDA(cudaGetDevice(&dev));
    if (!(attr_done.load() & (1u << dev))) {
        CB_CUDA(cudaFuncSetAttribute(k_actor_tail, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM));
        attr_done.fetch_or(1u << dev);
    }
    int clusters = h.n;
    const int max_clusters = num_sms / csize;
    if (clusters > max_clusters) clusters = max_clusters;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(clusters * csize); cfg.blockDim = dim3(AT_THREADS); cfg.dynamicSmemBytes = AT_SMEM; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int na = 0;
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = csize; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    ++na;
    if (pdl_enabled()) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr; cfg.numAttrs = na;
    (void)cudaLaunchKernelEx(&cfg, k_actor_tail, A);
    CB_LAUNCH_CHECK();
    return 0;
}

}  // namespace cb
