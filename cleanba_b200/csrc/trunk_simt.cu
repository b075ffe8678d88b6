// CUDA-core (fp32 FFMA) kernels of the IMPALA-ResNet trunk: frame unpack, max-pool forward/backward and
// the reference implementation of the 3x3 convolutions (forward, dgrad, wgrad).  The conv kernels here
// are the in-library cross-check for the tcgen05 kernels (conv_umma.cu) and run on the same layouts.
// Reference semantics: cleanba/cleanba_ppo.py:149-189 (Network / ConvSequence / ResidualBlock).
#include "common.cuh"
#include "kernels.h"

namespace cb {

// ------------------------------------------------------------------------------------------------
// uint8 NCHW frame stack -> fp16 chunk plane (8 channels: 4 real + 4 zero), borders zero.
// cleanba_ppo.py:180-181 (transpose + /255; the 1/255 is folded into the conv epilogue, 0..255 are exact in fp16).
// One block per (image, group of UNPACK_ROWS padded rows): the channel rows are staged through shared memory with
// coalesced 4-byte reads, then every thread emits 16-byte pixels (consecutive threads -> consecutive pixels).
constexpr int UNPACK_ROWS = 8;
__global__ void __launch_bounds__(256) k_unpack_frames(const uint8_t* __restrict__ obs, const int* __restrict__ idx, int n,
                                                       f16* __restrict__ out_hi, const cb_rollout_cursor* __restrict__ cursor,
                                                       const StepPtrs* __restrict__ ind) {
    griddep_launch();
    griddep_wait();
    if (ind) { obs = static_cast<const uint8_t*>(ind->p[0]); idx = static_cast<const int*>(ind->p[1]); }
    if (cursor) obs = reinterpret_cast<const uint8_t*>(cursor->obs) + (long long)cursor->row * cursor->obs_row_stride;
    constexpr int H = 84, W = 84, Wp = 85, Hp = 85, GROUPS = (Hp + UNPACK_ROWS - 1) / UNPACK_ROWS;   // shared borders (common.cuh)
    const int img = blockIdx.x / GROUPS;
    const int yp0 = (blockIdx.x % GROUPS) * UNPACK_ROWS;
    __shared__ uint32_t srow[UNPACK_ROWS][4][W / 4];
    const long long src = idx ? (long long)idx[img] : (long long)img;
    const uint8_t* base = obs + src * (4LL * H * W);
    for (int t = threadIdx.x; t < UNPACK_ROWS * 4 * (W / 4); t += blockDim.x) {
        const int w4 = t % (W / 4), c = (t / (W / 4)) % 4, ry = t / (4 * (W / 4));
        const int yp = yp0 + ry;
        if (yp >= 1 && yp <= H)
            srow[ry][c][w4] = *reinterpret_cast<const uint32_t*>(base + ((long long)c * H + (yp - 1)) * W + w4 * 4);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < UNPACK_ROWS * Wp; t += blockDim.x) {
        const int ry = t / Wp, xp = t - ry * Wp;
        const int yp = yp0 + ry;
        if (yp >= Hp) break;
        uint4 o = make_uint4(0, 0, 0, 0);
        if (yp >= 1 && yp <= H && xp >= 1 && xp <= W) {
            const int x = xp - 1;
            const uint8_t* s = reinterpret_cast<const uint8_t*>(&srow[ry][0][0]);
            const float c0 = s[0 * W + x], c1 = s[1 * W + x], c2 = s[2 * W + x], c3 = s[3 * W + x];
            const __half2 h01 = __floats2half2_rn(c0, c1), h23 = __floats2half2_rn(c2, c3);
            o.x = *reinterpret_cast<const uint32_t*>(&h01);
            o.y = *reinterpret_cast<const uint32_t*>(&h23);
        }
        const long long q = (long long)img * (Hp * Wp) + (long long)yp * Wp + xp;
        *reinterpret_cast<uint4*>(out_hi + q * 8) = o;
    }
}

int launch_unpack(const uint8_t* obs, const int* idx, int n, f16* out_hi, cudaStream_t st, const cb_rollout_cursor* cursor,
                  const StepPtrs* ind) {
    const int groups = (85 + UNPACK_ROWS - 1) / UNPACK_ROWS;
    launch_pdl(k_unpack_frames, dim3(n * groups), dim3(256), 0, st, obs, idx, n, out_hi, cursor, ind);
    CB_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Reference 3x3 conv (forward or dgrad): one thread per (flat pixel, output chunk of 8 channels).
__global__ void __launch_bounds__(128) k_conv_simt(ConvArgs a) {
    extern __shared__ float sw[];  // [9][cin_real][8] weights of this output chunk
    const int oc = blockIdx.y;
    const int cin = a.cin_real;
    for (int t = threadIdx.x; t < 9 * cin * 8; t += blockDim.x) {
        int e = t % 8, ci = (t / 8) % cin, tap = t / (8 * cin);
        int co = oc * 8 + e;
        float w;
        if (!a.transpose) w = a.w[((long long)tap * a.w_cin + ci) * a.w_cout + co];
        else w = a.w[((long long)(8 - tap) * a.w_cin + co) * a.w_cout + ci];  // dX[ci'] = sum G[co'] W[flip][ci'][co']
        sw[t] = w;
    }
    __syncthreads();
    long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    if (q < a.g.NP && interior(a.g, q)) {
        for (int tap = 0; tap < 9; ++tap) {
            int d = (tap / 3 - 1) * a.g.Wp + (tap % 3 - 1);
            for (int jc = 0; jc < a.cin_chunks; ++jc) {
                float x[8];
                load_planes8(a.in, ((long long)jc * a.in.plane_px + q + d) * 8, x);
                int nci = min(8, cin - jc * 8);
                for (int c = 0; c < nci; ++c) {
                    const float* wr = sw + ((tap * cin) + jc * 8 + c) * 8;
#pragma unroll
                    for (int e = 0; e < 8; ++e) acc[e] = fmaf(x[c], wr[e], acc[e]);
                }
            }
        }
    }
    conv_epilogue_store(a.ep, a.g, q, oc, acc);
}

int launch_conv_simt(const ConvArgs& a, cudaStream_t st) {
    dim3 grid((unsigned)((a.g.NP + 127) / 128), a.cout / 8);
    size_t smem = (size_t)9 * a.cin_real * 8 * sizeof(float);
    k_conv_simt<<<grid, 128, smem, st>>>(a);
    CB_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------------
// max_pool 3x3 stride 2 SAME, -inf padding, pad_lo = 0 (84->42, 42->21) or 1 (21->11)  (cleanba_ppo.py:168)
// in : carrier planes of the conv output      out: raw + rectified carrier planes on the pooled grid, and
// (training) the arg-max window slot 0..8 of every pooled element as one byte, first maximum in row-major window
// order (XLA select_and_scatter with a `ge` select), which is all the backward pass needs.
__global__ void k_pool_fwd(Planes in, ConvGeom gi, ConvGeom go, int pad_lo, int chunks, Planes out, Planes out_relu,
                           uint8_t* __restrict__ amax) {
    griddep_launch();
    griddep_wait();
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= go.NP * chunks) return;
    int jc = (int)(t / go.NP);
    long long q = t % go.NP;
    int img = (int)(q / go.P);
    int r = (int)(q % go.P);
    int yp = r / go.Wp, xp = r % go.Wp;
    float v[8];
    int am[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { v[e] = 0.f; am[e] = 15; }
    if (yp >= 1 && yp <= go.H && xp >= 1 && xp <= go.W) {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = -INFINITY;
        int i = yp - 1, j = xp - 1;
        for (int dy = 0; dy < 3; ++dy) {
            int y = 2 * i - pad_lo + dy;
            if (y < 0 || y >= gi.H) continue;
            for (int dx = 0; dx < 3; ++dx) {
                int x = 2 * j - pad_lo + dx;
                if (x < 0 || x >= gi.W) continue;
                long long qi = (long long)img * gi.P + (long long)(y + 1) * gi.Wp + (x + 1);
                float o[8];
                load_planes8(in, ((long long)jc * in.plane_px + qi) * 8, o);
#pragma unroll
                for (int e = 0; e < 8; ++e)
                    if (o[e] > v[e]) { v[e] = o[e]; am[e] = dy * 3 + dx; }
            }
        }
    }
    store_planes8(out, ((long long)jc * out.plane_px + q) * 8, v);
    float rl[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) rl[e] = fmaxf(v[e], 0.f);
    store_planes8(out_relu, ((long long)jc * out_relu.plane_px + q) * 8, rl);
    if (amax) {
        uint2 pk;
        pk.x = am[0] | (am[1] << 8) | (am[2] << 16) | (am[3] << 24);
        pk.y = am[4] | (am[5] << 8) | (am[6] << 16) | (am[7] << 24);
        *reinterpret_cast<uint2*>(amax + ((long long)jc * go.NP + q) * 8) = pk;
    }
}

int launch_pool_fwd(Planes in, ConvGeom gi, ConvGeom go, int pad_lo, int chunks, Planes out, Planes out_relu,
                    uint8_t* amax, cudaStream_t st) {
    long long total = go.NP * chunks;
    launch_pdl(k_pool_fwd, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, in, gi, go, pad_lo, chunks, out, out_relu, amax);
    CB_LAUNCH_CHECK();
    return 0;
}

// Backward of the pool in gather form (deterministic, no atomics).  One thread owns the 2x2 input pixels that are the
// slots (0..1, 0..1) of window (i, j); they can only be the arg-max of the four windows (i-1..i, j-1..j), so a thread loads
// 4 windows (arg-max bytes + gradient) for 4 output pixels (the one-pixel-per-thread form loaded 4 windows per pixel).
// Contributions are added in window order (i-1,j-1), (i-1,j), (i,j-1), (i,j).
//   amax: bytes from the forward pass;  dpool: gradient planes on the pooled grid;  out: gradient planes on the input grid
// grid = (quad blocks of one image, image * chunks); quads cover the whole padded grid so the border ring is zero-filled too.
__global__ void __launch_bounds__(256) k_pool_bwd(const uint8_t* __restrict__ amax, Planes dpool,
                                                  ConvGeom gi, ConvGeom go, int pad_lo, int chunks, int KQ, Planes out) {
    griddep_launch();
    griddep_wait();
    const int img = blockIdx.y / chunks, jc = blockIdx.y % chunks;
    const long long pbase = (long long)jc * out.plane_px;
    if (img == gi.n - 1 && blockIdx.x == 0) {
        // zero-fill the planes up to the 128-pixel tile boundary (wgrad / dgrad read whole blocks)
        const long long np_pad = (gi.NP + 127) / 128 * 128;
        const uint4 z = make_uint4(0, 0, 0, 0);
        for (long long q = gi.NP + threadIdx.x; q < np_pad; q += blockDim.x) {
            *reinterpret_cast<uint4*>(out.hi + (pbase + q) * 8) = z;
            if (out.mid) *reinterpret_cast<uint4*>(out.mid + (pbase + q) * 8) = z;
        }
    }
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= KQ * KQ) return;
    const int k = t / KQ, l = t - k * KQ;
    const int i = k - 1 + pad_lo, j = l - 1 + pad_lo;              // window whose first two rows / columns are this quad
    const int yp0 = 2 * k - 1 + pad_lo, xp0 = 2 * l - 1 + pad_lo;  // padded coordinates of the quad's first pixel
    // the four windows: w = di * 2 + dj  <->  (i - 1 + di, j - 1 + dj)
    uint2 pk[4];
    float d[4][8];
    const long long obase = ((long long)jc * go.NP + (long long)img * go.P) * 8;                  // arg-max bytes (no guards)
    const long long dbase = ((long long)jc * dpool.plane_px + (long long)img * go.P) * 8;       // gradient planes
#pragma unroll
    for (int w = 0; w < 4; ++w) {
        const int wi = i - 1 + (w >> 1), wj = j - 1 + (w & 1);
        pk[w] = make_uint2(0x0f0f0f0fu, 0x0f0f0f0fu);
#pragma unroll
        for (int e = 0; e < 8; ++e) d[w][e] = 0.f;
        if (wi >= 0 && wi < go.H && wj >= 0 && wj < go.W) {
            const long long po = (long long)((wi + 1) * go.Wp + (wj + 1)) * 8;
            pk[w] = *reinterpret_cast<const uint2*>(amax + obase + po);
            load_planes8(dpool, dbase + po, d[w]);
        }
    }
    // slot of pixel (a, b) of the quad inside window w (di, dj): row slot = a + 2 * (1 - di) (valid if <= 2), same for columns
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        const int yp = yp0 + a;
        if (yp < 0 || yp >= gi.Hp) continue;
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const int xp = xp0 + b;
            if (xp < 0 || xp >= gi.Wp) continue;
            float g[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) g[e] = 0.f;
            if (yp >= 1 && yp <= gi.H && xp >= 1 && xp <= gi.W) {
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    const int sy = a + 2 * (1 - (w >> 1)), sx = b + 2 * (1 - (w & 1));
                    if (sy > 2 || sx > 2) continue;
                    const int slot = sy * 3 + sx;
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int am = ((e < 4 ? pk[w].x : pk[w].y) >> (8 * (e & 3))) & 0xff;
                        g[e] += (am == slot) ? d[w][e] : 0.f;
                    }
                }
            }
            store_planes8(out, (pbase + (long long)img * gi.P + yp * gi.Wp + xp) * 8, g);
        }
    }
}

int launch_pool_bwd(const uint8_t* amax, Planes dpool, ConvGeom gi, ConvGeom go, int pad_lo, int chunks, Planes out,
                    cudaStream_t st) {
    const int KQ = (gi.Hp + 2 - pad_lo) / 2;   // quads per image side: padded rows 2k - 1 + pad_lo, 2k + pad_lo
    dim3 grid((KQ * KQ + 255) / 256, gi.n * chunks);
    launch_pdl(k_pool_bwd, grid, dim3(256), 0, st, amax, dpool, gi, go, pad_lo, chunks, KQ, out);
    CB_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------------
// First ConvSequence, backward: max-pool backward FUSED with the weight gradient of the frame conv (cleanba_ppo.py:167-168).
// The gradient w.r.t. the 84x84x16 conv output is non-zero only at the arg-max pixel of every pooled element and feeds
// nothing but this weight gradient (the frames need no dX), so it is never materialised:
//     dW[ky][kx][ci][co] = 1/255 * sum_{img, pooled (i,j)} g[img][i][j][co] * X[img][y* + ky - 1][x* + kx - 1][ci]
//     db[co]             =         sum g[img][i][j][co]                        (y*, x*) = arg-max pixel of (i, j, co)
// X is exact in bf16 (frames 0..255) and the products are accumulated in fp32 registers, so this is also more precise than the
// 2-plane tensor-core wgrad it replaces; HBM traffic drops from ~1.2 MB to ~0.27 MB per image (no gradient planes at 84x84).
// One block per image (persistent over images): the 4 real channels of the padded frame are staged in shared memory
// (8 bytes per pixel); a half-warp of 16 threads owns one pooled pixel (thread = output channel, 36 + 1 accumulators that
// live across all images of the block).  One partial per block, reduced in a fixed order by k_partial_reduce.
constexpr int PW0_THREADS = 256;
constexpr int PW0_HP = 85, PW0_P = PW0_HP * PW0_HP, PW0_HO = 42, PW0_WPO = 43, PW0_PO = PW0_WPO * PW0_WPO;   // shared borders (common.cuh)
constexpr int PW0_OUT = 37 * 16;      // 36 (tap, ci) x 16 co weight gradients + 16 bias gradients per partial
constexpr int PW0_RB = 3;                                   // pooled rows per staged band
constexpr int PW0_BANDS = PW0_HO / PW0_RB;                  // 14 bands per image
constexpr int PW0_BPIX = PW0_RB * PW0_HO;                   // 126 pooled pixels per band
constexpr int PW0_STAGED = PW0_P + PW0_HP + 1;               // the frame plus the zero row / pixel that follow it (taps of the last row)
constexpr int PW0_FRAME = (PW0_STAGED + 1) / 2 * 2 * 8;       // the staged frame (4 fp16 channels per pixel), 16-byte multiple
constexpr int PW0_SMEM = PW0_FRAME + PW0_BPIX * 16 * 4 + PW0_BPIX * 16;   // + band gradients (fp32) + band arg-max bytes
static_assert(PW0_HO % PW0_RB == 0 && PW0_BPIX * 2 <= PW0_THREADS, "one staging item (pixel, chunk) per thread");

__global__ void __launch_bounds__(PW0_THREADS, 3) k_pool_bwd_wgrad0(const uint8_t* __restrict__ amax, Planes dpool,
                                                                     const f16* __restrict__ x_hi, int n, long long go_NP,
                                                                     float* __restrict__ partial) {
    extern __shared__ __align__(16) uint8_t pw_smem[];
    uint2* sx = reinterpret_cast<uint2*>(pw_smem);
    float* sg = reinterpret_cast<float*>(pw_smem + PW0_FRAME);                       // [band pixel][16 channels]
    uint8_t* sam = pw_smem + PW0_FRAME + PW0_BPIX * 16 * 4;                          // [band pixel][16 channels]
    griddep_launch();
    griddep_wait();
    const int co = threadIdx.x & 15, grp = threadIdx.x >> 4;
    float acc[36];
#pragma unroll
    for (int k = 0; k < 36; ++k) acc[k] = 0.f;
    float accb = 0.f;
    // Staging item of this thread: (band pixel sp, chunk sc).  The pooled gradient planes and the arg-max bytes of a band are
    // fetched with coalesced 16 / 8-byte loads one band AHEAD (registers), combined to fp32 and parked in shared memory, so the
    // per-(pixel, channel) reads of the accumulation loop are shared-memory reads (the carrier made them three scalar global
    // loads per element).
    const bool stager = threadIdx.x < PW0_BPIX * 2;
    const int spx = threadIdx.x >> 1, sc = threadIdx.x & 1;
    uint4 r_hi = make_uint4(0, 0, 0, 0), r_mid = make_uint4(0, 0, 0, 0);
    uint2 r_am = make_uint2(0, 0);
    auto fetch = [&](int img, int band) {
        if (!stager) return;
        const int row = band * PW0_RB + spx / PW0_HO, col = spx % PW0_HO;
        const long long q = (long long)img * PW0_PO + (long long)(row + 1) * PW0_WPO + (col + 1);
        r_hi = *reinterpret_cast<const uint4*>(dpool.hi + ((long long)sc * dpool.plane_px + q) * 8);
        r_mid = *reinterpret_cast<const uint4*>(dpool.mid + ((long long)sc * dpool.plane_px + q) * 8);
        r_am = *reinterpret_cast<const uint2*>(amax + ((long long)sc * go_NP + q) * 8);
    };
    const int nwork = ((n - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x) * PW0_BANDS;     // (image, band) items of this block
    if (nwork > 0) fetch(blockIdx.x, 0);
    for (int w = 0; w < nwork; ++w) {
        const int img = blockIdx.x + (w / PW0_BANDS) * gridDim.x, band = w % PW0_BANDS;
        __syncthreads();                                    // the previous band (and frame) is no longer read
        if (band == 0) {
            const uint4* src = reinterpret_cast<const uint4*>(x_hi + (long long)img * PW0_P * 8);
            for (int q = threadIdx.x; q < PW0_STAGED; q += PW0_THREADS) {
                const uint4 v = src[q];
                sx[q] = make_uint2(v.x, v.y);
            }
        }
        if (stager) {
            float gh[8], gm[8];
            unpack8h(r_hi, gh);
            unpack8h(r_mid, gm);
            float4* d = reinterpret_cast<float4*>(sg + spx * 16 + sc * 8);
            d[0] = make_float4(fmaf(gm[0], MID_INV, gh[0]), fmaf(gm[1], MID_INV, gh[1]), fmaf(gm[2], MID_INV, gh[2]), fmaf(gm[3], MID_INV, gh[3]));
            d[1] = make_float4(fmaf(gm[4], MID_INV, gh[4]), fmaf(gm[5], MID_INV, gh[5]), fmaf(gm[6], MID_INV, gh[6]), fmaf(gm[7], MID_INV, gh[7]));
            *reinterpret_cast<uint2*>(sam + spx * 16 + sc * 8) = r_am;
        }
        __syncthreads();
        if (w + 1 < nwork) fetch(blockIdx.x + ((w + 1) / PW0_BANDS) * gridDim.x, (w + 1) % PW0_BANDS);   // in flight during the FMAs
        constexpr int PSTEP = PW0_THREADS / 16;
        for (int p = grp; p < PW0_BPIX; p += PSTEP) {       // uniform trip count per warp (126 = 7 * 16 + 14: the tail is per half-warp)
            const int i = band * PW0_RB + p / PW0_HO, j = p % PW0_HO;
            const int am = sam[p * 16 + co];
            const float gc = sg[p * 16 + co];
            const int dy = (am * 11) >> 5, dx = am - 3 * dy;   // am = dy * 3 + dx, 0..8
            // tap (ky, kx) of arg-max pixel (yp, xp) = (2i + dy + 1, 2j + dx + 1) reads padded pixel (yp + ky - 1, xp + kx - 1)
            const uint2* w0 = sx + (2 * i + dy) * PW0_HP + (2 * j + dx);
            accb += gc;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const uint2 v = w0[ky * PW0_HP + kx];
                    float* a = acc + (ky * 3 + kx) * 4;
                    const float2 x01 = h2_to_f2(v.x), x23 = h2_to_f2(v.y);
                    a[0] = fmaf(x01.x, gc, a[0]);
                    a[1] = fmaf(x01.y, gc, a[1]);
                    a[2] = fmaf(x23.x, gc, a[2]);
                    a[3] = fmaf(x23.y, gc, a[3]);
                }
        }
    }
    // the two pixel groups of a warp, then the 8 warps in a fixed order
#pragma unroll
    for (int k = 0; k < 36; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], 16);
    accb += __shfl_xor_sync(0xffffffffu, accb, 16);
    __syncthreads();
    float* sp = reinterpret_cast<float*>(pw_smem);     // [8][PW0_OUT]
    const int warp = threadIdx.x >> 5;
    if ((threadIdx.x & 31) < 16) {
#pragma unroll
        for (int k = 0; k < 36; ++k) sp[warp * PW0_OUT + k * 16 + co] = acc[k];
        sp[warp * PW0_OUT + 36 * 16 + co] = accb;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < PW0_OUT; t += PW0_THREADS) {
        float s = sp[t];
#pragma unroll
        for (int w = 1; w < 8; ++w) s += sp[w * PW0_OUT + t];
        partial[(long long)blockIdx.x * PW0_OUT + t] = s;
    }
}

// out[i] = sum_b partial[b][i] in a fixed order: warp w of 32 adds partials w, w + 32, ..., then the 32 warp sums are added in
// order.  The first nw outputs are scaled and go to dw, the rest to db.
__global__ void __launch_bounds__(1024) k_partial_reduce(const float* __restrict__ partial, int nparts, int count, int nw, float scale,
                                                         const float* __restrict__ inv_scale, float* __restrict__ dw, float* __restrict__ db) {
    __shared__ float sm[32][32];
    griddep_launch();
    griddep_wait();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + lane;
    float s = 0.f;
    if (i < count) {
#pragma unroll 4
        for (int b = w; b < nparts; b += 32) s += partial[(long long)b * count + i];
    }
    sm[w][lane] = s;
    __syncthreads();
    if (w == 0 && i < count) {
        float t = sm[0][lane];
#pragma unroll
        for (int k = 1; k < 32; ++k) t += sm[k][lane];
        const float inv = inv_scale ? *inv_scale : 1.f;      // the gradient planes carry the loss scale
        if (i < nw) dw[i] = t * scale * inv; else db[i - nw] = t * inv;
    }
}

int launch_pool_bwd_wgrad0(const uint8_t* amax, Planes dpool, const f16* x_hi, ConvGeom gi, ConvGeom go, float scale,
                           const float* inv_scale, float* dw, float* db, float* partial, int num_sms, cudaStream_t st) {
    CB_CHECK(gi.H == 84 && gi.W == 84 && go.H == PW0_HO && go.W == PW0_HO, "pool_bwd_wgrad0: first-stage geometry (84 -> 42) only");
    static std::atomic<unsigned> attr_done{0};
    int dev = 0;
    CB_CUDA(cudaGetDevice(&dev));
    if (!(attr_done.load() & (1u << dev))) {
        CB_CUDA(cudaFuncSetAttribute(k_pool_bwd_wgrad0, cudaFuncAttributeMaxDynamicSharedMemorySize, PW0_SMEM));
        attr_done.fetch_or(1u << dev);
    }
    const int grid = gi.n < 3 * num_sms ? gi.n : 3 * num_sms;
    launch_pdl(k_pool_bwd_wgrad0, dim3(grid), dim3(PW0_THREADS), (size_t)PW0_SMEM, st, amax, dpool, x_hi, gi.n, go.NP, partial);
    CB_LAUNCH_CHECK();
    launch_pdl(k_partial_reduce, dim3((PW0_OUT + 31) / 32), dim3(1024), 0, st, (const float*)partial, grid, PW0_OUT, 36 * 16, scale, inv_scale, dw, db);
    CB_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Reference wgrad: dW[tap][ci][co] = sum_q X[q + d_tap][ci] * G[q][co],  db[co] = sum_q G[q][co].
// Each block reduces a slab of TP-pixel tiles into registers, then writes one partial; a second kernel sums
// the partials in a fixed order (deterministic).
constexpr int WG_TP = 64;
constexpr int WG_MAXP = 36;   // (tap, ci) pairs per thread: 288 pairs / 8 groups

__global__ void __launch_bounds__(256) k_wgrad_simt(WgradArgs a, float* __restrict__ partial) {
    extern __shared__ float sm[];
    const int cin = a.cin_real, cout = a.cout;
    const int Wp = a.g.Wp;
    const int win = WG_TP + 2 * Wp + 2;
    float* sx = sm;                 // [win][cin]
    float* sg = sm + win * cin;     // [WG_TP][cout]
    const int co = threadIdx.x % cout;
    const int grp = threadIdx.x / cout;
    const int ngrp = 256 / cout;
    const int npairs = 9 * cin;
    float acc[WG_MAXP];
#pragma unroll
    for (int k = 0; k < WG_MAXP; ++k) acc[k] = 0.f;
    float accb = 0.f;
    const long long ntiles = (a.g.NP + WG_TP - 1) / WG_TP;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        long long q0 = tile * WG_TP;
        __syncthreads();
        for (int t = threadIdx.x; t < win * a.cin_chunks; t += 256) {
            int jc = t / win, p = t % win;
            long long q = q0 - Wp - 1 + p;  // guard zones make this in-bounds
            float x[8];
            load_planes8(a.x, ((long long)jc * a.x.plane_px + q) * 8, x);
            for (int e = 0; e < 8; ++e) {
                int c = jc * 8 + e;
                if (c < cin) sx[p * cin + c] = x[e];
            }
        }
        for (int t = threadIdx.x; t < WG_TP * (cout / 8); t += 256) {
            int jc = t / WG_TP, p = t % WG_TP;
            long long q = q0 + p;
            float gg[8];
            if (q < a.g.NP) {
                load_planes8(a.gy, ((long long)jc * a.gy.plane_px + q) * 8, gg);
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) gg[e] = 0.f;
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) sg[p * cout + jc * 8 + e] = gg[e];
        }
        __syncthreads();
        for (int p = 0; p < WG_TP; ++p) {
            float gv = sg[p * cout + co];
            if (grp == 0) accb += gv;
#pragma unroll
            for (int k = 0; k < WG_MAXP; ++k) {
                int pair = grp + k * ngrp;
                if (pair < npairs) {
                    int tap = pair / cin, ci = pair - tap * cin;
                    int pp = p + (tap / 3) * Wp + (tap % 3);
                    acc[k] = fmaf(sx[pp * cin + ci], gv, acc[k]);
                }
            }
        }
    }
    float* out = partial + (long long)blockIdx.x * (npairs * cout + cout);
#pragma unroll
    for (int k = 0; k < WG_MAXP; ++k) {
        int pair = grp + k * ngrp;
        if (pair < npairs) out[pair * cout + co] = acc[k];
    }
    if (grp == 0) out[npairs * cout + co] = accb;
}

// out[i] = sum_b partial[b][i]  (fixed order);  first nw entries -> dW (HWIO), last cout entries -> db
__global__ void k_wgrad_reduce(const float* __restrict__ partial, int nblocks, int nw, int cout, float scale,
                               const float* __restrict__ inv_scale, float* __restrict__ dw, float* __restrict__ db) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nw + cout) return;
    float s = 0.f;
    for (int b = 0; b < nblocks; ++b) s += partial[(long long)b * (nw + cout) + i];
    const float inv = inv_scale ? *inv_scale : 1.f;
    if (i < nw) dw[i] = s * scale * inv;
    else db[i - nw] = s * inv;
}

int launch_wgrad_simt(const WgradArgs& a, float* partial, int max_blocks, cudaStream_t st) {
    long long ntiles = (a.g.NP + WG_TP - 1) / WG_TP;
    int nb = (int)(ntiles < max_blocks ? ntiles : max_blocks);
    int win = WG_TP + 2 * a.g.Wp + 2;
    size_t smem = ((size_t)win * a.cin_real + (size_t)WG_TP * a.cout) * sizeof(float);
    k_wgrad_simt<<<nb, 256, smem, st>>>(a, partial);
    CB_LAUNCH_CHECK();
    int nw = 9 * a.cin_real * a.cout;
    k_wgrad_reduce<<<(nw + a.cout + 255) / 256, 256, 0, st>>>(partial, nb, nw, a.cout, a.scale, a.inv_scale, a.dw, a.db);
    CB_LAUNCH_CHECK();
    return 0;
}

}  // namespace cb
