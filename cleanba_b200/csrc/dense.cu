// Dense 3872 -> 256 layer of the trunk (cleanba/cleanba_ppo.py:185-188): forward, dW/db and dX.
// fp32 CUDA-core tiled GEMMs (64x64x32 tiles, 4x4 register micro-tiles) that read / write the chunk-plane layout
// directly, so the NHWC (h,w,c) flatten of the reference is an index map, never a copy.
#include "common.cuh"
#include "kernels.h"

namespace cb {

constexpr int FH = 11, FW = 11, FC = 32, FP = 12 * 12, FWp = 12;   // final feature map (shared borders, common.cuh)
constexpr int DK = FH * FW * FC;                                    // 3872
constexpr int TM = 64, TN = 64, TK = 32, LDS_ = 68;

__device__ __forceinline__ long long feat_pixel(int b, int pix /* h*11+w */) {
    int h = pix / FW, w = pix - h * FW;
    return (long long)b * FP + (long long)(h + 1) * FWp + (w + 1);
}

__device__ __forceinline__ void tile_fma(const float (*As)[LDS_], const float (*Bs)[LDS_], int ty, int tx, float acc[4][4]) {
#pragma unroll 8
    for (int kk = 0; kk < TK; ++kk) {
        float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
        float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
        float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
}

// load 8 channels (one chunk) of sample b, feature pixel `pix` as fp32
__device__ __forceinline__ void load_feat8(const Planes& x, int b, int pix, int chunk, float* f) {
    load_planes8(x, ((long long)chunk * x.plane_px + feat_pixel(b, pix)) * 8, f);
}

// ---------------- forward: part[z][b][j] = sum_{k in split z} x[b][k] W[k][j]
__global__ void __launch_bounds__(256) k_dense_fwd(DenseArgs a, int pix_per_split, float* __restrict__ part) {
    __shared__ __align__(16) float As[TK][LDS_];
    __shared__ __align__(16) float Bs[TK][LDS_];
    const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
    const int pix0 = blockIdx.z * pix_per_split;
    const int pix1 = min(FH * FW, pix0 + pix_per_split);
    const int ty = threadIdx.x / 16, tx = threadIdx.x % 16;
    float acc[4][4] = {};
    for (int pix = pix0; pix < pix1; ++pix) {
        __syncthreads();
        {   // A tile: 64 samples x 32 channels of this pixel
            int s = threadIdx.x / 4, chunk = threadIdx.x % 4;
            float f[8];
            if (m0 + s < a.n) load_feat8(a.x, m0 + s, pix, chunk, f);
            else {
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] = 0.f;
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) As[chunk * 8 + e][s] = f[e];
        }
        for (int t = threadIdx.x; t < TK * TN / 4; t += 256) {
            int kk = t / (TN / 4), c4 = t % (TN / 4);
            float4 w = *reinterpret_cast<const float4*>(a.w + (long long)(pix * FC + kk) * HIDDEN + n0 + c4 * 4);
            *reinterpret_cast<float4*>(&Bs[kk][c4 * 4]) = w;
        }
        __syncthreads();
        tile_fma(As, Bs, ty, tx, acc);
    }
    float* out = part + (long long)blockIdx.z * a.n * HIDDEN;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int b = m0 + ty * 4 + i;
        if (b < a.n)
            *reinterpret_cast<float4*>(out + (long long)b * HIDDEN + n0 + tx * 4) =
                make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    }
}

__global__ void k_dense_finish(const float* __restrict__ part, int nsplit, int n, const float* __restrict__ bias,
                               float* __restrict__ hidden) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)n * HIDDEN) return;
    float s = 0.f;
    for (int z = 0; z < nsplit; ++z) s += part[(long long)z * n * HIDDEN + i];
    s += bias[i % HIDDEN];
    hidden[i] = fmaxf(s, 0.f);
}

int dense_fwd_splits(int n) { return n <= 256 ? 11 : (n <= 1024 ? 4 : 1); }

int launch_dense_fwd(const DenseArgs& a, float* part, cudaStream_t st) {
    int ns = dense_fwd_splits(a.n);
    int pps = (FH * FW + ns - 1) / ns;
    dim3 grid((a.n + TM - 1) / TM, HIDDEN / TN, ns);
    k_dense_fwd<<<grid, 256, 0, st>>>(a, pps, part);
    CB_LAUNCH_CHECK();
    long long tot = (long long)a.n * HIDDEN;
    k_dense_finish<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(part, ns, a.n, a.b, a.hidden);
    CB_LAUNCH_CHECK();
    return 0;
}

// ---------------- dW[k][j] = sum_b x[b][k] dpre[b][j];  one block per (2 feature pixels, 64 outputs)
__global__ void __launch_bounds__(256) k_dense_bwd_w(DenseArgs a, const float* __restrict__ dpre, float* __restrict__ dw) {
    __shared__ __align__(16) float As[TK][LDS_];   // [sample][k]
    __shared__ __align__(16) float Bs[TK][LDS_];   // [sample][j]
    const int k0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
    const int ty = threadIdx.x / 16, tx = threadIdx.x % 16;
    float acc[4][4] = {};
    for (int b0 = 0; b0 < a.n; b0 += TK) {
        __syncthreads();
        {   // 32 samples x 64 k (= 2 pixels x 4 chunks x 8)
            int s = threadIdx.x / 8, sub = threadIdx.x % 8;
            int pix = k0 / FC + sub / 4, chunk = sub % 4;
            float f[8];
            if (b0 + s < a.n && pix < FH * FW) load_feat8(a.x, b0 + s, pix, chunk, f);
            else {
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] = 0.f;
            }
            *reinterpret_cast<float4*>(&As[s][sub * 8]) = make_float4(f[0], f[1], f[2], f[3]);
            *reinterpret_cast<float4*>(&As[s][sub * 8 + 4]) = make_float4(f[4], f[5], f[6], f[7]);
        }
        for (int t = threadIdx.x; t < TK * TN / 4; t += 256) {
            int s = t / (TN / 4), c4 = t % (TN / 4);
            float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
            if (b0 + s < a.n) d = *reinterpret_cast<const float4*>(dpre + (long long)(b0 + s) * HIDDEN + n0 + c4 * 4);
            *reinterpret_cast<float4*>(&Bs[s][c4 * 4]) = d;
        }
        __syncthreads();
        tile_fma(As, Bs, ty, tx, acc);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int k = k0 + ty * 4 + i;
        if (k < DK)
            *reinterpret_cast<float4*>(dw + (long long)k * HIDDEN + n0 + tx * 4) =
                make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    }
}

__global__ void k_colsum(const float* __restrict__ d, int n, int cols, float* __restrict__ out) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= cols) return;
    float s = 0.f;
    for (int b = 0; b < n; ++b) s += d[(long long)b * cols + j];
    out[j] = s;
}

int launch_dense_bwd_w(const DenseArgs& a, const float* dpre, float* dw, float* db, cudaStream_t st) {
    dim3 grid((DK + TM - 1) / TM, HIDDEN / TN);
    k_dense_bwd_w<<<grid, 256, 0, st>>>(a, dpre, dw);
    CB_LAUNCH_CHECK();
    k_colsum<<<(HIDDEN + 63) / 64, 64, 0, st>>>(dpre, a.n, HIDDEN, db);
    CB_LAUNCH_CHECK();
    return 0;
}

// ---------------- dX[b][k] = S * sum_j dpre[b][j] W[k][j], gated by the forward relu (x > 0), written as carrier planes
// (S = *gscale: the loss scale every gradient tensor of the trunk carries)
__global__ void __launch_bounds__(256) k_dense_bwd_x(DenseArgs a, const float* __restrict__ dpre, const float* __restrict__ gscale,
                                                     Planes out) {
    __shared__ __align__(16) float As[TK][LDS_];   // [j][sample]
    __shared__ __align__(16) float Bs[TK][LDS_];   // [j][k]
    const int m0 = blockIdx.x * TM, k0 = blockIdx.y * TN;
    const int ty = threadIdx.x / 16, tx = threadIdx.x % 16;
    float acc[4][4] = {};
    for (int j0 = 0; j0 < HIDDEN; j0 += TK) {
        __syncthreads();
        for (int t = threadIdx.x; t < TM * TK; t += 256) {
            int s = t / TK, jj = t % TK;
            As[jj][s] = (m0 + s < a.n) ? dpre[(long long)(m0 + s) * HIDDEN + j0 + jj] : 0.f;
        }
        for (int t = threadIdx.x; t < TN * TK; t += 256) {
            int kk = t / TK, jj = t % TK;
            Bs[jj][kk] = (k0 + kk < DK) ? a.w[(long long)(k0 + kk) * HIDDEN + j0 + jj] : 0.f;
        }
        __syncthreads();
        tile_fma(As, Bs, ty, tx, acc);
    }
    const float S = *gscale;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int b = m0 + ty * 4 + i;
        int k = k0 + tx * 4;
        if (b >= a.n || k >= DK) continue;
        int pix = k / FC, c = k % FC, chunk = c / 8, e0 = c % 8;   // e0 in {0, 4}
        long long q = feat_pixel(b, pix);
        long long poff = ((long long)chunk * a.x.plane_px + q) * 8 + e0;
        const uint2 m = *reinterpret_cast<const uint2*>(a.x.hi + poff);
        float v[4];
        v[0] = h_pos(m.x & 0xffffu) ? acc[i][0] * S : 0.f;
        v[1] = h_pos(m.x >> 16) ? acc[i][1] * S : 0.f;
        v[2] = h_pos(m.y & 0xffffu) ? acc[i][2] * S : 0.f;
        v[3] = h_pos(m.y >> 16) ? acc[i][3] * S : 0.f;
        uint2 h, md;
        split_f16x2(v[0], v[1], h.x, md.x);
        split_f16x2(v[2], v[3], h.y, md.y);
        long long ooff = ((long long)chunk * out.plane_px + q) * 8 + e0;
        *reinterpret_cast<uint2*>(out.hi + ooff) = h;
        *reinterpret_cast<uint2*>(out.mid + ooff) = md;
    }
}

int launch_dense_bwd_x(const DenseArgs& a, const float* dpre, const float* gscale, Planes out, cudaStream_t st) {
    // borders of the gradient tensors must be zero: clear, then fill the interior
    long long NP = (long long)a.n * FP;
    for (int c = 0; c < FC / 8; ++c) {
        size_t npr = (size_t)((NP + 127) / 128 * 128);   // zero up to the 128-pixel tile boundary (wgrad reads it)
        CB_CUDA(cudaMemsetAsync(out.hi + (long long)c * out.plane_px * 8, 0, npr * 8 * sizeof(f16), st));
        CB_CUDA(cudaMemsetAsync(out.mid + (long long)c * out.plane_px * 8, 0, npr * 8 * sizeof(f16), st));
    }
    dim3 grid((a.n + TM - 1) / TM, (DK + TN - 1) / TN);
    k_dense_bwd_x<<<grid, 256, 0, st>>>(a, dpre, gscale, out);
    CB_LAUNCH_CHECK();
    return 0;
}

}  // namespace cb
