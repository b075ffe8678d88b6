// Nature-CNN trunk (cleanba/legacy_scripts/cleanba_ppo_envpool_impala_atari_wrapper_naturecnn.py:143-178) behind the same
// context / C ABI as the IMPALA-ResNet:  x/255 -> Conv(32, 8x8, s4, VALID) -> relu -> Conv(64, 4x4, s2, VALID) -> relu ->
// Conv(64, 3x3, s1, VALID) -> relu -> flatten (h, w, c) -> Dense(512) -> relu.   cb_config.model = CB_MODEL_NATURE.
//
// Every layer is a GEMM over an im2col matrix held as carrier row planes (gemm_umma.cu): the strided VALID convolutions do
// not have the contiguous shifted windows of the ResNet's 3x3 SAME convs, so the patches are gathered once per layer
// (16-byte channel chunks; the frame patches straight from the uint8 frames, exact in one fp16 plane) and the same matrix
// serves the forward GEMM and the weight-gradient GEMM.  The dense layer is the 7x7 "convolution" of the last feature map
// (its im2col is the NHWC flatten).  dgrad = GEMM with the transposed weights into im2col space + a col2im gather that also
// applies the relu gate.  No residuals, no pooling.  Activations / gradients: fp16x2 carriers, gradients loss-scaled (common.cuh).
#include <stdio.h>
#include <string.h>

#include "ctx.h"

namespace cb {

struct NatLayer {
    int KH, stride, Cin, Cout, Hin, Hout, K;      // square kernels / maps; K = KH * KH * Cin
    long long off_b, off_w;                       // flat parameter offsets
    f16 *w_fwd, *w_dg;                            // packed weight images (forward; transposed for dgrad)
    int NBf, NBd;                                 // N block of the forward / dgrad GEMM
};

struct NatureNet {
    NatLayer L[4];                                // conv1, conv2, conv3, dense (as a 7x7 conv with a 1x1 output)
    long long rpad[4];                            // padded rows of layer l's im2col / output (max_batch * Hout^2, to 128)
    f16 *A_hi[4], *A_mid[4];                      // im2col matrices [K / 8][rpad][8]  (A[0]: frames, hi only)
    f16 *act_hi[3], *act_mid[3];                  // relu'd outputs of conv1..3  [Cout / 8][rpad][8]
    f16 *dA_hi[4], *dA_mid[4];                    // dgrad outputs in im2col space (layers 1..3)
    f16 *g_hi[3], *g_mid[3];                      // gradients w.r.t. the pre-relu outputs of conv1..3 (loss-scaled)
    f16 *dp_hi, *dp_mid;                          // loss-scaled dpre as row planes [512 / 8][rpad[3]][8]
    float* col_part;                              // column-sum partials
};

static const int kNatKH[4] = {8, 4, 3, 7}, kNatStride[4] = {4, 2, 1, 1}, kNatCin[4] = {4, 32, 64, 64}, kNatCout[4] = {32, 64, 64, 512};
static const int kNatHin[4] = {84, 20, 9, 7}, kNatHout[4] = {20, 9, 7, 1};

std::vector<Leaf> nature_leaves(int A) {
    std::vector<Leaf> L;
    long long off = 0;
    auto add = [&](const std::string& n, std::initializer_list<int> shp) {
        Leaf l;
        l.name = n; l.offset = off; l.ndim = (int)shp.size();
        int i = 0;
        for (int s : shp) l.shape[i++] = s;
        for (; i < 4; ++i) l.shape[i] = 1;
        off += l.size();
        L.push_back(l);
    };
    for (int l = 0; l < 3; ++l) {
        std::string p = "network_params/params/Conv_" + std::to_string(l);
        add(p + "/bias", {kNatCout[l]});
        add(p + "/kernel", {kNatKH[l], kNatKH[l], kNatCin[l], kNatCout[l]});
    }
    add("network_params/params/Dense_0/bias", {512});
    add("network_params/params/Dense_0/kernel", {3136, 512});
    add("actor_params/params/Dense_0/bias", {A});
    add("actor_params/params/Dense_0/kernel", {512, A});
    add("critic_params/params/Dense_0/bias", {1});
    add("critic_params/params/Dense_0/kernel", {512, 1});
    return L;
}

// ------------------------------------------------------------------------------------------------ kernels
// Frame patches: A0[(ky * 8 + kx) * 4 + c][r] = frame[c][4 oy + ky][4 ox + kx], r = (img, oy, ox).  One thread per (r, ky): the
// 8 x-positions of the 4 channels are 4 aligned 8-byte reads; the thread writes the 4 chunks (2 pixels x 4 channels each).
__global__ void __launch_bounds__(256) k_nat_im2col_frames(const uint8_t* __restrict__ obs, const int* __restrict__ idx, long long R,
                                                           long long rpad, long long rows_used, f16* __restrict__ out,
                                                           const cb_rollout_cursor* __restrict__ cursor,
                                                           const StepPtrs* __restrict__ ind) {
    griddep_launch();
    griddep_wait();
    if (ind) { obs = static_cast<const uint8_t*>(ind->p[0]); idx = static_cast<const int*>(ind->p[1]); }
    if (cursor) obs = reinterpret_cast<const uint8_t*>(cursor->obs) + (long long)cursor->row * cursor->obs_row_stride;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rows_used * 8) return;
    const long long r = t % rows_used;
    const int ky = (int)(t / rows_used);
    uint4 o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = make_uint4(0, 0, 0, 0);
    if (r < R) {
        const int img = (int)(r / 400), p = (int)(r % 400), oy = p / 20, ox = p % 20;
        const long long src = idx ? (long long)idx[img] : (long long)img;
        const uint8_t* base = obs + src * (4LL * 84 * 84) + (long long)(4 * oy + ky) * 84 + 4 * ox;
        __align__(8) uint8_t px[4][8];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const uint2 v = make_uint2(*reinterpret_cast<const uint32_t*>(base + (long long)c * 84 * 84),
                                       *reinterpret_cast<const uint32_t*>(base + (long long)c * 84 * 84 + 4));
            *reinterpret_cast<uint2*>(px[c]) = v;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {      // chunk j: pixels kx = 2j, 2j + 1
            const __half2 a = __floats2half2_rn((float)px[0][2 * j], (float)px[1][2 * j]), b = __floats2half2_rn((float)px[2][2 * j], (float)px[3][2 * j]);
            const __half2 c2 = __floats2half2_rn((float)px[0][2 * j + 1], (float)px[1][2 * j + 1]), d = __floats2half2_rn((float)px[2][2 * j + 1], (float)px[3][2 * j + 1]);
            o[j] = make_uint4(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b), *reinterpret_cast<const uint32_t*>(&c2),
                              *reinterpret_cast<const uint32_t*>(&d));
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(out + ((long long)(ky * 4 + j) * rpad + r) * 8) = o[j];
}

// Patches of a carrier tensor: dst[(ky * KH + kx) * Cc + c8][r] = src[c8][(img, s oy + ky, s ox + kx)]  (16-byte chunks, both planes)
__global__ void __launch_bounds__(256) k_nat_im2col(const f16* __restrict__ s_hi, const f16* __restrict__ s_mid, long long s_rpad,
                                                    f16* __restrict__ d_hi, f16* __restrict__ d_mid, long long d_rpad, long long rows_used,
                                                    long long R, int KH, int stride, int Cc, int Hin, int Hout) {
    griddep_launch();
    griddep_wait();
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int Kc = KH * KH * Cc;
    if (t >= rows_used * Kc) return;
    const long long r = t % rows_used;
    const int kc = (int)(t / rows_used);
    uint4 h = make_uint4(0, 0, 0, 0), m = make_uint4(0, 0, 0, 0);
    if (r < R) {
        const int tap = kc / Cc, c8 = kc % Cc, ky = tap / KH, kx = tap % KH;
        const int P = Hout * Hout;
        const int img = (int)(r / P), p = (int)(r % P), oy = p / Hout, ox = p % Hout;
        const long long q = ((long long)img * Hin + (stride * oy + ky)) * Hin + (stride * ox + kx);
        h = *reinterpret_cast<const uint4*>(s_hi + ((long long)c8 * s_rpad + q) * 8);
        m = *reinterpret_cast<const uint4*>(s_mid + ((long long)c8 * s_rpad + q) * 8);
    }
    *reinterpret_cast<uint4*>(d_hi + ((long long)kc * d_rpad + r) * 8) = h;
    *reinterpret_cast<uint4*>(d_mid + ((long long)kc * d_rpad + r) * 8) = m;
}

// col2im in gather form (deterministic) + relu gate: g[c8][q] = (act[c8][q] > 0) * sum over the taps whose window covers q of
// dA[(tap, c8)][(img, oy, ox)].   q = (img, y, x) on the layer's INPUT grid.
__global__ void __launch_bounds__(256) k_nat_col2im(const f16* __restrict__ a_hi, const f16* __restrict__ a_mid, long long a_rpad,
                                                    const f16* __restrict__ mask_hi, f16* __restrict__ g_hi, f16* __restrict__ g_mid,
                                                    long long g_rpad, long long rows_used, long long Rin, int KH, int stride, int Cc, int Hin,
                                                    int Hout) {
    griddep_launch();
    griddep_wait();
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rows_used * Cc) return;
    const long long q = t % rows_used;
    const int c8 = (int)(t / rows_used);
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = 0.f;
    Planes o;
    o.hi = g_hi; o.mid = g_mid; o.plane_px = g_rpad;
    if (q < Rin) {
        const int P = Hin * Hin;
        const int img = (int)(q / P), p = (int)(q % P), y = p / Hin, x = p % Hin;
        Planes A;
        A.hi = const_cast<f16*>(a_hi); A.mid = const_cast<f16*>(a_mid); A.plane_px = a_rpad;
        for (int ky = 0; ky < KH; ++ky) {
            const int ty = y - ky;
            if (ty < 0 || ty % stride) continue;
            const int oy = ty / stride;
            if (oy >= Hout) continue;
            for (int kx = 0; kx < KH; ++kx) {
                const int tx = x - kx;
                if (tx < 0 || tx % stride) continue;
                const int ox = tx / stride;
                if (ox >= Hout) continue;
                const long long r = ((long long)img * Hout + oy) * Hout + ox;
                float d[8];
                load_planes8(A, ((long long)((ky * KH + kx) * Cc + c8) * a_rpad + r) * 8, d);
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] += d[e];
            }
        }
        gate8h(*reinterpret_cast<const uint4*>(mask_hi + ((long long)c8 * g_rpad + q) * 8), v);
    }
    store_planes8(o, ((long long)c8 * g_rpad + q) * 8, v);
}

// dpre [n][512] fp32 -> loss-scaled carrier row planes [64][rpad][8] (rows >= n zero)
__global__ void __launch_bounds__(256) k_nat_dpre_planes(const float* __restrict__ dpre, int n, long long rpad, long long rows_used,
                                                         const float* __restrict__ gscale, f16* __restrict__ hi, f16* __restrict__ mid) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rows_used * 64) return;
    const long long r = t % rows_used;
    const int jc = (int)(t / rows_used);
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = 0.f;
    if (r < n) {
        const float S = gscale[0];
        const float4* p = reinterpret_cast<const float4*>(dpre + r * 512 + jc * 8);
        const float4 x = p[0], y = p[1];
        v[0] = x.x * S; v[1] = x.y * S; v[2] = x.z * S; v[3] = x.w * S; v[4] = y.x * S; v[5] = y.y * S; v[6] = y.z * S; v[7] = y.w * S;
    }
    Planes o;
    o.hi = hi; o.mid = mid; o.plane_px = rpad;
    store_planes8(o, ((long long)jc * rpad + r) * 8, v);
}

// bias gradients: db[c] = inv * sum_r g[c / 8][r][c % 8], two deterministic stages (NAT_CS slices per chunk)
constexpr int NAT_CS = 64;
__global__ void __launch_bounds__(256) k_nat_colsum_partial(const f16* __restrict__ g_hi, const f16* __restrict__ g_mid, long long rpad, long long R,
                                                            float* __restrict__ part) {
    __shared__ float red[256][8];
    const int c8 = blockIdx.y, slice = blockIdx.x;
    const long long per = (R + NAT_CS - 1) / NAT_CS, lo = slice * per, hi_r = lo + per < R ? lo + per : R;
    float s[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) s[e] = 0.f;
    Planes G;
    G.hi = const_cast<f16*>(g_hi); G.mid = const_cast<f16*>(g_mid); G.plane_px = rpad;
    for (long long r = lo + threadIdx.x; r < hi_r; r += 256) {
        float d[8];
        load_planes8(G, ((long long)c8 * rpad + r) * 8, d);
#pragma unroll
        for (int e = 0; e < 8; ++e) s[e] += d[e];
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) red[threadIdx.x][e] = s[e];
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o)
#pragma unroll
            for (int e = 0; e < 8; ++e) red[threadIdx.x][e] += red[threadIdx.x + o][e];
        __syncthreads();
    }
    if (threadIdx.x < 8) part[((long long)c8 * NAT_CS + slice) * 8 + threadIdx.x] = red[0][threadIdx.x];
}
__global__ void k_nat_colsum_final(const float* __restrict__ part, int C, const float* __restrict__ inv_scale, float* __restrict__ db) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float s = 0.f;
    for (int i = 0; i < NAT_CS; ++i) s += part[((long long)(c / 8) * NAT_CS + i) * 8 + (c % 8)];
    db[c] = s * (inv_scale ? *inv_scale : 1.f);
}

// ------------------------------------------------------------------------------------------------ host side
static inline long long pad128(long long r) { return (r + 127) / 128 * 128; }

static int alloc_planes(cb_ctx* c, f16** hi, f16** mid, long long chunks, long long rpad, bool two = true) {
    void* p;
    if (dev_alloc(c, &p, (size_t)chunks * rpad * 16)) return -1;
    *hi = (f16*)p;
    if (two) {
        if (dev_alloc(c, &p, (size_t)chunks * rpad * 16)) return -1;
        *mid = (f16*)p;
    } else if (mid) {
        *mid = nullptr;
    }
    return 0;
}

int nature_create(cb_ctx* c) {
    NatureNet* N = new NatureNet();
    memset(N, 0, sizeof(*N));
    c->nat = N;
    const long long mb = c->cfg.max_batch;
    for (int l = 0; l < 4; ++l) {
        NatLayer& L = N->L[l];
        L.KH = kNatKH[l]; L.stride = kNatStride[l]; L.Cin = kNatCin[l]; L.Cout = kNatCout[l]; L.Hin = kNatHin[l]; L.Hout = kNatHout[l];
        L.K = L.KH * L.KH * L.Cin;
        L.off_b = c->leaves[2 * l].offset; L.off_w = c->leaves[2 * l + 1].offset;
        L.NBf = l == 0 ? 32 : (l == 3 ? 128 : 64);
        L.NBd = 64;
        N->rpad[l] = pad128(mb * L.Hout * L.Hout);
        void* p;
        if (dev_alloc(c, &p, gemm_pack_elems(L.K, L.Cout, 0, L.NBf) * sizeof(f16))) return -1;
        L.w_fwd = (f16*)p;
        if (l > 0 && c->cfg.train) {
            if (dev_alloc(c, &p, gemm_pack_elems(L.K, L.Cout, 1, L.NBd) * sizeof(f16))) return -1;
            L.w_dg = (f16*)p;
        }
        if (alloc_planes(c, &N->A_hi[l], &N->A_mid[l], L.K / 8, N->rpad[l], l > 0)) return -1;
        if (l < 3 && alloc_planes(c, &N->act_hi[l], &N->act_mid[l], L.Cout / 8, N->rpad[l])) return -1;
        if (c->cfg.train) {
            if (l > 0 && alloc_planes(c, &N->dA_hi[l], &N->dA_mid[l], L.K / 8, N->rpad[l])) return -1;
            if (l < 3 && alloc_planes(c, &N->g_hi[l], &N->g_mid[l], L.Cout / 8, N->rpad[l])) return -1;
        }
    }
    if (c->cfg.train) {
        if (alloc_planes(c, &N->dp_hi, &N->dp_mid, 64, N->rpad[3])) return -1;
        void* p;
        if (dev_alloc(c, &p, (size_t)64 * NAT_CS * 8 * sizeof(float))) return -1;
        N->col_part = (float*)p;
    }
    c->off_dense_b = c->leaves[6].offset; c->off_dense_w = c->leaves[7].offset;
    c->off_actor_b = c->leaves[8].offset; c->off_actor_w = c->leaves[9].offset;
    c->off_critic_b = c->leaves[10].offset; c->off_critic_w = c->leaves[11].offset;
    return 0;
}

void nature_destroy(cb_ctx* c) {
    delete c->nat;
    c->nat = nullptr;
}

int nature_refresh_weights(cb_ctx* c, cudaStream_t st) {
    NatureNet* N = c->nat;
    ProfScope ps(c, "pack_weights", 0, 1684128.0 * (4 + 8), st);
    for (int l = 0; l < 4; ++l) {
        NatLayer& L = N->L[l];
        if (launch_pack_gemm(c->params + L.off_w, L.K, L.Cout, 0, L.NBf, L.w_fwd, st)) return -1;
        if (L.w_dg && launch_pack_gemm(c->params + L.off_w, L.K, L.Cout, 1, L.NBd, L.w_dg, st)) return -1;
    }
    return 0;
}

static inline unsigned blocks_for(long long items) { return (unsigned)((items + 255) / 256); }

int nature_forward(cb_ctx* c, const uint8_t* obs, const int* idx, int n, cudaStream_t st) {
    NatureNet* N = c->nat;
    for (int l = 0; l < 4; ++l) {
        NatLayer& L = N->L[l];
        const long long R = (long long)n * L.Hout * L.Hout, Rp = pad128(R);
        const long long rp = N->rpad[l];
        {
            char name[64];
            snprintf(name, sizeof(name), "im2col@%d", l);
            const double moved = l == 0 ? (double)n * 28224.0 + (double)Rp * L.K * 2 : (double)Rp * L.K * 4 * 2;
            ProfScope ps(c, name, 0, moved, st, l == 0 ? (double)n * 28224.0 : 4.0 * n * L.Hin * L.Hin * L.Cin);
            if (l == 0) {
                launch_pdl(k_nat_im2col_frames, dim3(blocks_for(Rp * 8)), dim3(256), 0, st, obs, idx, R, rp, Rp, N->A_hi[0], c->cursor, c->ind);
            } else {
                launch_pdl(k_nat_im2col, dim3(blocks_for(Rp * (L.K / 8))), dim3(256), 0, st, (const f16*)N->act_hi[l - 1], (const f16*)N->act_mid[l - 1],
                           N->rpad[l - 1], N->A_hi[l], N->A_mid[l], rp, Rp, R, L.KH, L.stride, L.Cin / 8, L.Hin, L.Hout);
            }
            CB_LAUNCH_CHECK();
        }
        GemmArgs g;
        memset(&g, 0, sizeof(g));
        g.a.hi = N->A_hi[l]; g.a.mid = N->A_mid[l]; g.a.rpad = rp;
        g.K = L.K; g.R = R; g.Rpad = Rp; g.wp = L.w_fwd; g.N = L.Cout;
        g.bias = c->params + L.off_b; g.acc_scale = l == 0 ? 1.0f / 255.0f : 1.0f; g.relu = 1;
        if (l < 3) { g.out_hi = N->act_hi[l]; g.out_mid = N->act_mid[l]; g.out_rpad = rp; }
        else { g.out_f32 = c->hidden; g.out_ld = 512; }
        char name[64];
        snprintf(name, sizeof(name), "nat_gemm_fwd<k%d,n%d>", L.K, L.Cout);
        ProfScope ps(c, name, 2.0 * R * L.K * L.Cout, (double)Rp * L.K * (l == 0 ? 2 : 4) + (double)Rp * L.Cout * 4, st,
                     4.0 * n * L.Hin * L.Hin * (l == 0 ? 1 : L.Cin) + 4.0 * R * L.Cout);
        if (launch_gemm_umma(g, L.NBf, c->num_sms, st)) return -1;
    }
    return 0;
}

static int nat_colsum(cb_ctx* c, const f16* hi, const f16* mid, long long rpad, long long R, int C, float* db, cudaStream_t st) {
    NatureNet* N = c->nat;
    k_nat_colsum_partial<<<dim3(NAT_CS, C / 8), 256, 0, st>>>(hi, mid, rpad, R, N->col_part);
    CB_LAUNCH_CHECK();
    k_nat_colsum_final<<<(C + 127) / 128, 128, 0, st>>>(N->col_part, C, c->gscale + 1, db);
    CB_LAUNCH_CHECK();
    return 0;
}

int nature_backward(cb_ctx* c, int n, float* grads, cudaStream_t st) {
    NatureNet* N = c->nat;
    if (launch_loss_scale(c->dpre, (long long)n * 512, c->gs_work, c->gscale, st)) return -1;
    {
        const long long Rp = pad128(n);
        k_nat_dpre_planes<<<blocks_for(Rp * 64), 256, 0, st>>>(c->dpre, n, N->rpad[3], Rp, c->gscale, N->dp_hi, N->dp_mid);
        CB_LAUNCH_CHECK();
    }
    const long long wg_cap = c->wg_cap;              // floats in c->wg_partial
    for (int l = 3; l >= 0; --l) {
        NatLayer& L = N->L[l];
        const long long R = (long long)n * L.Hout * L.Hout, Rp = pad128(R), rp = N->rpad[l];
        const f16* ghi = l == 3 ? N->dp_hi : N->g_hi[l];
        const f16* gmid = l == 3 ? N->dp_mid : N->g_mid[l];
        {   // dW = A^T G / S (x 1/255 for the frame conv: the forward folds x / 255 into its epilogue), db = colsum(G) / S
            GemmWgradArgs w;
            memset(&w, 0, sizeof(w));
            w.a.hi = N->A_hi[l]; w.a.mid = N->A_mid[l]; w.a.rpad = rp;
            w.g.hi = ghi; w.g.mid = gmid; w.g.rpad = rp;
            w.K = L.K; w.N = L.Cout; w.Rpad = Rp; w.scale = l == 0 ? 1.0f / 255.0f : 1.0f; w.inv_scale = c->gscale + 1;
            w.dw = grads + L.off_w;
            char name[64];
            snprintf(name, sizeof(name), "nat_gemm_wgrad<k%d,n%d>", L.K, L.Cout);
            ProfScope ps(c, name, 2.0 * R * L.K * L.Cout, (double)Rp * L.K * (l == 0 ? 2 : 4) + (double)Rp * L.Cout * 4, st,
                         4.0 * n * L.Hin * L.Hin * (l == 0 ? 1 : L.Cin) + 4.0 * R * L.Cout);
            if (launch_gemm_wgrad_umma(w, c->wg_partial, wg_cap, c->num_sms, st)) return -1;
            if (nat_colsum(c, ghi, gmid, rp, R, L.Cout, grads + L.off_b, st)) return -1;
        }
        if (l == 3 && c->milestone) CB_CUDA(cudaEventRecordWithFlags(c->milestone, st, c->capturing ? cudaEventRecordExternal : cudaEventRecordDefault));   // dense + head gradients are final (the tail of the flat vector)
        if (l == 0) break;
        {   // dA = G W^T (im2col space), then col2im + relu gate -> gradient w.r.t. the previous layer's pre-relu output
            GemmArgs g;
            memset(&g, 0, sizeof(g));
            g.a.hi = ghi; g.a.mid = gmid; g.a.rpad = rp;
            g.K = L.Cout; g.R = R; g.Rpad = Rp; g.wp = L.w_dg; g.N = L.K; g.acc_scale = 1.f;
            g.out_hi = N->dA_hi[l]; g.out_mid = N->dA_mid[l]; g.out_rpad = rp;
            char name[64];
            snprintf(name, sizeof(name), "nat_gemm_dgrad<k%d,n%d>", L.Cout, L.K);
            ProfScope ps(c, name, 2.0 * R * L.K * L.Cout, (double)Rp * L.Cout * 4 + (double)Rp * L.K * 4, st, 4.0 * R * L.Cout + 4.0 * n * L.Hin * L.Hin * L.Cin);
            if (launch_gemm_umma(g, L.NBd, c->num_sms, st)) return -1;
        }
        {
            NatLayer& P = N->L[l - 1];
            const long long Rin = (long long)n * L.Hin * L.Hin, Rinp = pad128(Rin);
            char name[64];
            snprintf(name, sizeof(name), "col2im@%d", l);
            ProfScope ps(c, name, 0, (double)Rp * L.K * 4 + (double)Rinp * L.Cin * 6, st, 4.0 * Rin * L.Cin * 2);
            launch_pdl(k_nat_col2im, dim3(blocks_for(Rinp * (L.Cin / 8))), dim3(256), 0, st, (const f16*)N->dA_hi[l], (const f16*)N->dA_mid[l], rp,
                       (const f16*)N->act_hi[l - 1], N->g_hi[l - 1], N->g_mid[l - 1], N->rpad[l - 1], Rinp, Rin, L.KH, L.stride, L.Cin / 8, L.Hin,
                       L.Hout);
            CB_LAUNCH_CHECK();
            (void)P;
        }
    }
    return 0;
}

}  // namespace cb
