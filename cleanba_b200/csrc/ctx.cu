// Context, buffer management, forward / backward orchestration and the C ABI (include/cleanba_b200.h).
// Model: IMPALA-ResNet channels (16,32,32), hidden 256, linear actor/critic  (cleanba/cleanba_ppo.py:149-203).
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <math.h>
#include <atomic>
#include <map>
#include <string>
#include <vector>

#include "ctx.h"

namespace cb {

// PDL pays for launch-latency-bound chains (the actor's n = 60 step: 0.236 -> 0.193 ms) and costs ~1% on the learner's
// large minibatches (early CTAs of the next kernel compete with the running one), so it is enabled per call by batch size
// (n <= CLEANBA_PDL_MAX_BATCH, default 1024: covers the actor step and the 630-frame IMPALA minibatch, +3% on config 3).
static thread_local bool g_pdl_scope = false;
bool pdl_enabled() {
    static const int mode = [] { const char* e = getenv("CLEANBA_PDL"); return e ? atoi(e) : 1; }();   // 0 off, 1 auto, 2 always
    return mode == 2 || (mode == 1 && g_pdl_scope);
}
std::atomic<long long> g_launches{0};   // every kernel launch of this library (CB_LAUNCH_CHECK increments it)
static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ------------------------------------------------------------------------------------------------ model tables
static const int kStageCin[3] = {4, 16, 32};
static const int kStageC[3] = {16, 32, 32};
static const int kStageHin[3] = {84, 42, 21};
static const int kStageHout[3] = {42, 21, 11};
static const int kStagePadLo[3] = {0, 0, 1};
constexpr int kFlat = 11 * 11 * 32;

static std::vector<Leaf> build_leaves(int A) {
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
    for (int s = 0; s < 3; ++s) {
        std::string p = "network_params/params/ConvSequence_" + std::to_string(s);
        add(p + "/Conv_0/bias", {kStageC[s]});
        add(p + "/Conv_0/kernel", {3, 3, kStageCin[s], kStageC[s]});
        for (int r = 0; r < 2; ++r)
            for (int k = 0; k < 2; ++k) {
                std::string q = p + "/ResidualBlock_" + std::to_string(r) + "/Conv_" + std::to_string(k);
                add(q + "/bias", {kStageC[s]});
                add(q + "/kernel", {3, 3, kStageC[s], kStageC[s]});
            }
    }
    add("network_params/params/Dense_0/bias", {HIDDEN});
    add("network_params/params/Dense_0/kernel", {kFlat, HIDDEN});
    add("actor_params/params/Dense_0/bias", {A});
    add("actor_params/params/Dense_0/kernel", {HIDDEN, A});
    add("critic_params/params/Dense_0/bias", {1});
    add("critic_params/params/Dense_0/kernel", {HIDDEN, 1});
    return L;
}

}  // namespace cb

using namespace cb;

namespace cb {

int dev_alloc(cb_ctx* c, void** p, size_t bytes, bool zero) {
    CB_CUDA(cudaMalloc(p, bytes));
    c->allocs.push_back(*p);
    if (zero) CB_CUDA(cudaMemset(*p, 0, bytes));
    return 0;
}

static long long plane_px_for(int max_batch, int H) {
    long long np = (long long)max_batch * (H + 1) * (H + 1) + (H + 3);      // + the zero row after the last image
    np = (np + 127) / 128 * 128;
    return GUARD + np + GUARD;
}

// two = false: single-plane tensor (the unpacked frames)
static int alloc_act(cb_ctx* c, Act& a, int C, int H, bool two = true) {
    a.C = C; a.H = H;
    const int chunks = (C + 7) / 8;
    a.pl.plane_px = plane_px_for(c->cfg.max_batch, H);
    size_t bytes = (size_t)chunks * a.pl.plane_px * 8 * sizeof(f16);
    void* p;
    if (dev_alloc(c, &p, bytes)) return -1;
    a.pl.hi = (f16*)p + (long long)GUARD * 8;
    if (two) {
        if (dev_alloc(c, &p, bytes)) return -1;
        a.pl.mid = (f16*)p + (long long)GUARD * 8;
    }
    return 0;
}

static double f32_once(const ConvGeom& g, int channels) { return 4.0 * g.n * g.H * g.W * channels; }   // unpadded fp32 tensor
static double planes_bytes(const ConvGeom& g, int chunks, bool two = true) { return (double)g.NP * chunks * 8 * (two ? 4 : 2); }
static double planes_bytes(const ConvGeom& g, int chunks, const Planes& p) { return p.hi ? planes_bytes(g, chunks, p.mid != nullptr) : 0.0; }

static int refresh_weights(cb_ctx* c, cudaStream_t st) {
    if (c->nat) return nature_refresh_weights(c, st);
    ProfScope ps(c, "pack_weights", 0, 1089232.0 * (4 + 8), st);
    if (launch_pack_conv(c->pack_dev, 15, st)) return -1;
    if (c->wd_fwd) return launch_pack_dense(c->params + c->off_dense_w, c->wd_fwd, c->wd_dx, st);
    return 0;
}
static DenseUmmaArgs dense_umma_args(cb_ctx* c, int n) {
    DenseUmmaArgs u;
    memset(&u, 0, sizeof(u));
    u.n = n; u.npad = (n + 127) / 128 * 128; u.NP = (long long)n * 144;
    u.ft_hi = c->ft[0]; u.ft_mid = c->ft[1]; u.ft_lo = c->ft[2];
    u.dp_hi = c->dpT[0]; u.dp_mid = c->dpT[1];
    u.w_fwd = c->wd_fwd; u.w_dx = c->wd_dx;
    u.bias = c->params + c->off_dense_b; u.hidden = c->hidden; u.part = c->dense_part;
    return u;
}

static ConvArgs conv_args(cb_ctx* c, int layer, const ConvGeom& g, const Act& in, bool transpose) {
    const ConvLayer& L = c->conv[layer];
    ConvArgs a;
    memset(&a, 0, sizeof(a));
    a.g = g;
    a.in = in.pl;
    a.w = c->params + L.off_w;
    a.w_cin = L.cin; a.w_cout = L.cout;
    a.transpose = transpose ? 1 : 0;
    if (!transpose) {
        a.cin_real = L.cin; a.cin_chunks = (L.cin + 7) / 8; a.cout = L.cout;
        a.wp = L.fwd;
    } else {
        a.cin_real = L.cout; a.cin_chunks = L.cout / 8; a.cout = L.cin;
        a.wp = L.dg;
    }
    a.ep.acc_scale = 1.f;
    return a;
}

static int run_conv(cb_ctx* c, const ConvArgs& a, cudaStream_t st) {
    char name[96];
    snprintf(name, sizeof(name), "%s<cin%d,cout%d>@%dx%d", a.transpose ? "conv_dgrad" : "conv_fwd", a.cin_real, a.cout, a.g.H, a.g.W);
    const double flops = 2.0 * a.g.n * a.g.H * a.g.W * 9.0 * a.cin_real * a.cout;
    double bytes = planes_bytes(a.g, a.cin_chunks, a.in) + planes_bytes(a.g, a.cout / 8, a.ep.out) + planes_bytes(a.g, a.cout / 8, a.ep.out_r) +
                   planes_bytes(a.g, a.cout / 8, a.ep.res);
    if (a.ep.bits_in) bytes += (double)a.g.NP * (a.cout / 8);
    else if (a.ep.mask_hi) bytes += planes_bytes(a.g, a.cout / 8, false);
    if (a.ep.bits_out) bytes += (double)a.g.NP * (a.cout / 8);
    // algorithmic: input, output and residual tensors once as unpadded fp32 (the relu gate of dgrad is one BIT per element: not counted)
    const double abytes = f32_once(a.g, a.cin_real) + f32_once(a.g, a.cout) * (1 + (a.ep.res.hi ? 1 : 0));
    ProfScope ps(c, name, flops, bytes, st, abytes);
    if (c->cfg.conv_backend == CB_CONV_SIMT) return launch_conv_simt(a, st);
    return launch_conv_umma(a, c->num_sms, st);
}

// Weight gradients feed nothing but the optimizer, so they run on the context's side stream beside the dgrad chain (the caller's
// stream): fork = "the gradient tensor gy is complete on `main`".  The one buffer hazard (the dgrad of the block's second conv
// re-writes gB while the wgrad of its first conv may still read it) is closed by wgrad_done_before().
// While profiling (cb_profile) everything runs on the caller's stream, so that the event brackets time each kernel ALONE.
static bool side_active(const cb_ctx* c) { return c->side && !c->prof_on; }
static int fork_side(cb_ctx* c, cudaStream_t main_st) {
    if (!side_active(c)) return 0;
    CB_CUDA(cudaEventRecord(c->ev_fork, main_st));
    CB_CUDA(cudaStreamWaitEvent(c->side, c->ev_fork, 0));
    return 0;
}
static int wgrad_done_before(cb_ctx* c, cudaStream_t main_st) {    // main waits for everything queued on the side stream so far
    if (!side_active(c)) return 0;
    CB_CUDA(cudaEventRecord(c->ev_join, c->side));
    CB_CUDA(cudaStreamWaitEvent(main_st, c->ev_join, 0));
    return 0;
}

static int run_wgrad(cb_ctx* c, int layer, const ConvGeom& g, const Act& x, const Act& gy, float* grads, cudaStream_t main_st) {
    if (fork_side(c, main_st)) return -1;
    cudaStream_t st = side_active(c) ? c->side : main_st;
    const ConvLayer& L = c->conv[layer];
    WgradArgs w;
    w.g = g; w.x = x.pl; w.cin_chunks = (L.cin + 7) / 8; w.cin_real = L.cin; w.gy = gy.pl; w.cout = L.cout;
    w.dw = grads + L.off_w; w.db = grads + L.off_b;
    w.scale = (layer == 0) ? (1.0f / 255.0f) : 1.0f;
    w.inv_scale = c->gscale + 1;
    char name[96];
    snprintf(name, sizeof(name), "conv_wgrad<cin%d,cout%d>@%dx%d", L.cin, L.cout, g.H, g.W);
    ProfScope ps(c, name, 2.0 * g.n * g.H * g.W * 9.0 * L.cin * L.cout,
                 planes_bytes(g, w.cin_chunks, x.pl) + planes_bytes(g, L.cout / 8, gy.pl), st, f32_once(g, L.cin) + f32_once(g, L.cout));
    if (c->cfg.conv_backend == CB_CONV_SIMT) return launch_wgrad_simt(w, c->wg_partial, 296, st);
    return launch_wgrad_umma(w, c->wg_partial, c->num_sms, st);
}

// The zero row after the last image of a batch of n (common.cuh) is dirty whenever a larger batch went through the context's
// buffers since the last batch of n: clear it before any kernel of a batch of a different size reads it.
static int clear_trailing_rows(cb_ctx* c, int n, cudaStream_t st) {
    if (c->clean_n == n || !c->trail_count) return 0;
    if (launch_clear_trailing_rows(c->trail_dev, c->trail_count, n, st)) return -1;
    c->clean_n = n;
    return 0;
}

// Network.__call__ (cleanba_ppo.py:178-189) on n frames -> c->hidden [n,256]
static int trunk_forward(cb_ctx* c, const uint8_t* obs, const int* idx, int n, cudaStream_t st) {
    CB_CHECK(n > 0 && n <= c->cfg.max_batch, "batch %d outside (0, max_batch=%d]", n, c->cfg.max_batch);
    c->last_n = n;
    if (!c->capturing && clear_trailing_rows(c, n, st)) return -1;
    if (c->nat) {
        static const int nat_pdl_max = [] { const char* e = getenv("CLEANBA_PDL_MAX_BATCH"); return e ? atoi(e) : 1024; }();
        g_pdl_scope = n <= nat_pdl_max;
        return nature_forward(c, obs, idx, n, st);
    }
    static const int pdl_max = [] { const char* e = getenv("CLEANBA_PDL_MAX_BATCH"); return e ? atoi(e) : 1024; }();
    g_pdl_scope = n <= pdl_max;
    {
        ProfScope ps(c, "unpack_frames", 0, (double)n * (28224.0 + 85.0 * 85 * 16), st);
        if (launch_unpack(obs, idx, n, c->st[0].x.pl.hi, st, c->cursor, c->ind)) return -1;
    }
    // Opt-in (cb_set_actor_tail): ConvSequence 1 and 2 as ONE persistent kernel, one thread-block cluster per frame
    // (actor_fused.cu).  Bit-identical to the per-layer launches; measured SLOWER at the rollout batch, so off by default.
    const int tail_cluster = c->actor_tail;
    const bool tail = tail_cluster > 0 && c->fuse12 && c->cfg.conv_backend == CB_CONV_TCGEN05 && n <= 128 && !c->prof_on;
    ActorTailHost th;
    th.n = n;
    for (int s = 0; s < 3; ++s) {
        Stage& S = c->st[s];
        const ConvGeom gi = make_geom(n, kStageHin[s], kStageHin[s]);
        const ConvGeom go = make_geom(n, kStageHout[s], kStageHout[s]);
        const int base = s * 5;
        const int C = kStageC[s];
        const bool fused = (s == 0) ? c->fuse0 : c->fuse12;       // sequence conv + max-pool in ONE tcgen05 kernel
        const bool in_tail = tail && s >= 1;
        if (fused && s == 0) {
            // frame conv + max-pool in one kernel: the 84x84x16 conv output never reaches HBM (conv_umma.cu)
            ConvArgs a = conv_args(c, 0, gi, S.x, false);
            a.ep.bias = c->params + c->conv[0].off_b;
            a.ep.acc_scale = 1.0f / 255.0f;                       // x / 255.0 (cleanba_ppo.py:181) folded into the epilogue
            ProfScope ps(c, "conv0_pool_fwd@84", 2.0 * n * 84 * 84 * 9.0 * 4 * 16,
                         planes_bytes(gi, 1, false) + 2 * planes_bytes(go, 2) + (S.amax ? (double)go.NP * 16 : 0.0), st,
                         (double)n * 28224.0 + f32_once(go, 16));      // uint8 frames in, pooled fp32 out
            if (launch_conv0_pool_umma(a, S.p.pl, S.pr.pl, S.amax, S.bits_pr, c->num_sms, st)) return -1;
        } else if (fused) {
            // sequence conv + max-pool in one kernel (conv_umma.cu: k_conv_pool_umma)
            ConvArgs a = conv_args(c, base, gi, S.x, false);
            a.ep.bias = c->params + c->conv[base].off_b;
            char name[96];
            snprintf(name, sizeof(name), "conv_pool_fwd<cin%d,cout%d>@%d", a.cin_real, a.cout, gi.H);
            ProfScope ps(c, name, 2.0 * n * gi.H * gi.W * 9.0 * a.cin_real * a.cout,
                         planes_bytes(gi, a.cin_chunks, a.in) + 2 * planes_bytes(go, a.cout / 8) + (S.amax ? (double)go.NP * a.cout : 0.0), st,
                         f32_once(gi, a.cin_real) + f32_once(go, a.cout));
            if (in_tail) {
                th.pool[s - 1] = a; th.pool_go[s - 1] = go; th.pad_lo[s - 1] = kStagePadLo[s];
                th.pool_out[s - 1] = S.p.pl; th.pool_out_r[s - 1] = S.pr.pl;
            } else if (launch_conv_pool_umma(a, go, kStagePadLo[s], S.p.pl, S.pr.pl, S.amax, S.bits_pr, c->num_sms, st)) return -1;
        } else {
            // x = nn.Conv(channels)(x)                                          (cleanba_ppo.py:167)
            ConvArgs a = conv_args(c, base + 0, gi, S.x, false);
            a.ep.bias = c->params + c->conv[base].off_b;
            a.ep.acc_scale = (s == 0) ? (1.0f / 255.0f) : 1.0f;   // x / 255.0 (cleanba_ppo.py:181) folded into the epilogue
            a.ep.out = S.y.pl;
            if (run_conv(c, a, st)) return -1;
            // x = nn.max_pool(x, (3,3), strides=(2,2), padding="SAME")          (cleanba_ppo.py:168)
            ProfScope ps(c, "pool_fwd@" + std::to_string(kStageHin[s]), 0,
                         planes_bytes(gi, C / 8) + 2 * planes_bytes(go, C / 8) + (S.amax ? (double)go.NP * C : 0.0), st,
                         f32_once(gi, C) + f32_once(go, C));
            if (launch_pool_fwd(S.y.pl, gi, go, kStagePadLo[s], C / 8, S.p.pl, S.pr.pl, S.amax, st)) return -1;
        }
        {   // ResidualBlock 0: x + Conv(relu(Conv(relu(x))))                    (cleanba_ppo.py:153-159)
            ConvArgs a = conv_args(c, base + 1, go, S.pr, false);
            a.ep.bias = c->params + c->conv[base + 1].off_b;
            a.ep.out_r = S.a0.pl; a.ep.bits_out = S.bits_a0;
            if (in_tail) th.conv[(s - 1) * 4 + 0] = a;
            else if (run_conv(c, a, st)) return -1;
            ConvArgs b = conv_args(c, base + 2, go, S.a0, false);
            b.ep.bias = c->params + c->conv[base + 2].off_b;
            b.ep.res = S.p.pl; b.ep.out = S.b0.pl; b.ep.out_r = S.b0r.pl; b.ep.bits_out = S.bits_b0r;
            if (in_tail) th.conv[(s - 1) * 4 + 1] = b;
            else if (run_conv(c, b, st)) return -1;
        }
        {   // ResidualBlock 1; its output feeds the next ConvSequence un-rectified, or the final nn.relu (cleanba_ppo.py:184)
            ConvArgs a = conv_args(c, base + 3, go, S.b0r, false);
            a.ep.bias = c->params + c->conv[base + 3].off_b;
            a.ep.out_r = S.a1.pl; a.ep.bits_out = S.bits_a1;
            if (in_tail) th.conv[(s - 1) * 4 + 2] = a;
            else if (run_conv(c, a, st)) return -1;
            ConvArgs b = conv_args(c, base + 4, go, S.a1, false);
            b.ep.bias = c->params + c->conv[base + 4].off_b;
            b.ep.res = S.b0.pl;
            if (s < 2) b.ep.out = S.out.pl; else b.ep.out_r = S.out.pl;
            if (s == 2 && c->wd_fwd) {   // sample-minor copy of the final features for the tcgen05 dense layer
                b.ep.ft_hi = c->ft[0]; b.ep.ft_mid = c->ft[1]; b.ep.ft_lo = c->ft[2];
                b.ep.ft_npad = (n + 127) / 128 * 128; b.ep.ft_pixpad = 124;
            }
            if (in_tail) th.conv[(s - 1) * 4 + 3] = b;
            else if (run_conv(c, b, st)) return -1;
        }
    }
    if (tail && launch_actor_tail(th, tail_cluster, c->num_sms, st)) return -1;
    DenseArgs d;
    d.n = n; d.x = c->st[2].out.pl; d.w = c->params + c->off_dense_w; d.b = c->params + c->off_dense_b; d.hidden = c->hidden;
    ProfScope ps(c, "dense_fwd", 2.0 * n * kFlat * HIDDEN, (double)n * kFlat * 6 + (double)kFlat * HIDDEN * 6, st,
                 4.0 * ((double)n * kFlat + (double)kFlat * HIDDEN + (double)n * HIDDEN));
    if (c->wd_fwd) return launch_dense_fwd_umma(dense_umma_args(c, n), st);
    return launch_dense_fwd(d, c->dense_part, st);
}

// Backward of the trunk given c->dpre (gradient w.r.t. the pre-relu dense output); writes all trunk gradients.
static int trunk_backward(cb_ctx* c, int n, float* grads, cudaStream_t st) {
    if (c->nat) return nature_backward(c, n, grads, st);
    // per-minibatch power-of-two loss scale of the fp16 gradient carriers (device side, no host sync)
    if (launch_loss_scale(c->dpre, (long long)n * c->HID, c->gs_work, c->gscale, st)) return -1;
    DenseArgs d;
    d.n = n; d.x = c->st[2].out.pl; d.w = c->params + c->off_dense_w; d.b = c->params + c->off_dense_b; d.hidden = c->hidden;
    if (c->wd_fwd) {
        ProfScope ps(c, "dense_bwd", 4.0 * n * kFlat * HIDDEN, (double)n * kFlat * 12 + (double)kFlat * HIDDEN * 8, st,
                     4.0 * (2.0 * n * kFlat + 2.0 * kFlat * HIDDEN + (double)n * HIDDEN));
        DenseUmmaArgs u = dense_umma_args(c, n);
        u.gscale = c->gscale; u.out = c->st[2].gA.pl;
        if (launch_dpre_transpose(c->dpre, n, u.npad, c->dpT[0], c->dpT[1], st)) return -1;
        if (launch_dense_bwd_umma(u, c->dpre, grads + c->off_dense_w, grads + c->off_dense_b, c->dense_part, st)) return -1;
    } else {
        {
            ProfScope ps(c, "dense_bwd_w", 2.0 * n * kFlat * HIDDEN, (double)n * kFlat * 4 + (double)kFlat * HIDDEN * 4, st);
            if (launch_dense_bwd_w(d, c->dpre, grads + c->off_dense_w, grads + c->off_dense_b, st)) return -1;
        }
        {
            ProfScope ps(c, "dense_bwd_x", 2.0 * n * kFlat * HIDDEN, (double)n * kFlat * 6 + (double)kFlat * HIDDEN * 4, st);
            if (launch_dense_bwd_x(d, c->dpre, c->gscale, c->st[2].gA.pl, st)) return -1;
        }
    }
    // The flat gradient vector is [conv stages | dense | actor | critic]: its tail (the dense layer: 91% of the parameters) is
    // final here, before the conv backward starts -- the caller may start exchanging it now (cb_set_grad_milestone).
    // (inside a captured step the record is an EXTERNAL event node: every replay records the caller's event again)
    if (c->milestone) CB_CUDA(cudaEventRecordWithFlags(c->milestone, st, c->capturing ? cudaEventRecordExternal : cudaEventRecordDefault));
    for (int s = 2; s >= 0; --s) {
        Stage& S = c->st[s];
        const ConvGeom gi = make_geom(n, kStageHin[s], kStageHin[s]);
        const ConvGeom go = make_geom(n, kStageHout[s], kStageHout[s]);
        const int base = s * 5;
        const int C = kStageC[s];
        // ---- ResidualBlock 1: out = b0 + conv4(relu(conv3(relu(b0))))
        if (run_wgrad(c, base + 4, go, S.a1, S.gA, grads, st)) return -1;
        {
            ConvArgs a = conv_args(c, base + 4, go, S.gA, true);
            a.ep.mask_hi = S.a1.pl.hi; a.ep.mask_plane_px = S.a1.pl.plane_px; a.ep.bits_in = S.bits_a1; a.ep.out = S.gB.pl;
            if (run_conv(c, a, st)) return -1;
        }
        if (run_wgrad(c, base + 3, go, S.b0r, S.gB, grads, st)) return -1;
        {
            ConvArgs a = conv_args(c, base + 3, go, S.gB, true);
            a.ep.mask_hi = S.b0r.pl.hi; a.ep.mask_plane_px = S.b0r.pl.plane_px; a.ep.bits_in = S.bits_b0r; a.ep.res = S.gA.pl;
            a.ep.out = S.gC.pl;
            if (run_conv(c, a, st)) return -1;
        }
        // ---- ResidualBlock 0: b0 = p + conv2(relu(conv1(relu(p))))
        if (run_wgrad(c, base + 2, go, S.a0, S.gC, grads, st)) return -1;
        if (wgrad_done_before(c, st)) return -1;      // the next dgrad re-writes gB, which the wgrad of base + 3 reads
        {
            ConvArgs a = conv_args(c, base + 2, go, S.gC, true);
            a.ep.mask_hi = S.a0.pl.hi; a.ep.mask_plane_px = S.a0.pl.plane_px; a.ep.bits_in = S.bits_a0; a.ep.out = S.gB.pl;
            if (run_conv(c, a, st)) return -1;
        }
        if (run_wgrad(c, base + 1, go, S.pr, S.gB, grads, st)) return -1;
        {
            ConvArgs a = conv_args(c, base + 1, go, S.gB, true);
            a.ep.mask_hi = S.pr.pl.hi; a.ep.mask_plane_px = S.pr.pl.plane_px; a.ep.bits_in = S.bits_pr; a.ep.res = S.gC.pl;
            a.ep.out = S.gA.pl;    // gradient w.r.t. the pooled tensor (gA is free again)
            if (run_conv(c, a, st)) return -1;
        }
        // ---- max-pool backward, then the sequence conv
        if (s == 0 && c->fuse0) {
            // frames need no dX: the pooled gradient goes straight into the frame conv's weight gradient (trunk_simt.cu)
            if (fork_side(c, st)) return -1;
            cudaStream_t ws = side_active(c) ? c->side : st;
            ProfScope ps(c, "pool_bwd_wgrad0@84", 2.0 * n * 42 * 42 * 16 * 36, (double)n * (1849.0 * (16 + 64) + 7225.0 * 16), ws,
                         f32_once(go, 16) + (double)n * 28224.0);
            if (launch_pool_bwd_wgrad0(S.amax, S.gA.pl, S.x.pl.hi, gi, go, 1.0f / 255.0f, c->gscale + 1, grads + c->conv[0].off_w,
                                       grads + c->conv[0].off_b, c->wg_partial, c->num_sms, ws)) return -1;
            continue;
        }
        {
            ProfScope ps(c, "pool_bwd@" + std::to_string(kStageHin[s]), 0,
                         (double)go.NP * C * 5 + planes_bytes(gi, C / 8), st, f32_once(go, C) + f32_once(gi, C));
            if (launch_pool_bwd(S.amax, S.gA.pl, gi, go, kStagePadLo[s], C / 8, S.gBin.pl, st)) return -1;
        }
        if (run_wgrad(c, base + 0, gi, S.x, S.gBin, grads, st)) return -1;
        if (s > 0) {
            ConvArgs a = conv_args(c, base + 0, gi, S.gBin, true);
            a.ep.out = c->st[s - 1].gA.pl;
            if (run_conv(c, a, st)) return -1;
        }
    }
    return wgrad_done_before(c, st);      // every weight gradient is complete on the caller's stream
}

// ---- graphed gradient steps (cb_graph_steps) --------------------------------------------------------------------------------
// A cb_*_grad call is ~70 launches on two streams.  With graph mode on, the launches of a (shape, gradient buffer, coefficients)
// combination are captured ONCE -- the second time the combination is seen, so that every lazily created resource exists -- and
// replayed afterwards: per step the host enqueues one tiny launch (the step's pointers -> cb_ctx::step_dev, which the frame
// unpack and the loss head read instead of their frozen arguments) and one cudaGraphLaunch.
template <class Body>
static int graphed_step(cb_ctx* c, const StepGraph& key, const StepPtrs& ptrs, cudaStream_t st, Body body) {
    StepGraph* g = nullptr;
    for (auto& e : c->graphs)
        if (e.kind == key.kind && e.n == key.n && e.T1 == key.T1 && e.B == key.B && e.grads == key.grads && e.c0 == key.c0 &&
            e.c1 == key.c1 && e.c2 == key.c2 && e.milestone == key.milestone) { g = &e; break; }
    if (!g) {                                   // first sight: run eagerly (warms every launcher), remember the key
        if (c->graphs.size() >= 64) {           // a caller that keeps changing shapes: drop the oldest entries
            for (int i = 0; i < 32; ++i) if (c->graphs[i].exec) cudaGraphExecDestroy(c->graphs[i].exec);
            c->graphs.erase(c->graphs.begin(), c->graphs.begin() + 32);
        }
        c->graphs.push_back(key);
        return body(st);
    }
    if (launch_set_step_ptrs(c->step_dev, ptrs, st)) return -1;
    if (clear_trailing_rows(c, key.n, st)) return -1;      // the replayed trunk_forward cannot do it
    if (!g->exec) {
        // captured on a private stream (the caller's may be the legacy default stream, which cannot capture); replayed on `st`
        CB_CUDA(cudaStreamBeginCapture(c->cap, cudaStreamCaptureModeThreadLocal));
        c->capturing = true; c->ind = c->step_dev;
        const long long l0 = g_launches.load();
        const int rc = body(c->cap);
        c->capturing = false; c->ind = nullptr;
        cudaGraph_t graph = nullptr;
        const cudaError_t e = cudaStreamEndCapture(c->cap, &graph);
        if (rc) { if (graph) cudaGraphDestroy(graph); (void)cudaGetLastError(); return -1; }   // body() set the error text
        CB_CUDA(e);
        g->launches = g_launches.load() - l0;
        g_launches.fetch_sub(g->launches);      // captured, not launched yet
        const cudaError_t ei = cudaGraphInstantiate(&g->exec, graph, 0);
        cudaGraphDestroy(graph);
        CB_CUDA(ei);
    }
    CB_CUDA(cudaGraphLaunch(g->exec, st));
    g_launches.fetch_add(g->launches);
    c->graph_replays += 1;
    return 0;
}

static int ppo_grad_body(cb_ctx* c, const uint8_t* obs, const int32_t* idx, int mb, const int32_t* actions, const float* logprobs,
                         const float* advantages, const float* returns, float clip_coef, float ent_coef, float vf_coef,
                         float* grads, float* stats, cudaStream_t st) {
    if (trunk_forward(c, obs, idx, mb, st)) return -1;
    PpoHeadArgs h;
    h.n = mb; h.num_actions = c->A; h.hid = c->HID; h.hidden = c->hidden;
    h.wa = c->params + c->off_actor_w; h.ba = c->params + c->off_actor_b;
    h.wc = c->params + c->off_critic_w; h.bc = c->params + c->off_critic_b;
    h.idx = idx; h.actions = actions; h.old_logprobs = logprobs; h.advantages = advantages; h.returns = returns;
    h.clip_coef = clip_coef; h.ent_coef = ent_coef; h.vf_coef = vf_coef;
    h.dpre = c->dpre; h.dlogits = c->dlogits; h.terms = c->terms; h.stats = stats; h.wgrad_scratch = c->wg_partial;
    h.dwa = grads + c->off_actor_w; h.dba = grads + c->off_actor_b; h.dwc = grads + c->off_critic_w; h.dbc = grads + c->off_critic_b;
    h.ind = c->ind;
    {
        ProfScope ps(c, "ppo_loss_head", 6.0 * mb * HIDDEN * (c->A + 1), (double)mb * (2 * HIDDEN * 4 + 92 + 76), st);
        if (launch_ppo_head(h, st)) return -1;
    }
    return trunk_backward(c, mb, grads, st);
}

static int impala_grad_body(cb_ctx* c, const uint8_t* obs, const int32_t* idx, int T1, int B, const int32_t* actions,
                            const float* behaviour_logits, const float* rewards, const uint8_t* dones, const uint8_t* firststeps,
                            float gamma, float vf_coef, float ent_coef, float* grads, float* stats, cudaStream_t st) {
    const int n = T1 * B;
    if (trunk_forward(c, obs, idx, n, st)) return -1;
    ImpalaHeadArgs h;
    h.T1 = T1; h.B = B; h.num_actions = c->A; h.hid = c->HID; h.hidden = c->hidden;
    h.wa = c->params + c->off_actor_w; h.ba = c->params + c->off_actor_b;
    h.wc = c->params + c->off_critic_w; h.bc = c->params + c->off_critic_b;
    h.idx = idx; h.actions = actions; h.behaviour_logits = behaviour_logits; h.rewards = rewards; h.dones = dones;
    h.firststeps = firststeps; h.gamma = gamma; h.vf_coef = vf_coef; h.ent_coef = ent_coef;
    h.logits_scratch = c->logits_scratch; h.cell_scratch = c->cell_scratch; h.dpre = c->dpre; h.dlogits = c->dlogits;
    h.stats = stats; h.wgrad_scratch = c->wg_partial;
    h.dwa = grads + c->off_actor_w; h.dba = grads + c->off_actor_b; h.dwc = grads + c->off_critic_w; h.dbc = grads + c->off_critic_b;
    h.ind = c->ind;
    {
        ProfScope ps(c, "vtrace_loss_head", 6.0 * n * HIDDEN * (c->A + 1), (double)n * (2 * HIDDEN * 4 + 157 + 76), st);
        if (launch_impala_head(h, st)) return -1;
    }
    return trunk_backward(c, n, grads, st);
}

static int copy_any(void* dst, const void* src, size_t bytes, cudaStream_t st) {
    CB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, st));
    return 0;
}

}  // namespace cb

// ================================================================================================ C ABI
#pragma GCC visibility push(default)
extern "C" {

const char* cb_last_error(void) { return g_err; }
int cb_version(void) { return 1; }

static std::vector<Leaf> model_leaves(int model, int A) { return model == CB_MODEL_NATURE_CNN ? nature_leaves(A) : build_leaves(A); }
long long cb_num_params_model(int model, int num_actions) {
    auto L = model_leaves(model, num_actions);
    return L.back().offset + L.back().size();
}
int cb_num_leaves_model(int model) { return (int)model_leaves(model, 18).size(); }
long long cb_num_params(int num_actions) { return cb_num_params_model(CB_MODEL_IMPALA_RESNET, num_actions); }
int cb_num_leaves(void) { return 36; }
int cb_leaf_info(int index, int num_actions, char* name, int name_cap, long long* offset, int* ndim, int* shape) {
    return cb_leaf_info_model(CB_MODEL_IMPALA_RESNET, index, num_actions, name, name_cap, offset, ndim, shape);
}
int cb_leaf_info_model(int model, int index, int num_actions, char* name, int name_cap, long long* offset, int* ndim, int* shape) {
    auto L = model_leaves(model, num_actions);
    CB_CHECK(index >= 0 && index < (int)L.size(), "leaf index %d out of range", index);
    if (name && name_cap > 0) snprintf(name, name_cap, "%s", L[index].name.c_str());
    if (offset) *offset = L[index].offset;
    if (ndim) *ndim = L[index].ndim;
    if (shape) for (int i = 0; i < 4; ++i) shape[i] = L[index].shape[i];
    return 0;
}

void cb_destroy(cb_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->cfg.device);
    for (auto& g : c->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    for (void* p : c->allocs) cudaFree(p);
    if (c->side) cudaStreamDestroy(c->side);
    if (c->cap) cudaStreamDestroy(c->cap);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    if (c->nat) nature_destroy(c);
    delete c;
}

int cb_create(const cb_config* cfg, cb_ctx** out) {
    CB_CHECK(cfg && out, "null argument");
    CB_CHECK(cfg->max_batch > 0, "max_batch must be positive");
    CB_CHECK(cfg->num_actions > 0 && cfg->num_actions <= MAX_ACTIONS, "num_actions must be in [1,%d]", MAX_ACTIONS);
    int ndev = 0;
    CB_CUDA(cudaGetDeviceCount(&ndev));
    CB_CHECK(cfg->device >= 0 && cfg->device < ndev, "device %d not available (%d CUDA devices)", cfg->device, ndev);
    CB_CUDA(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    CB_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
    CB_CHECK(prop.major == 10, "libcleanba_b200 needs an sm_100a (Blackwell B200) device, found sm_%d%d", prop.major, prop.minor);
    cb_ctx* c = new cb_ctx();
    c->cfg = *cfg;
    c->num_sms = prop.multiProcessorCount;
    c->A = cfg->num_actions;
    c->fuse0 = cfg->conv_backend == CB_CONV_TCGEN05 && !getenv("CLEANBA_NO_FUSE0");
    c->fuse12 = cfg->conv_backend == CB_CONV_TCGEN05 && !getenv("CLEANBA_NO_FUSE12");
    const bool nature = cfg->model == CB_MODEL_NATURE_CNN;
    if (cfg->model != CB_MODEL_IMPALA_RESNET && !nature) { set_error("unknown model %d", cfg->model); delete c; return -1; }
    if (nature && cfg->conv_backend != CB_CONV_TCGEN05) { set_error("the Nature-CNN trunk has no CUDA-core cross-check backend"); delete c; return -1; }
    c->HID = nature ? 512 : HIDDEN;
    c->leaves = model_leaves(cfg->model, c->A);
    c->nparam = c->leaves.back().offset + c->leaves.back().size();
    bool ok = false;
    do {
        void* p;
        if (dev_alloc(c, &p, c->nparam * sizeof(float))) break;
        c->params = (float*)p;
        if (cfg->train) {
            if (dev_alloc(c, &p, c->nparam * sizeof(float))) break;
            c->m = (float*)p;
            if (dev_alloc(c, &p, c->nparam * sizeof(float))) break;
            c->v = (float*)p;
            if (dev_alloc(c, &p, OPT_BLOCKS * sizeof(float))) break;
            c->opt_partials = (float*)p;
        }
        bool fail = false;
        if (nature && nature_create(c)) break;
        // conv layer table + packed weights
        std::vector<PackLayer> pl(15);
        for (int s = 0; s < 3 && !fail && !nature; ++s)
            for (int k = 0; k < 5 && !fail; ++k) {
                int li = s * 5 + k;
                ConvLayer& L = c->conv[li];
                L.cin = (k == 0) ? kStageCin[s] : kStageC[s];
                L.cout = kStageC[s];
                L.off_b = c->leaves[s * 10 + k * 2].offset;
                L.off_w = c->leaves[s * 10 + k * 2 + 1].offset;
                long long ef = packed_conv_elems((L.cin + 7) / 8, L.cout);
                if (dev_alloc(c, &p, ef * sizeof(f16))) { fail = true; break; }
                L.fwd = (f16*)p;
                L.dg = nullptr;
                if (li != 0) {
                    long long ed = packed_conv_elems(L.cout / 8, L.cin);
                    if (dev_alloc(c, &p, ed * sizeof(f16))) { fail = true; break; }
                    L.dg = (f16*)p;
                }
                pl[li].w = c->params + L.off_w; pl[li].cin = L.cin; pl[li].cout = L.cout;
                pl[li].fwd = L.fwd; pl[li].dg = L.dg;
            }
        if (fail) break;
        if (!nature) {
            if (dev_alloc(c, &p, 15 * sizeof(PackLayer))) break;
            c->pack_dev = (PackLayer*)p;
            if (cudaMemcpy(c->pack_dev, pl.data(), 15 * sizeof(PackLayer), cudaMemcpyHostToDevice) != cudaSuccess) {
                set_error("cudaMemcpy(pack table) failed");
                break;
            }
            c->off_dense_b = c->leaves[30].offset; c->off_dense_w = c->leaves[31].offset;
            c->off_actor_b = c->leaves[32].offset; c->off_actor_w = c->leaves[33].offset;
            c->off_critic_b = c->leaves[34].offset; c->off_critic_w = c->leaves[35].offset;
        }
        // activations
        for (int s = 0; s < 3 && !fail && !nature; ++s) {
            Stage& S = c->st[s];
            const int C = kStageC[s], Hin = kStageHin[s], Ho = kStageHout[s];
            const bool fused = (s == 0) ? c->fuse0 : c->fuse12;
            if (s == 0) fail |= alloc_act(c, S.x, 8, Hin, false) != 0;
            if (!fused) fail |= alloc_act(c, S.y, C, Hin) != 0;
            fail |= alloc_act(c, S.p, C, Ho) != 0;
            fail |= alloc_act(c, S.pr, C, Ho) != 0;
            fail |= alloc_act(c, S.a0, C, Ho) != 0;
            fail |= alloc_act(c, S.b0, C, Ho) != 0;
            fail |= alloc_act(c, S.b0r, C, Ho) != 0;
            fail |= alloc_act(c, S.a1, C, Ho) != 0;
            fail |= alloc_act(c, S.out, C, Ho) != 0;
            if (s < 2 && !fail) c->st[s + 1].x = S.out;
            if (cfg->train) {
                void* ap;
                if (dev_alloc(c, &ap, (size_t)cfg->max_batch * (Ho + 1) * (Ho + 1) * C)) { fail = true; break; }
                S.amax = (uint8_t*)ap;
                static const bool bits_on = [] { const char* e = getenv("CLEANBA_GATE_BITS"); return !e || atoi(e) != 0; }();
                if (bits_on && cfg->conv_backend == CB_CONV_TCGEN05) {      // relu gates as bits: [pixels to the tile boundary][C / 8] bytes
                    const size_t nb = (size_t)((((long long)cfg->max_batch * (Ho + 1) * (Ho + 1) + 127) / 128) * 128) * (C / 8);
                    uint8_t** dst[4] = {&S.bits_a0, &S.bits_b0r, &S.bits_a1, fused ? &S.bits_pr : nullptr};
                    for (auto d : dst) {
                        if (!d) continue;
                        if (dev_alloc(c, &ap, nb)) { fail = true; break; }
                        *d = (uint8_t*)ap;
                    }
                }
                fail |= alloc_act(c, S.gA, C, Ho) != 0;
                fail |= alloc_act(c, S.gB, C, Ho) != 0;
                fail |= alloc_act(c, S.gC, C, Ho) != 0;
                if (!(s == 0 && c->fuse0)) fail |= alloc_act(c, S.gBin, C, Hin) != 0;
            }
        }
        if (fail) break;
        if (!nature) {      // table of every activation / gradient plane for clear_trailing_rows
            std::vector<TrailRow> rows;
            for (int s = 0; s < 3; ++s) {
                Stage& S = c->st[s];
                const Act* acts[] = {s == 0 ? &S.x : nullptr, &S.y, &S.p, &S.pr, &S.a0, &S.b0, &S.b0r, &S.a1, &S.out, &S.gA, &S.gB, &S.gC, &S.gBin};
                for (const Act* a : acts) {
                    if (!a || !a->pl.hi) continue;
                    const ConvGeom g = make_geom(1, a->H, a->H);
                    for (int j = 0; j < (a->C + 7) / 8; ++j)
                        for (f16* base : {a->pl.hi, a->pl.mid})
                            if (base) rows.push_back(TrailRow{base + (long long)j * a->pl.plane_px * 8, g.P, g.Wp});
                }
            }
            if (dev_alloc(c, &p, rows.size() * sizeof(TrailRow))) break;
            c->trail_dev = (TrailRow*)p;
            c->trail_count = (int)rows.size();
            if (cudaMemcpy(c->trail_dev, rows.data(), rows.size() * sizeof(TrailRow), cudaMemcpyHostToDevice) != cudaSuccess) {
                set_error("cudaMemcpy(trailing-row table) failed");
                break;
            }
        }
        const size_t mb = (size_t)cfg->max_batch;
        if (!nature && cfg->conv_backend == CB_CONV_TCGEN05 && !getenv("CLEANBA_DENSE_SIMT")) {
            c->npad_max = (cfg->max_batch + 127) / 128 * 128;
            bool bad = false;
            for (int i = 0; i < 3 && !bad; ++i) {
                if (dev_alloc(c, &p, dense_featT_elems(c->npad_max) * sizeof(bf16))) { bad = true; break; }
                c->ft[i] = (bf16*)p;
            }
            if (bad) break;
            if (dev_alloc(c, &p, dense_pack_fwd_elems() * sizeof(bf16))) break;
            c->wd_fwd = (bf16*)p;
            if (dev_alloc(c, &p, dense_pack_dx_elems() * sizeof(bf16))) break;
            c->wd_dx = (bf16*)p;
            if (cfg->train) {
                for (int i = 0; i < 2 && !bad; ++i) {
                    if (dev_alloc(c, &p, dense_dpreT_elems(c->npad_max) * sizeof(bf16))) { bad = true; break; }
                    c->dpT[i] = (bf16*)p;
                }
                if (bad) break;
            }
            if (dense_umma_init()) break;
        }
        if (dev_alloc(c, &p, mb * c->HID * sizeof(float))) break;
        c->hidden = (float*)p;
        // split-K partial sums of the dense forward (dense_umma.cu: 11 splits up to 384 samples, 4 up to 1920, else none) and of
        // the SIMT path (dense.cu: 11 up to 256, 4 up to 1024): size for the worst batch <= max_batch
        size_t part = mb * HIDDEN;
        if (part < (size_t)11 * (mb < 384 ? mb : 384) * HIDDEN) part = (size_t)11 * (mb < 384 ? mb : 384) * HIDDEN;
        if (part < (size_t)4 * (mb < 1920 ? mb : 1920) * HIDDEN) part = (size_t)4 * (mb < 1920 ? mb : 1920) * HIDDEN;
        if (part < (size_t)64 * HIDDEN) part = (size_t)64 * HIDDEN;       // column-sum partials of the dense bias gradient
        if (dev_alloc(c, &p, part * sizeof(float))) break;
        c->dense_part = (float*)p;
        if (dev_alloc(c, &p, 2 * sizeof(uint32_t))) break;
        c->subkey = (uint32_t*)p;
        if (dev_alloc(c, &p, 2 * sizeof(uint32_t))) break;
        c->key_tmp = (uint32_t*)p;
        if (cfg->train) {
            if (dev_alloc(c, &p, mb * c->HID * sizeof(float))) break;
            c->dpre = (float*)p;
            if (dev_alloc(c, &p, mb * (MAX_ACTIONS + 1) * sizeof(float))) break;
            c->dlogits = (float*)p;
            if (dev_alloc(c, &p, mb * 8 * sizeof(float))) break;
            c->terms = (float*)p;
            if (dev_alloc(c, &p, mb * (MAX_ACTIONS + 1) * sizeof(float))) break;
            c->logits_scratch = (float*)p;
            if (dev_alloc(c, &p, mb * 8 * sizeof(float))) break;
            c->cell_scratch = (float*)p;
            // partial sums of the weight-gradient kernels (<= 3.2 M floats for the conv shapes) and of the head weight gradients
            // (one slice of (HID + 1) x (A + 1) floats per 16 samples: grows with max_batch)
            c->wg_cap = 4LL * 1024 * 1024;
            const long long head_need = ((long long)mb / 16 + 1) * (c->HID + 1) * (c->A + 1);
            if (c->wg_cap < head_need) c->wg_cap = head_need;
            if (dev_alloc(c, &p, (size_t)c->wg_cap * sizeof(float))) break;
            c->wg_partial = (float*)p;
            if (dev_alloc(c, &p, 2 * sizeof(float))) break;
            c->gscale = (float*)p;
            if (dev_alloc(c, &p, 2 * sizeof(unsigned))) break;
            c->gs_work = (unsigned*)p;
            if (dev_alloc(c, &p, sizeof(StepPtrs))) break;
            c->step_dev = (StepPtrs*)p;
            static const bool side_on = [] { const char* e = getenv("CLEANBA_WGRAD_SIDE"); return !e || atoi(e) != 0; }();
            if (side_on && !nature && cfg->conv_backend == CB_CONV_TCGEN05) {
                if (cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking) != cudaSuccess ||
                    cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
                    cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming) != cudaSuccess) {
                    set_error("cannot create the weight-gradient side stream");
                    break;
                }
            }
        }
        ok = true;
    } while (0);
    if (!ok) {
        cb_destroy(c);
        return -1;
    }
    *out = c;
    return 0;
}

int cb_set_params(cb_ctx* c, const float* src, cb_stream stream) {
    CB_CHECK(c && src, "null argument");
    CB_CUDA(cudaSetDevice(c->cfg.device));
    if (copy_any(c->params, src, c->nparam * sizeof(float), (cudaStream_t)stream)) return -1;
    return refresh_weights(c, (cudaStream_t)stream);
}
int cb_get_params(cb_ctx* c, float* dst, cb_stream stream) {
    CB_CHECK(c && dst, "null argument");
    CB_CUDA(cudaSetDevice(c->cfg.device));
    return copy_any(dst, c->params, c->nparam * sizeof(float), (cudaStream_t)stream);
}
float* cb_params_ptr(cb_ctx* c) { return c ? c->params : nullptr; }
int cb_refresh_weights(cb_ctx* c, cb_stream stream) {
    CB_CHECK(c, "null argument");
    CB_CUDA(cudaSetDevice(c->cfg.device));
    return refresh_weights(c, (cudaStream_t)stream);
}
int cb_publish_params(cb_ctx* dst, cb_ctx* src, cb_stream stream) {
    CB_CHECK(dst && src, "null argument");
    CB_CHECK(dst->nparam == src->nparam, "parameter count mismatch");
    CB_CUDA(cudaSetDevice(dst->cfg.device));
    if (dst->cfg.device == src->cfg.device) {
        if (copy_any(dst->params, src->params, dst->nparam * sizeof(float), (cudaStream_t)stream)) return -1;
    } else {
        CB_CUDA(cudaMemcpyPeerAsync(dst->params, dst->cfg.device, src->params, src->cfg.device, dst->nparam * sizeof(float),
                                    (cudaStream_t)stream));
    }
    return refresh_weights(dst, (cudaStream_t)stream);
}
int cb_get_opt_state(cb_ctx* c, float* m, float* v, long long* count, cb_stream stream) {
    CB_CHECK(c && c->cfg.train, "not a learner context");
    CB_CUDA(cudaSetDevice(c->cfg.device));
    if (m && copy_any(m, c->m, c->nparam * sizeof(float), (cudaStream_t)stream)) return -1;
    if (v && copy_any(v, c->v, c->nparam * sizeof(float), (cudaStream_t)stream)) return -1;
    if (count) *count = c->opt_count;
    return 0;
}
int cb_set_opt_state(cb_ctx* c, const float* m, const float* v, long long count, cb_stream stream) {
    CB_CHECK(c && c->cfg.train, "not a learner context");
    CB_CUDA(cudaSetDevice(c->cfg.device));
    if (m && copy_any(c->m, m, c->nparam * sizeof(float), (cudaStream_t)stream)) return -1;
    if (v && copy_any(c->v, v, c->nparam * sizeof(float), (cudaStream_t)stream)) return -1;
    c->opt_count = count;
    return 0;
}

int cb_actor_step(cb_ctx* c, const uint8_t* obs, int n, uint32_t* key, int32_t* action, float* logprob, float* value,
                  float* logits, cb_stream stream) {
    CB_CHECK(c && obs && key && action, "null argument");
    CB_CUDA(cudaSetDevice(c->cfg.device));
    cudaStream_t st = (cudaStream_t)stream;
    if (trunk_forward(c, obs, nullptr, n, st)) return -1;
    if (launch_split_key(key, c->subkey, st)) return -1;   // key, subkey = jax.random.split(key)
    ProfScope ps(c, "actor_head", 2.0 * n * c->HID * (c->A + 1), (double)n * (c->HID * 4 + 12), st);
    return launch_actor_head(c->hidden, n, c->A, c->params + c->off_actor_w, c->params + c->off_actor_b,
                             c->params + c->off_critic_w, c->params + c->off_critic_b, c->subkey, logits, value, action,
                             logprob, st, nullptr, c->HID);
}

int cb_actor_step_cursor(cb_ctx* c, cb_rollout_cursor* cursor, int n, uint32_t* key, cb_stream stream) {
    CB_CHECK(c && cursor && key, "null argument");
    CB_CUDA(cudaSetDevice(c->cfg.device));
    cudaStream_t st = (cudaStream_t)stream;
    // key, subkey = jax.random.split(key); the same kernel advances cursor->row, which unpack and the head then read
    if (launch_split_key(key, c->subkey, st, cursor)) return -1;
    c->cursor = cursor;
    const int rc = trunk_forward(c, nullptr, nullptr, n, st);
    c->cursor = nullptr;
    if (rc) return -1;
    ProfScope ps(c, "actor_head", 2.0 * n * c->HID * (c->A + 1), (double)n * (c->HID * 4 + 12), st);
    return launch_actor_head(c->hidden, n, c->A, c->params + c->off_actor_w, c->params + c->off_actor_b,
                             c->params + c->off_critic_w, c->params + c->off_critic_b, c->subkey, nullptr, nullptr,
                             reinterpret_cast<int*>(c->dense_part), nullptr, st, cursor, c->HID);
}

// forward-only heads (no sampling): reuse the actor head kernel with a scratch action buffer
int cb_policy_value(cb_ctx* c, const uint8_t* obs, const int32_t* idx, int n, float* logits, float* value, cb_stream stream) {
    CB_CHECK(c && obs, "null argument");
    CB_CUDA(cudaSetDevice(c->cfg.device));
    cudaStream_t st = (cudaStream_t)stream;
    if (trunk_forward(c, obs, idx, n, st)) return -1;
    // scratch for the (unused) sampled actions: the dense partial buffer is free after the forward
    int* scratch_act = reinterpret_cast<int*>(c->dense_part);
    return launch_actor_head(c->hidden, n, c->A, c->params + c->off_actor_w, c->params + c->off_actor_b,
                             c->params + c->off_critic_w, c->params + c->off_critic_b, c->subkey, logits, value,
                             scratch_act, nullptr, st, nullptr, c->HID);
}

int cb_gae(cb_ctx* c, const float* rewards, const float* values, const uint8_t* dones, const float* next_value,
           const uint8_t* next_done, int T, int B, float gamma, float gae_lambda, int num_groups, float* adv, float* ret,
           cb_stream stream) {
    CB_CHECK(c && rewards && values && dones && next_value && next_done && adv && ret, "null argument");
    CB_CHECK(T > 0 && B > 0, "empty rollout T=%d B=%d", T, B);
    CB_CUDA(cudaSetDevice(c->cfg.device));
    ProfScope ps(c, "gae_advnorm_scan", 0, 17.0 * T * B, (cudaStream_t)stream);
    return launch_gae(rewards, values, dones, next_value, next_done, T, B, gamma, gae_lambda, num_groups, adv, ret,
                      (cudaStream_t)stream);
}

int cb_split_key(cb_ctx* c, uint32_t* key, uint32_t* subkey, cb_stream stream) {
    CB_CHECK(c && key && subkey, "null argument");
    CB_CUDA(cudaSetDevice(c->cfg.device));
    return launch_split_key(key, subkey, (cudaStream_t)stream);
}

int cb_permutation(cb_ctx* c, const uint32_t* key, int n, int32_t* out, cb_stream stream) {
    CB_CHECK(c && key && out, "null argument");
    CB_CHECK(n >= 0, "negative n");
    if (n == 0) return 0;
    CB_CUDA(cudaSetDevice(c->cfg.device));
    cudaStream_t st = (cudaStream_t)stream;
    if (n > c->perm_cap) {
        // growth replaces the scratch buffers: the old ones are released once the stream has drained
        if (c->perm_cap) {
            CB_CUDA(cudaStreamSynchronize(st));
            for (void* old : {(void*)c->perm_tmp, (void*)c->sort_keys, (void*)c->perm_rank}) {
                for (size_t i = 0; i < c->allocs.size(); ++i)
                    if (c->allocs[i] == old) { c->allocs.erase(c->allocs.begin() + i); break; }
                cudaFree(old);
            }
        }
        void* p;
        if (dev_alloc(c, &p, (size_t)n * sizeof(int))) return -1;
        c->perm_tmp = (int*)p;
        if (dev_alloc(c, &p, (size_t)n * sizeof(uint32_t))) return -1;
        c->sort_keys = (uint32_t*)p;
        if (dev_alloc(c, &p, (size_t)n * sizeof(int))) return -1;
        c->perm_rank = (int*)p;
        c->perm_cap = n;
    }
    // num_rounds = ceil(3 ln n / ln(2^32 - 1))  (jax._src.random._shuffle)
    int rounds = (int)ceil(3.0 * log((double)(n > 1 ? n : 1)) / log(4294967295.0));
    if (copy_any(c->key_tmp, key, 2 * sizeof(uint32_t), st)) return -1;
    ProfScope ps(c, "permutation", 0, 12.0 * n, st);
    return launch_permutation(c->key_tmp, n, rounds, out, c->perm_tmp, c->sort_keys, c->perm_rank, c->subkey, st);
}

int cb_ppo_grad(cb_ctx* c, const uint8_t* obs, const int32_t* idx, int mb, const int32_t* actions, const float* logprobs,
                const float* advantages, const float* returns, float clip_coef, float ent_coef, float vf_coef, float* grads,
                float* stats, cb_stream stream) {
    CB_CHECK(c && obs && actions && logprobs && advantages && returns && grads && stats, "null argument");
    CB_CHECK(c->cfg.train, "cb_ppo_grad needs a learner context (train=1)");
    CB_CHECK(mb > 0 && mb <= c->cfg.max_batch, "batch %d outside (0, max_batch=%d]", mb, c->cfg.max_batch);
    CB_CUDA(cudaSetDevice(c->cfg.device));
    cudaStream_t st = (cudaStream_t)stream;
    auto body = [&](cudaStream_t s) {
        return ppo_grad_body(c, obs, idx, mb, actions, logprobs, advantages, returns, clip_coef, ent_coef, vf_coef, grads, stats, s);
    };
    if (!c->graph_on || c->prof_on) return body(st);
    StepGraph key;
    key.kind = 0; key.n = mb; key.T1 = key.B = 0; key.grads = grads; key.c0 = clip_coef; key.c1 = ent_coef; key.c2 = vf_coef;
    key.milestone = c->milestone;
    StepPtrs p = {{obs, idx, actions, logprobs, advantages, returns, stats, nullptr}};
    return graphed_step(c, key, p, st, body);
}

int cb_impala_grad(cb_ctx* c, const uint8_t* obs, const int32_t* idx, int T1, int B, const int32_t* actions,
                   const float* behaviour_logits, const float* rewards, const uint8_t* dones, const uint8_t* firststeps,
                   float gamma, float vf_coef, float ent_coef, float* grads, float* stats, cb_stream stream) {
    CB_CHECK(c && obs && actions && behaviour_logits && rewards && dones && firststeps && grads && stats, "null argument");
    CB_CHECK(c->cfg.train, "cb_impala_grad needs a learner context (train=1)");
    CB_CHECK(T1 >= 2 && B >= 1, "need T+1 >= 2 rows and B >= 1 columns");
    CB_CHECK((long long)T1 * B <= c->cfg.max_batch, "batch %lld outside (0, max_batch=%d]", (long long)T1 * B, c->cfg.max_batch);
    CB_CUDA(cudaSetDevice(c->cfg.device));
    cudaStream_t st = (cudaStream_t)stream;
    auto body = [&](cudaStream_t s) {
        return impala_grad_body(c, obs, idx, T1, B, actions, behaviour_logits, rewards, dones, firststeps, gamma, vf_coef, ent_coef,
                                grads, stats, s);
    };
    if (!c->graph_on || c->prof_on) return body(st);
    StepGraph key;
    key.kind = 1; key.n = T1 * B; key.T1 = T1; key.B = B; key.grads = grads; key.c0 = gamma; key.c1 = vf_coef; key.c2 = ent_coef;
    key.milestone = c->milestone;
    StepPtrs p = {{obs, idx, actions, behaviour_logits, rewards, dones, stats, firststeps}};
    return graphed_step(c, key, p, st, body);
}

static int optimizer_step(cb_ctx* c, const float* const* grads, int num_grads, float grad_scale, float lr, float max_norm,
                          float* norm_out, cb_stream stream);

int cb_optimizer_step(cb_ctx* c, const float* grads, float grad_scale, float lr, float max_norm, float* norm_out,
                      cb_stream stream) {
    CB_CHECK(c && grads, "null argument");
    return optimizer_step(c, &grads, 1, grad_scale, lr, max_norm, norm_out, stream);
}

int cb_optimizer_step_peers(cb_ctx* c, const float* const* grads, int num_grads, float grad_scale, float lr, float max_norm,
                            float* norm_out, cb_stream stream) {
    CB_CHECK(c && grads, "null argument");
    CB_CHECK(num_grads >= 1 && num_grads <= OPT_MAX_PEERS, "num_grads must be in [1,%d]", OPT_MAX_PEERS);
    for (int k = 0; k < num_grads; ++k) CB_CHECK(grads[k], "null gradient buffer %d", k);
    return optimizer_step(c, grads, num_grads, grad_scale, lr, max_norm, norm_out, stream);
}

int cb_reduce_peers(cb_ctx* c, const float* const* grads, int num_grads, float* out, cb_stream stream) {
    CB_CHECK(c && grads && out, "null argument");
    CB_CHECK(num_grads >= 1 && num_grads <= OPT_MAX_PEERS, "num_grads must be in [1,%d]", OPT_MAX_PEERS);
    CB_CUDA(cudaSetDevice(c->cfg.device));
    OptArgs o;
    memset(&o, 0, sizeof(o));
    o.n = c->nparam; o.ng = num_grads;
    for (int k = 0; k < num_grads; ++k) { CB_CHECK(grads[k], "null gradient buffer %d", k); o.gp[k] = grads[k]; }
    CB_CHECK((reinterpret_cast<uintptr_t>(out) & 15) == 0, "out must be 16-byte aligned");
    ProfScope ps(c, "grad_reduce_peers", 0, (double)c->nparam * 4 * (num_grads + 1), (cudaStream_t)stream);
    return launch_reduce_peers(o, out, (cudaStream_t)stream);
}

int cb_grad_accumulate(cb_ctx* c, float* acc, const float* grads, int mini_step, cb_stream stream) {
    CB_CHECK(c && acc && grads && mini_step >= 0, "bad argument");
    CB_CUDA(cudaSetDevice(c->cfg.device));
    return launch_grad_accumulate(acc, grads, c->nparam, mini_step, (cudaStream_t)stream);
}

int cb_memcpy_2d(void* dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t width_bytes, size_t rows, cb_stream stream) {
    CB_CHECK(dst && src, "null argument");
    CB_CHECK(width_bytes <= dst_pitch && width_bytes <= src_pitch, "row width %zu exceeds a pitch (%zu, %zu)", width_bytes, dst_pitch, src_pitch);
    if (!width_bytes || !rows) return 0;
    // device -> (peer) device with 16-byte granularity: copy with the SMs of the source GPU (stores over NVLink); anything else
    // (host memory, odd sizes) goes through the DMA engines
    cudaPointerAttributes sa, da;
    const bool okq = cudaPointerGetAttributes(&sa, src) == cudaSuccess && cudaPointerGetAttributes(&da, dst) == cudaSuccess;
    if (!okq) (void)cudaGetLastError();
    const bool vec = ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src) | dst_pitch | src_pitch | width_bytes) & 15) == 0;
    static const bool use_kernel = [] { const char* e = getenv("CLEANBA_COPY_KERNEL"); return !e || atoi(e) != 0; }();
    if (use_kernel && okq && vec && sa.type == cudaMemoryTypeDevice && da.type == cudaMemoryTypeDevice) {
        CB_CUDA(cudaSetDevice(sa.device));      // the stream must belong to the source device (the actor's copy streams do)
        return launch_copy_2d(dst, dst_pitch, src, src_pitch, width_bytes, rows, (cudaStream_t)stream);
    }
    CB_CUDA(cudaMemcpy2DAsync(dst, dst_pitch, src, src_pitch, width_bytes, rows, cudaMemcpyDefault, (cudaStream_t)stream));
    return 0;
}

int cb_hidden_width(cb_ctx* c) { return c ? c->HID : 0; }

int cb_set_grad_milestone(cb_ctx* c, void* cuda_event, long long* tail_offset) {
    CB_CHECK(c, "null argument");
    c->milestone = (cudaEvent_t)cuda_event;
    if (tail_offset) *tail_offset = c->off_dense_b;      // flax order: network Dense_0/bias is the first leaf after the conv stages
    return 0;
}

int cb_set_actor_tail(cb_ctx* c, int cluster_size) {
    CB_CHECK(c, "null argument");
    CB_CHECK(cluster_size >= 0 && cluster_size <= 2, "cluster_size must be 0 (off), 1 or 2");
    CB_CHECK(!cluster_size || (!c->nat && c->cfg.conv_backend == CB_CONV_TCGEN05), "the persistent tail exists for the tcgen05 IMPALA-ResNet trunk only");
    c->actor_tail = cluster_size;
    return 0;
}

int cb_graph_steps(cb_ctx* c, int enable) {
    CB_CHECK(c, "null argument");
    CB_CHECK(!enable || c->cfg.train, "cb_graph_steps needs a learner context (train=1)");
    CB_CUDA(cudaSetDevice(c->cfg.device));
    if (enable && !c->cap) CB_CUDA(cudaStreamCreateWithFlags(&c->cap, cudaStreamNonBlocking));
    c->graph_on = enable != 0;
    return 0;
}

long long cb_graph_replays(cb_ctx* c) { return c ? c->graph_replays : 0; }

int cb_set_sm_budget(cb_ctx* c, int num_sms) {
    CB_CHECK(c, "null argument");
    CB_CUDA(cudaSetDevice(c->cfg.device));
    cudaDeviceProp prop;
    CB_CUDA(cudaGetDeviceProperties(&prop, c->cfg.device));
    CB_CHECK(num_sms >= 1 && num_sms <= prop.multiProcessorCount, "num_sms must be in [1,%d]", prop.multiProcessorCount);
    c->num_sms = num_sms;
    return 0;
}

int cb_enable_peer_access(cb_ctx* c, int peer_device) {
    CB_CHECK(c, "null argument");
    if (peer_device == c->cfg.device) return 0;
    CB_CUDA(cudaSetDevice(c->cfg.device));
    int can = 0;
    CB_CUDA(cudaDeviceCanAccessPeer(&can, c->cfg.device, peer_device));
    CB_CHECK(can, "device %d cannot access peer device %d", c->cfg.device, peer_device);
    cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) { (void)cudaGetLastError(); return 0; }
    CB_CUDA(e);
    return 0;
}

static int optimizer_step(cb_ctx* c, const float* const* grads, int num_grads, float grad_scale, float lr, float max_norm,
                          float* norm_out, cb_stream stream) {
    CB_CHECK(c->cfg.train, "cb_optimizer_step needs a learner context (train=1)");
    CB_CUDA(cudaSetDevice(c->cfg.device));
    cudaStream_t st = (cudaStream_t)stream;
    OptArgs o;
    memset(&o, 0, sizeof(o));
    o.n = c->nparam; o.p = c->params; o.g = grads[0]; o.m = c->m; o.v = c->v;
    o.ng = num_grads;
    for (int k = 0; k < num_grads; ++k) o.gp[k] = grads[k];
    o.grad_scale = grad_scale; o.max_norm = max_norm; o.lr = lr;
    o.partials = c->opt_partials; o.norm_out = norm_out;
    c->opt_count += 1;
    if (c->cfg.algo == CB_ALGO_PPO) {
        o.kind = 0; o.b1 = 0.9f; o.b2 = 0.999f; o.eps = 1e-5f;
        o.bc1 = 1.0f - powf(0.9f, (float)c->opt_count);
        o.bc2 = 1.0f - powf(0.999f, (float)c->opt_count);
    } else {
        o.kind = 1; o.b1 = 0.f; o.b2 = 0.99f; o.eps = 0.01f; o.bc1 = o.bc2 = 1.f;
    }
    {
        ProfScope ps(c, c->cfg.algo == CB_ALGO_PPO ? "clip_adam" : "clip_rmsprop", 0,
                     (double)c->nparam * (c->cfg.algo == CB_ALGO_PPO ? 28 : 20), st);
        if (launch_optimizer(o, st)) return -1;
    }
    return refresh_weights(c, st);
}

long long cb_launch_count(void) { return g_launches.load(); }

int cb_profile(cb_ctx* c, int enable) {
    CB_CHECK(c, "null argument");
    CB_CUDA(cudaSetDevice(c->cfg.device));
    for (auto& r : c->prof_recs) { c->prof_pool.push_back(r.a); c->prof_pool.push_back(r.b); }
    c->prof_recs.clear();
    c->prof_on = enable != 0;
    return 0;
}

int cb_profile_report(cb_ctx* c, char* buf, int cap) {
    CB_CHECK(c && buf && cap > 2, "null argument");
    CB_CUDA(cudaSetDevice(c->cfg.device));
    CB_CUDA(cudaDeviceSynchronize());
    std::map<std::string, ProfAgg> agg;
    for (auto& r : c->prof_recs) {
        float ms = 0.f;
        CB_CUDA(cudaEventElapsedTime(&ms, r.a, r.b));
        ProfAgg& a = agg[r.name];
        a.launches += r.launches; a.records += 1; a.ms += ms; a.flops += r.flops; a.bytes += r.bytes; a.abytes += r.abytes;
    }
    std::string out = "[";
    bool first = true;
    for (auto& kv : agg) {
        char line[384];
        snprintf(line, sizeof(line), "%s{\"name\": \"%s\", \"calls\": %lld, \"records\": %lld, \"ms\": %.6f, \"flops\": %.6e, \"bytes\": %.6e, \"abytes\": %.6e}",
                 first ? "" : ", ", kv.first.c_str(), kv.second.launches, kv.second.records, kv.second.ms, kv.second.flops, kv.second.bytes,
                 kv.second.abytes);
        out += line;
        first = false;
    }
    out += "]";
    CB_CHECK((int)out.size() + 1 <= cap, "profile report needs %d bytes", (int)out.size() + 1);
    memcpy(buf, out.c_str(), out.size() + 1);
    return 0;
}

long long cb_debug_tensor(cb_ctx* c, const char* name, float* host_out, long long cap) {
    CB_CHECK(c && name && host_out, "null argument");
    CB_CUDA(cudaSetDevice(c->cfg.device));
    CB_CUDA(cudaDeviceSynchronize());
    const int n = c->last_n;
    if (!strcmp(name, "hidden") || !strcmp(name, "dpre")) {
        const float* src = !strcmp(name, "hidden") ? c->hidden : c->dpre;
        CB_CHECK(src, "tensor %s not allocated", name);
        long long cnt = (long long)n * c->HID;
        CB_CHECK(cnt <= cap, "buffer too small");
        CB_CUDA(cudaMemcpy(host_out, src, cnt * sizeof(float), cudaMemcpyDeviceToHost));
        return cnt;
    }
    CB_CHECK(!c->nat, "only \"hidden\" and \"dpre\" are exposed for the Nature-CNN trunk");
    CB_CHECK(strlen(name) >= 4 && (name[0] == 's' || name[0] == 'g') && name[2] == '.', "bad tensor name %s", name);
    int s = name[1] - '0';
    CB_CHECK(s >= 0 && s < 3, "bad stage in %s", name);
    Stage& S = c->st[s];
    const char* f = name + 3;
    const Act* a = nullptr;
    float unscale = 1.f;
    if (name[0] == 's') {
        if (!strcmp(f, "x")) a = &S.x;
        else if (!strcmp(f, "y")) {
            CB_CHECK(S.y.pl.hi, "tensor s%d.y is not materialised (conv fused with its max-pool)", s);
            a = &S.y;
        }
        else if (!strcmp(f, "p")) a = &S.p;
        else if (!strcmp(f, "pr")) a = &S.pr;
        else if (!strcmp(f, "a0")) a = &S.a0;
        else if (!strcmp(f, "b0")) a = &S.b0;
        else if (!strcmp(f, "b0r")) a = &S.b0r;
        else if (!strcmp(f, "a1")) a = &S.a1;
        else if (!strcmp(f, "out")) a = &S.out;
    } else {
        if (!strcmp(f, "A")) a = &S.gA;
        else if (!strcmp(f, "B")) a = &S.gB;
        else if (!strcmp(f, "C")) a = &S.gC;
        else if (!strcmp(f, "Bin")) a = &S.gBin;
        float gs[2] = {1.f, 1.f};
        if (c->gscale) CB_CUDA(cudaMemcpy(gs, c->gscale, sizeof(gs), cudaMemcpyDeviceToHost));
        unscale = gs[1];                                  // gradient tensors carry the loss scale
    }
    CB_CHECK(a && a->pl.hi != nullptr, "tensor %s not available", name);
    const int H = a->H, C = a->C, Hp = H + 1, P = Hp * Hp, chunks = (C + 7) / 8;
    const long long NP = (long long)n * P;
    const long long cnt = (long long)n * H * H * C;
    CB_CHECK(cnt <= cap, "buffer too small (%lld > %lld)", cnt, cap);
    std::vector<float> tmp((size_t)chunks * NP * 8, 0.f);
    {
        std::vector<f16> h((size_t)NP * 8);
        for (int j = 0; j < chunks; ++j) {
            const f16* planes[2] = {a->pl.hi, a->pl.mid};
            for (int pi = 0; pi < 2; ++pi) {
                if (!planes[pi]) continue;
                CB_CUDA(cudaMemcpy(h.data(), planes[pi] + (long long)j * a->pl.plane_px * 8, h.size() * sizeof(f16), cudaMemcpyDeviceToHost));
                const float w = (pi == 0 ? 1.f : MID_INV) * unscale;
                for (size_t i = 0; i < h.size(); ++i) tmp[(size_t)j * NP * 8 + i] += __half2float(h[i]) * w;
            }
        }
    }
    for (int b = 0; b < n; ++b)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < H; ++x)
                for (int ch = 0; ch < C; ++ch) {
                    long long q = (long long)b * P + (long long)(y + 1) * Hp + (x + 1);
                    host_out[(((long long)b * H + y) * H + x) * C + ch] = tmp[((size_t)(ch / 8) * NP + q) * 8 + ch % 8];
                }
    return cnt;
}

}  // extern "C"
#pragma GCC visibility pop
