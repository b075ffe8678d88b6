// Actor / critic heads, threefry Gumbel-max sampling and the fused loss heads (forward + backward).
//   actor   : get_action_and_value  cleanba/cleanba_ppo.py:245-261, get_action cleanba/cleanba_impala.py:287-301
//   PPO     : get_logprob_entropy_value + ppo_loss + its gradient  cleanba/cleanba_ppo.py:516-530,562-577,590
//   IMPALA  : impala_loss (V-trace, rlax 0.1.5 semantics) + its gradient  cleanba/cleanba_impala.py:557-597
#include "common.cuh"
#include "kernels.h"
#include "prng.cuh"

namespace cb {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Shared-memory copy of the two head matrices: [256][A] actor kernel, [256] critic kernel, biases.
struct HeadSmem {
    float* wa;   // [256 * A]
    float* wc;   // [256]
    float* ba;   // [A]
    float bc;
};

template <int HID>
__device__ __forceinline__ HeadSmem load_head_smem(float* sm, const float* wa, const float* ba, const float* wc,
                                                   const float* bc, int A) {
    HeadSmem h;
    h.wa = sm; h.wc = sm + HID * A; h.ba = h.wc + HID;
    for (int t = threadIdx.x; t < HID * A; t += blockDim.x) h.wa[t] = wa[t];
    for (int t = threadIdx.x; t < HID; t += blockDim.x) h.wc[t] = wc[t];
    for (int t = threadIdx.x; t < A; t += blockDim.x) h.ba[t] = ba[t];
    h.bc = bc[0];
    __syncthreads();
    return h;
}
static inline size_t head_smem_bytes(int HID, int A) { return (size_t)(HID * A + HID + A) * sizeof(float); }

// One warp computes the A logits and the value of one sample.  Lane a (< A) returns logit a; every lane returns value.
// The summation order is fixed, so the actor and the learner see bit-identical logits for identical hidden rows.
template <int HID>
__device__ __forceinline__ void head_forward(const HeadSmem& h, const float* hid /*[256]*/, int A, int lane,
                                             float hreg[HID / 32], float& mylogit, float& value) {
#pragma unroll
    for (int i = 0; i < HID / 32; ++i) hreg[i] = hid[lane + 32 * i];
    mylogit = -INFINITY;
    for (int a = 0; a < A; ++a) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < HID / 32; ++i) s = fmaf(hreg[i], h.wa[(lane + 32 * i) * A + a], s);
        s = warp_sum(s);
        if (lane == a) mylogit = s + h.ba[a];
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < HID / 32; ++i) s = fmaf(hreg[i], h.wc[lane + 32 * i], s);
    value = warp_sum(s) + h.bc;
}

// ------------------------------------------------------------------------------------------------
__global__ void k_split_key(uint32_t* key, uint32_t* subkey, cb_rollout_cursor* cursor) {
    griddep_launch();
    griddep_wait();
    if (threadIdx.x == 0) {
        if (cursor) cursor->row += 1;          // first kernel of a cursor step: the row every later kernel of the step uses
        uint32_t k0 = key[0], k1 = key[1], nk0, nk1, sk0, sk1;
        jax_split2(k0, k1, nk0, nk1, sk0, sk1);
        key[0] = nk0; key[1] = nk1; subkey[0] = sk0; subkey[1] = sk1;
    }
}
int launch_split_key(uint32_t* key_inout, uint32_t* subkey_out, cudaStream_t st, cb_rollout_cursor* cursor) {
    launch_pdl(k_split_key, dim3(1), dim3(32), 0, st, key_inout, subkey_out, cursor);
    CB_LAUNCH_CHECK();
    return 0;
}

// Sampling head: u = uniform(subkey, (n, A)); action = argmax(logits - log(-log u)) (first index on ties);
// logprob = log_softmax(logits)[action]; value.  One warp per sample.
template <int HID>
__global__ void __launch_bounds__(256) k_actor_head(const float* __restrict__ hidden, int n, int A, const float* wa,
                                                    const float* ba, const float* wc, const float* bc,
                                                    const uint32_t* __restrict__ subkey, float* logits_out,
                                                    float* value_out, int* action_out, float* logprob_out,
                                                    const cb_rollout_cursor* __restrict__ cursor) {
    extern __shared__ float sm[];
    griddep_launch();
    HeadSmem h = load_head_smem<HID>(sm, wa, ba, wc, bc, A);   // master parameters: not written by the preceding kernels
    griddep_wait();
    if (cursor) {   // outputs go to row cursor->row of the rollout storages
        const long long r = (long long)cursor->row * cursor->out_row_stride;
        action_out = reinterpret_cast<int*>(cursor->action) + r;
        logprob_out = cursor->logprob ? reinterpret_cast<float*>(cursor->logprob) + r : nullptr;
        value_out = cursor->value ? reinterpret_cast<float*>(cursor->value) + r : nullptr;
        logits_out = cursor->logits ? reinterpret_cast<float*>(cursor->logits) + r * A : nullptr;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const uint32_t k0 = subkey[0], k1 = subkey[1];
    const uint32_t total = (uint32_t)n * (uint32_t)A;
    for (int b = blockIdx.x * nwarp + warp; b < n; b += gridDim.x * nwarp) {
        float hreg[HID / 32], logit, value;
        head_forward<HID>(h, hidden + (long long)b * HID, A, lane, hreg, logit, value);
        float pert = -INFINITY;
        if (lane < A) {
            uint32_t bits = jax_random_bits_elem(k0, k1, (uint32_t)b * A + lane, total);
            float u = jax_bits_to_uniform(bits);
            pert = logit - logf(-logf(u));
        }
        // argmax with first-index tie break (a NaN-free input is assumed, as in the reference)
        float best = pert;
        int besti = lane < A ? lane : 0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            float ob = __shfl_xor_sync(0xffffffffu, best, o);
            int oi = __shfl_xor_sync(0xffffffffu, besti, o);
            if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
        }
        float m = warp_max(lane < A ? logit : -INFINITY);
        float se = warp_sum(lane < A ? expf(logit - m) : 0.f);
        float logp = logit - m - logf(se);
        float lp_a = __shfl_sync(0xffffffffu, logp, besti);
        if (logits_out && lane < A) logits_out[(long long)b * A + lane] = logit;
        if (lane == 0) {
            action_out[b] = besti;
            if (logprob_out) logprob_out[b] = lp_a;
            if (value_out) value_out[b] = value;
        }
    }
}

int launch_actor_head(const float* hidden, int n, int A, const float* wa, const float* ba, const float* wc,
                      const float* bc, const uint32_t* subkey, float* logits_out, float* value_out, int* action_out,
                      float* logprob_out, cudaStream_t st, const cb_rollout_cursor* cursor, int hid) {
    int blocks = (n + 7) / 8;
    if (blocks > 296) blocks = 296;
    CB_CHECK(hid == 256 || hid == 512, "heads: hidden width %d not built (256 | 512)", hid);
    if (hid == 256)
        launch_pdl(k_actor_head<256>, dim3(blocks), dim3(256), (size_t)head_smem_bytes(256, A), st, hidden, n, A, wa, ba, wc, bc, subkey,
                   logits_out, value_out, action_out, logprob_out, cursor);
    else
        launch_pdl(k_actor_head<512>, dim3(blocks), dim3(256), (size_t)head_smem_bytes(512, A), st, hidden, n, A, wa, ba, wc, bc, subkey,
                   logits_out, value_out, action_out, logprob_out, cursor);
    CB_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------------
// dpre[k] = (sum_a dl[a] Wa[k][a] + dv Wc[k]) * (hidden[k] > 0)   for the 8 k's of this lane
template <int HID>
__device__ __forceinline__ void head_backward_hidden(const HeadSmem& h, int A, int lane, const float hreg[HID / 32], float dl,
                                                     float dv, float* dpre_row) {
    float acc[HID / 32];
#pragma unroll
    for (int i = 0; i < HID / 32; ++i) acc[i] = dv * h.wc[lane + 32 * i];
    for (int a = 0; a < A; ++a) {
        float d = __shfl_sync(0xffffffffu, dl, a);
#pragma unroll
        for (int i = 0; i < HID / 32; ++i) acc[i] = fmaf(d, h.wa[(lane + 32 * i) * A + a], acc[i]);
    }
#pragma unroll
    for (int i = 0; i < HID / 32; ++i) dpre_row[lane + 32 * i] = hreg[i] > 0.f ? acc[i] : 0.f;
}

// PPO loss head, forward + backward, one warp per minibatch sample.
template <int HID>
__global__ void __launch_bounds__(256) k_ppo_head(PpoHeadArgs a) {
    extern __shared__ float sm[];
    const int A = a.num_actions;
    if (a.ind) {
        a.idx = static_cast<const int*>(a.ind->p[1]); a.actions = static_cast<const int*>(a.ind->p[2]);
        a.old_logprobs = static_cast<const float*>(a.ind->p[3]); a.advantages = static_cast<const float*>(a.ind->p[4]);
        a.returns = static_cast<const float*>(a.ind->p[5]);
    }
    HeadSmem h = load_head_smem<HID>(sm, a.wa, a.ba, a.wc, a.bc, A);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const float inv_n = 1.f / (float)a.n;
    for (int b = blockIdx.x * nwarp + warp; b < a.n; b += gridDim.x * nwarp) {
        float hreg[HID / 32], logit, value;
        head_forward<HID>(h, a.hidden + (long long)b * HID, A, lane, hreg, logit, value);
        const int src = a.idx ? a.idx[b] : b;
        const int act = min(max(a.actions[src], 0), A - 1);   // defensive: never index with a corrupt action
        const float oldlp = a.old_logprobs[src], adv = a.advantages[src], ret = a.returns[src];
        float m = warp_max(lane < A ? logit : -INFINITY);
        float ex = lane < A ? expf(logit - m) : 0.f;
        float se = warp_sum(ex);
        float logp = lane < A ? (logit - m - logf(se)) : 0.f;
        float p = ex / se;
        float ent = -warp_sum(lane < A ? p * logp : 0.f);
        float newlp = __shfl_sync(0xffffffffu, logp, act);
        float logratio = newlp - oldlp;
        float ratio = expf(logratio);
        float lo = 1.f - a.clip_coef, hi = 1.f + a.clip_coef;
        float clipped = fminf(fmaxf(ratio, lo), hi);
        float pg1 = -adv * ratio, pg2 = -adv * clipped;
        float pg = fmaxf(pg1, pg2);
        // d pg / d newlogprob  (jnp.maximum / jnp.clip sub-gradients; ties inside the clip range sum to -adv*ratio)
        float dpg;
        if (ratio >= lo && ratio <= hi) dpg = -adv * ratio;
        else dpg = (pg1 > pg2) ? -adv * ratio : 0.f;
        float verr = value - ret;
        float c_lp = dpg * inv_n;
        float dl = 0.f;
        if (lane < A) dl = c_lp * ((lane == act ? 1.f : 0.f) - p) + a.ent_coef * inv_n * p * (logp + ent);
        float dv = a.vf_coef * verr * inv_n;
        head_backward_hidden<HID>(h, A, lane, hreg, dl, dv, a.dpre + (long long)b * HID);
        if (lane < A) a.dlogits[(long long)b * (A + 1) + lane] = dl;
        if (lane == 0) {
            a.dlogits[(long long)b * (A + 1) + A] = dv;
            float* t = a.terms + (long long)b * 5;
            t[0] = pg; t[1] = 0.5f * verr * verr; t[2] = ent; t[3] = (ratio - 1.f) - logratio; t[4] = 0.f;
        }
    }
}

// stats = mean over samples of the per-sample terms (fixed summation order)
__global__ void __launch_bounds__(256) k_ppo_stats(const float* __restrict__ terms, int n, float ent_coef, float vf_coef,
                                                   float* __restrict__ stats, const StepPtrs* __restrict__ ind) {
    if (ind) stats = static_cast<float*>(const_cast<void*>(ind->p[6]));
    __shared__ float red[4][256];
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    for (int b = threadIdx.x; b < n; b += 256)
        for (int j = 0; j < 4; ++j) s[j] += terms[(long long)b * 5 + j];
    for (int j = 0; j < 4; ++j) red[j][threadIdx.x] = s[j];
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o)
            for (int j = 0; j < 4; ++j) red[j][threadIdx.x] += red[j][threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        float inv = 1.f / (float)n;
        float pg = red[0][0] * inv, vl = red[1][0] * inv, en = red[2][0] * inv, kl = red[3][0] * inv;
        stats[0] = pg - ent_coef * en + vl * vf_coef;
        stats[1] = pg; stats[2] = vl; stats[3] = en; stats[4] = kl;
    }
}

// Head weight gradients: dWa[k][a] = sum_b hidden[b][k] dl[b][a]; column A of dl is dvalue (critic); k == 256 is the bias.
// Two deterministic stages: every block reduces a slice of HW_SLICE samples (hidden tile staged in shared memory, one
// thread per hidden unit), then a second kernel adds the slices in a fixed order.
constexpr int HW_SLICE = 16;      // samples per slice: sh[16][512] floats = 32 KB of static shared memory at the widest trunk
template <int HID>
__global__ void __launch_bounds__(HID) k_head_wgrad_partial(const float* __restrict__ hidden, const float* __restrict__ dl,
                                                               int n, int A, float* __restrict__ partial) {
    __shared__ float sh[HW_SLICE][HID];
    __shared__ float sd[HW_SLICE][MAX_ACTIONS + 1];
    const int b0 = blockIdx.x * HW_SLICE, k = threadIdx.x;
    const int nb = min(HW_SLICE, n - b0);
    for (int b = 0; b < nb; ++b) sh[b][k] = hidden[(long long)(b0 + b) * HID + k];
    for (int t = threadIdx.x; t < nb * (A + 1); t += HID) sd[t / (A + 1)][t % (A + 1)] = dl[(long long)b0 * (A + 1) + t];
    __syncthreads();
    float* out = partial + (long long)blockIdx.x * (HID + 1) * (A + 1);
    for (int a = 0; a <= A; ++a) {
        float s = 0.f;
        for (int b = 0; b < nb; ++b) s = fmaf(sh[b][k], sd[b][a], s);
        out[k * (A + 1) + a] = s;
    }
    if (k <= A) {   // bias row
        float s = 0.f;
        for (int b = 0; b < nb; ++b) s += sd[b][k];
        out[HID * (A + 1) + k] = s;
    }
}
template <int HID>
__global__ void __launch_bounds__(128) k_head_wgrad_reduce(const float* __restrict__ partial, int nslices, int A, float* dwa,
                                                           float* dba, float* dwc, float* dbc) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (HID + 1) * (A + 1)) return;
    int k = t / (A + 1), a = t % (A + 1);
    float s = 0.f;
    for (int i = 0; i < nslices; ++i) s += partial[(long long)i * (HID + 1) * (A + 1) + t];
    if (k < HID) { if (a < A) dwa[k * A + a] = s; else dwc[k] = s; }
    else { if (a < A) dba[a] = s; else dbc[0] = s; }
}
template <int HID>
static int launch_head_wgrad(const float* hidden, const float* dl, int n, int A, float* scratch, float* dwa, float* dba,
                             float* dwc, float* dbc, cudaStream_t st) {
    int nslices = (n + HW_SLICE - 1) / HW_SLICE;
    k_head_wgrad_partial<HID><<<nslices, HID, 0, st>>>(hidden, dl, n, A, scratch);
    CB_LAUNCH_CHECK();
    int total = (HID + 1) * (A + 1);
    k_head_wgrad_reduce<HID><<<(total + 127) / 128, 128, 0, st>>>(scratch, nslices, A, dwa, dba, dwc, dbc);
    CB_LAUNCH_CHECK();
    return 0;
}

template <int HID>
static int launch_ppo_head_t(const PpoHeadArgs& a, cudaStream_t st) {
    int blocks = (a.n + 7) / 8;
    if (blocks > 592) blocks = 592;
    k_ppo_head<HID><<<blocks, 256, head_smem_bytes(HID, a.num_actions), st>>>(a);
    CB_LAUNCH_CHECK();
    k_ppo_stats<<<1, 256, 0, st>>>(a.terms, a.n, a.ent_coef, a.vf_coef, a.stats, a.ind);
    CB_LAUNCH_CHECK();
    return launch_head_wgrad<HID>(a.hidden, a.dlogits, a.n, a.num_actions, a.wgrad_scratch, a.dwa, a.dba, a.dwc, a.dbc, st);
}
int launch_ppo_head(const PpoHeadArgs& a, cudaStream_t st) {
    CB_CHECK(a.hid == 256 || a.hid == 512, "heads: hidden width %d not built (256 | 512)", a.hid);
    return a.hid == 256 ? launch_ppo_head_t<256>(a, st) : launch_ppo_head_t<512>(a, st);
}

// ------------------------------------------------------------------------------------------------
// IMPALA / V-trace loss head (a minibatch is [T1 = T+1, B] frames ordered f = t*B + b), three launches of one kernel:
//   phases = 1: logits + value of every frame            (one warp per frame, all SMs)
//   phases = 2: per-cell terms, the V-trace scan along T, loss scalars, d/dlogits, d/dvalue   (ONE block: the scan is a
//               sequential recurrence per column and the whole minibatch is <= a few thousand cells)
//   phases = 4: gradient w.r.t. the pre-relu dense output (one warp per frame, all SMs)
template <int HID>
__global__ void __launch_bounds__(1024) k_impala_head(ImpalaHeadArgs a, int phases) {
    extern __shared__ float sm[];
    if (a.ind) {
        a.idx = static_cast<const int*>(a.ind->p[1]); a.actions = static_cast<const int*>(a.ind->p[2]);
        a.behaviour_logits = static_cast<const float*>(a.ind->p[3]); a.rewards = static_cast<const float*>(a.ind->p[4]);
        a.dones = static_cast<const uint8_t*>(a.ind->p[5]); a.stats = static_cast<float*>(const_cast<void*>(a.ind->p[6]));
        a.firststeps = static_cast<const uint8_t*>(a.ind->p[7]);
    }
    const int A = a.num_actions, T1 = a.T1, B = a.B, T = T1 - 1;
    HeadSmem h = load_head_smem<HID>(sm, a.wa, a.ba, a.wc, a.bc, A);
    float* red = h.ba + A + 1;   // [3][1024]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int nf = T1 * B, nc = T * B;
    if (phases & 1) {
        // phase 0: logits + value of every frame
        for (int f = blockIdx.x * nwarp + warp; f < nf; f += gridDim.x * nwarp) {
            float hreg[HID / 32], logit, value;
            head_forward<HID>(h, a.hidden + (long long)f * HID, A, lane, hreg, logit, value);
            if (lane < A) a.logits_scratch[(long long)f * (A + 1) + lane] = logit;
            if (lane == 0) a.logits_scratch[(long long)f * (A + 1) + A] = value;
        }
    }
    if (phases & 4) {
        // phase 4: gradient w.r.t. the pre-relu dense output
        for (int f = blockIdx.x * nwarp + warp; f < nf; f += gridDim.x * nwarp) {
            float hreg[HID / 32];
#pragma unroll
            for (int i = 0; i < HID / 32; ++i) hreg[i] = a.hidden[(long long)f * HID + lane + 32 * i];
            float dl = lane < A ? a.dlogits[(long long)f * (A + 1) + lane] : 0.f;
            float dv = a.dlogits[(long long)f * (A + 1) + A];
            head_backward_hidden<HID>(h, A, lane, hreg, dl, dv, a.dpre + (long long)f * HID);
        }
    }
    if (!(phases & 2) || blockIdx.x != 0) return;
    // phase 1: per cell log pi(a), rho, entropy
    for (int c = threadIdx.x; c < nc; c += blockDim.x) {
        const int src = a.idx ? a.idx[c] : c;
        const int act = min(max(a.actions[src], 0), A - 1);
        const float* z = a.logits_scratch + (long long)c * (A + 1);
        const float* mu = a.behaviour_logits + (long long)src * A;
        float m = -INFINITY, mm = -INFINITY;
        for (int j = 0; j < A; ++j) { m = fmaxf(m, z[j]); mm = fmaxf(mm, mu[j]); }
        float se = 0.f, sm_ = 0.f;
        for (int j = 0; j < A; ++j) { se += expf(z[j] - m); sm_ += expf(mu[j] - mm); }
        float lse = m + logf(se), lsm = mm + logf(sm_);
        float ent = 0.f;
        for (int j = 0; j < A; ++j) {
            float lp = z[j] - lse, p = expf(z[j] - m) / se;
            if (p > 0.f) ent -= p * lp;
        }
        float lpa = z[act] - lse, lma = mu[act] - lsm;
        float* cs = a.cell_scratch + (long long)c * 8;
        cs[0] = lpa; cs[1] = expf(lpa - lma); cs[2] = ent; cs[3] = lse;
    }
    __syncthreads();
    // phase 2: V-trace backward scan per column (rlax.vtrace + vtrace_td_error_and_advantage, lambda = 1, clips = 1)
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        float err = 0.f;
        float target_next = 0.f;
        for (int t = T - 1; t >= 0; --t) {
            int c = t * B + b;
            const int src = a.idx ? a.idx[c] : c;
            float r = a.rewards[src];
            float disc = (1.f - (float)a.dones[src]) * a.gamma;
            float v_tm1 = a.logits_scratch[(long long)c * (A + 1) + A];
            float v_t = a.logits_scratch[(long long)(c + B) * (A + 1) + A];
            float* cs = a.cell_scratch + (long long)c * 8;
            float cr = fminf(1.f, cs[1]);
            float td = cr * (r + disc * v_t - v_tm1);
            err = td + disc * cr * err;
            float q_boot = (t == T - 1) ? v_t : target_next;
            float q = r + disc * q_boot;
            cs[4] = err;                   // errors (value gradient flows through -v_tm1 only)
            cs[5] = cr * (q - v_tm1);      // pg_advantage
            target_next = err + v_tm1;
        }
    }
    __syncthreads();
    // phase 3: per cell loss terms and d/dlogits, d/dvalue
    float s_pg = 0.f, s_bl = 0.f, s_en = 0.f;
    for (int c = threadIdx.x; c < nf; c += blockDim.x) {
        float* dl = a.dlogits + (long long)c * (A + 1);
        if (c >= nc) {
            for (int j = 0; j <= A; ++j) dl[j] = 0.f;
            continue;
        }
        const int src = a.idx ? a.idx[c] : c;
        const int act = min(max(a.actions[src], 0), A - 1);
        const float mask = 1.f - (float)a.firststeps[src];
        const float* z = a.logits_scratch + (long long)c * (A + 1);
        const float* cs = a.cell_scratch + (long long)c * 8;
        float lpa = cs[0], ent = cs[2], lse = cs[3], err = cs[4], adv = cs[5];
        s_pg += -lpa * adv * mask;
        s_bl += 0.5f * err * err * mask;
        s_en += -ent * mask;
        for (int j = 0; j < A; ++j) {
            float lp = z[j] - lse, p = expf(lp);
            float g = -adv * ((j == act ? 1.f : 0.f) - p) + a.ent_coef * p * (lp + ent);
            dl[j] = mask * g;
        }
        dl[A] = -a.vf_coef * err * mask;
    }
    red[threadIdx.x] = s_pg; red[1024 + threadIdx.x] = s_bl; red[2048 + threadIdx.x] = s_en;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            red[threadIdx.x] += red[threadIdx.x + o];
            red[1024 + threadIdx.x] += red[1024 + threadIdx.x + o];
            red[2048 + threadIdx.x] += red[2048 + threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        float pg = red[0], bl = red[1024], en = red[2048];
        a.stats[0] = pg + a.vf_coef * bl + a.ent_coef * en;
        a.stats[1] = pg; a.stats[2] = bl; a.stats[3] = en;
    }
}

template <int HID>
static int launch_impala_head_t(const ImpalaHeadArgs& a, cudaStream_t st) {
    size_t smem = head_smem_bytes(HID, a.num_actions) + (1 + 3 * 1024) * sizeof(float);
    if (smem > 48 * 1024) {      // the 512-wide trunk needs the opt-in limit (once per device, outside graph capture)
        static std::atomic<unsigned> attr_done{0};
        int dev = 0;
        CB_CUDA(cudaGetDevice(&dev));
        if (!(attr_done.load() & (1u << dev))) {
            CB_CUDA(cudaFuncSetAttribute(k_impala_head<HID>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
            attr_done.fetch_or(1u << dev);
        }
    }
    const int nf = a.T1 * a.B;
    int blocks = (nf + 7) / 8;                       // 8 warps (frames) per block
    if (blocks > 592) blocks = 592;
    k_impala_head<HID><<<blocks, 256, smem, st>>>(a, 1);
    CB_LAUNCH_CHECK();
    k_impala_head<HID><<<1, 1024, smem, st>>>(a, 2);
    CB_LAUNCH_CHECK();
    k_impala_head<HID><<<blocks, 256, smem, st>>>(a, 4);
    CB_LAUNCH_CHECK();
    return launch_head_wgrad<HID>(a.hidden, a.dlogits, a.T1 * a.B, a.num_actions, a.wgrad_scratch, a.dwa, a.dba, a.dwc, a.dbc, st);
}
int launch_impala_head(const ImpalaHeadArgs& a, cudaStream_t st) {
    CB_CHECK(a.hid == 256 || a.hid == 512, "heads: hidden width %d not built (256 | 512)", a.hid);
    return a.hid == 256 ? launch_impala_head_t<256>(a, st) : launch_impala_head_t<512>(a, st);
}

}  // namespace cb
