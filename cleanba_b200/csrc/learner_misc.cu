// Learner-side scalar kernels: GAE scan + advantage normalisation, the minibatch permutation and the fused
// global-norm-clip + Adam / RMSProp step.
//   GAE        : compute_gae / compute_gae_once   cleanba/cleanba_ppo.py:532-560
//   adv norm   : cleanba/cleanba_ppo.py:592-595
//   shuffle    : jax.random.permutation in update_epoch  cleanba/cleanba_ppo.py:599-615
//   optimizers : optax chain cleanba/cleanba_ppo.py:492-500, rmsprop_pytorch_style cleanba/cleanba_impala.py:152-188
#include "common.cuh"
#include "kernels.h"
#include "prng.cuh"

namespace cb {

// ------------------------------------------------------------------------------------------------
// GAE: one warp per env column.  The 32 lanes load 4 consecutive time steps each (all loads in flight at once),
// then the recurrence  A[t] = delta[t] + (gamma*lambda*nonterminal[t+1]) * A[t+1]  is carried lane to lane with warp
// shuffles, in exactly the sequential order (and with the un-fused fp32 mul/add) of the reference scan, so the
// raw advantages are bit-identical to the CPU restatement.  One block per advantage-normalisation column group.
__global__ void __launch_bounds__(1024) k_gae(const float* __restrict__ rewards, const float* __restrict__ values,
                                              const uint8_t* __restrict__ dones, const float* __restrict__ next_value,
                                              const uint8_t* __restrict__ next_done, int T, int B, float gamma,
                                              float gamma_lambda, int num_groups, int normalize, float* __restrict__ adv,
                                              float* __restrict__ ret) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int Bg = B / num_groups;
    const int col0 = blockIdx.x * Bg;
    for (int cb_ = warp; cb_ < Bg; cb_ += nwarp) {
        const int b = col0 + cb_;
        float carry = 0.f;  // A[T] = 0
        for (int end = T; end > 0; end -= 128) {
            const int base = end - 128;  // may be negative: those slots are invalid
            float delta[4], coef[4], val[4];
            bool ok[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                int t = base + lane * 4 + i;
                ok[i] = (t >= 0);
                delta[i] = 0.f; coef[i] = 0.f; val[i] = 0.f;
                if (ok[i]) {
                    float r = rewards[(long long)t * B + b];
                    float v = values[(long long)t * B + b];
                    float vn = (t + 1 < T) ? values[(long long)(t + 1) * B + b] : next_value[b];
                    float dn = (t + 1 < T) ? (float)dones[(long long)(t + 1) * B + b] : (float)next_done[b];
                    float nn = __fsub_rn(1.0f, dn);
                    // delta = reward + gamma * nextvalues * nextnonterminal - curvalues
                    delta[i] = __fsub_rn(__fadd_rn(r, __fmul_rn(__fmul_rn(gamma, vn), nn)), v);
                    coef[i] = __fmul_rn(gamma_lambda, nn);
                    val[i] = v;
                }
            }
            float a_out[4] = {0.f, 0.f, 0.f, 0.f};
            float a_next = carry;
            for (int L = 31; L >= 0; --L) {
                float a_run = a_next;
                if (lane == L) {
#pragma unroll
                    for (int i = 3; i >= 0; --i)
                        if (ok[i]) {
                            a_run = __fadd_rn(delta[i], __fmul_rn(coef[i], a_run));
                            a_out[i] = a_run;
                        }
                }
                a_next = __shfl_sync(0xffffffffu, a_run, L);
            }
            carry = a_next;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                int t = base + lane * 4 + i;
                if (ok[i]) {
                    adv[(long long)t * B + b] = a_out[i];
                    ret[(long long)t * B + b] = __fadd_rn(a_out[i], val[i]);
                }
            }
        }
    }
    if (!normalize) return;
    // per-group mean / population std over (T x Bg) elements, fixed summation order
    __shared__ float red[1024];
    __shared__ float s_mean, s_std;
    __syncthreads();
    const int cnt = T * Bg;
    float s = 0.f;
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) s += adv[(long long)(i / Bg) * B + col0 + (i % Bg)];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) s_mean = red[0] / (float)cnt;
    __syncthreads();
    const float mean = s_mean;
    s = 0.f;
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
        float d = adv[(long long)(i / Bg) * B + col0 + (i % Bg)] - mean;
        s += d * d;
    }
    __syncthreads();
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) s_std = sqrtf(red[0] / (float)cnt);
    __syncthreads();
    const float denom = s_std + 1e-8f;
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
        long long o = (long long)(i / Bg) * B + col0 + (i % Bg);
        adv[o] = (adv[o] - mean) / denom;
    }
}

int launch_gae(const float* rewards, const float* values, const uint8_t* dones, const float* next_value,
               const uint8_t* next_done, int T, int B, float gamma, float lambda, int num_groups, float* adv, float* ret,
               cudaStream_t st) {
    int normalize = num_groups > 0;
    int groups = normalize ? num_groups : 1;
    CB_CHECK(B % groups == 0, "gae: B=%d not divisible by num_groups=%d", B, groups);
    float gl = (float)((double)gamma * (double)lambda);  // python-float product, then cast (cleanba_ppo.py:538)
    k_gae<<<groups, 1024, 0, st>>>(rewards, values, dones, next_value, next_done, T, B, gamma, gl, groups, normalize, adv, ret);
    CB_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------------
// jax.random.permutation(key, n): `rounds` rounds of { key, sub = split(key); keys = random_bits(sub, n);
// x = stable_sort_by_key(keys, x) }.  The stable sort is a rank-by-counting pass (n <= 65536 here, n^2 compares are
// microseconds on 148 SMs) which is deterministic and needs no scratch beyond the key array.
__global__ void k_iota(int* x, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = i;
}
__global__ void k_random_bits(const uint32_t* __restrict__ subkey, int n, uint32_t* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = jax_random_bits_elem(subkey[0], subkey[1], (uint32_t)i, (uint32_t)n);
}
// Stable rank-by-counting sort of n <= a few 10^4 (key, index) pairs: rank[i] = #{j : key[j] < key[i] or (key[j] == key[i] and j < i)}.
// The n^2 comparisons are split over blockIdx.y key segments (integer atomicAdd: exact, order independent) so that the grid
// covers the GPU (n = 15,360: 60 x 8 blocks instead of 60; 0.22 -> 0.04 ms per round); k_scatter_by_rank then permutes.
constexpr int RANK_SEGS = 8;
__global__ void __launch_bounds__(256) k_rank_count(const uint32_t* __restrict__ keys, int n, int* __restrict__ rank) {
    __shared__ uint32_t tile[2048];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t ki = i < n ? keys[i] : 0u;
    const int seg = (n + RANK_SEGS - 1) / RANK_SEGS;
    const int jlo = blockIdx.y * seg, jhi = min(n, jlo + seg);
    int r = 0;
    for (int j0 = jlo; j0 < jhi; j0 += 2048) {
        __syncthreads();
        for (int t = threadIdx.x; t < 2048; t += 256) tile[t] = (j0 + t < jhi) ? keys[j0 + t] : 0xffffffffu;
        __syncthreads();
        const int lim = min(2048, jhi - j0);
        for (int t = 0; t < lim; ++t) {
            const uint32_t kj = tile[t];
            r += (kj < ki) || (kj == ki && (j0 + t) < i);
        }
    }
    if (i < n && r) atomicAdd(rank + i, r);
}
__global__ void k_scatter_by_rank(const int* __restrict__ rank, const int* __restrict__ xin, int* __restrict__ xout, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) xout[rank[i]] = xin[i];
}

int launch_permutation(uint32_t* key_inout, int n, int rounds, int* out, int* tmp, uint32_t* sort_keys, int* rank, uint32_t* subkey,
                       cudaStream_t st) {
    int blocks = (n + 255) / 256;
    int* cur = (rounds % 2 == 0) ? out : tmp;   // after `rounds` swaps the result lands in `out`
    int* nxt = (rounds % 2 == 0) ? tmp : out;
    k_iota<<<blocks, 256, 0, st>>>(cur, n);
    CB_LAUNCH_CHECK();
    for (int r = 0; r < rounds; ++r) {
        if (launch_split_key(key_inout, subkey, st)) return -1;
        k_random_bits<<<blocks, 256, 0, st>>>(subkey, n, sort_keys);
        CB_LAUNCH_CHECK();
        CB_CUDA(cudaMemsetAsync(rank, 0, (size_t)n * sizeof(int), st));
        k_rank_count<<<dim3(blocks, RANK_SEGS), 256, 0, st>>>(sort_keys, n, rank);
        CB_LAUNCH_CHECK();
        k_scatter_by_rank<<<blocks, 256, 0, st>>>(rank, cur, nxt, n);
        CB_LAUNCH_CHECK();
        int* t = cur; cur = nxt; nxt = t;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Optimizer: pass 1 writes OPT_BLOCKS partial sums of (g * grad_scale)^2; pass 2 re-reduces them in a fixed order,
// applies clip_by_global_norm and the Adam / RMSProp update on the flat parameter vector.
// gradient element i: the local buffer, or the fixed-order sum over the replicas' buffers (peer memory)
__device__ __forceinline__ float opt_grad(const OptArgs& a, long long i) {
    if (a.ng <= 1) return a.g[i];
    float g = a.gp[0][i];
#pragma unroll 1
    for (int k = 1; k < a.ng; ++k) g += a.gp[k][i];
    return g;
}

__global__ void __launch_bounds__(256) k_sumsq(OptArgs a) {
    __shared__ float red[256];
    const long long n = a.n;
    long long per = (n + gridDim.x - 1) / gridDim.x;
    long long lo = (long long)blockIdx.x * per, hi = lo + per < n ? lo + per : n;
    float s = 0.f;
    for (long long i = lo + threadIdx.x; i < hi; i += 256) {
        float x = opt_grad(a, i) * a.grad_scale;
        s = fmaf(x, x, s);
    }
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) a.partials[blockIdx.x] = red[0];
}

__global__ void __launch_bounds__(256) k_opt_apply(OptArgs a) {
    __shared__ float red[512];
    __shared__ float s_norm;
    for (int t = threadIdx.x; t < 512; t += 256) red[t] = t < OPT_BLOCKS ? a.partials[t] : 0.f;
    __syncthreads();
    for (int o = 256; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        s_norm = sqrtf(red[0]);
        if (blockIdx.x == 0 && a.norm_out) a.norm_out[0] = s_norm;
    }
    __syncthreads();
    const float norm = s_norm;
    const bool clip = !(norm < a.max_norm);
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < a.n; i += (long long)gridDim.x * 256) {
        float g = opt_grad(a, i) * a.grad_scale;
        if (clip) g = (g / norm) * a.max_norm;      // optax.clip_by_global_norm: (g / norm) * max_norm
        float p = a.p[i];
        if (a.kind == 0) {
            float m = a.b1 * a.m[i] + (1.f - a.b1) * g;
            float v = a.b2 * a.v[i] + (1.f - a.b2) * g * g;
            a.m[i] = m; a.v[i] = v;
            float mhat = m / a.bc1, vhat = v / a.bc2;
            float u = mhat / (sqrtf(vhat) + a.eps);
            a.p[i] = p - a.lr * u;
        } else {
            float nu = a.b2 * a.v[i] + (1.f - a.b2) * g * g;
            a.v[i] = nu;
            float u = g / (sqrtf(nu) + a.eps);
            a.p[i] = p - a.lr * u;
        }
    }
}

// Loss scale of one minibatch's backward pass: S = 2^k with max |dpre| * S in [32, 64), chosen on the device (no host sync).
// The trunk's gradient tensors are fp16x2 carriers; S keeps them in fp16's range with >= 2^10 of head room for growth on the
// way down the trunk and >= 22 significant bits for every element above 2^-19 of the largest (common.cuh).  max is order
// independent, so the result is deterministic.  work = {max bits, finished-block counter}, both left at zero.
__global__ void __launch_bounds__(256) k_loss_scale(const float* __restrict__ dpre, long long count, unsigned* __restrict__ work,
                                                    float* __restrict__ gscale) {
    __shared__ float red[256];
    float m = 0.f;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < count; i += (long long)gridDim.x * 256) m = fmaxf(m, fabsf(dpre[i]));
    red[threadIdx.x] = m;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] = fmaxf(red[threadIdx.x], red[threadIdx.x + o]);
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        atomicMax(&work[0], __float_as_uint(red[0]));          // non-negative floats order like their bit patterns
        __threadfence();
        if (atomicAdd(&work[1], 1u) == gridDim.x - 1) {
            const float mx = __uint_as_float(atomicExch(&work[0], 0u));
            work[1] = 0u;
            int e = 0;
            float S = 1.f;
            if (mx > 0.f && mx < INFINITY) {
                (void)frexpf(mx, &e);                          // mx = f * 2^e, f in [0.5, 1)
                int k = 6 - e;
                k = k < -100 ? -100 : (k > 100 ? 100 : k);
                S = ldexpf(1.f, k);
            }
            gscale[0] = S;
            gscale[1] = 1.f / S;
        }
    }
}
int launch_loss_scale(const float* dpre, long long count, unsigned* work, float* gscale, cudaStream_t st) {
    k_loss_scale<<<64, 256, 0, st>>>(dpre, count, work, gscale);
    CB_LAUNCH_CHECK();
    return 0;
}

// Strided block copy by the SMs: `rows` rows of width16 16-byte words, src row pitch spitch16, dst row pitch dpitch16 (in words).
// dst may live on a PEER GPU: the stores go out over NVLink at several hundred GB/s, where cudaMemcpy2DAsync's DMA path
// measured 107 GB/s for the payload's 0.5 MB rows (profiles/r02_v2_config4_*).
__global__ void __launch_bounds__(256) k_copy_2d(uint4* __restrict__ dst, long long dpitch16, const uint4* __restrict__ src, long long spitch16,
                                                 long long width16, long long rows) {
    const long long total = width16 * rows;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const long long r = i / width16, c = i - r * width16;
        dst[r * dpitch16 + c] = src[r * spitch16 + c];
    }
}
int launch_copy_2d(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t rows, cudaStream_t st) {
    const long long total = (long long)(width / 16) * (long long)rows;
    long long blocks = (total + 255) / 256 / 8;
    if (blocks < 1) blocks = 1;
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_copy_2d<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<uint4*>(dst), (long long)(dpitch / 16), reinterpret_cast<const uint4*>(src),
                                                (long long)(spitch / 16), (long long)(width / 16), (long long)rows);
    CB_LAUNCH_CHECK();
    return 0;
}

// optax.MultiSteps accumulation (cleanba_ppo.py:492-500 with gradient_accumulation_steps > 1): acc <- acc + (g - acc) / (mini_step + 1)
__global__ void __launch_bounds__(256) k_grad_accumulate(float* __restrict__ acc, const float* __restrict__ g, long long n, float inv) {
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        const float a = acc[i];
        acc[i] = a + (g[i] - a) * inv;
    }
}
int launch_grad_accumulate(float* acc, const float* g, long long n, int mini_step, cudaStream_t st) {
    k_grad_accumulate<<<OPT_BLOCKS, 256, 0, st>>>(acc, g, n, 1.0f / (float)(mini_step + 1));
    CB_LAUNCH_CHECK();
    return 0;
}

// Shared-border layout (common.cuh): the zero row AFTER the last image of a batch of n is the first row of image n in a
// larger batch, so it holds stale data whenever the context has seen a larger batch.  One block per plane clears it.
__global__ void __launch_bounds__(128) k_clear_trailing_rows(const TrailRow* __restrict__ rows, int n) {
    const TrailRow r = rows[blockIdx.x];
    uint4* dst = reinterpret_cast<uint4*>(r.base + ((long long)n * r.P) * 8);
    for (int i = threadIdx.x; i < r.Wp + 1; i += blockDim.x) dst[i] = make_uint4(0, 0, 0, 0);
}
int launch_clear_trailing_rows(const TrailRow* rows, int count, int n, cudaStream_t st) {
    if (count <= 0) return 0;
    k_clear_trailing_rows<<<count, 128, 0, st>>>(rows, n);
    CB_LAUNCH_CHECK();
    return 0;
}

// graphed learner step: publish this step's pointers to the table the captured kernels read (launch arguments are copied at launch
// time, so the host may reuse `v` at once)
__global__ void k_set_step_ptrs(StepPtrs* __restrict__ dst, StepPtrs v) {
    if (threadIdx.x < 8) dst->p[threadIdx.x] = v.p[threadIdx.x];
}
int launch_set_step_ptrs(StepPtrs* dst, const StepPtrs& v, cudaStream_t st) {
    k_set_step_ptrs<<<1, 32, 0, st>>>(dst, v);
    CB_LAUNCH_CHECK();
    return 0;
}

// out[i] = gp[0][i] + gp[1][i] + ... (fixed order, peer memory): the in-process stage of a two-level gradient exchange
__global__ void __launch_bounds__(256) k_reduce_peers(OptArgs a, float* __restrict__ out) {
    for (long long i = ((long long)blockIdx.x * 256 + threadIdx.x) * 4; i < a.n; i += (long long)gridDim.x * 256 * 4) {
        if (i + 4 <= a.n) {
            float4 s = *reinterpret_cast<const float4*>(a.gp[0] + i);
#pragma unroll 1
            for (int k = 1; k < a.ng; ++k) {
                const float4 t = *reinterpret_cast<const float4*>(a.gp[k] + i);
                s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
            }
            *reinterpret_cast<float4*>(out + i) = s;
        } else {
            for (long long j = i; j < a.n; ++j) {
                float s = a.gp[0][j];
                for (int k = 1; k < a.ng; ++k) s += a.gp[k][j];
                out[j] = s;
            }
        }
    }
}
int launch_reduce_peers(const OptArgs& a, float* out, cudaStream_t st) {
    k_reduce_peers<<<OPT_BLOCKS, 256, 0, st>>>(a, out);
    CB_LAUNCH_CHECK();
    return 0;
}

int launch_optimizer(const OptArgs& a, cudaStream_t st) {
    k_sumsq<<<OPT_BLOCKS, 256, 0, st>>>(a);
    CB_LAUNCH_CHECK();
    k_opt_apply<<<OPT_BLOCKS, 256, 0, st>>>(a);
    CB_LAUNCH_CHECK();
    return 0;
}

}  // namespace cb
