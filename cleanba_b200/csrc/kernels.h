// Internal launcher declarations (host side) for libcleanba_b200.
#pragma once
#include "common.cuh"
#include "../../include/cleanba_b200.h"

namespace cb {

// Per-step pointers of a GRAPHED learner step (cb_graph_steps): the captured kernels read them from this device-side table,
// which one tiny launch rewrites before every replay, instead of from their (frozen) launch arguments.
//   p[0] obs  p[1] idx  p[2] actions  p[3] old logprobs | behaviour logits  p[4] advantages | rewards  p[5] returns | dones
//   p[6] stats  p[7] firststeps
struct StepPtrs { const void* p[8]; };

struct WgradArgs {
    ConvGeom g;
    Planes x;            // forward input planes of the conv
    int cin_chunks, cin_real;
    Planes gy;           // gradient planes w.r.t. the conv output
    int cout;
    float* dw;           // [3][3][cin_real][cout] fp32 (HWIO, flax order)
    float* db;           // [cout]
    float scale;         // applied to dW only (1/255 for the first conv, whose forward folds x/255 into the epilogue)
    const float* inv_scale;   // device scalar 1 / (loss scale carried by gy), applied to dW and db; null = 1
};

// trunk_simt.cu
int launch_unpack(const uint8_t* obs, const int* idx, int n, f16* out_hi, cudaStream_t st, const cb_rollout_cursor* cursor = nullptr,
                  const StepPtrs* ind = nullptr);
int launch_conv_simt(const ConvArgs& a, cudaStream_t st);
int launch_pool_fwd(Planes in, ConvGeom gi, ConvGeom go, int pad_lo, int chunks, Planes out, Planes out_relu,
                    uint8_t* amax, cudaStream_t st);
int launch_pool_bwd(const uint8_t* amax, Planes dpool, ConvGeom gi, ConvGeom go, int pad_lo, int chunks, Planes out,
                    cudaStream_t st);
int launch_wgrad_simt(const WgradArgs& a, float* partial, int max_blocks, cudaStream_t st);
// first ConvSequence: pool backward fused with the frame conv's weight gradient (no 84x84 gradient tensor)
int launch_pool_bwd_wgrad0(const uint8_t* amax, Planes dpool, const f16* x_hi, ConvGeom gi, ConvGeom go, float scale,
                           const float* inv_scale, float* dw, float* db, float* partial, int num_sms, cudaStream_t st);

// conv_umma.cu (tcgen05 + TMA bulk copies)
int launch_conv_umma(const ConvArgs& a, int num_sms, cudaStream_t st);
int launch_wgrad_umma(const WgradArgs& a, float* partial, int num_sms, cudaStream_t st);
int umma_conv_smem_bytes(int cin_chunks, int cout, int Wp);
// frame conv fused with its max-pool (84x84x4 -> pooled 42x42x16: stream, relu'd planes, arg-max bytes)
int launch_conv0_pool_umma(const ConvArgs& a, Planes out, Planes out_r, uint8_t* amax, uint8_t* bits, int num_sms, cudaStream_t st);
// sequence conv of the second / third ConvSequence fused with its max-pool (same outputs on the pooled grid `go`)
int launch_conv_pool_umma(const ConvArgs& a, ConvGeom go, int pad_lo, Planes out, Planes out_r, uint8_t* amax, uint8_t* bits, int num_sms,
                          cudaStream_t st);

// dense.cu
struct DenseArgs {
    int n;                       // samples
    Planes x;                    // final feature planes (relu'd), 4 chunks on the 12x12 padded grid (shared borders)
    const float* w;              // [3872][256] fp32 master (k = (h*11+w)*32 + c)
    const float* b;              // [256]
    float* hidden;               // [n][256]  relu(x W + b)
};
int dense_fwd_splits(int n);
int launch_dense_fwd(const DenseArgs& a, float* part, cudaStream_t st);
// dpre: [n][256] gradient w.r.t. the pre-relu dense output.
int launch_dense_bwd_w(const DenseArgs& a, const float* dpre, float* dw, float* db, cudaStream_t st);
// dX -> gradient w.r.t. the (pre-relu) trunk output, masked by the forward relu, times the loss scale *gscale, as carrier planes
int launch_dense_bwd_x(const DenseArgs& a, const float* dpre, const float* gscale, Planes out, cudaStream_t st);

// dense_umma.cu (tcgen05 dense layer on the sample-minor copies)
struct DenseUmmaArgs {
    int n, npad;                 // samples, samples rounded up to 128
    long long NP;                // n * 144 (flat pixels of the 11x11 grid with its shared zero borders)
    const bf16 *ft_hi, *ft_mid, *ft_lo;   // featT[chunk 4][pixel 124][npad][8]
    const bf16 *dp_hi, *dp_mid;           // dpreT[256/8][npad][8]
    const bf16 *w_fwd, *w_dx;             // packed weights (k_pack_dense)
    const float* bias;
    float* hidden;               // forward out [n][256]
    float* part;                 // forward partial sums [psplit][n][256] when the pixels are split over CTAs
    const float* gscale;         // device scalar: loss scale S applied to dX (the trunk's gradient tensors carry it)
    Planes out;                  // dX planes out (fp16x2 carrier)
};
int dense_umma_init();
int launch_pack_dense(const float* w, bf16* fwd, bf16* dx, cudaStream_t st);
long long dense_pack_fwd_elems();
long long dense_pack_dx_elems();
long long dense_featT_elems(int npad);
long long dense_dpreT_elems(int npad);
int launch_dense_fwd_umma(const DenseUmmaArgs& a, cudaStream_t st);
int launch_dense_bwd_umma(const DenseUmmaArgs& a, const float* dpre, float* dw, float* db, float* scratch, cudaStream_t st);
int launch_dpre_transpose(const float* dpre, int n, int npad, bf16* dp_hi, bf16* dp_mid, cudaStream_t st);

// gemm_umma.cu (generic tcgen05 GEMMs on carrier "row planes": the Nature-CNN trunk)
struct RowPlanes {               // X[R, K] as plane[K / 8][rpad][8], hi / mid carrier planes (mid null: exact single plane)
    const f16* hi;
    const f16* mid;
    long long rpad;
};
struct GemmArgs {                // C[R, N] = A[R, K] * W[K, N]
    RowPlanes a;
    int K;                       // multiple of 64
    long long R, Rpad;           // valid rows; rows padded to 128 (outputs of padded rows are written as zeros)
    const f16* wp;               // packed weights (launch_pack_gemm) for the N block size used
    int N;                       // output columns (multiple of the N block)
    const float* bias;           // [N] or null
    float acc_scale;
    int relu;
    f16 *out_hi, *out_mid;       // carrier planes out [N / 8][out_rpad][8], or null
    long long out_rpad;
    float* out_f32;              // row-major fp32 out [R][out_ld], or null
    int out_ld;
};
struct GemmWgradArgs {           // dW[K, N] = scale * inv_scale * A[R, K]^T * G[R, N]
    RowPlanes a, g;
    int K, N;
    long long Rpad;
    float scale;
    const float* inv_scale;      // device scalar or null
    float* dw;                   // [K][N]
};
int launch_gemm_umma(const GemmArgs& a, int NB, int num_sms, cudaStream_t st);
int launch_gemm_wgrad_umma(const GemmWgradArgs& a, float* partial, long long partial_cap, int num_sms, cudaStream_t st);
long long gemm_pack_elems(int Kl, int Nl, int transpose, int NB);
int launch_pack_gemm(const float* w, int Kl, int Nl, int transpose, int NB, f16* out, cudaStream_t st);

// heads.cu
int launch_split_key(uint32_t* key_inout, uint32_t* subkey_out, cudaStream_t st, cb_rollout_cursor* cursor = nullptr);
int launch_actor_head(const float* hidden, int n, int num_actions, const float* wa, const float* ba, const float* wc,
                      const float* bc, const uint32_t* subkey, float* logits_out, float* value_out, int* action_out,
                      float* logprob_out, cudaStream_t st, const cb_rollout_cursor* cursor = nullptr, int hid = HIDDEN);
struct PpoHeadArgs {
    int n, num_actions;
    int hid;                     // width of `hidden` (256 | 512)
    const float* hidden;         // [n][hid]
    const float *wa, *ba, *wc, *bc;
    const int* idx;              // minibatch sample indices into the flat [T*B] fields (null = identity)
    const int* actions;          // flat fields of the whole update
    const float* old_logprobs;
    const float* advantages;
    const float* returns;
    float clip_coef, ent_coef, vf_coef;
    float* dpre;                 // [n][256] gradient w.r.t. pre-relu dense output
    float* dlogits;              // [n][num_actions + 1] (last column = dvalue) scratch for the head weight grads
    float* terms;                // [n][5] per-sample loss terms scratch
    float* stats;                // [5] loss, pg_loss, v_loss, entropy, approx_kl
    float* wgrad_scratch;        // [ceil(n/32)][257][A+1] partial head weight gradients
    float *dwa, *dba, *dwc, *dbc;
    const StepPtrs* ind;         // graphed step: idx / actions / old_logprobs / advantages / returns / stats come from here
};
int launch_ppo_head(const PpoHeadArgs& a, cudaStream_t st);
struct ImpalaHeadArgs {
    int T1, B, num_actions;      // T1 = T + 1 rows; frames are ordered f = t * B + b
    int hid;                     // width of `hidden` (256 | 512)
    const float* hidden;
    const float *wa, *ba, *wc, *bc;
    const int* idx;              // [T1*B] indices into the flat [T1*Bl] fields (null = identity)
    const int* actions;
    const float* behaviour_logits;   // flat [T1*Bl][A]
    const float* rewards;
    const uint8_t* dones;
    const uint8_t* firststeps;
    float gamma, vf_coef, ent_coef;
    float* logits_scratch;       // [T1*B][A + 1] logits and value
    float* cell_scratch;         // [T1*B][8]
    float* dpre;
    float* dlogits;
    float* stats;                // [4] total, pg, baseline, entropy
    float* wgrad_scratch;
    float *dwa, *dba, *dwc, *dbc;
    const StepPtrs* ind;         // graphed step: the per-step pointers come from here
};
int launch_impala_head(const ImpalaHeadArgs& a, cudaStream_t st);

// learner_misc.cu
int launch_gae(const float* rewards, const float* values, const uint8_t* dones, const float* next_value,
               const uint8_t* next_done, int T, int B, float gamma, float lambda, int num_groups, float* adv, float* ret,
               cudaStream_t st);
int launch_permutation(uint32_t* key_inout, int n, int rounds, int* out, int* tmp, uint32_t* sort_keys, int* rank,
                       uint32_t* subkey, cudaStream_t st);
struct OptArgs {
    int kind;                    // 0 = Adam, 1 = RMSProp (PyTorch style)
    long long n;
    float* p;
    const float* g;
    float* m;                    // Adam first moment (unused for RMSProp)
    float* v;                    // Adam second moment / RMSProp nu
    float grad_scale;            // 1 / L for the pmean over learner devices
    float max_norm, lr, b1, b2, eps, bc1, bc2;
    float* partials;             // [OPT_BLOCKS]
    float* norm_out;             // [1] global norm (after grad_scale), optional
    // Fused gradient exchange: when ng > 1 the gradient is the fixed-order sum gp[0][i] + gp[1][i] + ... of the flat gradient
    // buffers of ALL learner replicas, read straight from peer memory over NVLink (g is ignored); every replica evaluates the
    // same sum in the same order, so the replicas stay bit-identical without a separate allreduce pass.
    const float* gp[8];
    int ng;
};
constexpr int OPT_MAX_PEERS = 8;
constexpr int OPT_BLOCKS = 296;
int launch_optimizer(const OptArgs& a, cudaStream_t st);
// gscale[0] = S, gscale[1] = 1 / S for this minibatch's gradient tensors (work: 2 zero-initialised words)
int launch_loss_scale(const float* dpre, long long count, unsigned* work, float* gscale, cudaStream_t st);
int launch_copy_2d(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t rows, cudaStream_t st);
int launch_grad_accumulate(float* acc, const float* g, long long n, int mini_step, cudaStream_t st);
int launch_set_step_ptrs(StepPtrs* dst, const StepPtrs& v, cudaStream_t st);
struct TrailRow { f16* base; int P, Wp; };      // one 8-channel plane (flat pixel 0), pixels per image, row length
int launch_clear_trailing_rows(const TrailRow* rows, int count, int n, cudaStream_t st);

// actor_fused.cu: ConvSequence 1 and 2 of the actor's forward pass (ten convs, two pools) in one persistent kernel, one
// thread-block cluster per frame.  The ConvArgs are exactly what the per-layer launches would get.
struct ActorTailHost {
    int n;
    ConvArgs pool[2];          // the sequence convs (fused with their pools): 16 -> 32 at 42x42, 32 -> 32 at 21x21
    ConvGeom pool_go[2];       // pooled grids
    int pad_lo[2];
    Planes pool_out[2], pool_out_r[2];
    ConvArgs conv[8];          // residual-block convs, in execution order
};
int launch_actor_tail(const ActorTailHost& h, int csize, int num_sms, cudaStream_t st);
int launch_reduce_peers(const OptArgs& a, float* out, cudaStream_t st);   // needs n, gp, ng only

// pack.cu
struct PackLayer {
    const float* w;              // HWIO master
    int cin, cout;               // real channel counts
    f16* fwd;                    // packed forward image ([hi|mid] stacked along N)
    f16* dg;                     // packed dgrad image (null for the first conv)
};
int launch_pack_conv(const PackLayer* layers_dev, int nlayers, cudaStream_t st);
long long packed_conv_elems(int cin_chunks, int cout);

}  // namespace cb
