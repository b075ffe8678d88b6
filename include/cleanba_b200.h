/* libcleanba_b200 -- C ABI of the B200-native Sebulba hot path.
 *
 * The reference (vwxyzjn/cleanba) has no FFI: its hot path is a set of jitted Python callables over pytrees
 * (SURVEY.md section 8b).  Each entry point below replaces one of those callables, or one fused piece of it, and
 * cites the reference lines it stands in for.  All pointers are plain device pointers unless stated otherwise, all
 * work is enqueued on the caller's CUDA stream (no hidden host synchronisation), every call returns 0 on success or
 * -1 with a thread-local message retrievable through cb_last_error().  Contexts are independent, so Python threads
 * may call concurrently (ctypes releases the GIL).
 *
 * Parameter vector: ONE flat fp32 buffer in flax tree order (jax.tree_util.tree_leaves of
 * AgentParams(network_params, actor_params, critic_params), cleanba/cleanba_ppo.py:206-210): per ConvSequence
 * Conv_0/{bias,kernel}, ResidualBlock_{0,1}/Conv_{0,1}/{bias,kernel} (kernels HWIO), then Dense_0/{bias,kernel[in,out]},
 * actor Dense_0, critic Dense_0.  cb_leaf_info() enumerates the leaves.  Gradients and optimizer moments use the
 * same flat layout, so the data-parallel gradient allreduce (jax.lax.pmean, cleanba_ppo.py:628) is a single
 * ncclAllReduce on one buffer issued by the host between cb_*_grad() and cb_optimizer_step().
 */
#ifndef CLEANBA_B200_H_
#define CLEANBA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cb_ctx cb_ctx;
typedef void* cb_stream; /* cudaStream_t */

enum { CB_ALGO_PPO = 0, CB_ALGO_IMPALA = 1 };        /* Adam (eps 1e-5) vs PyTorch-style RMSProp (eps .01, decay .99) */
enum { CB_CONV_TCGEN05 = 0, CB_CONV_SIMT = 1 };      /* tensor-core kernels (product) or the fp32 CUDA-core cross-check */
/* Trunk: the IMPALA-ResNet of cleanba_ppo.py:149-189 (channels 16,32,32; hidden 256) or the Nature-CNN of
 * legacy_scripts/cleanba_ppo_envpool_impala_atari_wrapper_naturecnn.py:143-178 (8x8 s4 / 4x4 s2 / 3x3 s1 VALID convs, hidden 512;
 * tcgen05 only).  The parameter vector follows the flax tree of the selected trunk (cb_leaf_info_model). */
enum { CB_MODEL_IMPALA_RESNET = 0, CB_MODEL_NATURE_CNN = 1 };

typedef struct cb_config {
    int device;       /* CUDA device ordinal */
    int algo;         /* CB_ALGO_* : selects the optimizer state */
    int max_batch;    /* largest number of frames in one forward / one minibatch */
    int train;        /* 0: actor / inference context, 1: learner (allocates backward workspace + optimizer state) */
    int num_actions;  /* 18 for full_action_space Atari (cleanba_ppo.py:135) */
    int conv_backend; /* CB_CONV_* */
    int model;        /* CB_MODEL_* */
} cb_config;

/* ---- lifecycle / errors ------------------------------------------------------------------------------------ */
const char* cb_last_error(void);
int cb_version(void);
int cb_create(const cb_config* cfg, cb_ctx** out);
int cb_hidden_width(cb_ctx* ctx);   /* width of the trunk's dense output: 256 (IMPALA-ResNet) or 512 (Nature-CNN) */
void cb_destroy(cb_ctx* ctx);

/* ---- parameters (AgentParams / TrainState, cleanba_ppo.py:206-210,485-502) ---------------------------------- */
long long cb_num_params(int num_actions);
int cb_num_leaves(void);
/* name_cap bytes of `name` receive the flax path; shape has up to 4 entries. */
int cb_leaf_info(int index, int num_actions, char* name, int name_cap, long long* offset, int* ndim, int* shape);
/* The same three queries for a given trunk (CB_MODEL_*); the functions above answer for CB_MODEL_IMPALA_RESNET. */
long long cb_num_params_model(int model, int num_actions);
int cb_num_leaves_model(int model);
int cb_leaf_info_model(int model, int index, int num_actions, char* name, int name_cap, long long* offset, int* ndim, int* shape);
/* src / dst may be host or device memory (cudaMemcpyDefault). cb_set_params also refreshes the packed bf16 weights. */
int cb_set_params(cb_ctx* ctx, const float* src, cb_stream stream);
int cb_get_params(cb_ctx* ctx, float* dst, cb_stream stream);
/* Device pointer to the flat master parameters (for NCCL broadcast / peer copies); call cb_refresh_weights after writing. */
float* cb_params_ptr(cb_ctx* ctx);
int cb_refresh_weights(cb_ctx* ctx, cb_stream stream);
/* Param publish learner -> actor (jax.device_put of the unreplicated params, cleanba_ppo.py:721-725): device-to-device
 * (peer) copy of the master vector into dst and refresh of dst's packed weights, on `stream`. */
int cb_publish_params(cb_ctx* dst, cb_ctx* src, cb_stream stream);
/* Optimizer state: m (Adam only) and v / nu, flat; count = number of optimizer steps taken. */
int cb_get_opt_state(cb_ctx* ctx, float* m, float* v, long long* count, cb_stream stream);
int cb_set_opt_state(cb_ctx* ctx, const float* m, const float* v, long long count, cb_stream stream);

/* ---- actor ------------------------------------------------------------------------------------------------- */
/* get_action_and_value (cleanba_ppo.py:245-261) / get_action (cleanba_impala.py:287-301).
 * obs: uint8 [n,4,84,84] on the device.  key: uint32[2] on the device, advanced in place (key, subkey = split(key)).
 * Outputs (device): action int32[n]; logprob, value float[n] (may be NULL); logits float[n,num_actions] (may be NULL). */
int cb_actor_step(cb_ctx* ctx, const uint8_t* obs, int n, uint32_t* key, int32_t* action, float* logprob, float* value,
                  float* logits, cb_stream stream);

/* Rollout-storage form of cb_actor_step (prepare_data without the stack, cleanba_ppo.py:276-278, 342-356): the step reads its n
 * frames from row `row` of a frame storage and writes action / logprob / value / logits into row `row` of the rollout storages
 * described by a cursor that lives in DEVICE memory.  The step itself advances cursor->row BEFORE using it (initialise it to
 * first_row - 1), so ONE captured CUDA graph serves every step of every rollout: per step the host only copies the frames into
 * their storage row and replays the graph -- no staging buffers, no per-field copies.  Null (0) output pointers are skipped. */
typedef struct cb_rollout_cursor {
    unsigned long long obs;        /* uint8 frames of this actor: row r = obs + r * obs_row_stride bytes, n*4*84*84 bytes each */
    unsigned long long action;     /* int32: row r = action + r * out_row_stride elements */
    unsigned long long logprob;    /* float or 0 */
    unsigned long long value;      /* float or 0 */
    unsigned long long logits;     /* float, rows out_row_stride * num_actions elements apart, or 0 */
    long long obs_row_stride;      /* bytes */
    long long out_row_stride;      /* elements */
    int row;
    int reserved;
} cb_rollout_cursor;
int cb_actor_step_cursor(cb_ctx* ctx, cb_rollout_cursor* cursor_dev, int n, uint32_t* key, cb_stream stream);
/* Network + heads only (bootstrap value in compute_gae, cleanba_ppo.py:550-552). idx (int32[n], may be NULL) gathers
 * frames obs[idx[i]]. logits / value may be NULL. */
int cb_policy_value(cb_ctx* ctx, const uint8_t* obs, const int32_t* idx, int n, float* logits, float* value, cb_stream stream);

/* ---- learner pieces ---------------------------------------------------------------------------------------- */
/* compute_gae (cleanba_ppo.py:532-560) fused with the per-column-group advantage normalisation (:592-595).
 * rewards, values: float [T,B]; dones: uint8 [T,B]; next_value float[B]; next_done uint8[B].
 * num_groups = num_minibatches (0 disables the normalisation).  adv, ret: float [T,B]. */
int cb_gae(cb_ctx* ctx, const float* rewards, const float* values, const uint8_t* dones, const float* next_value,
           const uint8_t* next_done, int T, int B, float gamma, float gae_lambda, int num_groups, float* adv, float* ret,
           cb_stream stream);
/* key, subkey = jax.random.split(key)  (cleanba_ppo.py:599); both uint32[2] on the device. */
int cb_split_key(cb_ctx* ctx, uint32_t* key, uint32_t* subkey, cb_stream stream);
/* out = jax.random.permutation(key, n) (cleanba_ppo.py:606); key uint32[2] on the device (not modified). */
int cb_permutation(cb_ctx* ctx, const uint32_t* key, int n, int32_t* out, cb_stream stream);
/* value_and_grad(ppo_loss) on one minibatch (cleanba_ppo.py:562-577,590,619-627).
 * obs: uint8 [N,4,84,84] (the whole update, flattened t*B+b); idx: int32[mb] rows of this minibatch (NULL = first mb
 * rows); actions int32[N]; logprobs, advantages, returns float[N] (all indexed through idx).
 * grads: float[num_params] out.  stats: float[5] out = loss, pg_loss, v_loss, entropy, approx_kl. */
int cb_ppo_grad(cb_ctx* ctx, const uint8_t* obs, const int32_t* idx, int mb, const int32_t* actions, const float* logprobs,
                const float* advantages, const float* returns, float clip_coef, float ent_coef, float vf_coef, float* grads,
                float* stats, cb_stream stream);
/* value_and_grad(impala_loss) on one minibatch of env columns (cleanba_impala.py:569-597,606-618).
 * Fields are the whole shard flattened [T1*Bl] (row t*Bl + col); idx: int32[T1*B] with idx[t*B+b] = t*Bl + col_b selects the
 * minibatch columns (NULL = identity, Bl == B).  obs uint8 [T1*Bl,4,84,84]; behaviour_logits float [T1*Bl,A];
 * dones / firststeps uint8.  stats: float[4] = total, pg_loss, baseline_loss, entropy_loss (sums). */
int cb_impala_grad(cb_ctx* ctx, const uint8_t* obs, const int32_t* idx, int T1, int B, const int32_t* actions,
                   const float* behaviour_logits, const float* rewards, const uint8_t* dones, const uint8_t* firststeps,
                   float gamma, float vf_coef, float ent_coef, float* grads, float* stats, cb_stream stream);
/* clip_by_global_norm + Adam / RMSProp on the flat vectors (cleanba_ppo.py:492-500,629; cleanba_impala.py:152-188).
 * grads are multiplied by grad_scale first (1/L turns an allreduce-sum into the pmean).  norm_out (device float[1],
 * may be NULL) receives the pre-clip global norm.  Refreshes the packed bf16 weights. */
int cb_optimizer_step(cb_ctx* ctx, const float* grads, float grad_scale, float lr, float max_norm, float* norm_out,
                      cb_stream stream);
/* optax.MultiSteps(every_k_schedule = gradient_accumulation_steps) (cleanba_ppo.py:492-500): running mean of the gradients of the k
 * mini-steps that make up one optimizer step, acc <- acc + (grads - acc) / (mini_step + 1), mini_step = 0 .. k-1 (mini_step 0
 * overwrites acc).  The optimizer step is then taken on acc. */
int cb_grad_accumulate(cb_ctx* ctx, float* acc, const float* grads, int mini_step, cb_stream stream);
/* The same step with the gradient exchange FUSED in: the gradient is the fixed-order sum grads[0] + ... + grads[n-1] of the
 * flat gradient buffers of all learner replicas of this process, read directly from peer memory over NVLink / NVSwitch inside
 * the norm and the update kernels -- `jax.lax.pmean(grads, "local_devices")` (cleanba_ppo.py:628) without a separate
 * allreduce pass or a host rendezvous.  Every replica calls this with the SAME pointer list, so all replicas apply bit-identical
 * updates.  The caller orders the streams with events: every replica's backward must have finished before any replica's step
 * starts, and every step must have finished before a replica's next backward overwrites its buffer.  grads is a HOST array of
 * 1..8 device pointers; peer devices must have been opened with cb_enable_peer_access. */
int cb_optimizer_step_peers(cb_ctx* ctx, const float* const* grads, int num_grads, float grad_scale, float lr, float max_norm,
                            float* norm_out, cb_stream stream);
/* Overlap of the gradient exchange with the backward pass.  The flat gradient vector is ordered [conv stages | dense | actor |
 * critic] and the backward pass produces it back to front: elements [*tail_offset, num_params) -- the dense layer and the heads,
 * 91% of the bytes -- are final before the conv backward starts.  With a milestone set, every cb_ppo_grad / cb_impala_grad
 * records `cuda_event` (a cudaEvent_t; NULL clears it) on its stream at that point, so the caller can run the allreduce of the
 * tail on a side stream under the conv backward and only the small head of the vector after the call
 * (`jax.lax.pmean(grads)`, cleanba_ppo.py:628, split in two collectives; every element is still reduced exactly once). */
int cb_set_grad_milestone(cb_ctx* ctx, void* cuda_event, long long* tail_offset);
/* CUDA-graph replay of the gradient step.  A cb_ppo_grad / cb_impala_grad call is ~70 kernel launches on two streams; with
 * graph mode on, the launches of one (shape, `grads` buffer, loss coefficients, milestone) combination are captured into a CUDA
 * graph the second time the combination is seen and replayed from then on: per call the host enqueues one small launch that
 * publishes the call's obs / idx / field / stats pointers to a device-side table (the captured frame-unpack and loss-head
 * kernels read them from there) plus one cudaGraphLaunch.  Results are identical to the un-graphed call.  The reference's
 * equivalent is the jit/pmap-compiled `update_minibatch` executable (cleanba_ppo.py:621-633, 656-660).  Profiling
 * (cb_profile) temporarily falls back to plain launches.  cb_graph_replays counts the replays (tests, bench). */
int cb_graph_steps(cb_ctx* ctx, int enable);
/* Experimental forward path for small batches (n <= 128): ConvSequence 1 and 2 of Network.__call__ (cleanba_ppo.py:178-189) --
 * ten convolutions and two max-pools -- run as ONE persistent kernel with one thread-block cluster of `cluster_size` CTAs
 * (1 | 2) per frame instead of ten launches; 0 restores the per-layer launches.  Results are bit-identical either way.  Off by
 * default: on B200 the per-layer launches are faster at the rollout batch. */
int cb_set_actor_tail(cb_ctx* ctx, int cluster_size);
long long cb_graph_replays(cb_ctx* ctx);
/* out = grads[0] + ... + grads[n-1] (fixed order, read from peer memory): the in-process stage of the gradient exchange when
 * the learner group ALSO spans processes (`--distributed` with several learner devices per process, cleanba_ppo.py:419-423,628):
 * one replica sums its process' replicas into `out`, ONE NCCL allreduce on `out` follows, and every replica then applies
 * cb_optimizer_step_peers on that single buffer.  Same stream-ordering rules as cb_optimizer_step_peers. */
int cb_reduce_peers(cb_ctx* ctx, const float* const* grads, int num_grads, float* out, cb_stream stream);
/* Strided block copy: `rows` rows of `width_bytes` from src (row pitch src_pitch) to dst (row pitch dst_pitch) on `stream`,
 * between any two of host-pinned / device / peer-device memory.  Device -> (peer) device copies with 16-byte granularity are done
 * by a copy kernel on the source GPU (`stream` must belong to it; stores over NVLink), everything else by cudaMemcpy2DAsync.  The actor -> learner
 * payload hand-off (`jax.device_put_sharded`, cleanba_ppo.py:357-363) uses it to move one learner's env-column block of the
 * [T, N, ...] rollout storage straight to that learner's GPU on the actor's copy stream, without a contiguous temporary. */
int cb_memcpy_2d(void* dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t width_bytes, size_t rows, cb_stream stream);
/* Size the persistent grids of this context for num_sms SMs instead of the whole device: for contexts whose streams live on an
 * SM partition (a CUDA green context), e.g. actor replicas on a small partition running beside the learner (the reference runs
 * actor and learner threads concurrently on one GPU in the a0-l0 topology, cleanba_ppo.py:669-686). */
int cb_set_sm_budget(cb_ctx* ctx, int num_sms);
/* cudaDeviceEnablePeerAccess from ctx's device to peer_device (no-op if already enabled or the same device). */
int cb_enable_peer_access(cb_ctx* ctx, int peer_device);

/* ---- measurement ---------------------------------------------------------------------------------------------- */
/* Number of kernels this library has launched in this process (all contexts, all threads). */
long long cb_launch_count(void);
/* Per-kernel timing: while enabled, every launcher call is bracketed by CUDA events on the caller's stream.
 * cb_profile_report synchronises the device and writes a JSON array of {"name","calls","ms","flops","bytes"} (sums
 * since cb_profile(ctx, 1); flops / bytes are ALGORITHMIC: 2*MACs of the real, unpadded operator and the tensors it
 * must read and write once). */
int cb_profile(cb_ctx* ctx, int enable);
int cb_profile_report(cb_ctx* ctx, char* json, int cap);

/* ---- diagnostics (tests only) ------------------------------------------------------------------------------ */
/* Copy an internal tensor of the last forward/backward to the host as dense NHWC fp32 [n,H,W,C] (or [n,256] for
 * "hidden").  Names: "s{0,1,2}.{x,y,p,a0,b0,a1,out}", gradients "g{0,1,2}.{A,B,C,Bin}", "hidden", "dpre".  The pre-pool
 * conv outputs "s*.y" (and "g0.Bin") exist on the CB_CONV_SIMT cross-check backend only: the tcgen05 backend fuses every
 * sequence conv with its max-pool and never materialises them.
 * Returns the number of floats written (<= cap) or -1. Synchronises the device. */
long long cb_debug_tensor(cb_ctx* ctx, const char* name, float* host_out, long long cap);

#ifdef __cplusplus
}
#endif
#endif /* CLEANBA_B200_H_ */
