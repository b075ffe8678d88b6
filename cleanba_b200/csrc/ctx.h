// Context of libcleanba_b200 (shared by ctx.cu and nature.cu): model tables, buffers, profiling brackets.
#pragma once
#include <map>
#include <string>
#include <vector>

#include "../../include/cleanba_b200.h"
#include "common.cuh"
#include "kernels.h"

namespace cb {

struct Leaf {
    std::string name;
    long long offset;
    int ndim;
    int shape[4];
    long long size() const {
        long long s = 1;
        for (int i = 0; i < ndim; ++i) s *= shape[i];
        return s;
    }
};

struct ConvLayer {
    int cin, cout;          // real channels
    long long off_b, off_w; // offsets in the flat parameter vector
    f16 *fwd, *dg;          // packed [hi|mid] weight images (forward / dgrad)
};

struct Act {                // one activation / gradient tensor: fp16x2 carrier planes (common.cuh)
    Planes pl = {nullptr, nullptr, 0};
    int C = 0, H = 0;
};

struct Stage {
    Act x;                  // input of the sequence conv (stage 0: frames, hi only; else the previous stage's raw output)
    Act y;                  // conv output before the pool (only when the conv is not fused with its pool)
    Act p, pr;              // pooled: raw (residual input of block 0) and rectified (operand of its first conv)
    uint8_t* amax = nullptr; // arg-max slots of the pool (learner contexts)
    uint8_t *bits_pr = nullptr, *bits_a0 = nullptr, *bits_b0r = nullptr, *bits_a1 = nullptr;   // relu gate bits (learner, tcgen05)
    Act a0;                 // relu(conv1(relu(p)))
    Act b0, b0r;            // p + conv2(a0): raw and rectified
    Act a1;                 // relu(conv3(relu(b0)))
    Act out;                // b0 + conv4(a1): raw for stages 0 / 1 (the next ConvSequence is fed un-rectified), rectified for stage 2
    Act gA, gB, gC, gBin;   // gradients (learner contexts), all scaled by the minibatch's loss scale
};

struct NatureNet;           // nature.cu

}  // namespace cb

// bytes  = what the kernel moves in THIS library's storage formats (carrier planes, padded grids);
// abytes = SURVEY 8(d) algorithmic bytes: every operand tensor of the operator read / written once as unpadded fp32.
struct ProfAgg { long long launches = 0, records = 0; double ms = 0, flops = 0, bytes = 0, abytes = 0; };
struct ProfRec { std::string name; cudaEvent_t a, b; double flops, bytes, abytes; int launches; };

// One captured gradient step (cb_graph_steps): everything that is baked into the graph is in the key; the per-step pointers
// are read through cb_ctx::step_dev.
struct StepGraph {
    int kind, n, T1, B;                     // 0 = cb_ppo_grad, 1 = cb_impala_grad
    const float* grads;
    float c0, c1, c2;
    void* milestone;
    cudaGraphExec_t exec = nullptr;         // null: seen once (ran eagerly, warm), captured on the next call
    long long launches = 0;
};

struct cb_ctx {
    cb_config cfg;
    bool prof_on = false;
    std::vector<ProfRec> prof_recs;
    std::vector<cudaEvent_t> prof_pool;
    int num_sms = 148;
    int A = 18;
    std::vector<cb::Leaf> leaves;
    long long nparam = 0;
    std::vector<void*> allocs;
    float *params = nullptr, *m = nullptr, *v = nullptr;
    long long opt_count = 0;
    cb::ConvLayer conv[15];
    cb::PackLayer* pack_dev = nullptr;
    long long off_dense_b, off_dense_w, off_actor_b, off_actor_w, off_critic_b, off_critic_w;
    cb::Stage st[3];
    float *hidden = nullptr, *dense_part = nullptr, *dpre = nullptr, *dlogits = nullptr, *terms = nullptr;
    float *logits_scratch = nullptr, *cell_scratch = nullptr, *wg_partial = nullptr, *opt_partials = nullptr;
    uint32_t* subkey = nullptr;
    uint32_t* key_tmp = nullptr;
    int* perm_tmp = nullptr;
    int* perm_rank = nullptr;
    uint32_t* sort_keys = nullptr;
    int perm_cap = 0;
    long long wg_cap = 0;                   // floats in wg_partial
    int last_n = 0;
    int clean_n = -1;                       // batch size whose trailing zero row is known to be clear (shared-border layout)
    cb::TrailRow* trail_dev = nullptr;      // every plane of every activation / gradient tensor
    int trail_count = 0;
    // tcgen05 dense layer (dense_umma.cu)
    cb::bf16 *ft[3] = {nullptr, nullptr, nullptr}, *dpT[2] = {nullptr, nullptr}, *wd_fwd = nullptr, *wd_dx = nullptr;
    int npad_max = 0;
    cudaStream_t side = nullptr;            // weight-gradient kernels run here, beside the dgrad chain on the caller's stream
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaEvent_t milestone = nullptr;        // recorded once the dense + head gradients of a cb_*_grad call are complete
    float* gscale = nullptr;                // device {S, 1 / S}: loss scale of the current minibatch's gradient tensors
    unsigned* gs_work = nullptr;            // scratch of k_loss_scale
    const cb_rollout_cursor* cursor = nullptr;   // set for the duration of a cb_actor_step_cursor call
    cb::NatureNet* nat = nullptr;           // Nature-CNN trunk (cfg.model == CB_MODEL_NATURE)
    int HID = 256;                          // width of the trunk's dense output (256 IMPALA-ResNet, 512 Nature-CNN)
    int actor_tail = 0;                     // cb_set_actor_tail: cluster size of the persistent ConvSequence 1+2 kernel (0 = per-layer launches)
    bool graph_on = false;                  // cb_graph_steps: cb_*_grad replays a captured CUDA graph
    bool capturing = false;
    const cb::StepPtrs* ind = nullptr;      // non-null while a gradient step is being captured (= step_dev)
    cb::StepPtrs* step_dev = nullptr;
    cudaStream_t cap = nullptr;             // capture stream of the graphed steps
    std::vector<StepGraph> graphs;
    long long graph_replays = 0;
    bool fuse0 = false;
    bool fuse12 = false;                    // second / third ConvSequence: conv + pool (forward) fused                     // first ConvSequence: conv + pool (forward) and pool + wgrad (backward) fused
};


namespace cb {

int dev_alloc(cb_ctx* c, void** p, size_t bytes, bool zero = true);

// CUDA-event bracket around one launcher call (only when profiling is enabled): per-kernel device time measured on the
// launching stream, with the kernel's algorithmic flops / bytes, for bench.py's roofline.
struct ProfScope {
    cb_ctx* c; cudaStream_t st; ProfRec r; bool on; long long l0;
    ProfScope(cb_ctx* c_, const std::string& name, double flops, double bytes, cudaStream_t st_, double abytes = -1.0)
        : c(c_), st(st_), on(c_->prof_on) {
        if (!on) return;
        r.abytes = abytes >= 0 ? abytes : bytes;
        auto get = [&]() { cudaEvent_t e; if (!c->prof_pool.empty()) { e = c->prof_pool.back(); c->prof_pool.pop_back(); } else cudaEventCreate(&e); return e; };
        r.name = name; r.flops = flops; r.bytes = bytes; r.a = get(); r.b = get();
        l0 = g_launches.load();
        cudaEventRecord(r.a, st);
    }
    ~ProfScope() {
        if (!on) return;
        cudaEventRecord(r.b, st);
        r.launches = (int)(g_launches.load() - l0);
        c->prof_recs.push_back(r);
    }
};
inline double f32_bytes(double elems) { return 4.0 * elems; }

// nature.cu: the Nature-CNN trunk behind the same context (cb_config.model == CB_MODEL_NATURE)
std::vector<Leaf> nature_leaves(int A);
int nature_create(cb_ctx* c);
void nature_destroy(cb_ctx* c);
int nature_refresh_weights(cb_ctx* c, cudaStream_t st);
int nature_forward(cb_ctx* c, const uint8_t* obs, const int* idx, int n, cudaStream_t st);   // -> c->hidden [n, 512]
int nature_backward(cb_ctx* c, int n, float* grads, cudaStream_t st);                        // from c->dpre [n, 512]

}  // namespace cb
