// Shared definitions for libcleanba_b200 (sm_100a only).
//
// Activation layout used by every trunk kernel ("chunk planes"):
//   A tensor with C channels (C % 8 == 0) over a batch of n images of H x W pixels is stored on a zero
//   padded grid Hp = H + 2, Wp = W + 2 (the SAME-padding ring of the 3x3 convs, cleanba_ppo.py:156,167),
//   flattened over (image, y', x') into NP = n * Hp * Wp "flat pixels", and split into C/8 planes of
//   8 channels:   plane[c / 8][flat pixel][c % 8].
//   * "planes"  = three bf16 arrays (hi, mid, lo) with x == hi + mid + lo to 24 significant bits (an exact
//     3-way split of the fp32 value); each plane has GUARD zero pixels before and after so a 3x3 tap window
//     of a 128-pixel tile is one contiguous, in-bounds byte range (a 1-D bulk TMA copy) and the tile itself
//     is an UMMA operand.
//   * "stream"  = one fp32 array [C/8][NP][8] (residual stream / pre-pool conv output / gradients).
//   Border pixels of every tensor are kept at exactly zero by the producing kernel.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace cb {

constexpr int GUARD = 256;          // zero pixels before / after every bf16 plane
constexpr int NUM_TAPS = 9;
constexpr int HIDDEN = 256;
constexpr int MAX_ACTIONS = 32;

typedef __nv_bfloat16 bf16;

struct Planes {           // bf16 hi/mid/lo chunk planes; pointers address flat pixel 0 (guard lies before it)
    bf16* hi;
    bf16* mid;            // 1 plane (hi only: exact-in-bf16 data, the unpacked uint8 frames), 2 planes (hi, mid: 16
    bf16* lo;             // significant bits, gradient tensors) or 3 planes (hi, mid, lo: 24 bits, forward activations)
    long long plane_px;   // pixels per plane INCLUDING both guards (plane stride = plane_px * 8 elements)
};

struct ConvGeom {
    int n, H, W;          // images, unpadded height / width
    int Hp, Wp, P;        // padded dims, P = Hp * Wp
    long long NP;         // n * P flat pixels
};

__host__ __device__ inline ConvGeom make_geom(int n, int H, int W) {
    ConvGeom g;
    g.n = n; g.H = H; g.W = W; g.Hp = H + 2; g.Wp = W + 2; g.P = g.Hp * g.Wp; g.NP = (long long)n * g.P;
    return g;
}

// Epilogue shared by the SIMT and the tcgen05 conv kernels (forward conv and dgrad):
//   v = acc * acc_scale + bias;  v *= (mask_hi > 0);  v += res;  border -> 0
//   out_s <- v ;  out planes <- split_bf16x3(relu ? max(v, 0) : v)
struct ConvEpilogue {
    const float* bias;        // [Cout] or null
    float acc_scale;          // 1/255 for the first conv (cleanba_ppo.py:181), else 1
    const bf16* mask_hi;      // planes (hi) of the forward activation whose sign gates the gradient, or null
    long long mask_plane_px;
    const float* res;         // fp32 stream added after the mask, or null
    float* out_s;             // fp32 stream out, or null
    Planes out;               // bf16 planes out (hi may be null)
    int relu;
    // optional second copy of the (relu'd) output in the sample-minor layout of the tcgen05 dense layer
    // (dense_umma.cu): featT[plane][chunk][pixel (H*W, no padding ring)][sample (ft_npad)][8]
    bf16 *ft_hi, *ft_mid, *ft_lo;
    int ft_npad, ft_pixpad;
};

struct ConvArgs {
    ConvGeom g;
    Planes in;                // input planes
    int cin_chunks;           // input channel chunks (8 channels each)
    int cin_real;             // real input channels (4 for the frame stack, else cin_chunks * 8)
    int cout;                 // output channels (16 / 32)
    const float* w;           // fp32 master kernel, HWIO [3][3][Cin_f][Cout_f] (SIMT path)
    int w_cin, w_cout;        // Cin_f, Cout_f of the master kernel
    int transpose;            // 0: forward conv; 1: dgrad (flipped taps, in/out channels swapped)
    const bf16* wp;           // packed bf16 UMMA weight image ([hi|mid|lo] stacked along N), see pack.cu
    ConvEpilogue ep;
};

// exact 3-way split: x == hi + mid + lo up to 24 significant bits (each residual is exact in fp32)
__device__ __forceinline__ void split_bf16(float x, bf16& hi, bf16& mid, bf16& lo) {
    hi = __float2bfloat16_rn(x);
    float r1 = x - __bfloat162float(hi);
    mid = __float2bfloat16_rn(r1);
    lo = __float2bfloat16_rn(r1 - __bfloat162float(mid));
}

__device__ __forceinline__ uint32_t pack_bf16x2(bf16 a, bf16 b) {
    return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}

__device__ __forceinline__ float bf16lo_to_f(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16hi_to_f(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

// 8 bf16 (one uint4) -> 8 floats
__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
    f[0] = bf16lo_to_f(v.x); f[1] = bf16hi_to_f(v.x);
    f[2] = bf16lo_to_f(v.y); f[3] = bf16hi_to_f(v.y);
    f[4] = bf16lo_to_f(v.z); f[5] = bf16hi_to_f(v.z);
    f[6] = bf16lo_to_f(v.w); f[7] = bf16hi_to_f(v.w);
}

// 8 consecutive channels of one pixel-chunk (element offset `off`) <- -> fp32
__device__ __forceinline__ void load_planes8(const Planes& p, long long off, float* f) {
    unpack8(*reinterpret_cast<const uint4*>(p.hi + off), f);
    if (p.mid) {
        float t[8];
        unpack8(*reinterpret_cast<const uint4*>(p.mid + off), t);
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] += t[e];
        if (p.lo) {
            unpack8(*reinterpret_cast<const uint4*>(p.lo + off), t);
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] += t[e];
        }
    }
}
__device__ __forceinline__ void store_planes8(const Planes& p, long long off, const float* v) {
    bf16 h[8], m[8], l[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) split_bf16(v[e], h[e], m[e], l[e]);
    *reinterpret_cast<uint4*>(p.hi + off) =
        make_uint4(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]), pack_bf16x2(h[4], h[5]), pack_bf16x2(h[6], h[7]));
    if (p.mid)
        *reinterpret_cast<uint4*>(p.mid + off) =
            make_uint4(pack_bf16x2(m[0], m[1]), pack_bf16x2(m[2], m[3]), pack_bf16x2(m[4], m[5]), pack_bf16x2(m[6], m[7]));
    if (p.lo)
        *reinterpret_cast<uint4*>(p.lo + off) =
            make_uint4(pack_bf16x2(l[0], l[1]), pack_bf16x2(l[2], l[3]), pack_bf16x2(l[4], l[5]), pack_bf16x2(l[6], l[7]));
}

// Is flat pixel q (q < NP) an interior (non-border) pixel of its image?
__device__ __forceinline__ bool interior(const ConvGeom& g, long long q) {
    int r = (int)(q % g.P);
    int y = r / g.Wp, x = r - y * g.Wp;
    return (y >= 1) & (y <= g.H) & (x >= 1) & (x <= g.W);
}

__device__ __forceinline__ void store_featT(bf16* ft_hi, bf16* ft_mid, bf16* ft_lo, int ft_npad, int ft_pixpad, const ConvGeom& g,
                                            long long q, int oc, const float* v) {
    const int b = (int)(q / g.P), r = (int)(q % g.P);
    const int y = r / g.Wp - 1, x = r % g.Wp - 1;
    const long long off = (((long long)oc * ft_pixpad + (y * g.W + x)) * ft_npad + b) * 8;
    float x8[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) x8[e] = fmaxf(v[e], 0.f);
    Planes p;
    p.hi = ft_hi; p.mid = ft_mid; p.lo = ft_lo; p.plane_px = 0;
    store_planes8(p, off, x8);
}

// Apply the epilogue to the 8 accumulators of (flat pixel q, output chunk oc) and store.
__device__ __forceinline__ void conv_epilogue_store(const ConvEpilogue& ep, const ConvGeom& g, long long q, int oc,
                                                    float* acc) {
    const bool tail = q >= g.NP;   // pixels past the last image (rest of the last 128-tile): planes get zeros
    const bool in = !tail && interior(g, q);
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = 0.f;
    if (in) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            float b = ep.bias ? ep.bias[oc * 8 + e] : 0.f;
            v[e] = acc[e] * ep.acc_scale + b;
        }
        if (ep.mask_hi) {
            uint4 m = *reinterpret_cast<const uint4*>(ep.mask_hi + ((long long)oc * ep.mask_plane_px + q) * 8);
            float mf[8];
            unpack8(m, mf);
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = mf[e] > 0.f ? v[e] : 0.f;
        }
        if (ep.res) {
            const float4* r = reinterpret_cast<const float4*>(ep.res + ((long long)oc * g.NP + q) * 8);
            float4 r0 = r[0], r1 = r[1];
            v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w;
            v[4] += r1.x; v[5] += r1.y; v[6] += r1.z; v[7] += r1.w;
        }
    }
    if (ep.out_s && !tail) {
        float4* o = reinterpret_cast<float4*>(ep.out_s + ((long long)oc * g.NP + q) * 8);
        o[0] = make_float4(v[0], v[1], v[2], v[3]);
        o[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
    if (ep.out.hi) {
        float x[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) x[e] = ep.relu ? fmaxf(v[e], 0.f) : v[e];
        store_planes8(ep.out, ((long long)oc * ep.out.plane_px + q) * 8, x);
    }
    if (ep.ft_hi && in) store_featT(ep.ft_hi, ep.ft_mid, ep.ft_lo, ep.ft_npad, ep.ft_pixpad, g, q, oc, v);
}

// Split epilogue for the tcgen05 kernels: the residual / gate operands of a tile are fetched into registers BEFORE the
// accumulator is waited for, so their global-memory latency overlaps the MMAs instead of following them.
template <int COUT>
struct EpiPrefetch {
    float res[COUT];
    uint4 mask[COUT / 8];
    bool in, tail;
};

template <int COUT>
__device__ __forceinline__ void epi_prefetch(const ConvEpilogue& ep, const ConvGeom& g, long long q, EpiPrefetch<COUT>& p) {
    p.tail = q >= g.NP;
    p.in = !p.tail && interior(g, q);
    if (!p.in) return;
#pragma unroll
    for (int oc = 0; oc < COUT / 8; ++oc) {
        if (ep.mask_hi) p.mask[oc] = *reinterpret_cast<const uint4*>(ep.mask_hi + ((long long)oc * ep.mask_plane_px + q) * 8);
        if (ep.res) {
            const float4* r = reinterpret_cast<const float4*>(ep.res + ((long long)oc * g.NP + q) * 8);
            float4 r0 = r[0], r1 = r[1];
            p.res[oc * 8 + 0] = r0.x; p.res[oc * 8 + 1] = r0.y; p.res[oc * 8 + 2] = r0.z; p.res[oc * 8 + 3] = r0.w;
            p.res[oc * 8 + 4] = r1.x; p.res[oc * 8 + 5] = r1.y; p.res[oc * 8 + 6] = r1.z; p.res[oc * 8 + 7] = r1.w;
        }
    }
}

template <int COUT>
__device__ __forceinline__ void epi_finish(const ConvEpilogue& ep, const ConvGeom& g, long long q, const float* acc,
                                           const EpiPrefetch<COUT>& p) {
#pragma unroll
    for (int oc = 0; oc < COUT / 8; ++oc) {
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = 0.f;
        if (p.in) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                float b = ep.bias ? ep.bias[oc * 8 + e] : 0.f;
                v[e] = acc[oc * 8 + e] * ep.acc_scale + b;
            }
            if (ep.mask_hi) {
                float mf[8];
                unpack8(p.mask[oc], mf);
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = mf[e] > 0.f ? v[e] : 0.f;
            }
            if (ep.res) {
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] += p.res[oc * 8 + e];
            }
        }
        if (ep.out_s && !p.tail) {
            float4* o = reinterpret_cast<float4*>(ep.out_s + ((long long)oc * g.NP + q) * 8);
            o[0] = make_float4(v[0], v[1], v[2], v[3]);
            o[1] = make_float4(v[4], v[5], v[6], v[7]);
        }
        if (ep.out.hi) {
            float x[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) x[e] = ep.relu ? fmaxf(v[e], 0.f) : v[e];
            store_planes8(ep.out, ((long long)oc * ep.out.plane_px + q) * 8, x);
        }
        if (ep.ft_hi && p.in) store_featT(ep.ft_hi, ep.ft_mid, ep.ft_lo, ep.ft_npad, ep.ft_pixpad, g, q, oc, v);
    }
}

// ----------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL).  Every kernel of the forward / backward chain starts with griddep_launch() (its
// successor in the stream may be scheduled as soon as all CTAs of this grid are running) and calls griddep_wait() before
// its first access to memory produced by earlier kernels; whatever precedes the wait (barrier init, TMEM allocation, the
// TMA load of the packed weights) overlaps the predecessor's tail.  Rules that keep this correct:
//   * a kernel launched through launch_pdl() MUST call griddep_wait() before touching dependent memory;
//   * the only global data read before the wait are the packed weights, and their writers (k_pack_conv / k_pack_dense,
//     cudaMemcpy) are launched normally and never trigger early, so they are complete before any successor starts;
//   * kernels launched with <<<>>> stay fully stream-ordered (they are barriers of the chain).
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ----------------------------------------------------------------------------------------------
// error plumbing (host)
void set_error(const char* fmt, ...);
}  // namespace cb
#include <atomic>
namespace cb {
extern std::atomic<long long> g_launches;
#define CB_CUDA(expr)                                                                                  \
    do {                                                                                               \
        cudaError_t _e = (expr);                                                                       \
        if (_e != cudaSuccess) {                                                                       \
            (void)cudaGetLastError(); /* clear the non-sticky error state */                           \
            cb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));       \
            return -1;                                                                                 \
        }                                                                                              \
    } while (0)
#define CB_CHECK(cond, ...)                                                                            \
    do {                                                                                               \
        if (!(cond)) {                                                                                 \
            cb::set_error(__VA_ARGS__);                                                                \
            return -1;                                                                                 \
        }                                                                                              \
    } while (0)
#define CB_LAUNCH_CHECK()                 \
    do {                                  \
        cb::g_launches.fetch_add(1);      \
        CB_CUDA(cudaGetLastError());      \
    } while (0)

bool pdl_enabled();   // ctx.cu: CLEANBA_PDL != "0"

template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    (void)cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);   // errors are picked up by CB_LAUNCH_CHECK
}

}  // namespace cb
