// Shared definitions for libcleanba_b200 (sm_100a only).
//
// Activation layout used by every trunk kernel ("chunk planes"):
//   A tensor with C channels (C % 8 == 0) over a batch of n images of H x W pixels is stored on a zero
//   padded grid (the SAME-padding ring of the 3x3 convs, cleanba_ppo.py:156,167) with SHARED borders: every image has ONE
//   zero row above it and every row ONE zero pixel before it, Hp = H + 1, Wp = W + 1.  In the flattened order
//   (image, y', x') the pixel after the last pixel of a row is the zero pixel that starts the next row, and the row after
//   the last row of an image is the zero row that starts the next image (after the last image: a zero row the context
//   keeps clear, ctx.cu clear_trailing_rows), so every 3x3 neighbour offset (ky - 1) * Wp + (kx - 1) lands on the right
//   pixel or on a zero.  Interior pixel (y, x) is flat pixel  image * P + (y + 1) * Wp + (x + 1),  P = Hp * Wp,
//   NP = n * P "flat pixels", split into C/8 planes of 8 channels:   plane[c / 8][flat pixel][c % 8].
//   (A full ring per image, (H + 2)(W + 2), costs 4.7 % / 9.3 % / 17 % more pixels at 42 / 21 / 11 -- bytes AND MMAs.)
//   Every tensor is stored ONCE, as the fp16x2 CARRIER: two fp16 arrays
//       hi  = fp16(x),     mid = fp16((x - hi) * 2^11)        x == hi + mid * 2^-11 to 22 significant bits
//   (4 bytes per element, the size of the fp32 value it stands for).  Each plane has GUARD zero pixels before and after, so a
//   3x3 tap window of a 128-pixel tile is one contiguous, in-bounds byte range (a 1-D bulk TMA copy) and the tile itself is a
//   tcgen05 operand: no separate fp32 copy of any activation or gradient exists (round 1 kept three bf16 planes AND an fp32
//   stream, 10 bytes per element).  Residual adds and relu gates read the same planes.  Tensors whose consumers need both
//   the raw value (residual) and the rectified value (next conv's operand) are stored twice (raw planes + relu'd planes).
//   Border pixels of every tensor are kept at exactly zero by the producing kernel.
//   Gradient tensors use the same carrier, multiplied by a per-minibatch power-of-two LOSS SCALE (fp16 range): the scale is
//   chosen on the device from max |dL/d(pre-relu hidden)| and divided out by the weight-gradient reductions (ctx.cu).
//   The unpacked uint8 frames (0..255) are exact in fp16: one plane (mid == nullptr).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace cb {

constexpr int GUARD = 256;          // zero pixels before / after every plane
constexpr int NUM_TAPS = 9;
constexpr int HIDDEN = 256;
constexpr int MAX_ACTIONS = 32;
constexpr float MID_SCALE = 2048.f;             // 2^11: the mid plane holds the fp16 rounding residual times this
constexpr float MID_INV = 1.f / 2048.f;

typedef __nv_bfloat16 bf16;         // dense layer operands (dense_umma.cu keeps the 3-way bf16 split of round 1)
typedef __half f16;                 // carrier element of every trunk tensor

struct Planes {           // fp16x2 carrier planes; pointers address flat pixel 0 (the guard lies before it)
    f16* hi;
    f16* mid;             // nullptr: single-plane tensor (the unpacked frames, exact in fp16)
    long long plane_px;   // pixels per plane INCLUDING both guards (plane stride = plane_px * 8 elements)
};

struct ConvGeom {
    int n, H, W;          // images, unpadded height / width
    int Hp, Wp, P;        // padded dims, P = Hp * Wp
    long long NP;         // n * P flat pixels
};

__host__ __device__ inline ConvGeom make_geom(int n, int H, int W) {
    ConvGeom g;
    g.n = n; g.H = H; g.W = W; g.Hp = H + 1; g.Wp = W + 1; g.P = g.Hp * g.Wp; g.NP = (long long)n * g.P;
    return g;
}

// Epilogue shared by the SIMT and the tcgen05 conv kernels (forward conv and dgrad):
//   v = acc * acc_scale + bias;  v *= (mask_hi > 0);  v += res;  border -> 0
//   out <- carrier(v) ;  out_r <- carrier(max(v, 0))
struct ConvEpilogue {
    const float* bias;        // [Cout] or null
    float acc_scale;          // 1/255 for the first conv (cleanba_ppo.py:181), else 1
    const f16* mask_hi;       // hi plane of the (rectified) forward activation whose sign gates the gradient, or null
    long long mask_plane_px;
    // relu gates as BITS (tcgen05 path): one byte per (flat pixel, 8-channel chunk), bit e = (value of channel chunk*8+e > 0),
    // written by the producer of a rectified tensor (bits_out) and read by the dgrad that is gated by it (bits_in): Cout / 8
    // bytes per pixel instead of the 2-byte-per-element hi plane
    uint8_t* bits_out;
    const uint8_t* bits_in;
    Planes res;               // tensor added after the mask (residual input / residual gradient); res.hi null = none
    Planes out;               // raw output planes (hi may be null)
    Planes out_r;             // rectified output planes (hi may be null)
    // optional third copy of the rectified output in the sample-minor bf16 layout of the tcgen05 dense layer
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
    const f16* wp;            // packed fp16 UMMA weight image ([hi|mid] stacked along N), see pack.cu
    ConvEpilogue ep;
};

// ---------------------------------------------------------------------------------------------- fp16x2 carrier
__device__ __forceinline__ void split_f16(float x, f16& hi, f16& mid) {
    hi = __float2half_rn(x);
    mid = __float2half_rn((x - __half2float(hi)) * MID_SCALE);      // x - hi is exact in fp32
}
// two values -> packed (hi, hi) and (mid, mid) words
__device__ __forceinline__ void split_f16x2(float a, float b, uint32_t& hi2, uint32_t& mid2) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(h);
    const __half2 m = __floats2half2_rn((a - hf.x) * MID_SCALE, (b - hf.y) * MID_SCALE);
    hi2 = *reinterpret_cast<const uint32_t*>(&h);
    mid2 = *reinterpret_cast<const uint32_t*>(&m);
}
__device__ __forceinline__ float2 h2_to_f2(uint32_t w) { return __half22float2(*reinterpret_cast<const __half2*>(&w)); }
// 8 fp16 (one uint4) -> 8 floats
__device__ __forceinline__ void unpack8h(const uint4& v, float* f) {
    float2 t;
    t = h2_to_f2(v.x); f[0] = t.x; f[1] = t.y;
    t = h2_to_f2(v.y); f[2] = t.x; f[3] = t.y;
    t = h2_to_f2(v.z); f[4] = t.x; f[5] = t.y;
    t = h2_to_f2(v.w); f[6] = t.x; f[7] = t.y;
}
// relu gate from the hi plane of a rectified activation: bit pattern of a positive, non-zero fp16 (0x0001 .. 0x7fff)
__device__ __forceinline__ bool h_pos(uint32_t bits16) { return ((bits16 - 1u) & 0xffffu) < 0x7fffu; }
__device__ __forceinline__ void gate8h(const uint4& m, float* v) {
    const uint32_t w[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (!h_pos(w[i] & 0xffffu)) v[2 * i] = 0.f;
        if (!h_pos(w[i] >> 16)) v[2 * i + 1] = 0.f;
    }
}

// 8 consecutive channels of one pixel-chunk (element offset `off`) <- -> fp32
__device__ __forceinline__ void load_planes8(const Planes& p, long long off, float* f) {
    unpack8h(*reinterpret_cast<const uint4*>(p.hi + off), f);
    if (p.mid) {
        float t[8];
        unpack8h(*reinterpret_cast<const uint4*>(p.mid + off), t);
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = fmaf(t[e], MID_INV, f[e]);      // exact: both terms fit 24 bits
    }
}
__device__ __forceinline__ void store_planes8(const Planes& p, long long off, const float* v) {
    uint4 h, m;
    split_f16x2(v[0], v[1], h.x, m.x);
    split_f16x2(v[2], v[3], h.y, m.y);
    split_f16x2(v[4], v[5], h.z, m.z);
    split_f16x2(v[6], v[7], h.w, m.w);
    *reinterpret_cast<uint4*>(p.hi + off) = h;
    if (p.mid) *reinterpret_cast<uint4*>(p.mid + off) = m;
}
__device__ __forceinline__ void store_planes8_zero(const Planes& p, long long off) {
    *reinterpret_cast<uint4*>(p.hi + off) = make_uint4(0, 0, 0, 0);
    if (p.mid) *reinterpret_cast<uint4*>(p.mid + off) = make_uint4(0, 0, 0, 0);
}

// ---------------------------------------------------------------------------------------------- bf16 x3 (dense layer operands)
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
// 8 floats -> hi / mid / lo bf16 arrays at element offset off (mid / lo optional)
__device__ __forceinline__ void store_bf16_split8(bf16* hi, bf16* mid, bf16* lo, long long off, const float* v) {
    bf16 h[8], m[8], l[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) split_bf16(v[e], h[e], m[e], l[e]);
    *reinterpret_cast<uint4*>(hi + off) =
        make_uint4(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]), pack_bf16x2(h[4], h[5]), pack_bf16x2(h[6], h[7]));
    if (mid)
        *reinterpret_cast<uint4*>(mid + off) =
            make_uint4(pack_bf16x2(m[0], m[1]), pack_bf16x2(m[2], m[3]), pack_bf16x2(m[4], m[5]), pack_bf16x2(m[6], m[7]));
    if (lo)
        *reinterpret_cast<uint4*>(lo + off) =
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
    store_bf16_split8(ft_hi, ft_mid, ft_lo, off, x8);
}

// Stores of one (pixel, chunk) of 8 epilogue values: raw planes, rectified planes, sample-minor dense copy.
__device__ __forceinline__ void epi_store8(const ConvEpilogue& ep, const ConvGeom& g, long long q, int oc, const float* v, bool in) {
    if (ep.out.hi) store_planes8(ep.out, ((long long)oc * ep.out.plane_px + q) * 8, v);
    if (ep.out_r.hi) {
        float x[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) x[e] = fmaxf(v[e], 0.f);
        store_planes8(ep.out_r, ((long long)oc * ep.out_r.plane_px + q) * 8, x);
    }
    if (ep.ft_hi && in) store_featT(ep.ft_hi, ep.ft_mid, ep.ft_lo, ep.ft_npad, ep.ft_pixpad, g, q, oc, v);
}

// Apply the epilogue to the 8 accumulators of (flat pixel q, output chunk oc) and store.  Pixels past the last image (rest of
// the last 128-tile) and border pixels get zeros.
__device__ __forceinline__ void conv_epilogue_store(const ConvEpilogue& ep, const ConvGeom& g, long long q, int oc,
                                                    float* acc) {
    const bool tail = q >= g.NP;
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
        if (ep.mask_hi) gate8h(*reinterpret_cast<const uint4*>(ep.mask_hi + ((long long)oc * ep.mask_plane_px + q) * 8), v);
        if (ep.res.hi) {
            float r[8];
            load_planes8(ep.res, ((long long)oc * ep.res.plane_px + q) * 8, r);
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] += r[e];
        }
    }
    epi_store8(ep, g, q, oc, v, in);
}

// Split epilogue for the tcgen05 kernels: the residual / gate operands of a tile are fetched into registers BEFORE the
// accumulator is waited for, so their global-memory latency overlaps the MMAs instead of following them.
template <int COUT>
struct EpiPrefetch {
    uint4 res_hi[COUT / 8], res_mid[COUT / 8];
    uint4 mask[COUT / 8];
    uint32_t mbits;
    bool in, tail;
};

// `in` / `tail` known to the caller (the persistent tcgen05 kernels carry the pixel's position from tile to tile instead of
// dividing the flat index again: the divisions were 16 % of the epilogue's instructions, profiles/r02_v9_*)
template <int COUT>
__device__ __forceinline__ void epi_prefetch_known(const ConvEpilogue& ep, long long q, bool in, bool tail, EpiPrefetch<COUT>& p) {
    p.tail = tail;
    p.in = in;
    if (!p.in) return;
    if (ep.bits_in) {
        if (COUT == 16) p.mbits = *reinterpret_cast<const uint16_t*>(ep.bits_in + q * 2);
        else p.mbits = *reinterpret_cast<const uint32_t*>(ep.bits_in + q * (COUT / 8));
    }
#pragma unroll
    for (int oc = 0; oc < COUT / 8; ++oc) {
        if (ep.mask_hi && !ep.bits_in) p.mask[oc] = *reinterpret_cast<const uint4*>(ep.mask_hi + ((long long)oc * ep.mask_plane_px + q) * 8);
        if (ep.res.hi) {
            const long long off = ((long long)oc * ep.res.plane_px + q) * 8;
            p.res_hi[oc] = *reinterpret_cast<const uint4*>(ep.res.hi + off);
            p.res_mid[oc] = *reinterpret_cast<const uint4*>(ep.res.mid + off);
        }
    }
}

template <int COUT>
__device__ __forceinline__ void epi_prefetch(const ConvEpilogue& ep, const ConvGeom& g, long long q, EpiPrefetch<COUT>& p) {
    const bool tail = q >= g.NP;
    epi_prefetch_known<COUT>(ep, q, !tail && interior(g, q), tail, p);
}

// STAGED_BIAS: `bias` = COUT floats (zeros when the layer has none) staged once per CTA, e.g. in shared memory; else ep.bias
template <int COUT, bool STAGED_BIAS = false>
__device__ __forceinline__ void epi_finish(const ConvEpilogue& ep, const ConvGeom& g, long long q, const float* acc,
                                           const EpiPrefetch<COUT>& p, const float* bias = nullptr) {
    uint32_t obits = 0;
#pragma unroll
    for (int oc = 0; oc < COUT / 8; ++oc) {
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = 0.f;
        if (p.in) {
            if (STAGED_BIAS) {
                const float4 b0 = *reinterpret_cast<const float4*>(bias + oc * 8), b1 = *reinterpret_cast<const float4*>(bias + oc * 8 + 4);
                const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = acc[oc * 8 + e] * ep.acc_scale + b[e];
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    float b = ep.bias ? ep.bias[oc * 8 + e] : 0.f;
                    v[e] = acc[oc * 8 + e] * ep.acc_scale + b;
                }
            }
            if (ep.bits_in) {
#pragma unroll
                for (int e = 0; e < 8; ++e)
                    if (!((p.mbits >> (oc * 8 + e)) & 1u)) v[e] = 0.f;
            } else if (ep.mask_hi) {
                gate8h(p.mask[oc], v);
            }
            if (ep.res.hi) {
                float rh[8], rm[8];
                unpack8h(p.res_hi[oc], rh);
                unpack8h(p.res_mid[oc], rm);
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] += fmaf(rm[e], MID_INV, rh[e]);
            }
        }
        if (ep.bits_out) {
#pragma unroll
            for (int e = 0; e < 8; ++e) obits |= (v[e] > 0.f ? 1u : 0u) << (oc * 8 + e);
        }
        epi_store8(ep, g, q, oc, v, p.in);
    }
    if (ep.bits_out) {      // rows up to the 128-pixel tile boundary exist (zero bits for border / tail pixels)
        if (COUT == 16) *reinterpret_cast<uint16_t*>(ep.bits_out + q * 2) = (uint16_t)obits;
        else *reinterpret_cast<uint32_t*>(ep.bits_out + q * (COUT / 8)) = obits;
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
