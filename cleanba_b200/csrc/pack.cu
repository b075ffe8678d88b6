// Packs the fp32 master conv kernels (HWIO, flax order) into the fp16 carrier images consumed by the tcgen05 conv kernels
// (conv_umma.cu): per K=16 step a K-major SWIZZLE_NONE B tile  [kc(2)][n2(2*Cout)][8]  where the N rows are the two carrier
// planes of the weights stacked: rows [0,Cout) = hi = fp16(w), rows [Cout,2Cout) = mid = fp16((w - hi) * 2^11), so that
// B(n2,k) sits at n2*16 + (k/8)*2*Cout*16 + (k%8)*2 bytes (LBO = 2*Cout*16, SBO = 128) and an MMA with N = 2*Cout or Cout
// uses the hi|mid or the hi planes of the same image (forward and dgrad images alike).
//   forward : step = tap * (Cin/16) + pair,  k -> ci = pair*16 + k,  B[n=co][k] = W[tap][ci][co]
//   dgrad   : conv of the output gradient with flipped taps and swapped channels:
//             step = tap' * (Cout/16) + pair, k -> co = pair*16 + k, B[n=ci][k] = W[8 - tap'][ci][co]
//   frames  : the 4(+4 zero)-channel first conv pairs two taps per step (see k_conv_umma):
//             steps 0..2 = taps (s,0)|(s,1), step 3 = taps (0,2)|(1,2), step 4 = tap (2,2)|zero.
#include "common.cuh"
#include "kernels.h"

namespace cb {

__global__ void k_pack_conv(const PackLayer* __restrict__ layers) {
    const PackLayer L = layers[blockIdx.y >> 1];
    const int variant = blockIdx.y & 1;   // 0 forward, 1 dgrad
    f16* dst = variant ? L.dg : L.fwd;
    if (!dst) return;
    const int kin = variant ? L.cout : L.cin;     // contraction channels
    const int nout = variant ? L.cin : L.cout;    // output channels of this conv
    const bool frames = (kin < 8);
    const int chunks = frames ? 1 : kin / 8;
    const int steps = frames ? 5 : 9 * (chunks / 2);
    const int n3tot = 2 * nout;                   // [W_hi | W_mid] rows
    const long long total = (long long)steps * 2 * n3tot * 8;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        int k8 = (int)(e % 8);
        int n3 = (int)((e / 8) % n3tot);
        int kc = (int)((e / (8LL * n3tot)) % 2);
        int step = (int)(e / (16LL * n3tot));
        int plane = n3 / nout, n = n3 % nout;
        float w = 0.f;
        if (frames) {
            int tap = -1;
            if (step < 3) tap = step * 3 + kc;
            else if (step == 3) tap = kc == 0 ? 2 : 5;
            else if (kc == 0) tap = 8;
            if (tap >= 0 && k8 < L.cin) w = L.w[((long long)tap * L.cin + k8) * L.cout + n];
        } else {
            int tap = step / (chunks / 2), pair = step % (chunks / 2);
            int kch = pair * 16 + kc * 8 + k8;
            if (!variant) w = L.w[((long long)tap * L.cin + kch) * L.cout + n];
            else w = L.w[((long long)(8 - tap) * L.cin + n) * L.cout + kch];
        }
        f16 h, m;
        split_f16(w, h, m);
        dst[e] = plane == 0 ? h : m;
    }
}

int launch_pack_conv(const PackLayer* layers_dev, int nlayers, cudaStream_t st) {
    dim3 grid(8, nlayers * 2);
    k_pack_conv<<<grid, 256, 0, st>>>(layers_dev);
    CB_LAUNCH_CHECK();
    return 0;
}

}  // namespace cb
