// JAX 0.4.8 threefry2x32 PRNG on the device (bit-exact restatement; see oracle/threefry.py for the KATs).
// Reference call sites: cleanba/cleanba_ppo.py:256-257 (split + uniform), :599,606 (split + permutation).
#pragma once
#include <stdint.h>

namespace cb {

__host__ __device__ __forceinline__ uint32_t rotl32(uint32_t x, int d) { return (x << d) | (x >> (32 - d)); }

// Threefry-2x32, 20 rounds (Random123), key (k0,k1), counter (x0,x1).
__host__ __device__ __forceinline__ void threefry2x32(uint32_t k0, uint32_t k1, uint32_t x0, uint32_t x1, uint32_t& y0,
                                                      uint32_t& y1) {
    const uint32_t ks0 = k0, ks1 = k1, ks2 = k0 ^ k1 ^ 0x1BD11BDAu;
    x0 += ks0; x1 += ks1;
#define CB_TF_R(r) { x0 += x1; x1 = rotl32(x1, r); x1 ^= x0; }
    CB_TF_R(13) CB_TF_R(15) CB_TF_R(26) CB_TF_R(6)
    x0 += ks1; x1 += ks2 + 1u;
    CB_TF_R(17) CB_TF_R(29) CB_TF_R(16) CB_TF_R(24)
    x0 += ks2; x1 += ks0 + 2u;
    CB_TF_R(13) CB_TF_R(15) CB_TF_R(26) CB_TF_R(6)
    x0 += ks0; x1 += ks1 + 3u;
    CB_TF_R(17) CB_TF_R(29) CB_TF_R(16) CB_TF_R(24)
    x0 += ks1; x1 += ks2 + 4u;
    CB_TF_R(13) CB_TF_R(15) CB_TF_R(26) CB_TF_R(6)
    x0 += ks2; x1 += ks0 + 5u;
#undef CB_TF_R
    y0 = x0; y1 = x1;
}

// Element e of jax `random_bits(key, 32, shape)` with `total` elements: threefry_2x32(key, iota(total)) splits the
// (zero padded to even) counter array into halves x0 = count[:h], x1 = count[h:] and concatenates the outputs.
__host__ __device__ __forceinline__ uint32_t jax_random_bits_elem(uint32_t k0, uint32_t k1, uint32_t e, uint32_t total) {
    const uint32_t h = (total + 1u) >> 1;
    uint32_t y0, y1;
    if (e < h) {
        uint32_t c1 = e + h;
        if (c1 >= total) c1 = 0u;  // the pad element
        threefry2x32(k0, k1, e, c1, y0, y1);
        return y0;
    }
    threefry2x32(k0, k1, e - h, e, y0, y1);
    return y1;
}

// jax.random.uniform float32: mantissa bits | 1.0, minus 1.0 -> [0, 1) on a 2^-23 grid.
__host__ __device__ __forceinline__ float jax_bits_to_uniform(uint32_t bits) {
    uint32_t w = (bits >> 9) | 0x3F800000u;
#ifdef __CUDA_ARCH__
    return __uint_as_float(w) - 1.0f;
#else
    union { uint32_t u; float f; } c; c.u = w; return c.f - 1.0f;
#endif
}

// key, subkey = jax.random.split(key): threefry_2x32(key, [0,1,2,3]).reshape(2,2)
__host__ __device__ __forceinline__ void jax_split2(uint32_t k0, uint32_t k1, uint32_t& nk0, uint32_t& nk1, uint32_t& sk0,
                                                    uint32_t& sk1) {
    uint32_t a0, a1, b0, b1;
    threefry2x32(k0, k1, 0u, 2u, a0, a1);   // outputs for counters (0,2): y0 -> elem 0, y1 -> elem 2
    threefry2x32(k0, k1, 1u, 3u, b0, b1);   // outputs for counters (1,3): y0 -> elem 1, y1 -> elem 3
    nk0 = a0; nk1 = b0; sk0 = a1; sk1 = b1;
}

}  // namespace cb
