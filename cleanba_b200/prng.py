"""Host-side JAX-compatible PRNG key derivation (threefry2x32, 20 rounds) used once at start-up to derive the actor /
learner keys exactly as the reference does: key = PRNGKey(seed); key, network_key, actor_key, critic_key = split(key, 4)
(cleanba/cleanba_ppo.py:468-469).  Per-step key splitting and sampling run on the device (csrc/prng.cuh)."""
import numpy as np

_M = 0xFFFFFFFF


def _rotl(x, d):
    return ((x << d) | (x >> (32 - d))) & _M


def _threefry2x32(k0, k1, x0, x1):
    ks = (k0, k1, k0 ^ k1 ^ 0x1BD11BDA)
    rot = ((13, 15, 26, 6), (17, 29, 16, 24))
    x0 = (x0 + ks[0]) & _M
    x1 = (x1 + ks[1]) & _M
    for i in range(5):
        for r in rot[i % 2]:
            x0 = (x0 + x1) & _M
            x1 = _rotl(x1, r) ^ x0
        x0 = (x0 + ks[(i + 1) % 3]) & _M
        x1 = (x1 + ks[(i + 2) % 3] + i + 1) & _M
    return x0, x1


def prng_key(seed: int) -> np.ndarray:
    return np.array([(seed >> 32) & _M, seed & _M], dtype=np.uint32)


def split(key, num: int = 2) -> np.ndarray:
    """jax.random.split: threefry_2x32(key, iota(2*num)).reshape(num, 2)."""
    k0, k1 = int(key[0]), int(key[1])
    n = 2 * num
    out = [0] * n
    for i in range(num):
        y0, y1 = _threefry2x32(k0, k1, i, i + num)
        out[i], out[i + num] = y0, y1
    return np.array(out, dtype=np.uint32).reshape(num, 2)


def first_key(seed: int) -> np.ndarray:
    """The key handed to every actor thread and learner device (cleanba_ppo.py:469-470,677)."""
    return split(prng_key(seed), 4)[0]
