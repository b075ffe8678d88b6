"""JAX 0.4.8 threefry2x32 PRNG restated in numpy (uint32, bit-exact).

Reference call sites: `jax.random.PRNGKey/split` cleanba/cleanba_ppo.py:468-469, `split` + `uniform`
in the actor cleanba/cleanba_ppo.py:256-257 (IMPALA cleanba/cleanba_impala.py:298-299), `split` +
`permutation` in the learner cleanba/cleanba_ppo.py:599,606.  The algorithm itself lives in
jax 0.4.8 (`jax/_src/prng.py`, `jax/_src/random.py`; poetry.lock:1084), which is not vendored: this is
a restatement of the published algorithm, pinned by Random123 KATs and JAX doc constants.
"""
import math

import numpy as np

_U32 = np.uint32
_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))


def _rotl(x, d):
    return (x << _U32(d)) | (x >> _U32(32 - d))


def threefry2x32_block(k0, k1, x0, x1):
    """Threefry-2x32, 20 rounds, on arrays of counter words (Random123 reference algorithm)."""
    k0 = _U32(k0)
    k1 = _U32(k1)
    x0 = np.asarray(x0, dtype=_U32).copy()
    x1 = np.asarray(x1, dtype=_U32).copy()
    ks = (k0, k1, _U32(k0 ^ k1 ^ _U32(0x1BD11BDA)))
    with np.errstate(over="ignore"):
        x0 += ks[0]
        x1 += ks[1]
        for i in range(5):
            for r in _ROT[i % 2]:
                x0 += x1
                x1 = _rotl(x1, r)
                x1 ^= x0
            x0 += ks[(i + 1) % 3]
            x1 += ks[(i + 2) % 3] + _U32(i + 1)
    return x0, x1


def threefry_2x32(key, count):
    """`jax._src.prng.threefry_2x32`: hash a flat uint32 counter array with `key` (odd sizes padded)."""
    count = np.asarray(count, dtype=_U32).ravel()
    n = count.size
    if n % 2:
        count = np.concatenate([count, np.zeros(1, _U32)])
    h = count.size // 2
    y0, y1 = threefry2x32_block(key[0], key[1], count[:h], count[h:])
    return np.concatenate([y0, y1])[:n]


def PRNGKey(seed):
    seed = int(seed)
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=_U32)


def split(key, num=2):
    """`jax.random.split` (cleanba_ppo.py:256,469,599)."""
    return threefry_2x32(key, np.arange(2 * num, dtype=_U32)).reshape(num, 2)


def random_bits(key, shape):
    size = int(np.prod(shape)) if len(shape) else 1
    return threefry_2x32(key, np.arange(size, dtype=_U32)).reshape(shape)


def uniform(key, shape=()):
    """`jax.random.uniform` float32 in [0,1) on a 2^-23 grid (cleanba_ppo.py:257)."""
    bits = random_bits(key, shape)
    f = ((bits >> _U32(9)) | _U32(0x3F800000)).view(np.float32) - np.float32(1.0)
    return f.reshape(shape)


def permutation_rounds(n):
    return int(math.ceil(3 * math.log(max(1, n)) / math.log(np.iinfo(np.uint32).max)))


def permutation(key, n):
    """`jax.random.permutation(key, n)` index vector (cleanba_ppo.py:606): rounds of stable sort by
    fresh 32-bit keys.  `permutation(key, x)` for an array x equals `x[permutation(key, len(x))]`."""
    x = np.arange(n, dtype=np.int32)
    key = np.asarray(key, dtype=_U32)
    for _ in range(permutation_rounds(n)):
        key, sub = split(key)
        sort_keys = random_bits(sub, (n,))
        x = x[np.argsort(sort_keys, kind="stable")]
    return x
