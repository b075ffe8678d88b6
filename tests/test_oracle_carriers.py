"""Accuracy guarantees of the operand carriers (oracle/carriers.py): what the tensor-core kernels may assume."""
import numpy as np

from oracle import carriers as C


def _samples(rng, n):
    mant = rng.uniform(1.0, 2.0, n)
    expo = rng.integers(-20, 12, n)
    sign = rng.choice([-1.0, 1.0], n)
    x = (sign * mant * 2.0 ** expo).astype(np.float32)
    special = np.array([0.0, 1.0, -1.0, 255.0, 1.0 + 2 ** -23, 1.0 - 2 ** -24, 3.0 * 2 ** -20, 0.1, 1e-3, 65504.0], np.float32)
    return np.concatenate([x, special])


def test_bf16x3_is_an_exact_split_of_fp32():
    """x == hi + mid + lo bit for bit: three 8-bit planes with nearest rounding cover the 24-bit significand, every residual is
    exact in fp32 -- the claim of csrc/common.cuh::split_bf16 that the forward kernels rest on."""
    x = _samples(np.random.default_rng(0), 200_000)
    hi, mid, lo = C.split_bf16(x, 3)
    assert np.array_equal(C.join_bf16([hi, mid, lo]), x)
    assert np.array_equal((hi.astype(np.float64) + mid + lo).astype(np.float32), x)
    # planes shrink by >= 2^8 each: the dropped cross products (mid*lo, lo*lo, ...) are below 2^-24 relative
    nz = x != 0
    assert (np.abs(mid[nz]) <= np.abs(x[nz]) * 2.0 ** -8).all() and (np.abs(lo[nz]) <= np.abs(x[nz]) * 2.0 ** -16).all()


def test_bf16x2_carries_16_bits():
    x = _samples(np.random.default_rng(1), 200_000)
    hi, mid = C.split_bf16(x, 2)
    nz = x != 0
    rel = np.abs((hi.astype(np.float64) + mid)[nz] - x[nz]) / np.abs(x[nz])
    assert rel.max() <= 2.0 ** -17 and rel.max() > 2.0 ** -19     # 16 significant bits, and it really is that coarse


def test_frames_are_exact_in_one_plane():
    f = np.arange(256, dtype=np.float32)
    assert np.array_equal(C.split_bf16(f, 1)[0], f)               # uint8 frames: one bf16 (or fp16) plane, no residual
    hi, mid = C.split_f16x2(f)
    assert np.array_equal(hi, f) and not mid.any()


def test_f16x2_carries_22_bits_in_the_working_range():
    """|x| in [2^-14 * 2^11 ... 65504]: hi has 11 bits, the scaled residual another 11 -> relative error <= 2^-22; below that range the
    error is bounded ABSOLUTELY (fp16 subnormal spacing 2^-24 / 2^11), which is what matters for sums of products."""
    rng = np.random.default_rng(2)
    x = _samples(rng, 200_000)
    x = x[np.abs(x) <= 60000.0]
    hi, mid = C.split_f16x2(x)
    y = C.join_f16x2(hi, mid)
    big = np.abs(x) >= 2.0 ** -3
    rel = np.abs(y[big] - x[big]) / np.abs(x[big])
    assert rel.max() <= 2.0 ** -22
    err = np.abs(y - x.astype(np.float64))
    assert err[~big].max() <= 2.0 ** -25 * 1.0001                  # absolute: half the subnormal spacing of the scaled mid plane
    assert np.isfinite(hi).all() and np.isfinite(mid).all()


def test_f16x2_overflow_is_detectable():
    hi, mid = C.split_f16x2(np.array([7.0e4, -1.0e5], np.float32))
    assert np.isinf(hi).all()                                      # activations beyond fp16 range saturate to inf: kernels must flag it
