"""CPU restatements of integer / float tricks the CUDA kernels rely on, checked exhaustively (no GPU needed)."""
import numpy as np


def test_epilogue_row_from_float_multiply_is_exact():
    """k_conv_umma's epilogue (csrc/conv_umma.cu) gets the padded row of a pixel as trunc((float(r) + 0.5f) * (1.0f / Wp)) instead
    of r / Wp; the launcher admits images of up to 2^20 padded pixels.  Exhaustive over the geometries in use and over random
    (Wp, r) pairs up to that bound, in float32 exactly as the kernel evaluates it."""
    f32 = np.float32
    for W in (11, 21, 42, 84):                      # shared-border grids: Wp = W + 1, P = Wp * Wp
        Wp = W + 1
        r = np.arange(Wp * Wp, dtype=np.int64)
        y = ((r.astype(f32) + f32(0.5)) * (f32(1.0) / f32(Wp))).astype(np.int64)
        assert np.array_equal(y, r // Wp), W
    rng = np.random.default_rng(0)
    for _ in range(200):
        Wp = int(rng.integers(2, 1024))
        Hp = int(rng.integers(2, (1 << 20) // Wp + 1))
        P = Wp * Hp
        assert P <= (1 << 20)
        r = np.unique(np.concatenate([rng.integers(0, P, 4096), np.arange(0, P, Wp), np.arange(Wp - 1, P, Wp)]))
        y = ((r.astype(f32) + f32(0.5)) * (f32(1.0) / f32(Wp))).astype(np.int64)
        assert np.array_equal(y, r // Wp), (Wp, Hp)


def test_epilogue_position_recurrence():
    """r = q mod P carried from tile to tile: r += stride mod P with one conditional subtraction (both operands are < P)."""
    P, stride, q0 = 43 * 43, 2 * 296 * 128, 5 * 128 + 77
    dr, r = stride % P, q0 % P
    for k in range(2000):
        assert r == (q0 + k * stride) % P
        r += dr
        if r >= P:
            r -= P
