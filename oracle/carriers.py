"""Operand carriers of the tcgen05 kernels, restated on the CPU (test infrastructure, like the rest of oracle/).

The tensor cores multiply 16-bit operands; every fp32 activation / weight / gradient is therefore carried as a sum of
low-precision planes and the significant cross products are accumulated in fp32:

  bf16 x 3 : x = hi + mid + lo, each plane bf16 (round to nearest even of the running residual)  -- forward activations, weights
  bf16 x 2 : x ~ hi + mid                                                                        -- gradient tensors
  fp16 x 2 : x ~ hi + mid' / 2^11 with hi = fp16(x), mid' = fp16((x - hi) * 2^11)                 -- the round-2 candidate

These functions define the splits bit for bit (the CUDA `split_bf16` in csrc/common.cuh performs the same fp32 operations), so GPU
tests can compare planes exactly, and `tests/test_oracle_carriers.py` states the accuracy each carrier guarantees."""
import numpy as np
import torch

F16_MID_SCALE = 2048.0


def _rn(x: np.ndarray, dtype) -> np.ndarray:
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(dtype).to(torch.float32).numpy()


def split_bf16(x: np.ndarray, planes: int = 3):
    """-> list of `planes` fp32 arrays holding bf16-representable values; residuals are computed in fp32 like the kernels do."""
    x = np.asarray(x, np.float32)
    out, r = [], x
    for _ in range(planes):
        p = _rn(r, torch.bfloat16)
        out.append(p)
        r = (r - p).astype(np.float32)
    return out


def join_bf16(planes) -> np.ndarray:
    """fp32 sum in the order the kernels use when they read planes back (hi + mid + lo)."""
    s = planes[0].astype(np.float32)
    for p in planes[1:]:
        s = (s + p).astype(np.float32)
    return s


def split_f16x2(x: np.ndarray):
    """-> (hi, mid') as fp32 arrays of fp16-representable values; x ~ hi + mid' / 2^11."""
    x = np.asarray(x, np.float32)
    hi = _rn(x, torch.float16)
    mid = _rn(((x - hi).astype(np.float32) * np.float32(F16_MID_SCALE)).astype(np.float32), torch.float16)
    return hi, mid


def join_f16x2(hi, mid) -> np.ndarray:
    return (hi.astype(np.float64) + mid.astype(np.float64) / F16_MID_SCALE).astype(np.float64)
