"""Parameter vector layout and seeded initialisation (host side).

The flat fp32 vector is in flax tree order (see include/cleanba_b200.h).  Initialisers follow the reference's choices
(cleanba/cleanba_ppo.py:156,187,195,203): lecun-normal 3x3 convs, orthogonal(sqrt 2) dense, orthogonal(0.01) actor,
orthogonal(1) critic, zero biases.  Flax's RNG folding cannot be reproduced without JAX, so the draw comes from numpy
PCG64(seed); parity tests feed the same vector to the CUDA path and to the CPU oracle."""
import numpy as np

from . import lib as _lib


def leaves(num_actions: int = 18, model: int = 0):
    return _lib.leaves(num_actions, model)


def _orthogonal(rng, rows, cols, scale):
    a = rng.standard_normal((max(rows, cols), min(rows, cols)))
    q, r = np.linalg.qr(a)
    q = q * np.sign(np.diag(r))
    if rows < cols:
        q = q.T
    return (scale * q[:rows, :cols]).astype(np.float32)


def init_params(seed: int = 1, num_actions: int = 18, model: int = 0) -> np.ndarray:
    """model = lib.CB_MODEL_IMPALA_RESNET | lib.CB_MODEL_NATURE_CNN.  The Nature-CNN convs are orthogonal(sqrt 2) like its dense
    layer (cleanba/legacy_scripts/cleanba_ppo_envpool_impala_atari_wrapper_naturecnn.py:152,160,168)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = []
    for name, _, shape in leaves(num_actions, model):
        if name.endswith("bias"):
            out.append(np.zeros(shape, np.float32))
        elif len(shape) == 4 and model == _lib.CB_MODEL_NATURE_CNN:
            out.append(_orthogonal(rng, shape[0] * shape[1] * shape[2], shape[3], np.sqrt(2.0)).reshape(shape))
        elif len(shape) == 4:
            fan_in = shape[0] * shape[1] * shape[2]
            std = np.sqrt(1.0 / fan_in) / 0.87962566103423978   # truncated-normal variance correction (lecun_normal)
            w = rng.standard_normal(shape)
            bad = np.abs(w) > 2
            while bad.any():
                w[bad] = rng.standard_normal(int(bad.sum()))
                bad = np.abs(w) > 2
            out.append((w * std).astype(np.float32))
        else:
            scale = {"network_params": np.sqrt(2.0), "actor_params": 0.01, "critic_params": 1.0}[name.split("/")[0]]
            out.append(_orthogonal(rng, shape[0], shape[1], scale))
    return np.concatenate([x.ravel() for x in out])
