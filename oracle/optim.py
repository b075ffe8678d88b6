"""Optimizer chains of the reference, restated (numpy fp32 elementwise, flat parameter vector).

PPO:    optax.MultiSteps(k=1) o clip_by_global_norm(0.5) o inject_hyperparams(adam)(lr schedule, eps=1e-5)
        cleanba/cleanba_ppo.py:475-479,492-500 (optax 0.1.4 semantics).
IMPALA: clip_by_global_norm(40) o rmsprop_pytorch_style(decay .99, eps .01) cleanba/cleanba_impala.py:152-188,
        515-519,532-540.
PARITY UNPINNED vs optax (not installable here); anchored by closed-form checks in tests/test_oracle_optim.py.
"""
import numpy as np

F32 = np.float32


def linear_schedule(count: int, base_lr: float, steps_per_update: int, num_updates: int, anneal: bool = True) -> np.float32:
    """cleanba_ppo.py:475-479 (steps_per_update = num_minibatches*update_epochs) and
    cleanba_impala.py:515-519 (steps_per_update = num_minibatches); evaluated at the pre-increment count."""
    if not anneal:
        return F32(base_lr)
    frac = 1.0 - (count // steps_per_update) / num_updates
    return F32(base_lr * frac)


def global_norm(g: np.ndarray) -> np.float32:
    return F32(np.sqrt(np.sum(np.square(g.astype(F32)), dtype=F32)))


def clip_by_global_norm(g: np.ndarray, max_norm: float) -> np.ndarray:
    """optax.clip_by_global_norm: g if norm < c else (g / norm) * c."""
    n = global_norm(g)
    if n < F32(max_norm):
        return g.astype(F32)
    return (g.astype(F32) / n) * F32(max_norm)


class Adam:
    """optax.adam(b1=.9, b2=.999, eps=1e-5, eps_root=0) with bias correction."""

    def __init__(self, n: int, b1=0.9, b2=0.999, eps=1e-5):
        self.m = np.zeros(n, F32)
        self.v = np.zeros(n, F32)
        self.count = 0
        self.b1, self.b2, self.eps = F32(b1), F32(b2), F32(eps)

    def step(self, p: np.ndarray, g: np.ndarray, lr: float) -> np.ndarray:
        g = g.astype(F32)
        self.m = self.b1 * self.m + (F32(1) - self.b1) * g
        self.v = self.b2 * self.v + (F32(1) - self.b2) * g * g
        self.count += 1
        t = self.count
        bc1 = F32(1) - F32(np.power(np.float32(self.b1), np.float32(t)))
        bc2 = F32(1) - F32(np.power(np.float32(self.b2), np.float32(t)))
        mhat = self.m / bc1
        vhat = self.v / bc2
        u = mhat / (np.sqrt(vhat) + self.eps)
        return (p.astype(F32) - F32(lr) * u).astype(F32)


class RMSPropPyTorchStyle:
    """cleanba_impala.py:152-188: nu = d*nu + (1-d)*g^2; u = g / (sqrt(nu) + eps); p -= lr*u."""

    def __init__(self, n: int, decay=0.99, eps=0.01):
        self.nu = np.zeros(n, F32)
        self.count = 0
        self.decay, self.eps = F32(decay), F32(eps)

    def step(self, p: np.ndarray, g: np.ndarray, lr: float) -> np.ndarray:
        g = g.astype(F32)
        self.nu = self.decay * self.nu + (F32(1) - self.decay) * g * g
        self.count += 1
        u = g / (np.sqrt(self.nu) + self.eps)
        return (p.astype(F32) - F32(lr) * u).astype(F32)
