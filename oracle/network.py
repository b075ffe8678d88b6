"""IMPALA-ResNet trunk + actor/critic heads (torch CPU restatement).

Follows cleanba/cleanba_ppo.py:149-203 (identical in cleanba/cleanba_impala.py:192-246) with the
flax 0.6.8 layer semantics: `nn.Conv` = 3x3 cross-correlation, stride 1, SAME zero
padding, HWIO kernels, bias; `nn.max_pool(3x3, stride 2, SAME)` with -inf padding and XLA's SAME split
(lo = total//2, hi = total-lo: (0,1),(0,1),(1,1) for 84->42->21->11); flatten in NHWC (h,w,c) order;
`nn.Dense` kernels are [in,out].

Parameters are held as a flat list of leaves in *flax tree order* (jax.tree_util.tree_leaves of
AgentParams(network_params, actor_params, critic_params): dataclass field order, dict keys sorted,
so `bias` precedes `kernel`).  PARITY UNPINNED vs JAX (see oracle/__init__.py).
"""
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

CHANNELS = (16, 32, 32)
HIDDEN = 256
NUM_ACTIONS = 18
IN_CH, IN_H, IN_W = 4, 84, 84


def same_pool_pads(n: int, window: int = 3, stride: int = 2) -> Tuple[int, int]:
    """XLA SAME padding split for reduce_window."""
    out = -(-n // stride)
    total = max((out - 1) * stride + window - n, 0)
    lo = total // 2
    return lo, total - lo


def param_spec(channels: Sequence[int] = CHANNELS, hidden: int = HIDDEN, num_actions: int = NUM_ACTIONS,
               in_ch: int = IN_CH, in_hw: int = IN_H) -> List[Tuple[str, Tuple[int, ...]]]:
    """(name, shape) of every leaf in flax tree order (cleanba_ppo.py:206-210,298)."""
    spec = []
    cin, hw = in_ch, in_hw
    for s, c in enumerate(channels):
        p = f"network_params/params/ConvSequence_{s}"
        spec.append((f"{p}/Conv_0/bias", (c,)))
        spec.append((f"{p}/Conv_0/kernel", (3, 3, cin, c)))
        for r in range(2):
            for k in range(2):
                spec.append((f"{p}/ResidualBlock_{r}/Conv_{k}/bias", (c,)))
                spec.append((f"{p}/ResidualBlock_{r}/Conv_{k}/kernel", (3, 3, c, c)))
        cin = c
        hw = -(-hw // 2)
    flat = hw * hw * cin
    spec.append(("network_params/params/Dense_0/bias", (hidden,)))
    spec.append(("network_params/params/Dense_0/kernel", (flat, hidden)))
    spec.append(("actor_params/params/Dense_0/bias", (num_actions,)))
    spec.append(("actor_params/params/Dense_0/kernel", (hidden, num_actions)))
    spec.append(("critic_params/params/Dense_0/bias", (1,)))
    spec.append(("critic_params/params/Dense_0/kernel", (hidden, 1)))
    return spec


def nature_param_spec(num_actions: int = NUM_ACTIONS) -> List[Tuple[str, Tuple[int, ...]]]:
    """Leaves of the Nature-CNN agent in flax tree order
    (cleanba/legacy_scripts/cleanba_ppo_envpool_impala_atari_wrapper_naturecnn.py:143-193: Conv_0..2, Dense_0, actor, critic)."""
    p = "network_params/params"
    return [(f"{p}/Conv_0/bias", (32,)), (f"{p}/Conv_0/kernel", (8, 8, 4, 32)),
            (f"{p}/Conv_1/bias", (64,)), (f"{p}/Conv_1/kernel", (4, 4, 32, 64)),
            (f"{p}/Conv_2/bias", (64,)), (f"{p}/Conv_2/kernel", (3, 3, 64, 64)),
            (f"{p}/Dense_0/bias", (512,)), (f"{p}/Dense_0/kernel", (3136, 512)),
            ("actor_params/params/Dense_0/bias", (num_actions,)), ("actor_params/params/Dense_0/kernel", (512, num_actions)),
            ("critic_params/params/Dense_0/bias", (1,)), ("critic_params/params/Dense_0/kernel", (512, 1))]


def spec_for_size(n: int):
    """The two trunks have different parameter counts: pick the spec a flat vector belongs to."""
    for spec in (param_spec(), nature_param_spec()):
        if num_params(spec) == n:
            return spec
    raise ValueError(f"no known model has {n} parameters")


def num_params(spec=None) -> int:
    spec = spec or param_spec()
    return int(sum(int(np.prod(s)) for _, s in spec))


def _orthogonal(rng: np.random.Generator, rows: int, cols: int, scale: float) -> np.ndarray:
    a = rng.standard_normal((max(rows, cols), min(rows, cols)))
    q, r = np.linalg.qr(a)
    q = q * np.sign(np.diag(r))
    if rows < cols:
        q = q.T
    return (scale * q[:rows, :cols]).astype(np.float32)


def init_params(seed: int = 1, spec=None) -> np.ndarray:
    """Seeded initial parameters as ONE flat fp32 vector in flax tree order.

    Same distributions as the reference initialisers (cleanba_ppo.py:156,187,195,203: lecun-normal convs,
    orthogonal(sqrt2) dense, orthogonal(0.01) actor, orthogonal(1) critic, zero biases) drawn from
    numpy PCG64(seed).  Flax's own RNG folding cannot be reproduced without JAX, so parameters are an
    explicit seeded INPUT shared by the oracle and the CUDA path (SURVEY.md section 7 hard part 6).
    """
    spec = spec or param_spec()
    nature = any(name.endswith("/Conv_2/kernel") and "ConvSequence" not in name for name, _ in spec)
    rng = np.random.Generator(np.random.PCG64(seed))
    leaves = []
    for name, shape in spec:
        if name.endswith("bias"):
            leaves.append(np.zeros(shape, np.float32))
        elif len(shape) == 4 and nature:
            # Nature-CNN convs: orthogonal(sqrt 2) on the [kh*kw*cin, cout] matrix (naturecnn.py:152,160,168)
            leaves.append(_orthogonal(rng, shape[0] * shape[1] * shape[2], shape[3], np.sqrt(2.0)).reshape(shape))
        elif len(shape) == 4:
            fan_in = shape[0] * shape[1] * shape[2]
            # lecun_normal = truncated normal(+-2 sigma) with variance 1/fan_in
            std = np.sqrt(1.0 / fan_in) / 0.87962566103423978
            w = rng.standard_normal(shape)
            bad = np.abs(w) > 2
            while bad.any():
                w[bad] = rng.standard_normal(int(bad.sum()))
                bad = np.abs(w) > 2
            leaves.append((w * std).astype(np.float32))
        else:
            scale = {"network_params": np.sqrt(2.0), "actor_params": 0.01, "critic_params": 1.0}[name.split("/")[0]]
            leaves.append(_orthogonal(rng, shape[0], shape[1], scale))
    return np.concatenate([l.ravel() for l in leaves])


def unflatten(flat, spec=None) -> Dict[str, torch.Tensor]:
    """Views of the flat vector as named leaves (works for numpy arrays and torch tensors)."""
    spec = spec or spec_for_size(int(flat.shape[0]))
    out, off = {}, 0
    for name, shape in spec:
        n = int(np.prod(shape))
        out[name] = flat[off:off + n].reshape(shape)
        off += n
    assert off == flat.shape[0]
    return out


def _conv(x, kernel_hwio, bias):
    # flax nn.Conv, SAME, stride 1 (cleanba_ppo.py:156,158,167): cross-correlation, HWIO kernel.
    w = kernel_hwio.permute(3, 2, 0, 1)  # -> OIHW
    return F.conv2d(x, w, bias, stride=1, padding=1)


def _max_pool_same(x):
    # nn.max_pool(x,(3,3),strides=(2,2),padding="SAME") (cleanba_ppo.py:168), asymmetric -inf padding.
    ph = same_pool_pads(x.shape[2])
    pw = same_pool_pads(x.shape[3])
    x = F.pad(x, (pw[0], pw[1], ph[0], ph[1]), value=float("-inf"))
    return F.max_pool2d(x, kernel_size=3, stride=2)


def trunk_forward(p: Dict[str, torch.Tensor], obs_u8: torch.Tensor, channels=CHANNELS, record: dict = None) -> torch.Tensor:
    """Network.__call__ (cleanba_ppo.py:178-189).  obs_u8: [b,4,84,84] uint8 (NCHW) -> hidden [b,256].

    The reference transposes to NHWC; torch computes in NCHW, which is the same arithmetic.  Only the
    flatten order (h,w,c) (cleanba_ppo.py:185) needs an explicit permute."""
    dt = p["network_params/params/Dense_0/kernel"].dtype
    if "network_params/params/Conv_0/kernel" in p:
        return nature_trunk_forward(p, obs_u8, record)

    def rec(name, t):
        # test hook: keep intermediates (NCHW) and, when differentiating, their gradients
        if record is not None:
            if t.requires_grad:
                t.retain_grad()
            record[name] = t
        return t

    x = obs_u8.to(dt) / 255.0
    for s in range(len(channels)):
        pre = f"network_params/params/ConvSequence_{s}"
        x = rec(f"s{s}.y", _conv(x, p[f"{pre}/Conv_0/kernel"], p[f"{pre}/Conv_0/bias"]))
        x = rec(f"s{s}.p", _max_pool_same(x))
        for r in range(2):
            inputs = x
            x = torch.relu(x)
            x = rec(f"s{s}.a{r}pre", _conv(x, p[f"{pre}/ResidualBlock_{r}/Conv_0/kernel"], p[f"{pre}/ResidualBlock_{r}/Conv_0/bias"]))
            x = torch.relu(x)
            x = _conv(x, p[f"{pre}/ResidualBlock_{r}/Conv_1/kernel"], p[f"{pre}/ResidualBlock_{r}/Conv_1/bias"])
            x = rec(f"s{s}.b{r}", x + inputs)
    x = torch.relu(x)
    x = x.permute(0, 2, 3, 1).reshape(x.shape[0], -1)  # NHWC flatten
    x = x @ p["network_params/params/Dense_0/kernel"] + p["network_params/params/Dense_0/bias"]
    return torch.relu(x)


def nature_trunk_forward(p: Dict[str, torch.Tensor], obs_u8: torch.Tensor, record: dict = None) -> torch.Tensor:
    """Nature-CNN Network.__call__ (cleanba/legacy_scripts/cleanba_ppo_envpool_impala_atari_wrapper_naturecnn.py:143-178):
    x/255 -> Conv(32, 8x8, s4, VALID) -> relu -> Conv(64, 4x4, s2, VALID) -> relu -> Conv(64, 3x3, s1, VALID) -> relu -> NHWC flatten
    -> Dense(512) -> relu.  obs_u8 [b,4,84,84] -> hidden [b,512]."""
    dt = p["network_params/params/Dense_0/kernel"].dtype
    x = obs_u8.to(dt) / 255.0
    for l, stride in enumerate((4, 2, 1)):
        w = p[f"network_params/params/Conv_{l}/kernel"].permute(3, 2, 0, 1)      # HWIO -> OIHW
        x = torch.relu(F.conv2d(x, w, p[f"network_params/params/Conv_{l}/bias"], stride=stride, padding=0))
        if record is not None:
            record[f"c{l}"] = x
    x = x.permute(0, 2, 3, 1).reshape(x.shape[0], -1)
    x = x @ p["network_params/params/Dense_0/kernel"] + p["network_params/params/Dense_0/bias"]
    return torch.relu(x)


def heads(p: Dict[str, torch.Tensor], hidden: torch.Tensor):
    """Actor / Critic dense heads (cleanba_ppo.py:192-203) -> logits [b,18], value [b]."""
    logits = hidden @ p["actor_params/params/Dense_0/kernel"] + p["actor_params/params/Dense_0/bias"]
    value = hidden @ p["critic_params/params/Dense_0/kernel"] + p["critic_params/params/Dense_0/bias"]
    return logits, value.squeeze(-1)


def forward(flat_params, obs_u8, dtype=torch.float32):
    """Convenience: numpy in, torch out (logits, value, hidden)."""
    fp = torch.as_tensor(np.asarray(flat_params)).to(dtype)
    p = unflatten(fp)
    obs = torch.as_tensor(np.asarray(obs_u8))
    hidden = trunk_forward(p, obs)
    logits, value = heads(p, hidden)
    return logits, value, hidden
