"""CPU study for the round-2 activation-carrier decision: how far is the trunk output from exact arithmetic when
conv / dense operands are carried as split low-precision planes and only the significant cross products are kept?

  bf16x3  : x = hi + mid + lo (3 x bf16, 24 bits), six products   -- what the CUDA kernels do today (three MMAs per K step)
  bf16x2  : x = hi + mid      (2 x bf16, 16 bits), three products
  f16x2s  : hi = fp16(x), mid = fp16((x - hi) * 2^11) (22 bits), three products, two MMAs per K step  -- the round-2 candidate
  fp32    : plain fp32 torch (accumulation-order noise only), for scale

Products and sums are evaluated in fp64, so the numbers isolate the CARRIER error.  Frames and parameters as in the tests
(seeded).  Prints the max error relative to the tensor's max (the metric of tests/test_gpu_parity.py) for hidden / logits / value
and the number of Gumbel-max action flips over all samples."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.nn.functional as F
from oracle import network as net, threefry as tf

torch.set_num_threads(os.cpu_count() or 1)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
D = torch.float64


def bf16(x):
    return x.to(torch.float32).to(torch.bfloat16).to(D)


def f16(x):
    return x.to(torch.float32).to(torch.float16).to(D)


def planes(x, kind):
    """-> list of (plane, scale) with x ~= sum(plane * scale)"""
    if kind == "bf16x3":
        h = bf16(x); m = bf16(x - h); l = bf16(x - h - m)
        return [(h, 1.0), (m, 1.0), (l, 1.0)]
    if kind == "bf16x2":
        h = bf16(x); m = bf16(x - h)
        return [(h, 1.0), (m, 1.0)]
    if kind == "f16x2s":
        h = f16(x); m = f16((x - h) * 2048.0)
        return [(h, 1.0), (m, 1.0 / 2048.0)]
    raise ValueError(kind)


KEEP = {"bf16x3": [(0, 0), (0, 1), (0, 2), (1, 0), (1, 1), (2, 0)], "bf16x2": [(0, 0), (0, 1), (1, 0)], "f16x2s": [(0, 0), (0, 1), (1, 0)]}


def conv(x, w_hwio, b, kind):
    w = w_hwio.permute(3, 2, 0, 1)
    if kind in ("exact", "fp32"):
        return F.conv2d(x, w, b, padding=1)
    xp, wp = planes(x, kind), planes(w, kind)
    y = 0
    for i, j in KEEP[kind]:
        y = y + F.conv2d(xp[i][0], wp[j][0], None, padding=1) * (xp[i][1] * wp[j][1])
    return y + b.view(1, -1, 1, 1)


def dense(x, w, b, kind):
    if kind in ("exact", "fp32"):
        return x @ w + b
    xp, wp = planes(x, kind), planes(w, kind)
    y = 0
    for i, j in KEEP[kind]:
        y = y + (xp[i][0] @ wp[j][0]) * (xp[i][1] * wp[j][1])
    return y + b


def forward(p, obs, kind):
    dt = torch.float32 if kind == "fp32" else D
    p = {k: v.to(dt) for k, v in p.items()}
    frames = obs.to(dt)
    x = None
    for s in range(3):
        pre = f"network_params/params/ConvSequence_{s}"
        if s == 0:   # frames are exact in every carrier; the 1/255 is applied to the accumulator (cleanba_ppo.py:181)
            if kind in ("exact", "fp32"):
                x = conv(frames / 255.0, p[f"{pre}/Conv_0/kernel"], p[f"{pre}/Conv_0/bias"], kind)
            else:
                wp = planes(p[f"{pre}/Conv_0/kernel"], kind)
                y = 0
                for (wpl, sc) in wp:
                    y = y + F.conv2d(frames, wpl.permute(3, 2, 0, 1), None, padding=1) * sc
                x = y / 255.0 + p[f"{pre}/Conv_0/bias"].view(1, -1, 1, 1)
        else:
            x = conv(x, p[f"{pre}/Conv_0/kernel"], p[f"{pre}/Conv_0/bias"], kind)
        x = net._max_pool_same(x)
        for r in range(2):
            inp = x
            x = conv(torch.relu(x), p[f"{pre}/ResidualBlock_{r}/Conv_0/kernel"], p[f"{pre}/ResidualBlock_{r}/Conv_0/bias"], kind)
            x = conv(torch.relu(x), p[f"{pre}/ResidualBlock_{r}/Conv_1/kernel"], p[f"{pre}/ResidualBlock_{r}/Conv_1/bias"], kind)
            x = x + inp
    x = torch.relu(x).permute(0, 2, 3, 1).reshape(x.shape[0], -1)
    h = torch.relu(dense(x, p["network_params/params/Dense_0/kernel"], p["network_params/params/Dense_0/bias"], kind))
    logits, value = net.heads(p, h)
    return h.to(D), logits.to(D), value.to(D)


rng = np.random.default_rng(11)
obs = torch.from_numpy(rng.integers(0, 256, (N, 4, 84, 84), dtype=np.uint8))
p = net.unflatten(torch.tensor(net.init_params(1)))
with torch.no_grad():
    ref = forward(p, obs, "exact")
    key = tf.split(tf.PRNGKey(1), 4)[0]
    u = torch.from_numpy(tf.uniform(key, (N, 18)).astype(np.float64))
    g = -torch.log(-torch.log(u))
    a_ref = (ref[1] + g).argmax(1)
    srt = torch.sort(ref[1] + g, dim=1, descending=True).values
    print(f"{N} frames; smallest Gumbel-max decision gap in the exact arithmetic: {float((srt[:, 0] - srt[:, 1]).min()):.3e}")
    print(f"{'carrier':8s} {'hidden':>10s} {'logits':>10s} {'value':>10s}  action flips")
    for kind in ("fp32", "bf16x3", "f16x2s", "bf16x2"):
        h, l, v = forward(p, obs, kind)
        e = [float((a - b).abs().max() / b.abs().max()) for a, b in zip((h, l, v), ref)]
        flips = int(((l + g).argmax(1) != a_ref).sum())
        print(f"{kind:8s} {e[0]:10.2e} {e[1]:10.2e} {e[2]:10.2e}  {flips}")
