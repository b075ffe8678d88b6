"""CPU study, part 2 (DESIGN.md section 7, round-2 plan): error of the PPO minibatch GRADIENT when every conv / dense operand of
the forward AND the backward pass is carried as split low-precision planes, against exact (fp64) arithmetic.

  current : forward x, w = bf16 x 3 (six products); backward G = bf16 x 2, W (dgrad) and X (wgrad) = bf16 x 2 (three products)
  round-2 : everything fp16 x 2 with the mid plane scaled by 2^11 (three products); gradients carry the loss scale S = 2^16
  fp32    : plain fp32 autograd (accumulation-order noise only)

Only the carrier rounding is emulated (products and sums in fp64); relu / max-pool gates are evaluated on the emulated forward, so
gate flips caused by forward differences show up in the gradient exactly as they do on the GPU."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.nn.functional as F
from torch.nn.grad import conv2d_input, conv2d_weight
from oracle import network as net, ppo as oppo

torch.set_num_threads(os.cpu_count() or 1)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 24
D = torch.float64
S_LOSS = 65536.0


def bf16(x): return x.to(torch.float32).to(torch.bfloat16).to(D)
def f16(x): return x.to(torch.float32).to(torch.float16).to(D)


def planes(x, kind):
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


K6 = [(0, 0), (0, 1), (0, 2), (1, 0), (1, 1), (2, 0)]
K3 = [(0, 0), (0, 1), (1, 0)]
SCHEMES = {  # forward (x kind, w kind, kept), backward (g kind, w kind, x kind, kept), loss scale
    "current": dict(fx="bf16x3", fw="bf16x3", fk=K6, bg="bf16x2", bw="bf16x2", bx="bf16x2", bk=K3, S=1.0),
    "round-2": dict(fx="f16x2s", fw="f16x2s", fk=K3, bg="f16x2s", bw="f16x2s", bx="f16x2s", bk=K3, S=S_LOSS),
}


def combine(ap, bp, kept, op):
    y = 0
    for i, j in kept:
        y = y + op(ap[i][0], bp[j][0]) * (ap[i][1] * bp[j][1])
    return y


class QConv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w_hwio, b, sch, frames):
        w = w_hwio.permute(3, 2, 0, 1).contiguous()
        ctx.save_for_backward(x, w)
        ctx.sch, ctx.frames = sch, frames
        xp = [(x, 1.0)] if frames else planes(x, sch["fx"])          # frames are exact in every carrier
        kept = [(0, j) for j in range(len(planes(w, sch["fw"])))] if frames else sch["fk"]
        return combine(xp, planes(w, sch["fw"]), kept, lambda a, c: F.conv2d(a, c, None, padding=1)) + b.view(1, -1, 1, 1)

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        sch = ctx.sch
        gp = [(p, s / sch["S"]) for p, s in planes(gy * sch["S"], sch["bg"])]
        dx = None
        if not ctx.frames:
            dx = combine(gp, planes(w, sch["bw"]), sch["bk"], lambda g, c: conv2d_input(x.shape, c, g, padding=1))
        xp = [(x, 1.0)] if ctx.frames else planes(x, sch["bx"])
        kept = [(0, 0), (0, 1)] if ctx.frames else sch["bk"]
        dw = combine(xp, gp, kept, lambda a, g: conv2d_weight(a, w.shape, g, padding=1))
        return dx, dw.permute(2, 3, 1, 0), gy.sum((0, 2, 3)), None, None


class QDense(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, sch):
        ctx.save_for_backward(x, w)
        ctx.sch = sch
        return combine(planes(x, sch["fx"]), planes(w, sch["fw"]), sch["fk"], lambda a, c: a @ c) + b

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        sch = ctx.sch
        gp = [(p, s / sch["S"]) for p, s in planes(gy * sch["S"], sch["bg"])]
        dx = combine(gp, planes(w, sch["bw"]), sch["bk"], lambda g, c: g @ c.t())
        dw = combine(planes(x, sch["bx"]), gp, sch["bk"], lambda a, g: a.t() @ g)
        return dx, dw, gy.sum(0), None


def trunk(p, obs, sch):
    if sch is None:
        return net.trunk_forward(p, obs)
    x = obs.to(D)
    for s in range(3):
        pre = f"network_params/params/ConvSequence_{s}"
        if s == 0:
            x = QConv.apply(x, p[f"{pre}/Conv_0/kernel"] / 255.0, p[f"{pre}/Conv_0/bias"], sch, True)   # 1/255 on the accumulator
        else:
            x = QConv.apply(x, p[f"{pre}/Conv_0/kernel"], p[f"{pre}/Conv_0/bias"], sch, False)
        x = net._max_pool_same(x)
        for r in range(2):
            inp = x
            x = QConv.apply(torch.relu(x), p[f"{pre}/ResidualBlock_{r}/Conv_0/kernel"], p[f"{pre}/ResidualBlock_{r}/Conv_0/bias"], sch, False)
            x = QConv.apply(torch.relu(x), p[f"{pre}/ResidualBlock_{r}/Conv_1/kernel"], p[f"{pre}/ResidualBlock_{r}/Conv_1/bias"], sch, False)
            x = x + inp
    x = torch.relu(x).permute(0, 2, 3, 1).reshape(x.shape[0], -1)
    return torch.relu(QDense.apply(x, p["network_params/params/Dense_0/kernel"], p["network_params/params/Dense_0/bias"], sch))


def loss_grad(flat, obs, actions, oldlp, adv, ret, sch, dtype=D):
    fp = torch.tensor(flat, dtype=dtype, requires_grad=True)
    p = net.unflatten(fp)
    hidden = trunk(p, obs, sch)
    logits, value = net.heads(p, hidden)
    lp_all = torch.log_softmax(logits, -1)
    newlp = lp_all.gather(1, actions.long()[:, None]).squeeze(1)
    ent = -(lp_all * torch.softmax(logits, -1)).sum(-1)
    loss, *_ = oppo.ppo_loss_from_heads(newlp, ent, value, oldlp.to(dtype), adv.to(dtype), ret.to(dtype), 0.1, 0.01, 0.5)
    loss.backward()
    return float(loss.detach()), fp.grad.detach().to(D)


rng = np.random.default_rng(5)
obs = torch.from_numpy(rng.integers(0, 256, (N, 4, 84, 84), dtype=np.uint8))
actions = torch.from_numpy(rng.integers(0, 18, N).astype(np.int32))
oldlp = torch.full((N,), float(np.log(1 / 18)))
adv = torch.from_numpy(rng.standard_normal(N).astype(np.float32))
ret = torch.from_numpy(rng.standard_normal(N).astype(np.float32))
flat = net.init_params(1).astype(np.float64)
l0, g0 = loss_grad(flat, obs, actions, oldlp, adv, ret, None)
print(f"{N} samples; exact loss {l0:.8f}, |grad| {float(g0.norm()):.4e}")
print(f"{'scheme':8s} {'loss rel err':>13s} {'grad rel err':>13s}")
l, g = loss_grad(flat.astype(np.float32), obs, actions, oldlp, adv, ret, None, dtype=torch.float32)
print(f"{'fp32':8s} {abs(l - l0) / abs(l0):13.2e} {float((g - g0).norm() / g0.norm()):13.2e}")
for name, sch in SCHEMES.items():
    l, g = loss_grad(flat, obs, actions, oldlp, adv, ret, sch)
    print(f"{name:8s} {abs(l - l0) / abs(l0):13.2e} {float((g - g0).norm() / g0.norm()):13.2e}")
