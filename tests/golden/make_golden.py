"""Generates tests/golden/oracle_golden.npz FROM THE ORACLE (regression pins; see oracle/__init__.py for what is and is
not pinned against JAX).  Run:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import impala as oimpala, network as net, optim, ppo as oppo, threefry as tf  # noqa: E402


def build():
    torch.set_num_threads(1)          # fixed reduction order
    rng = np.random.default_rng(2024)
    params = net.init_params(1)
    key = tf.split(tf.PRNGKey(1), 4)[0]
    obs = rng.integers(0, 256, (4, 4, 84, 84), dtype=np.uint8)
    _, action, logprob, value, key2, logits = oppo.get_action_and_value(params, obs, key)
    T, B = 12, 8
    r = rng.choice([-1.0, 0.0, 1.0], size=(T, B)).astype(np.float32)
    v = (rng.standard_normal((T, B)) * 0.5).astype(np.float32)
    d = rng.random((T, B)) < 0.1
    nv = (rng.standard_normal(B) * 0.5).astype(np.float32)
    nd = rng.random(B) < 0.2
    adv, ret = oppo.compute_gae(r, v, d, nv, nd)
    acts = rng.integers(0, 18, 4).astype(np.int32)
    a4 = rng.standard_normal(4).astype(np.float32)
    r4 = rng.standard_normal(4).astype(np.float32)
    stats, grad = oppo.ppo_loss_and_grad(params, obs, acts, logprob, a4, r4)
    g = (rng.standard_normal(64) * 0.1).astype(np.float32)
    p = rng.standard_normal(64).astype(np.float32)
    adam = optim.Adam(64)
    p1 = adam.step(p, optim.clip_by_global_norm(g, 0.5), 2.5e-4)
    p2 = adam.step(p1, optim.clip_by_global_norm(g * 0.1, 0.5), 2.5e-4)
    rms = optim.RMSPropPyTorchStyle(64)
    q1 = rms.step(p, optim.clip_by_global_norm(g, 40.0), 6e-4)
    vt = torch.tensor(rng.standard_normal((5, 3)), dtype=torch.float64)
    rt = torch.tensor(rng.standard_normal((4, 3)), dtype=torch.float64)
    disc = torch.tensor((rng.random((4, 3)) > 0.1) * 0.99, dtype=torch.float64)
    rho = torch.tensor(np.exp(rng.standard_normal((4, 3)) * 0.3), dtype=torch.float64)
    err, pg_adv, q = oimpala.vtrace_td_error_and_advantage(vt[:-1], vt[1:], rt, disc, rho)
    return dict(
        key=key, key_after=key2, obs=obs, action=action, logprob=logprob, value=value, logits=logits,
        perm_2048_head=tf.permutation(key, 2048)[:64], perm_15360_sum=np.int64((tf.permutation(key, 15360)[:1000].astype(np.int64) * np.arange(1000)).sum()),
        gae_r=r, gae_v=v, gae_d=d, gae_nv=nv, gae_nd=nd, gae_adv=adv, gae_ret=ret, gae_norm=oppo.normalize_advantages(adv, 4),
        ppo_actions=acts, ppo_adv=a4, ppo_ret=r4, ppo_stats=stats, ppo_grad_norm=np.float64(np.linalg.norm(grad)), ppo_grad_head=grad[:16],
        opt_g=g, opt_p=p, adam_p1=p1, adam_p2=p2, rms_p1=q1,
        vt_v=vt.numpy(), vt_r=rt.numpy(), vt_disc=disc.numpy(), vt_rho=rho.numpy(), vt_err=err.numpy(), vt_pgadv=pg_adv.numpy(), vt_q=q.numpy())


if __name__ == "__main__":
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_golden.npz")
    np.savez_compressed(out, **build())
    print("wrote", out, os.path.getsize(out), "bytes")
