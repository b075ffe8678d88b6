"""The oracle against its committed golden vectors (tests/golden/oracle_golden.npz, made by tests/golden/make_golden.py)."""
import os

import numpy as np
import torch

from oracle import impala as oimpala, network as net, optim, ppo as oppo, threefry as tf

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_golden.npz"))


def test_actor_golden():
    _, a, lp, v, key2, logits = oppo.get_action_and_value(net.init_params(1), G["obs"], G["key"])
    assert np.array_equal(a, G["action"]) and key2.tolist() == G["key_after"].tolist()
    np.testing.assert_allclose(logits, G["logits"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(v, G["value"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(lp, G["logprob"], rtol=1e-6)


def test_permutation_golden():
    assert np.array_equal(tf.permutation(G["key"], 2048)[:64], G["perm_2048_head"])
    p = tf.permutation(G["key"], 15360)
    assert np.int64((p[:1000].astype(np.int64) * np.arange(1000)).sum()) == G["perm_15360_sum"]


def test_gae_golden():
    adv, ret = oppo.compute_gae(G["gae_r"], G["gae_v"], G["gae_d"], G["gae_nv"], G["gae_nd"])
    assert np.array_equal(adv, G["gae_adv"]) and np.array_equal(ret, G["gae_ret"])
    np.testing.assert_allclose(oppo.normalize_advantages(adv, 4), G["gae_norm"], rtol=1e-6, atol=1e-7)


def test_ppo_loss_golden():
    stats, grad = oppo.ppo_loss_and_grad(net.init_params(1), G["obs"], G["ppo_actions"], G["logprob"], G["ppo_adv"], G["ppo_ret"])
    np.testing.assert_allclose(stats[:4], G["ppo_stats"][:4], rtol=1e-5)
    np.testing.assert_allclose(np.linalg.norm(grad), G["ppo_grad_norm"], rtol=1e-4)


def test_optimizer_golden():
    adam = optim.Adam(64)
    p1 = adam.step(G["opt_p"], optim.clip_by_global_norm(G["opt_g"], 0.5), 2.5e-4)
    p2 = adam.step(p1, optim.clip_by_global_norm(G["opt_g"] * 0.1, 0.5), 2.5e-4)
    assert np.array_equal(p1, G["adam_p1"]) and np.array_equal(p2, G["adam_p2"])
    rms = optim.RMSPropPyTorchStyle(64)
    assert np.array_equal(rms.step(G["opt_p"], optim.clip_by_global_norm(G["opt_g"], 40.0), 6e-4), G["rms_p1"])


def test_vtrace_golden():
    vt = torch.tensor(G["vt_v"])
    err, adv, q = oimpala.vtrace_td_error_and_advantage(vt[:-1], vt[1:], torch.tensor(G["vt_r"]), torch.tensor(G["vt_disc"]), torch.tensor(G["vt_rho"]))
    np.testing.assert_allclose(err.numpy(), G["vt_err"], atol=1e-12)
    np.testing.assert_allclose(adv.numpy(), G["vt_pgadv"], atol=1e-12)
    np.testing.assert_allclose(q.numpy(), G["vt_q"], atol=1e-12)
