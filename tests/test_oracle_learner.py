"""Closed-form anchors for the unpinned parts of the oracle."""
import numpy as np
import torch

from oracle import impala, network as net, optim, ppo


def test_param_count_and_order():
    spec = net.param_spec()
    assert net.num_params(spec) == 1_094_115 and len(spec) == 36
    assert spec[0][0].endswith("ConvSequence_0/Conv_0/bias") and spec[1][1] == (3, 3, 4, 16)
    assert spec[30][0] == "network_params/params/Dense_0/bias" and spec[31][1] == (3872, 256)
    assert spec[33][1] == (256, 18) and spec[35][1] == (256, 1)


def test_pool_asymmetric_padding():
    assert [net.same_pool_pads(n) for n in (84, 42, 21)] == [(0, 1), (0, 1), (1, 1)]
    x = torch.full((1, 1, 84, 84), -5.0)
    x[0, 0, 83, 83] = 7.0  # a one-hot at the last row/col must survive the (0,1) padding
    y = net._max_pool_same(x)
    assert y.shape == (1, 1, 42, 42) and y[0, 0, 41, 41] == 7.0 and (y.flatten()[:-1] == -5.0).all()
    x = torch.full((1, 1, 21, 21), -5.0)
    x[0, 0, 0, 0] = 3.0
    y = net._max_pool_same(x)
    assert y.shape == (1, 1, 11, 11) and y[0, 0, 0, 0] == 3.0 and y[0, 0, 0, 1] == -5.0


def test_flatten_order_is_hwc():
    # a one-hot at (h,w,c) of the last feature map must hit dense row (h*11+w)*32+c
    x = torch.zeros(1, 32, 11, 11)
    h, w, c = 3, 7, 5
    x[0, c, h, w] = 1.0
    flat = x.permute(0, 2, 3, 1).reshape(1, -1)
    assert flat[0, (h * 11 + w) * 32 + c] == 1.0 and flat.sum() == 1.0


def test_gae_closed_forms():
    rng = np.random.default_rng(0)
    T, B = 16, 5
    r = rng.standard_normal((T, B)).astype(np.float32)
    v = rng.standard_normal((T, B)).astype(np.float32)
    nv = rng.standard_normal(B).astype(np.float32)
    z = np.zeros((T, B), bool)
    # lambda = 0 -> A = delta
    adv, ret = ppo.compute_gae(r, v, z, nv, np.zeros(B, bool), 0.99, 0.0)
    vn = np.concatenate([v[1:], nv[None]])
    np.testing.assert_allclose(adv, r + np.float32(0.99) * vn - v, rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(ret, adv + v, rtol=0, atol=0)
    # lambda = 1, no dones -> discounted return minus value
    adv, _ = ppo.compute_gae(r, v, z, nv, np.zeros(B, bool), 0.9, 1.0)
    G = nv.astype(np.float64)
    for t in range(T - 1, -1, -1):
        G = r[t] + 0.9 * G
        np.testing.assert_allclose(adv[t], G - v[t], rtol=1e-4, atol=1e-4)
    # a done at t+1 cuts the bootstrap
    d = z.copy()
    d[5] = True
    adv, _ = ppo.compute_gae(r, v, d, nv, np.zeros(B, bool), 0.99, 0.95)
    np.testing.assert_allclose(adv[4], r[4] - v[4], rtol=1e-6, atol=1e-6)


def test_advantage_normalisation_groups():
    rng = np.random.default_rng(1)
    a = rng.standard_normal((8, 12)).astype(np.float32) * 3 + 1
    n = ppo.normalize_advantages(a, 4)
    for g in range(4):
        blk = n[:, g * 3:(g + 1) * 3]
        assert abs(blk.mean()) < 1e-6 and abs(blk.std() - 1) < 1e-5


def test_vtrace_on_policy_is_td1():
    rng = np.random.default_rng(2)
    T, B = 6, 3
    v = torch.tensor(rng.standard_normal((T + 1, B)), dtype=torch.float64)
    r = torch.tensor(rng.standard_normal((T, B)), dtype=torch.float64)
    disc = torch.full((T, B), 0.99, dtype=torch.float64)
    rho = torch.ones(T, B, dtype=torch.float64)
    errors, pg_adv, _ = impala.vtrace_td_error_and_advantage(v[:-1], v[1:], r, disc, rho)
    # rho = 1: target[t] = r[t] + gamma*target[t+1], target[T] = v[T]
    target = v[-1].clone()
    for t in range(T - 1, -1, -1):
        nxt = target
        target = r[t] + 0.99 * target
        np.testing.assert_allclose(errors[t].numpy(), (target - v[t]).numpy(), atol=1e-12)
        np.testing.assert_allclose(pg_adv[t].numpy(), (r[t] + 0.99 * nxt - v[t]).numpy(), atol=1e-12)
    # T = 1 reduces to clipped-rho TD
    e1, _, _ = impala.vtrace_td_error_and_advantage(v[:1], v[1:2], r[:1], disc[:1], torch.full((1, B), 0.5, dtype=torch.float64))
    np.testing.assert_allclose(e1.numpy(), (0.5 * (r[:1] + 0.99 * v[1:2] - v[:1])).numpy(), atol=1e-12)


def test_optimizers_closed_forms():
    g = np.array([0.3, -0.2, 0.0, 1e-3], np.float32)
    assert np.array_equal(optim.clip_by_global_norm(g, 10.0), g)  # identity when norm < c
    c = optim.clip_by_global_norm(g * 100, 0.5)
    np.testing.assert_allclose(optim.global_norm(c), 0.5, rtol=1e-6)
    # Adam step 1: u = g / (|g| + eps * ...) ~ sign(g)
    p = np.zeros(4, np.float32)
    a = optim.Adam(4)
    p1 = a.step(p, g, 1e-2)
    want = -1e-2 * g / (np.abs(g) + 1e-5)
    np.testing.assert_allclose(p1, want, rtol=1e-4, atol=1e-9)
    # RMSProp step 1: nu = .01 g^2 -> u = g / (.1|g| + .01)
    rms = optim.RMSPropPyTorchStyle(4)
    p1 = rms.step(p, g, 1e-2)
    np.testing.assert_allclose(p1, -1e-2 * g / (0.1 * np.abs(g) + 0.01), rtol=1e-5, atol=1e-9)
    # schedules (cleanba_ppo.py:475-479, cleanba_impala.py:515-519)
    assert optim.linear_schedule(0, 2.5e-4, 16, 100) == np.float32(2.5e-4)
    assert optim.linear_schedule(15, 2.5e-4, 16, 100) == np.float32(2.5e-4)
    np.testing.assert_allclose(optim.linear_schedule(16, 2.5e-4, 16, 100), 2.5e-4 * 0.99, rtol=1e-6)


def test_ppo_loss_first_minibatch_identities():
    # with behaviour logprobs == new logprobs the ratio is 1: pg_loss = -mean(A), approx_kl = 0
    rng = np.random.default_rng(3)
    fp = net.init_params(1)
    obs = rng.integers(0, 256, (6, 4, 84, 84), dtype=np.uint8)
    act = rng.integers(0, 18, 6).astype(np.int32)
    p = net.unflatten(torch.tensor(fp))
    with torch.no_grad():
        lp, ent, val = ppo.logprob_entropy_value(p, torch.as_tensor(obs), torch.as_tensor(act))
    adv = rng.standard_normal(6).astype(np.float32)
    ret = rng.standard_normal(6).astype(np.float32)
    stats, g = ppo.ppo_loss_and_grad(fp, obs, act, lp.numpy(), adv, ret)
    np.testing.assert_allclose(stats[1], -adv.mean(), rtol=1e-5, atol=1e-6)
    assert abs(stats[4]) < 1e-6 and np.isfinite(g).all() and g.shape == fp.shape
    np.testing.assert_allclose(stats[3], ent.mean().item(), rtol=1e-6)
    assert 0 < stats[3] <= np.log(18) + 1e-5
    # fp64 autograd agrees with fp32 autograd (sanity of the gradient oracle)
    stats64, g64 = ppo.ppo_loss_and_grad(fp, obs, act, lp.numpy(), adv, ret, dtype=torch.float64)
    np.testing.assert_allclose(stats, stats64, rtol=1e-5, atol=1e-6)
    rel = np.linalg.norm(g - g64) / np.linalg.norm(g64)
    assert rel < 1e-4, rel


def test_multi_shard_update_matches_manual_mean():
    rng = np.random.default_rng(4)
    fp = net.init_params(1)
    T, B = 2, 8

    def shard():
        return ppo.Shard(obs=rng.integers(0, 256, (T, B, 4, 84, 84), dtype=np.uint8), dones=rng.random((T, B)) < 0.1,
                         actions=rng.integers(0, 18, (T, B)).astype(np.int32),
                         logprobs=np.full((T, B), -2.89, np.float32),
                         values=(rng.standard_normal((T, B)) * 0.1).astype(np.float32),
                         rewards=rng.integers(-1, 2, (T, B)).astype(np.float32),
                         next_obs=rng.integers(0, 256, (B, 4, 84, 84), dtype=np.uint8), next_done=np.zeros(B, bool))
    shards = [shard(), shard()]
    from oracle import threefry as tf
    key = tf.split(tf.PRNGKey(1), 4)[0]
    L = ppo.PPOLearner(fp, ppo.PPOConfig(update_epochs=1, num_minibatches=4))
    rec = []
    stats, k2 = L.update(shards, key, record=rec)
    assert len(rec) == 4 and np.isfinite(stats).all() and (k2 != key).any()
    assert optim.global_norm(rec[0]["grad"]) <= 0.5 + 1e-6
    assert not np.array_equal(rec[-1]["params"], fp)


def test_gradient_accumulation_is_optax_multisteps():
    """gradient_accumulation_steps = k (cleanba_ppo.py:78,492-500,607): the shuffled batch is cut into num_minibatches * k mini-steps;
    optax.MultiSteps keeps the running mean of the k mini-step gradients and applies clip + Adam to it on the k-th.  Checked against an
    explicit computation with the primitives (one update epoch, 1 minibatch, k = 2)."""
    import numpy as np
    from oracle import network as net, optim, ppo as oppo, threefry as tf
    rng = np.random.default_rng(5)
    T, B = 2, 4
    shard = oppo.Shard(obs=rng.integers(0, 256, (T, B, 4, 84, 84), dtype=np.uint8), dones=np.zeros((T, B), bool),
                       actions=rng.integers(0, 18, (T, B)).astype(np.int32), logprobs=np.full((T, B), np.log(1 / 18), np.float32),
                       values=(rng.standard_normal((T, B)) * 0.1).astype(np.float32), rewards=rng.choice([-1.0, 0.0, 1.0], size=(T, B)).astype(np.float32),
                       next_obs=rng.integers(0, 256, (B, 4, 84, 84), dtype=np.uint8), next_done=np.zeros(B, bool))
    p0 = net.init_params(1)
    key = tf.split(tf.PRNGKey(3), 4)[0]
    cfg = oppo.PPOConfig(update_epochs=1, num_minibatches=1, gradient_accumulation_steps=2, num_updates=10, norm_adv=False)
    L = oppo.PPOLearner(p0, cfg)
    stats, _ = L.update([shard], key)
    assert L.opt.count == 1
    adv, ret = oppo.PPOLearner(p0, cfg).prepare(shard)
    _, sub = tf.split(key)
    idx = tf.permutation(sub, T * B).reshape(2, -1)
    flat = lambda x: x.reshape((-1,) + x.shape[2:])
    gs = [oppo.ppo_loss_and_grad(p0, flat(shard.obs)[ii], flat(shard.actions)[ii], flat(shard.logprobs)[ii], adv.reshape(-1)[ii], ret.reshape(-1)[ii])[1]
          for ii in idx]
    acc = gs[0] + (gs[1] - gs[0]) / np.float32(2)
    want = optim.Adam(p0.size).step(p0, optim.clip_by_global_norm(acc.astype(np.float32), 0.5), optim.linear_schedule(0, 2.5e-4, 1, 10))
    assert np.abs(L.params - want).max() < 1e-7
