"""GPU parity of the Nature-CNN trunk (cleanba/legacy_scripts/cleanba_ppo_envpool_impala_atari_wrapper_naturecnn.py:143-178)
behind the same C ABI (cb_config.model = CB_MODEL_NATURE_CNN): forward, sampled actor step, PPO / IMPALA minibatch gradients and a
whole pinned PPO update against the CPU oracle.  Same bars as the IMPALA-ResNet tests (tests/test_gpu_parity.py)."""
import numpy as np
import pytest
import torch

from _pin import pin_hook
from oracle import impala as oimpala
from oracle import network as net
from oracle import ppo as oppo
from oracle import threefry as tf
from test_gpu_parity import _diag, _frames, _leafwise, _relerr

pytestmark = pytest.mark.gpu
NATURE = 1


@pytest.fixture(scope="module")
def agent():
    from cleanba_b200 import agent as ag
    return ag


@pytest.fixture(scope="module")
def params():
    return net.init_params(3, net.nature_param_spec())


def test_nature_param_layout_matches_oracle(agent, params):
    from cleanba_b200 import lib, params as P
    spec = net.nature_param_spec()
    got = lib.leaves(18, NATURE)
    assert [(n, tuple(s)) for n, _, s in got] == [(n, tuple(s)) for n, s in spec]
    assert np.array_equal(P.init_params(3, 18, NATURE), params)
    ctx = agent.Context("cuda:0", max_batch=4, model=NATURE)
    assert ctx.num_params == params.size == 1693875 and ctx.hidden_width == 512
    ctx.set_params(params)
    assert np.array_equal(ctx.get_params().cpu().numpy(), params)
    ctx.close()


@pytest.mark.parametrize("n", [1, 5, 60, 129])
def test_nature_forward(agent, params, n):
    rng = np.random.default_rng(100 + n)
    pool = _frames(rng, n + 7)
    idx = rng.permutation(n + 7)[:n].astype(np.int32)
    ctx = agent.Context("cuda:0", max_batch=160, model=NATURE)
    ctx.set_params(params)
    logits, value = ctx.policy_value(torch.from_numpy(pool).to(ctx.device), torch.from_numpy(idx).to(ctx.device))
    torch.cuda.synchronize()
    with torch.no_grad():
        ol, ov, oh = net.forward(params, pool[idx])
    errs = dict(hidden=_relerr(ctx.debug_tensor("hidden", (n, 512)), oh.numpy()), logits=_relerr(logits.cpu().numpy(), ol.numpy()),
                value=_relerr(value.cpu().numpy(), ov.numpy()))
    _diag(f"nature_forward_n{n}", **errs)
    assert errs["hidden"] < 2e-5 and errs["logits"] < 1e-4 and errs["value"] < 1e-4, errs      # forward bar 2e-5 (rel. to max), outputs 1e-4
    ctx.close()


def test_nature_actor_step_actions_bit_exact(agent, params):
    rng = np.random.default_rng(113)
    N = 60
    ctx = agent.Context("cuda:0", max_batch=N, model=NATURE)
    ctx.set_params(params)
    key = tf.split(tf.PRNGKey(1), 4)[0]
    kt = agent.key_tensor(key, ctx.device)
    okey = key
    for step in range(3):
        obs = _frames(rng, N)
        _, sk = tf.split(okey)
        u = tf.uniform(sk, (N, 18))
        action, logprob, value, logits = ctx.actor_step(torch.from_numpy(obs).to(ctx.device), kt, True, True)
        _, oa, olp, ov, okey, ologits = oppo.get_action_and_value(params, obs, okey)
        assert agent.key_numpy(kt).tolist() == okey.tolist()
        a = action.cpu().numpy()
        assert np.array_equal(a, oppo.gumbel_argmax(logits.cpu().numpy(), u)), "sampling head is not bit-exact"
        lerr = np.abs(logits.cpu().numpy() - ologits).max()
        with np.errstate(divide="ignore"):
            pert = ologits - np.log(-np.log(u))
        top2 = np.sort(pert, axis=1)[:, -2:]
        assert not ((a != oa) & (top2[:, 1] - top2[:, 0] > 4 * lerr)).any(), "action differs where the oracle's decision is not a near-tie"
        assert np.array_equal(a, oa)
        assert _relerr(logprob.cpu().numpy(), olp) < 1e-4 and _relerr(value.cpu().numpy(), ov) < 1e-4
    ctx.close()


@pytest.mark.parametrize("mb", [8, 70, 200])
def test_nature_ppo_grad_matches_autograd(agent, params, mb):
    from cleanba_b200 import lib
    rng = np.random.default_rng(114 + mb)
    N = mb + 9
    obs = _frames(rng, N)
    actions = rng.integers(0, 18, N).astype(np.int32)
    oldlp = (np.log(1 / 18) + rng.standard_normal(N) * 0.05).astype(np.float32)
    adv = rng.standard_normal(N).astype(np.float32)
    ret = rng.standard_normal(N).astype(np.float32)
    idx = rng.permutation(N)[:mb].astype(np.int32)
    ctx = agent.Context("cuda:0", max_batch=mb, train=True, model=NATURE)
    ctx.set_params(params)
    dev = ctx.device
    grads = torch.zeros(ctx.num_params, dtype=torch.float32, device=dev)
    stats = torch.zeros(5, dtype=torch.float32, device=dev)
    tt = lambda x: torch.from_numpy(x).to(dev)
    ctx.ppo_grad(tt(obs), tt(idx), mb, tt(actions), tt(oldlp), tt(adv), tt(ret), 0.1, 0.01, 0.5, grads, stats)
    torch.cuda.synchronize()
    ostats, og = oppo.ppo_loss_and_grad(params, obs[idx], actions[idx], oldlp[idx], adv[idx], ret[idx], dtype=torch.float64)
    g = grads.cpu().numpy().astype(np.float64)
    st = stats.cpu().numpy()
    lw = _leafwise(g, og, lib.leaves(18, NATURE))
    tot = float(np.linalg.norm(g - og) / np.linalg.norm(og))
    serr = [abs(st[i] - ostats[i]) / max(abs(ostats[i]), 1e-6) for i in range(4)]
    _diag(f"nature_ppo_grad_mb{mb}", total=tot, stats_relerr=serr, worst_leaf=max(lw, key=lw.get), worst=max(lw.values()))
    assert max(serr) < 1e-4, (st, ostats)                     # losses: 1e-4 relative
    assert tot < 1e-3 and max(lw.values()) < 1e-2, (tot, lw)  # gradient: 1e-3 (norm), 1e-2 per leaf
    g2 = torch.zeros_like(grads)
    ctx.ppo_grad(tt(obs), tt(idx), mb, tt(actions), tt(oldlp), tt(adv), tt(ret), 0.1, 0.01, 0.5, g2, stats)
    assert torch.equal(grads, g2), "not deterministic"
    ctx.close()


def test_nature_impala_grad_matches_autograd(agent, params):
    from cleanba_b200 import lib
    rng = np.random.default_rng(115)
    T1, Bl, B = 6, 8, 4
    obs = rng.integers(0, 256, (T1, Bl, 4, 84, 84), dtype=np.uint8)
    a = rng.integers(0, 18, (T1, Bl)).astype(np.int32)
    mu = (rng.standard_normal((T1, Bl, 18)) * 0.3).astype(np.float32)
    r = rng.choice([-1.0, 0.0, 1.0], size=(T1, Bl)).astype(np.float32)
    d = rng.random((T1, Bl)) < 0.15
    fs = rng.random((T1, Bl)) < 0.15
    cols = np.arange(4, 8)
    idx = (np.arange(T1)[:, None] * Bl + cols[None, :]).astype(np.int32).ravel()
    ctx = agent.Context("cuda:0", max_batch=T1 * B, algo=1, train=True, model=NATURE)
    ctx.set_params(params)
    tt = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(ctx.device)
    grads = torch.zeros(ctx.num_params, dtype=torch.float32, device=ctx.device)
    stats = torch.zeros(4, dtype=torch.float32, device=ctx.device)
    ctx.impala_grad(tt(obs.reshape(-1, 4, 84, 84)), tt(idx), T1, B, tt(a.ravel()), tt(mu.reshape(-1, 18)), tt(r.ravel()),
                    tt(d.ravel()), tt(fs.ravel()), 0.99, 0.5, 0.01, grads, stats)
    torch.cuda.synchronize()
    ostats, og = oimpala.impala_loss_and_grad(params, obs[:, cols], a[:, cols], mu[:, cols], r[:, cols], d[:, cols], fs[:, cols], dtype=torch.float64)
    g = grads.cpu().numpy().astype(np.float64)
    st = stats.cpu().numpy()
    tot = float(np.linalg.norm(g - og) / np.linalg.norm(og))
    serr = [abs(st[i] - ostats[i]) / max(abs(ostats[i]), 1e-6) for i in range(4)]
    lw = _leafwise(g, og, lib.leaves(18, NATURE))
    _diag("nature_impala_grad", total=tot, stats_relerr=serr, worst_leaf=max(lw, key=lw.get), worst=max(lw.values()))
    assert max(serr) < 1e-4 and tot < 1e-3, (st, ostats, tot)
    ctx.close()


def test_nature_ppo_update_every_step_pinned_to_oracle(agent, params):
    from cleanba_b200.learner import PPOHyper, PPOLearner
    rng = np.random.default_rng(117)
    T, B = 8, 8
    shard = oppo.Shard(obs=rng.integers(0, 256, (T, B, 4, 84, 84), dtype=np.uint8), dones=rng.random((T, B)) < 0.1,
                       actions=rng.integers(0, 18, (T, B)).astype(np.int32),
                       logprobs=(np.log(1 / 18) + rng.standard_normal((T, B)) * 0.01).astype(np.float32),
                       values=(rng.standard_normal((T, B)) * 0.1).astype(np.float32),
                       rewards=rng.choice([-1.0, 0.0, 1.0], size=(T, B)).astype(np.float32),
                       next_obs=rng.integers(0, 256, (B, 4, 84, 84), dtype=np.uint8), next_done=rng.random(B) < 0.1)
    key = tf.split(tf.PRNGKey(1), 4)[0]
    ol = oppo.PPOLearner(params, oppo.PPOConfig(update_epochs=2, num_updates=10))
    record = []
    ostats, okey = ol.update([shard], key, record=record)
    L = PPOLearner("cuda:0", PPOHyper(update_epochs=2, num_updates=10), T=T, Bl=B, model=NATURE)
    L.ctx.set_params(params)
    diag = []
    L.step_hook = pin_hook(record, diag, "ppo")
    tt = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(L.ctx.device)
    kt = agent.key_tensor(key, L.ctx.device)
    stats = L.update(tt(shard.obs), tt(shard.dones), tt(shard.actions), tt(shard.logprobs), tt(shard.values), tt(shard.rewards),
                     tt(shard.next_obs), tt(shard.next_done), kt)
    assert agent.key_numpy(kt).tolist() == okey.tolist() and len(diag) == 16
    serr = [abs(float(stats[i]) - ostats[i]) / max(abs(ostats[i]), 1e-6) for i in range(4)]
    _diag("nature_ppo_update_pinned", steps=diag, mean_stats_relerr=serr)
    assert max(serr) < 1e-4, (stats, ostats)
