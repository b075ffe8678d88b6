"""GPU parity: the CUDA hot path (through the C ABI, via cleanba_b200.agent) against the CPU oracle.

Bars (BASELINE.json north_star): integer outputs (actions, permutations, PRNG keys) bit-exact; floating point
(values, log-probs, losses) within 1e-4 relative.  Tolerances are written next to every assertion.  Diagnostics of
every comparison are also dumped to gpurun_out/parity_diag.json so a failing run can be read after the fact.
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import impala as oimpala
from oracle import network as net
from oracle import optim as ooptim
from oracle import ppo as oppo
from oracle import threefry as tf

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DIAG = {}


def _diag(name, **kw):
    path = os.path.join(ROOT, "gpurun_out", "parity_diag.json")
    if not DIAG and os.path.exists(path):
        try:
            DIAG.update(json.load(open(path)))
        except Exception:
            pass
    DIAG[name] = {k: (float(v) if isinstance(v, (int, float, np.floating, np.integer)) else v) for k, v in kw.items()}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "parity_diag.json"), "w") as f:
        json.dump(DIAG, f, indent=1, sort_keys=True)


def _relerr(a, b):
    """max |a-b| relative to the largest magnitude of the reference tensor (robust to zeros)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.fixture(scope="module")
def agent():
    from cleanba_b200 import agent as ag
    return ag


@pytest.fixture(scope="module")
def params():
    return net.init_params(1)


BACKENDS = [pytest.param(1, id="simt"), pytest.param(0, id="tcgen05")]


def _frames(rng, n):
    return rng.integers(0, 256, (n, 4, 84, 84), dtype=np.uint8)


# --------------------------------------------------------------------------------------------- PRNG / permutation
def test_split_and_permutation_bit_exact(agent):
    ctx = agent.Context("cuda:0", max_batch=4)
    key = tf.split(tf.PRNGKey(1), 4)[0]
    kt = agent.key_tensor(key, ctx.device)
    sub = ctx.split_key(kt)
    nk, sk = tf.split(key)
    assert agent.key_numpy(kt).tolist() == nk.tolist() and agent.key_numpy(sub).tolist() == sk.tolist()
    for n in (1, 2, 7, 512, 2048, 5120, 15360):
        kt = agent.key_tensor(key, ctx.device)
        got = ctx.permutation(kt, n).cpu().numpy()
        want = tf.permutation(key, n)
        assert np.array_equal(got, want), f"permutation mismatch at n={n}"
        assert agent.key_numpy(kt).tolist() == key.tolist()  # key is not consumed
    ctx.close()


# --------------------------------------------------------------------------------------------- GAE
@pytest.mark.parametrize("T,B,groups", [(128, 120, 4), (128, 16, 4), (20, 8, 2), (1, 4, 1), (300, 12, 4), (130, 40, 4)])
def test_gae_and_normalisation(agent, T, B, groups):
    ctx = agent.Context("cuda:0", max_batch=4)
    rng = np.random.default_rng(T * 1000 + B)
    r = rng.choice([-1.0, 0.0, 1.0], size=(T, B), p=[0.05, 0.9, 0.05]).astype(np.float32)
    v = (rng.standard_normal((T, B)) * 0.5).astype(np.float32)
    d = rng.random((T, B)) < 0.02
    nv = (rng.standard_normal(B) * 0.5).astype(np.float32)
    nd = rng.random(B) < 0.1
    dev = ctx.device
    tt = lambda x: torch.from_numpy(x).to(dev)
    adv, ret = ctx.gae(tt(r), tt(v), tt(d), tt(nv), tt(nd), 0.99, 0.95, 0)
    oadv, oret = oppo.compute_gae(r, v, d, nv, nd, 0.99, 0.95)
    # the scan replays the reference's operation order with un-fused fp32 ops: bit-exact
    assert np.array_equal(adv.cpu().numpy(), oadv), "raw GAE advantages are not bit-exact"
    assert np.array_equal(ret.cpu().numpy(), oret)
    adv_n, _ = ctx.gae(tt(r), tt(v), tt(d), tt(nv), tt(nd), 0.99, 0.95, groups)
    on = oppo.normalize_advantages(oadv, groups)
    err = _relerr(adv_n.cpu().numpy(), on)
    _diag(f"gae_norm_{T}x{B}", relerr=err)
    assert err < 1e-5  # reduction order differs from numpy's pairwise sum
    ctx.close()


# --------------------------------------------------------------------------------------------- forward
NAMES = {"y": "y", "p": "p", "b0": "b0"}


@pytest.mark.parametrize("backend", BACKENDS)
def test_forward_intermediates(agent, params, backend):
    rng = np.random.default_rng(11)
    n = 5
    obs = _frames(rng, n)
    ctx = agent.Context("cuda:0", max_batch=8, conv_backend=backend)
    ctx.set_params(params)
    logits, value = ctx.policy_value(torch.from_numpy(obs).to(ctx.device))
    torch.cuda.synchronize()
    rec = {}
    p = net.unflatten(torch.tensor(params))
    with torch.no_grad():
        hidden = net.trunk_forward(p, torch.from_numpy(obs), record=rec)
        ol, ov = net.heads(p, hidden)
    H = [84, 42, 21]
    Ho = [42, 21, 11]
    C = [16, 32, 32]
    errs = {}
    for s in range(3):
        for name, oname, hh in (("y", f"s{s}.y", H[s]), ("p", f"s{s}.p", Ho[s]), ("b0", f"s{s}.b0", Ho[s])):
            if name == "y" and backend == 0:
                continue   # the tcgen05 path fuses every sequence conv with its max-pool: s{s}.y is never materialised
            got = ctx.debug_tensor(f"s{s}.{name}", (n, hh, hh, C[s]))
            want = rec[oname].permute(0, 2, 3, 1).numpy()
            errs[f"s{s}.{name}"] = _relerr(got, want)
        got = ctx.debug_tensor(f"s{s}.a0", (n, Ho[s], Ho[s], C[s]))
        errs[f"s{s}.a0"] = _relerr(got, torch.relu(rec[f"s{s}.a0pre"]).permute(0, 2, 3, 1).numpy())
        got = ctx.debug_tensor(f"s{s}.a1", (n, Ho[s], Ho[s], C[s]))
        errs[f"s{s}.a1"] = _relerr(got, torch.relu(rec[f"s{s}.a1pre"]).permute(0, 2, 3, 1).numpy())
        want = rec[f"s{s}.b1"]
        if s == 2:
            want = torch.relu(want)
        errs[f"s{s}.out"] = _relerr(ctx.debug_tensor(f"s{s}.out", (n, Ho[s], Ho[s], C[s])), want.permute(0, 2, 3, 1).numpy())
    errs["hidden"] = _relerr(ctx.debug_tensor("hidden", (n, 256)), hidden.numpy())
    errs["logits"] = _relerr(logits.cpu().numpy(), ol.numpy())
    errs["value"] = _relerr(value.cpu().numpy(), ov.numpy())
    _diag(f"forward_backend{backend}", **errs)
    bad = {k: v for k, v in errs.items() if not v < 2e-5}
    assert not bad, f"forward mismatch (rel-to-max error, bar 2e-5): {bad}"
    ctx.close()


@pytest.mark.parametrize("backend", BACKENDS)
def test_forward_gather_and_batch_sizes(agent, params, backend):
    """idx gathers frames; ragged batch sizes (1, 3, 60, 129) exercise partial tiles."""
    rng = np.random.default_rng(12)
    pool = _frames(rng, 140)
    ctx = agent.Context("cuda:0", max_batch=140, conv_backend=backend)
    ctx.set_params(params)
    dev_pool = torch.from_numpy(pool).to(ctx.device)
    with torch.no_grad():
        ol, ov, _ = net.forward(params, pool)
    for n in (1, 3, 60, 129):
        idx = rng.permutation(140)[:n].astype(np.int32)
        logits, value = ctx.policy_value(dev_pool, torch.from_numpy(idx).to(ctx.device))
        e1 = _relerr(logits.cpu().numpy(), ol.numpy()[idx])
        e2 = _relerr(value.cpu().numpy(), ov.numpy()[idx])
        _diag(f"gather_backend{backend}_n{n}", logits=e1, value=e2)
        assert e1 < 1e-4 and e2 < 1e-4, (n, e1, e2)
    ctx.close()


# --------------------------------------------------------------------------------------------- actor
@pytest.mark.parametrize("backend", BACKENDS)
def test_actor_step_actions_bit_exact(agent, params, backend):
    rng = np.random.default_rng(13)
    N = 60
    ctx = agent.Context("cuda:0", max_batch=N, conv_backend=backend)
    ctx.set_params(params)
    key = tf.split(tf.PRNGKey(1), 4)[0]
    kt = agent.key_tensor(key, ctx.device)
    okey = key
    flips = 0
    near = 0
    for step in range(4):
        obs = _frames(rng, N)
        _, sk = tf.split(okey)          # the subkey this step will draw its uniforms from
        u = tf.uniform(sk, (N, 18))
        action, logprob, value, logits = ctx.actor_step(torch.from_numpy(obs).to(ctx.device), kt, True, True)
        _, oa, olp, ov, okey, ologits = oppo.get_action_and_value(params, obs, okey)
        assert agent.key_numpy(kt).tolist() == okey.tolist(), "PRNG key stream diverged"
        a = action.cpu().numpy()
        # (1) the sampling head in isolation: same logits in -> identical integer actions out
        assert np.array_equal(a, oppo.gumbel_argmax(logits.cpu().numpy(), u)), "sampling head is not bit-exact"
        # (2) end to end against the fp32 oracle: a flip is only tolerated where the oracle's own top-2 Gumbel gap is
        # below the measured logit error (none expected at these sizes)
        lerr = np.abs(logits.cpu().numpy() - ologits).max()
        with np.errstate(divide="ignore"):
            pert = ologits - np.log(-np.log(u))
        top2 = np.sort(pert, axis=1)[:, -2:]
        gap = top2[:, 1] - top2[:, 0]
        mism = a != oa
        flips += int(mism.sum())
        near += int((gap < 4 * lerr).sum())
        assert not (mism & (gap > 4 * lerr)).any(), "action differs where the oracle's decision is not a near-tie"
        assert _relerr(logprob.cpu().numpy(), olp) < 1e-4      # bar: 1e-4 relative
        assert _relerr(value.cpu().numpy(), ov) < 1e-4
        assert _relerr(logits.cpu().numpy(), ologits) < 1e-4
    _diag(f"actor_backend{backend}", flips=flips, near_ties=near)
    assert flips == 0, f"{flips} action flips on near-ties (near-tie count {near})"
    ctx.close()


# --------------------------------------------------------------------------------------------- PPO gradient
def _leafwise(g, og, leaves):
    out = {}
    for name, off, shape in leaves:
        n = int(np.prod(shape))
        a, b = g[off:off + n], og[off:off + n]
        out[name] = float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
    return out


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("mb", [8, 70])
def test_ppo_grad_matches_autograd(agent, params, backend, mb):
    from cleanba_b200 import lib
    rng = np.random.default_rng(14 + mb)
    N = 2 * mb
    obs = _frames(rng, N)
    actions = rng.integers(0, 18, N).astype(np.int32)
    oldlp = (np.log(1 / 18) + rng.standard_normal(N) * 0.05).astype(np.float32)
    adv = rng.standard_normal(N).astype(np.float32)
    ret = rng.standard_normal(N).astype(np.float32)
    idx = rng.permutation(N)[:mb].astype(np.int32)
    ctx = agent.Context("cuda:0", max_batch=mb, train=True, conv_backend=backend)
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
    lw = _leafwise(g, og, lib.leaves())
    tot = float(np.linalg.norm(g - og) / np.linalg.norm(og))
    serr = [abs(st[i] - ostats[i]) / max(abs(ostats[i]), 1e-6) for i in range(4)]
    _diag(f"ppo_grad_backend{backend}_mb{mb}", total=tot, stats_relerr=serr, kl=[float(st[4]), float(ostats[4])],
          worst_leaf=max(lw, key=lw.get), worst=max(lw.values()))
    assert max(serr) < 1e-4, (st, ostats)                     # losses: 1e-4 relative
    assert abs(st[4] - ostats[4]) < 1e-5
    # Gradient bar.  relu / max-pool gates are discontinuous, so gradient parity needs the forward to agree to ~1e-6:
    # activations are carried as exact 3-way bf16 splits (24 significant bits) for this reason (DESIGN.md "precision").
    # Typical agreement with the fp64 autograd oracle is 2e-6 .. 1.5e-5 (see gpurun_out/parity_diag.json); a single gate
    # whose pre-activation sits within 1e-6 of zero flips once in a while -- it does so between the fp32 and the fp64
    # CPU oracle too -- and moves the gradient by a few 1e-4, hence the looser assertion.
    assert tot < 1e-3, f"gradient relative error {tot}; per-leaf {lw}"
    assert max(lw.values()) < 1e-2, lw
    # determinism: a second identical call gives bit-identical gradients
    g2 = torch.zeros_like(grads)
    ctx.ppo_grad(tt(obs), tt(idx), mb, tt(actions), tt(oldlp), tt(adv), tt(ret), 0.1, 0.01, 0.5, g2, stats)
    assert torch.equal(grads, g2)
    ctx.close()


@pytest.mark.parametrize("backend", BACKENDS)
def test_impala_grad_matches_autograd(agent, params, backend):
    from cleanba_b200 import lib
    rng = np.random.default_rng(15)
    T1, Bl, B = 6, 8, 4
    obs = rng.integers(0, 256, (T1, Bl, 4, 84, 84), dtype=np.uint8)
    a = rng.integers(0, 18, (T1, Bl)).astype(np.int32)
    mu = (rng.standard_normal((T1, Bl, 18)) * 0.3).astype(np.float32)
    r = rng.choice([-1.0, 0.0, 1.0], size=(T1, Bl)).astype(np.float32)
    d = rng.random((T1, Bl)) < 0.15
    fs = rng.random((T1, Bl)) < 0.15
    cols = np.arange(4, 8)
    idx = (np.arange(T1)[:, None] * Bl + cols[None, :]).astype(np.int32).ravel()
    ctx = agent.Context("cuda:0", max_batch=T1 * B, algo=1, train=True, conv_backend=backend)
    ctx.set_params(params)
    dev = ctx.device
    tt = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    grads = torch.zeros(ctx.num_params, dtype=torch.float32, device=dev)
    stats = torch.zeros(4, dtype=torch.float32, device=dev)
    ctx.impala_grad(tt(obs.reshape(-1, 4, 84, 84)), tt(idx), T1, B, tt(a.ravel()), tt(mu.reshape(-1, 18)), tt(r.ravel()),
                    tt(d.ravel()), tt(fs.ravel()), 0.99, 0.5, 0.01, grads, stats)
    torch.cuda.synchronize()
    ostats, og = oimpala.impala_loss_and_grad(params, obs[:, cols], a[:, cols], mu[:, cols], r[:, cols], d[:, cols],
                                              fs[:, cols], dtype=torch.float64)
    g = grads.cpu().numpy().astype(np.float64)
    st = stats.cpu().numpy()
    lw = _leafwise(g, og, lib.leaves())
    tot = float(np.linalg.norm(g - og) / np.linalg.norm(og))
    serr = [abs(st[i] - ostats[i]) / max(abs(ostats[i]), 1e-6) for i in range(4)]
    _diag(f"impala_grad_backend{backend}", total=tot, stats_relerr=serr, worst_leaf=max(lw, key=lw.get), worst=max(lw.values()))
    assert max(serr) < 1e-4, (st, ostats)
    assert tot < 1e-3, f"gradient relative error {tot}; per-leaf {lw}"   # see test_ppo_grad_matches_autograd
    ctx.close()


# --------------------------------------------------------------------------------------------- optimizers
@pytest.mark.parametrize("algo", [0, 1])
def test_optimizer_step(agent, params, algo):
    rng = np.random.default_rng(16)
    ctx = agent.Context("cuda:0", max_batch=4, algo=algo, train=True)
    ctx.set_params(params)
    p = params.copy()
    opt = ooptim.Adam(p.size) if algo == 0 else ooptim.RMSPropPyTorchStyle(p.size)
    max_norm = 0.5 if algo == 0 else 40.0
    for step in range(3):
        g = (rng.standard_normal(p.size) * (1e-3 if step else 1.0)).astype(np.float32)  # step 0 clips, later ones do not
        lr = 2.5e-4 * (1 - step / 10)
        gt = torch.from_numpy(g).to(ctx.device)
        norm = torch.zeros(1, dtype=torch.float32, device=ctx.device)
        ctx.optimizer_step(gt * 2.0, 0.5, lr, max_norm, norm)   # grad_scale 1/L with L=2 replicas summed
        p = opt.step(p, ooptim.clip_by_global_norm(g, max_norm), lr)
        assert abs(norm.item() - ooptim.global_norm(g)) / ooptim.global_norm(g) < 1e-5
        err = float(np.abs(ctx.get_params().cpu().numpy() - p).max())
        _diag(f"opt{algo}_step{step}", maxabs=err)
        assert err < 2e-7, err          # one fp32 ulp of a ~1e-1 parameter after an lr ~1e-4 update
    ctx.close()


# --------------------------------------------------------------------------------------------- whole updates, step by step
from _pin import pin_hook as _pin_hook  # noqa: E402  (tests/_pin.py)


@pytest.mark.parametrize("backend", BACKENDS)
def test_ppo_update_every_step_pinned_to_oracle(agent, params, backend):
    """PPOLearner.update = single_device_update (cleanba_ppo.py:579-654): bootstrap value -> GAE -> normalisation -> 2 epochs x 4
    shuffled minibatches -> clip + Adam, with EVERY minibatch step started from the oracle's recorded state."""
    from cleanba_b200.learner import PPOLearner, PPOHyper
    rng = np.random.default_rng(17)
    T, B = 8, 8
    shard = oppo.Shard(obs=rng.integers(0, 256, (T, B, 4, 84, 84), dtype=np.uint8), dones=rng.random((T, B)) < 0.1,
                       actions=rng.integers(0, 18, (T, B)).astype(np.int32),
                       logprobs=(np.log(1 / 18) + rng.standard_normal((T, B)) * 0.01).astype(np.float32),
                       values=(rng.standard_normal((T, B)) * 0.1).astype(np.float32),
                       rewards=rng.choice([-1.0, 0.0, 1.0], size=(T, B)).astype(np.float32),
                       next_obs=rng.integers(0, 256, (B, 4, 84, 84), dtype=np.uint8), next_done=rng.random(B) < 0.1)
    key = tf.split(tf.PRNGKey(1), 4)[0]
    cfg = oppo.PPOConfig(update_epochs=2, num_updates=10)
    ol = oppo.PPOLearner(params, cfg)
    record = []
    ostats, okey = ol.update([shard], key, record=record)
    hyper = PPOHyper(update_epochs=2, num_updates=10)
    L = PPOLearner("cuda:0", hyper, T=T, Bl=B, conv_backend=backend)
    L.ctx.set_params(params)
    diag = []
    L.step_hook = _pin_hook(record, diag, "ppo")
    dev = L.ctx.device
    tt = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    kt = agent.key_tensor(key, dev)
    stats = L.update(tt(shard.obs), tt(shard.dones), tt(shard.actions), tt(shard.logprobs), tt(shard.values),
                     tt(shard.rewards), tt(shard.next_obs), tt(shard.next_done), kt)
    assert agent.key_numpy(kt).tolist() == okey.tolist()                  # same shuffles (bit-exact permutations)
    assert len(diag) == 16                                                # 8 steps x (grad, post)
    serr = [abs(float(stats[i]) - ostats[i]) / max(abs(ostats[i]), 1e-6) for i in range(4)]
    _diag(f"ppo_update_pinned_backend{backend}", steps=diag, mean_stats_relerr=serr)
    assert max(serr) < 1e-4, (stats, ostats)                              # the update's averaged scalars: 1e-4 relative
    assert abs(float(stats[4]) - ostats[4]) < 1e-5


@pytest.mark.parametrize("backend", BACKENDS)
def test_impala_update_every_step_pinned_to_oracle(agent, params, backend):
    """ImpalaLearner.update = single_device_update (cleanba_impala.py:599-639): 2 contiguous column-block minibatches -> V-trace
    loss -> clip + RMSProp, every step started from the oracle's recorded state."""
    from cleanba_b200.learner import ImpalaLearner, ImpalaHyper
    rng = np.random.default_rng(19)
    T1, Bl = 6, 8
    sh = oimpala.Shard(obs=rng.integers(0, 256, (T1, Bl, 4, 84, 84), dtype=np.uint8), dones=rng.random((T1, Bl)) < 0.15,
                       actions=rng.integers(0, 18, (T1, Bl)).astype(np.int32),
                       logitss=(rng.standard_normal((T1, Bl, 18)) * 0.3).astype(np.float32),
                       rewards=rng.choice([-1.0, 0.0, 1.0], size=(T1, Bl)).astype(np.float32), firststeps=rng.random((T1, Bl)) < 0.15)
    ol = oimpala.ImpalaLearner(params, oimpala.ImpalaConfig(num_minibatches=2, num_updates=10))
    record = []
    ostats = ol.update([sh], record=record)
    L = ImpalaLearner("cuda:0", ImpalaHyper(num_minibatches=2, num_updates=10), T1=T1, Bl=Bl, conv_backend=backend)
    L.ctx.set_params(params)
    diag = []
    L.step_hook = _pin_hook(record, diag, "impala")
    dev = L.ctx.device
    tt = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    stats = L.update(tt(sh.obs), tt(sh.dones), tt(sh.actions), tt(sh.logitss), tt(sh.rewards), tt(sh.firststeps))
    serr = [abs(float(stats[i]) - ostats[i]) / max(abs(ostats[i]), 1e-6) for i in range(4)]
    _diag(f"impala_update_pinned_backend{backend}", steps=diag, mean_stats_relerr=serr)
    assert len(diag) == 4 and max(serr) < 1e-4, (stats, ostats)


# --------------------------------------------------------------------------------------------- config shapes vs the oracle
@pytest.mark.slow
def test_ppo_grad_config2_shape_against_oracle(agent, params):
    """BASELINE config 2: one 3840-frame minibatch of cb_ppo_grad against the fp32 oracle (chunked autograd on the host cores,
    about a minute): loss scalars 1e-4 relative, gradient 1e-3."""
    from cleanba_b200 import lib
    rng = np.random.default_rng(41)
    mb = 3840
    obs = _frames(rng, mb)
    actions = rng.integers(0, 18, mb).astype(np.int32)
    oldlp = (np.log(1 / 18) + rng.standard_normal(mb) * 0.02).astype(np.float32)
    adv = rng.standard_normal(mb).astype(np.float32)
    ret = rng.standard_normal(mb).astype(np.float32)
    ctx = agent.Context("cuda:0", max_batch=mb, train=True)
    ctx.set_params(params)
    tt = lambda x: torch.from_numpy(x).to(ctx.device)
    grads = torch.zeros(ctx.num_params, dtype=torch.float32, device=ctx.device)
    stats = torch.zeros(5, dtype=torch.float32, device=ctx.device)
    ctx.ppo_grad(tt(obs), None, mb, tt(actions), tt(oldlp), tt(adv), tt(ret), 0.1, 0.01, 0.5, grads, stats)
    torch.cuda.synchronize()
    ostats, og = oppo.ppo_loss_and_grad_chunked(params, obs, actions, oldlp, adv, ret, chunk=256)
    g = grads.cpu().numpy().astype(np.float64)
    st = stats.cpu().numpy()
    lw = _leafwise(g, og, lib.leaves())
    tot = float(np.linalg.norm(g - og) / np.linalg.norm(og))
    serr = [abs(st[i] - ostats[i]) / max(abs(ostats[i]), 1e-6) for i in range(4)]
    _diag("ppo_grad_config2_mb3840_vs_oracle", total=tot, stats_relerr=serr, worst_leaf=max(lw, key=lw.get), worst=max(lw.values()))
    assert max(serr) < 1e-4, (st, ostats)
    assert tot < 1e-3, f"gradient relative error {tot}; per-leaf {lw}"
    ctx.close()


@pytest.mark.slow
def test_impala_grad_config3_shape_against_oracle(agent, params):
    """BASELINE config 3: one [T+1 = 21, B = 30] minibatch (630 frames) of cb_impala_grad against the fp32 oracle."""
    from cleanba_b200 import lib
    rng = np.random.default_rng(43)
    T1, Bl, B = 21, 120, 30
    obs = rng.integers(0, 256, (T1, Bl, 4, 84, 84), dtype=np.uint8)
    a = rng.integers(0, 18, (T1, Bl)).astype(np.int32)
    mu = (rng.standard_normal((T1, Bl, 18)) * 0.3).astype(np.float32)
    r = rng.choice([-1.0, 0.0, 1.0], size=(T1, Bl), p=[.05, .9, .05]).astype(np.float32)
    d = rng.random((T1, Bl)) < 0.02
    fs = rng.random((T1, Bl)) < 0.02
    cols = np.arange(60, 90)                             # the third minibatch's column block
    idx = (np.arange(T1)[:, None] * Bl + cols[None, :]).astype(np.int32).ravel()
    ctx = agent.Context("cuda:0", max_batch=T1 * B, algo=1, train=True)
    ctx.set_params(params)
    tt = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(ctx.device)
    grads = torch.zeros(ctx.num_params, dtype=torch.float32, device=ctx.device)
    stats = torch.zeros(4, dtype=torch.float32, device=ctx.device)
    ctx.impala_grad(tt(obs.reshape(-1, 4, 84, 84)), tt(idx), T1, B, tt(a.ravel()), tt(mu.reshape(-1, 18)), tt(r.ravel()),
                    tt(d.ravel()), tt(fs.ravel()), 0.99, 0.5, 0.01, grads, stats)
    torch.cuda.synchronize()
    ostats, og = oimpala.impala_loss_and_grad(params, obs[:, cols], a[:, cols], mu[:, cols], r[:, cols], d[:, cols], fs[:, cols])
    g = grads.cpu().numpy().astype(np.float64)
    st = stats.cpu().numpy()
    lw = _leafwise(g, og, lib.leaves())
    tot = float(np.linalg.norm(g - og) / np.linalg.norm(og))
    serr = [abs(st[i] - ostats[i]) / max(abs(ostats[i]), 1e-6) for i in range(4)]
    _diag("impala_grad_config3_21x30_vs_oracle", total=tot, stats_relerr=serr, worst_leaf=max(lw, key=lw.get), worst=max(lw.values()))
    assert max(serr) < 1e-4, (st, ostats)
    assert tot < 1e-3, f"gradient relative error {tot}; per-leaf {lw}"
    ctx.close()


# --------------------------------------------------------------------------------------------- parameter publish
def test_publish_params_equals_set_params(agent, params):
    """cb_publish_params (learner -> actor, cleanba_ppo.py:721-725; what bench.py and the Sebulba loop use after every update):
    an actor fed by publish_to() behaves bit-identically to one fed by set_params(get_params()) -- parameters, packed weight
    images (same logits) and the sampled step -- also after an optimizer step changed the learner's weights."""
    rng = np.random.default_rng(23)
    n = 16
    learner = agent.Context("cuda:0", max_batch=n, train=True)
    learner.set_params(params)
    a_pub = agent.Context("cuda:0", max_batch=n)
    a_set = agent.Context("cuda:0", max_batch=n)
    obs = torch.from_numpy(_frames(rng, n)).cuda()
    for rnd in range(2):
        if rnd == 1:      # move the learner's weights first: publish must carry the NEW master weights and re-pack them
            g = torch.from_numpy((rng.standard_normal(learner.num_params) * 1e-2).astype(np.float32)).cuda()
            learner.optimizer_step(g, 1.0, 2.5e-4, 0.5)
        learner.publish_to(a_pub)
        a_set.set_params(learner.get_params())
        torch.cuda.synchronize()
        assert torch.equal(a_pub.get_params(), learner.get_params())
        k1 = agent.key_tensor(np.array([3, 4], np.uint32), a_pub.device)
        k2 = agent.key_tensor(np.array([3, 4], np.uint32), a_set.device)
        o1 = a_pub.actor_step(obs, k1, True, True)
        o2 = a_set.actor_step(obs, k2, True, True)
        for x, y in zip(o1, o2):
            assert torch.equal(x, y), "publish_to and set_params actors differ"
        assert agent.key_numpy(k1).tolist() == agent.key_numpy(k2).tolist()
        if rnd == 1:
            assert not torch.equal(o1[3], first_logits), "the published weights did not change the actor's logits"
        first_logits = o1[3].clone()
    for c in (learner, a_pub, a_set):
        c.close()


# --------------------------------------------------------------------------------------------- golden fixtures
def test_cuda_against_committed_golden_fixtures(agent):
    """tests/golden/oracle_golden.npz (made by tests/golden/make_golden.py from the oracle) against the CUDA path."""
    G = np.load(os.path.join(ROOT, "tests", "golden", "oracle_golden.npz"))
    ctx = agent.Context("cuda:0", max_batch=8, train=True)
    ctx.set_params(net.init_params(1))
    dev = ctx.device
    tt = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    kt = agent.key_tensor(G["key"], dev)
    action, logprob, value, logits = ctx.actor_step(tt(G["obs"]), kt, True, True)
    assert np.array_equal(action.cpu().numpy(), G["action"]) and agent.key_numpy(kt).tolist() == G["key_after"].tolist()
    assert _relerr(logits.cpu().numpy(), G["logits"]) < 1e-4 and _relerr(value.cpu().numpy(), G["value"]) < 1e-4
    assert np.array_equal(ctx.permutation(agent.key_tensor(G["key"], dev), 2048).cpu().numpy()[:64], G["perm_2048_head"])
    adv, ret = ctx.gae(tt(G["gae_r"]), tt(G["gae_v"]), tt(G["gae_d"]), tt(G["gae_nv"]), tt(G["gae_nd"]), 0.99, 0.95, 0)
    assert np.array_equal(adv.cpu().numpy(), G["gae_adv"]) and np.array_equal(ret.cpu().numpy(), G["gae_ret"])
    grads = torch.zeros(ctx.num_params, device=dev)
    stats = torch.zeros(5, device=dev)
    ctx.ppo_grad(tt(G["obs"]), None, 4, tt(G["ppo_actions"]), tt(G["logprob"]), tt(G["ppo_adv"]), tt(G["ppo_ret"]), 0.1, 0.01, 0.5, grads, stats)
    st = stats.cpu().numpy()
    assert max(abs(st[i] - G["ppo_stats"][i]) / max(abs(G["ppo_stats"][i]), 1e-6) for i in range(4)) < 1e-4
    assert abs(float(grads.norm()) - float(G["ppo_grad_norm"])) / float(G["ppo_grad_norm"]) < 1e-3
    ctx.close()


# --------------------------------------------------------------------------------------------- full-size properties
def test_full_size_minibatch_properties(agent, params):
    """BASELINE config 2 sizes (minibatch 3840 of a 15360-sample update): size-independent properties instead of the
    (minutes-long) CPU oracle: determinism, agreement of the tcgen05 path with the fp32 CUDA-core path, invariance of
    the minibatch loss under a re-ordering of the minibatch, idx-gather == explicit gather."""
    rng = np.random.default_rng(31)
    N, mb = 4096, 3840
    obs = torch.from_numpy(rng.integers(0, 256, (N, 4, 84, 84), dtype=np.uint8)).cuda()
    actions = torch.from_numpy(rng.integers(0, 18, N).astype(np.int32)).cuda()
    oldlp = torch.from_numpy((np.log(1 / 18) + rng.standard_normal(N) * 0.02).astype(np.float32)).cuda()
    adv = torch.randn(N, device="cuda"); ret = torch.randn(N, device="cuda")
    idx = torch.from_numpy(rng.permutation(N)[:mb].astype(np.int32)).cuda()
    out = {}
    for backend in (0, 1):
        ctx = agent.Context("cuda:0", max_batch=mb, train=True, conv_backend=backend)
        ctx.set_params(params)
        g = torch.zeros(ctx.num_params, device="cuda"); s = torch.zeros(5, device="cuda")
        ctx.ppo_grad(obs, idx, mb, actions, oldlp, adv, ret, 0.1, 0.01, 0.5, g, s)
        out[backend] = (g.clone(), s.clone())
        if backend == 0:
            g2 = torch.zeros_like(g); s2 = torch.zeros_like(s)
            ctx.ppo_grad(obs, idx, mb, actions, oldlp, adv, ret, 0.1, 0.01, 0.5, g2, s2)
            assert torch.equal(g, g2) and torch.equal(s, s2), "not deterministic at full size"
            perm = torch.randperm(mb, device="cuda")
            ctx.ppo_grad(obs, idx[perm].contiguous(), mb, actions, oldlp, adv, ret, 0.1, 0.01, 0.5, g2, s2)
            assert _relerr(s2.cpu().numpy()[:4], s.cpu().numpy()[:4]) < 1e-5
            assert float((g2 - g).norm() / g.norm()) < 1e-4
            ii = idx.long()
            ctx.ppo_grad(obs[ii].contiguous(), None, mb, actions[ii].contiguous(), oldlp[ii].contiguous(), adv[ii].contiguous(),
                         ret[ii].contiguous(), 0.1, 0.01, 0.5, g2, s2)
            assert torch.equal(g, g2) and torch.equal(s, s2), "idx gather differs from an explicit gather"
        ctx.close()
    serr = _relerr(out[0][1].cpu().numpy()[:4], out[1][1].cpu().numpy()[:4])
    gerr = float((out[0][0] - out[1][0]).norm() / out[1][0].norm())
    _diag("full_size_mb3840", stats_tcgen05_vs_simt=serr, grad_tcgen05_vs_simt=gerr)
    assert serr < 1e-5 and gerr < 1e-3


def test_rollout_actor_rows_match_plain_actor_steps(agent, params):
    """cb_actor_step_cursor through agent.RolloutActor (one CUDA graph, outputs written straight into rows of a column-sliced
    [T, Bl, ...] rollout storage) is bit-identical to consecutive cb_actor_step calls: actions, log-probs, values, PRNG key,
    and the frames land in their storage rows; a second rollout re-points the cursor (first_row = 1, the IMPALA carry row)."""
    rng = np.random.default_rng(21)
    n, T, Bl = 6, 4, 16
    dev = torch.device("cuda:0")
    frames = [torch.from_numpy(_frames(rng, n)).pin_memory() for _ in range(2 * T)]
    ref = agent.Context(dev, max_batch=n); ref.set_params(params)
    kr = agent.key_tensor(np.array([11, 12], np.uint32), dev)
    want = [tuple(x.cpu().numpy() for x in ref.actor_step(f.to(dev), kr)[:3]) for f in frames[:2 * T - 1]]
    ctx = agent.Context(dev, max_batch=n); ctx.set_params(params)
    key = agent.key_tensor(np.array([11, 12], np.uint32), dev)
    ra = agent.RolloutActor(ctx, n, key)
    obs = torch.zeros(T, Bl, 4, 84, 84, dtype=torch.uint8, device=dev)
    act = torch.full((T, Bl), -1, dtype=torch.int32, device=dev)
    lp = torch.zeros(T, Bl, device=dev); val = torch.zeros(T, Bl, device=dev)
    c = slice(5, 5 + n)                                  # a column block in the middle of the storage
    k = 0
    for first_row in (0, 1):                             # second rollout: row 0 is a carried row, steps start at row 1
        ra.stream.wait_stream(torch.cuda.current_stream(dev))
        ra.begin(obs[:, c], act[:, c], lp[:, c], val[:, c], first_row=first_row)
        for t in range(first_row, T):
            ra.step(frames[k], t)
            ra.stream.synchronize()
            a, l, v = want[k]
            assert np.array_equal(act[t, c].cpu().numpy(), a), (first_row, t)
            assert np.array_equal(lp[t, c].cpu().numpy(), l) and np.array_equal(val[t, c].cpu().numpy(), v)
            assert torch.equal(obs[t, c].cpu(), frames[k])
            k += 1
    assert (act[:, :5] == -1).all() and (act[:, 5 + n:] == -1).all(), "wrote outside its storage columns"
    assert agent.key_numpy(key).tolist() == agent.key_numpy(kr).tolist()
    with pytest.raises(agent.CleanbaError):
        ra.step(frames[0], 0)                            # rows must be stepped in order
    ref.close(); ctx.close()


# --------------------------------------------------------------------------------------------- gradient accumulation
def test_gradient_accumulation_matches_optax_multisteps(agent, params):
    """gradient_accumulation_steps = 2 (optax.MultiSteps(every_k_schedule), cleanba_ppo.py:78,492-500,607): the shuffled batch is cut into
    num_minibatches * k mini-steps, the optimizer steps on the running mean of k mini-step gradients.  Every mini-step starts from
    the oracle's recorded parameters; the accumulated gradient is compared with the oracle's (1e-3), the parameters after each
    optimizer step with the oracle's optimizer applied to the replica's own accumulated gradient (2e-7)."""
    from cleanba_b200.learner import PPOHyper, PPOLearner
    rng = np.random.default_rng(51)
    T, B = 4, 8
    shard = oppo.Shard(obs=rng.integers(0, 256, (T, B, 4, 84, 84), dtype=np.uint8), dones=rng.random((T, B)) < 0.1,
                       actions=rng.integers(0, 18, (T, B)).astype(np.int32),
                       logprobs=(np.log(1 / 18) + rng.standard_normal((T, B)) * 0.01).astype(np.float32),
                       values=(rng.standard_normal((T, B)) * 0.1).astype(np.float32),
                       rewards=rng.choice([-1.0, 0.0, 1.0], size=(T, B)).astype(np.float32),
                       next_obs=rng.integers(0, 256, (B, 4, 84, 84), dtype=np.uint8), next_done=rng.random(B) < 0.1)
    key = tf.split(tf.PRNGKey(1), 4)[0]
    ol = oppo.PPOLearner(params, oppo.PPOConfig(update_epochs=1, num_minibatches=2, gradient_accumulation_steps=2, num_updates=10))
    rec = []
    ostats, okey = ol.update([shard], key, record=rec)
    assert [r.get("mini_step") for r in rec] == [0, None, 0, None]      # (mini-step, optimizer step) x 2
    L = PPOLearner("cuda:0", PPOHyper(update_epochs=1, num_minibatches=2, gradient_accumulation_steps=2, num_updates=10), T=T, Bl=B)
    L.ctx.set_params(params)
    assert L.mb == T * B // 4
    seen = []

    def hook(phase, k, LL):
        step = rec[k | 1]                                   # the optimizer step this mini-step belongs to
        if phase == "pre":
            LL.ctx.set_params(step["params_before"])
            LL.ctx.set_opt_state(step["m_before"], step["v_before"], step["count_before"])
            LL.opt_count = step["count_before"]
        elif phase == "grad" and "mini_step" in rec[k]:
            g = LL.grads.cpu().numpy().astype(np.float64)
            err = float(np.linalg.norm(g - rec[k]["raw_grad"]) / np.linalg.norm(rec[k]["raw_grad"]))
            seen.append(("mini", k, err))
            assert err < 1e-3, (k, err)
        elif phase == "post":
            acc = LL.acc.cpu().numpy()
            err = float(np.linalg.norm(acc.astype(np.float64) - step["raw_grad"]) / np.linalg.norm(step["raw_grad"]))
            opt = ooptim.Adam(acc.size)
            opt.m, opt.v, opt.count = step["m_before"].copy(), step["v_before"].copy(), step["count_before"]
            want = opt.step(step["params_before"], ooptim.clip_by_global_norm(acc, 0.5), step["lr"])
            perr = float(np.abs(LL.ctx.get_params().cpu().numpy() - want).max())
            seen.append(("step", k, err, perr))
            assert err < 1e-3 and perr < 2e-7, (k, err, perr)

    L.step_hook = hook
    dev = L.ctx.device
    tt = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    kt = agent.key_tensor(key, dev)
    stats = L.update(tt(shard.obs), tt(shard.dones), tt(shard.actions), tt(shard.logprobs), tt(shard.values), tt(shard.rewards),
                     tt(shard.next_obs), tt(shard.next_done), kt)
    assert agent.key_numpy(kt).tolist() == okey.tolist() and L.opt_count == 2
    assert [s[0] for s in seen] == ["mini", "step", "mini", "step"], seen
    serr = [abs(float(stats[i]) - ostats[i]) / max(abs(ostats[i]), 1e-2 * np.abs(ostats[:4]).max()) for i in range(4)]
    _diag("ppo_gradient_accumulation_k2", seen=[list(map(float, s[1:])) for s in seen], mean_stats_relerr=serr)
    assert max(serr) < 1e-4, (stats, ostats)


# --------------------------------------------------------------------------------------------- exchange / hand-off primitives
def test_reduce_peers_copy_columns_and_milestone(agent, params):
    """The small ABI pieces around the hot path: cb_reduce_peers (fixed-order sum of replica buffers), cb_memcpy_2d through
    agent.copy_columns (column block of a [T, N, ...] storage, no contiguous temporary) and cb_set_grad_milestone (the event fires
    inside the gradient call and the tail offset is the dense layer's first leaf)."""
    from cleanba_b200 import lib
    ctx = agent.Context("cuda:0", max_batch=8, train=True)
    ctx.set_params(params)
    dev = ctx.device
    g = [torch.randn(ctx.num_params, device=dev) for _ in range(3)]
    out = torch.empty(ctx.num_params, device=dev)
    ctx.reduce_peers(g, out)
    assert torch.equal(out, (g[0] + g[1]) + g[2])                      # the kernel's order: ((g0 + g1) + g2)
    T, N = 5, 12
    storage = torch.randint(0, 255, (T, N, 4, 84, 84), dtype=torch.uint8, device=dev)
    vals = torch.randn(T, N, device=dev)
    st = torch.cuda.Stream(dev)
    st.wait_stream(torch.cuda.current_stream(dev))
    for c in (slice(0, 4), slice(4, 12)):
        d1 = torch.empty((T, c.stop - c.start, 4, 84, 84), dtype=torch.uint8, device=dev)
        d2 = torch.empty((T, c.stop - c.start), device=dev)
        agent.copy_columns(d1, storage[:, c], st)
        agent.copy_columns(d2, vals[:, c], st)
        st.synchronize()
        assert torch.equal(d1, storage[:, c]) and torch.equal(d2, vals[:, c])
    ev = torch.cuda.Event()
    ev.record()
    tail = ctx.set_grad_milestone(ev)
    assert tail == dict((n, o) for n, o, _ in lib.leaves())["network_params/params/Dense_0/bias"]
    rng = np.random.default_rng(3)
    obs = torch.from_numpy(_frames(rng, 8)).to(dev)
    grads = torch.zeros(ctx.num_params, device=dev); stats = torch.zeros(5, device=dev)
    z = torch.zeros(8, device=dev)
    ctx.ppo_grad(obs, None, 8, torch.zeros(8, dtype=torch.int32, device=dev), z, z + 1, z, 0.1, 0.01, 0.5, grads, stats)
    ev.synchronize()                                                    # recorded by the call: completes without a device sync
    torch.cuda.synchronize()
    assert float(grads[tail:].abs().sum()) > 0
    ctx.set_grad_milestone(None)
    ctx.close()


@pytest.mark.slow
def test_forward_config4_minibatch_size(agent, params):
    """mb = 1280 (BASELINE config 4: 3 learner GPUs, 5,120 samples each) and its neighbours sit in the 4-way split-K range of the
    dense forward (1024 < n <= 1920): regression test for the partial-sum buffer, which round 1 sized for n <= 1024."""
    rng = np.random.default_rng(61)
    n = 1300
    obs = _frames(rng, n)
    ctx = agent.Context("cuda:0", max_batch=n, train=True)
    ctx.set_params(params)
    guard = torch.full((1 << 20,), 7.0, device=ctx.device)        # whatever the allocator placed nearby must stay intact
    logits, value = ctx.policy_value(torch.from_numpy(obs).to(ctx.device))
    torch.cuda.synchronize()
    with torch.no_grad():
        ol, ov, _ = net.forward(params, obs)
    e1, e2 = _relerr(logits.cpu().numpy(), ol.numpy()), _relerr(value.cpu().numpy(), ov.numpy())
    _diag("forward_n1300", logits=e1, value=e2)
    assert e1 < 1e-4 and e2 < 1e-4 and bool((guard == 7.0).all())
    grads = torch.zeros(ctx.num_params, device=ctx.device); stats = torch.zeros(5, device=ctx.device)
    z = torch.zeros(n, device=ctx.device)
    ctx.ppo_grad(torch.from_numpy(obs).to(ctx.device), None, 1280, torch.zeros(n, dtype=torch.int32, device=ctx.device), z, z + 1, z, 0.1, 0.01, 0.5, grads, stats)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(grads).all()) and bool(torch.isfinite(stats).all())
    ctx.close()


@pytest.mark.parametrize("model", [0, 1])
def test_graphed_gradient_step_equals_plain_launches(agent, params, model):
    """cb_graph_steps: the gradient step replayed as a captured CUDA graph (pointers read through the device-side table) gives
    BIT-identical gradients and loss scalars to the plain launches, with different obs / idx / field / stats pointers on every
    call, for both algorithms and both trunks; an external milestone event is recorded by every replay."""
    from oracle import network as onet
    p = params if model == 0 else onet.init_params(3, onet.nature_param_spec())
    rng = np.random.default_rng(77)
    mb, N = 24, 96
    dev = torch.device("cuda:0")
    tt = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    plain = agent.Context("cuda:0", max_batch=mb, train=True, model=model)
    graphed = agent.Context("cuda:0", max_batch=mb, train=True, model=model)
    ev = torch.cuda.Event(); ev.record()
    for c in (plain, graphed):
        c.set_params(p)
    graphed.graph_steps(True)
    graphed.set_grad_milestone(ev)
    gp, gg = (torch.zeros(plain.num_params, device=dev) for _ in range(2))
    sp, sg = torch.zeros(6, 5, device=dev), torch.zeros(6, 5, device=dev)
    l0 = None
    for k in range(6):
        obs = tt(_frames(rng, N)); idx = tt(rng.permutation(N)[:mb].astype(np.int32))
        f = [tt(rng.integers(0, 18, N).astype(np.int32)), tt((np.log(1 / 18) + rng.standard_normal(N) * 0.05).astype(np.float32)),
             tt(rng.standard_normal(N).astype(np.float32)), tt(rng.standard_normal(N).astype(np.float32))]
        plain.ppo_grad(obs, idx, mb, *f, 0.1, 0.01, 0.5, gp, sp[k])
        graphed.ppo_grad(obs, idx, mb, *f, 0.1, 0.01, 0.5, gg, sg[k])
        torch.cuda.synchronize()
        assert ev.query()
        assert torch.equal(gp, gg), (k, float((gp - gg).abs().max()))
        assert torch.equal(sp[k], sg[k])
    assert graphed.graph_replays == 5 and plain.graph_replays == 0      # call 0 ran eagerly, call 1 captured + replayed
    plain.close(); graphed.close()
    # IMPALA head: the [T1, B] column block moves between calls
    T1, Bl, B = 5, 8, 4
    plain = agent.Context("cuda:0", max_batch=T1 * B, algo=1, train=True, model=model)
    graphed = agent.Context("cuda:0", max_batch=T1 * B, algo=1, train=True, model=model)
    for c in (plain, graphed):
        c.set_params(p)
    graphed.graph_steps(True)
    sp, sg = torch.zeros(4, 4, device=dev), torch.zeros(4, 4, device=dev)
    for k in range(4):
        obs = tt(rng.integers(0, 256, (T1 * Bl, 4, 84, 84), dtype=np.uint8))
        cols = np.arange(4) + 4 * (k % 2)
        idx = tt((np.arange(T1)[:, None] * Bl + cols[None, :]).astype(np.int32).ravel())
        f = [tt(rng.integers(0, 18, T1 * Bl).astype(np.int32)), tt((rng.standard_normal((T1 * Bl, 18)) * 0.3).astype(np.float32)),
             tt(rng.choice([-1.0, 0.0, 1.0], size=T1 * Bl).astype(np.float32)), tt(rng.random(T1 * Bl) < 0.15), tt(rng.random(T1 * Bl) < 0.15)]
        plain.impala_grad(obs, idx, T1, B, *f, 0.99, 0.5, 0.01, gp, sp[k])
        graphed.impala_grad(obs, idx, T1, B, *f, 0.99, 0.5, 0.01, gg, sg[k])
        torch.cuda.synchronize()
        assert torch.equal(gp, gg), (k, float((gp - gg).abs().max()))
        assert torch.equal(sp[k], sg[k])
    assert graphed.graph_replays == 3
    # profiling falls back to plain launches and still reports kernels
    graphed.profile(True)
    graphed.impala_grad(obs, idx, T1, B, *f, 0.99, 0.5, 0.01, gg, sg[0])
    assert len(graphed.profile_report()) > 3 and graphed.graph_replays == 3
    graphed.profile(False)
    plain.close(); graphed.close()


@pytest.mark.parametrize("cluster", [1, 2])
@pytest.mark.parametrize("n", [1, 7, 60, 128])
def test_actor_persistent_tail_equals_per_layer_kernels(agent, params, n, cluster):
    """cb_set_actor_tail: ConvSequence 1 and 2 as ONE persistent kernel, one thread-block cluster (1 or 2 CTAs) per frame
    (actor_fused.cu), against the same ten layers as ten launches.  Same packed weights, same MMAs, same epilogues: logits and
    values must be BIT-identical, for a single frame, a ragged batch, the rollout batch and the largest fused batch."""
    rng = np.random.default_rng(300 + n)
    obs = torch.from_numpy(_frames(rng, n)).cuda()
    actor = agent.Context("cuda:0", max_batch=n, train=False)
    learner = agent.Context("cuda:0", max_batch=n, train=False)
    actor.set_actor_tail(cluster)
    for c in (actor, learner):
        c.set_params(params)
    for rep in range(3):                                        # repeated calls: barrier re-initialisation, buffer reuse
        la, va = actor.policy_value(obs)
        ll, vl = learner.policy_value(obs)
        torch.cuda.synchronize()
        assert torch.equal(la, ll), (rep, float((la - ll).abs().max()))
        assert torch.equal(va, vl)
        obs = torch.from_numpy(_frames(rng, n)).cuda()
    actor.close(); learner.close()


@pytest.mark.parametrize("graph", [False, True])
def test_shared_border_layout_survives_batch_size_changes(agent, params, graph):
    """Shared-border plane layout (csrc/common.cuh): the zero row after the last image of a batch of n is the first row of
    image n of a larger batch, so a context that alternates batch sizes must clear it (ctx.cu clear_trailing_rows) -- also
    when the larger batch ran as a graph replay.  Small batches computed in a context that keeps running larger ones must be
    BIT-identical to the same batches in a fresh context: logits / values and the full PPO gradient."""
    rng = np.random.default_rng(91)
    dev = torch.device("cuda:0")
    tt = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    big, small = 40, 7
    used = agent.Context("cuda:0", max_batch=big, train=True)
    used.set_params(params)
    if graph:
        used.graph_steps(True)

    def batch(n):
        return dict(obs=tt(_frames(rng, n)), idx=tt(rng.permutation(n).astype(np.int32)), act=tt(rng.integers(0, 18, n).astype(np.int32)),
                    lp=tt(np.full(n, np.log(1 / 18), np.float32)), adv=tt(rng.standard_normal(n).astype(np.float32)),
                    ret=tt(rng.standard_normal(n).astype(np.float32)))

    def grad(c, b, n):
        g = torch.zeros(c.num_params, device=dev); s = torch.zeros(5, device=dev)
        c.ppo_grad(b["obs"], b["idx"], n, b["act"], b["lp"], b["adv"], b["ret"], 0.1, 0.01, 0.5, g, s)
        return g, s

    gbuf = torch.zeros(used.num_params, device=dev); sbuf = torch.zeros(5, device=dev)
    for rep in range(3):
        for _ in range(3):                                  # larger batches fill every plane (the third one replays a graph)
            b = batch(big)
            used.ppo_grad(b["obs"], b["idx"], big, b["act"], b["lp"], b["adv"], b["ret"], 0.1, 0.01, 0.5, gbuf, sbuf)
        b = batch(small)
        fresh = agent.Context("cuda:0", max_batch=small, train=True)
        fresh.set_params(params)
        lu, vu = used.policy_value(b["obs"]); lf, vf = fresh.policy_value(b["obs"])
        assert torch.equal(lu, lf) and torch.equal(vu, vf), rep
        gu, su = grad(used, b, small); gf, sf = grad(fresh, b, small)
        assert torch.equal(gu, gf), (rep, float((gu - gf).abs().max()))
        assert torch.equal(su, sf)
        fresh.close()
    if graph:
        assert used.graph_replays >= 4
    used.close()
