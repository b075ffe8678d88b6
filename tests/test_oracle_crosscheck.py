"""Independent cross-checks of the oracle's restated third-party semantics (the reference ships no golden
vectors and JAX / optax / rlax cannot be installed here).  Each check compares the restatement with a DIFFERENT
implementation or a different derivation of the same published definition:

  * RMSProp "pytorch style" (cleanba_impala.py:152-188 says it mirrors torch.optim.RMSprop) vs torch.optim.RMSprop itself;
  * optax.adam's published update vs torch.optim.Adam (same formula, independent code);
  * flax nn.Conv SAME / HWIO cross-correlation vs an explicit numpy loop;
  * V-trace (rlax 0.1.5 recursion) vs the closed-form sum of the IMPALA paper (Espeholt et al. 2018, eq. 1);
  * GAE recursion (cleanba_ppo.py:532-560) vs the explicit sum of discounted TD errors of the GAE paper;
  * categorical log-prob / entropy (cleanba_ppo.py:516-530) vs torch.distributions.Categorical;
  * Gumbel-max sampling (cleanba_ppo.py:256-260) draws from softmax(logits) (chi-square test on 200k samples).
"""
import numpy as np
import torch

from oracle import impala, network as net, optim, ppo, threefry as tf


def test_rmsprop_matches_torch_rmsprop():
    rng = np.random.default_rng(0)
    p0 = rng.standard_normal(257).astype(np.float32)
    tp = torch.nn.Parameter(torch.tensor(p0.copy()))
    topt = torch.optim.RMSprop([tp], lr=6e-4, alpha=0.99, eps=0.01, momentum=0.0, centered=False)
    o = optim.RMSPropPyTorchStyle(p0.size, decay=0.99, eps=0.01)
    p = p0.copy()
    for _ in range(6):
        g = (rng.standard_normal(p0.size) * 10.0 ** rng.integers(-4, 1)).astype(np.float32)
        tp.grad = torch.tensor(g.copy())
        topt.step()
        p = o.step(p, g, 6e-4)
        np.testing.assert_allclose(p, tp.detach().numpy(), rtol=0, atol=2e-7)


def test_adam_matches_torch_adam():
    rng = np.random.default_rng(1)
    p0 = rng.standard_normal(301).astype(np.float32)
    tp = torch.nn.Parameter(torch.tensor(p0.copy()))
    topt = torch.optim.Adam([tp], lr=2.5e-4, betas=(0.9, 0.999), eps=1e-5)
    o = optim.Adam(p0.size, eps=1e-5)
    p = p0.copy()
    for _ in range(8):
        g = (rng.standard_normal(p0.size) * 10.0 ** rng.integers(-3, 1)).astype(np.float32)
        tp.grad = torch.tensor(g.copy())
        topt.step()
        p = o.step(p, g, 2.5e-4)
        # torch divides sqrt(v) by sqrt(bc2) and adds eps; optax adds eps to sqrt(v / bc2): equal up to fp32 rounding
        np.testing.assert_allclose(p, tp.detach().numpy(), rtol=0, atol=3e-7)


def test_conv_same_hwio_matches_explicit_loop():
    rng = np.random.default_rng(2)
    x = rng.standard_normal((2, 3, 6, 5))                       # NCHW
    w = rng.standard_normal((3, 3, 3, 4))                       # HWIO
    b = rng.standard_normal(4)
    got = net._conv(torch.tensor(x), torch.tensor(w), torch.tensor(b)).numpy()
    xp = np.pad(x, ((0, 0), (0, 0), (1, 1), (1, 1)))
    want = np.zeros((2, 4, 6, 5))
    for n in range(2):
        for o in range(4):
            for y in range(6):
                for xx in range(5):
                    acc = b[o]
                    for ky in range(3):
                        for kx in range(3):
                            for c in range(3):
                                acc += xp[n, c, y + ky, xx + kx] * w[ky, kx, c, o]
                    want[n, o, y, xx] = acc
    np.testing.assert_allclose(got, want, atol=1e-12)


def test_vtrace_matches_impala_paper_closed_form():
    """v_s = V(x_s) + sum_{t>=s} gamma^{t-s} (prod_{i=s}^{t-1} c_i) delta_t V, delta_t V = rho_t (r_t + gamma V(x_{t+1}) - V(x_t)),
    rho_t = min(rho_bar, pi/mu), c_i = lambda min(c_bar, pi/mu) with rho_bar = c_bar = lambda = 1 (cleanba_impala.py:585-587)."""
    rng = np.random.default_rng(3)
    T, B = 7, 4
    v = rng.standard_normal((T + 1, B))
    r = rng.standard_normal((T, B))
    disc = 0.99 * (rng.random((T, B)) > 0.2)                    # per-step discounts incl. terminations
    ratio = np.exp(rng.standard_normal((T, B)))                 # off-policy importance ratios on both sides of 1
    tt = lambda a: torch.tensor(a, dtype=torch.float64)
    errors, pg_adv, q = impala.vtrace_td_error_and_advantage(tt(v[:-1]), tt(v[1:]), tt(r), tt(disc), tt(ratio))
    rho = np.minimum(1.0, ratio)
    c = np.minimum(1.0, ratio)
    delta = rho * (r + disc * v[1:] - v[:-1])
    vs = np.zeros((T + 1, B))
    vs[T] = v[T]
    for s in range(T):
        acc = np.zeros(B)
        for t in range(s, T):
            w = np.ones(B)
            for i in range(s, t):
                w = w * disc[i] * c[i]
            acc += w * delta[t]
        vs[s] = v[s] + acc
    np.testing.assert_allclose(errors.numpy(), vs[:-1] - v[:-1], atol=1e-12)
    # policy-gradient advantage of the paper: rho_s (r_s + gamma v_{s+1} - V(x_s))
    np.testing.assert_allclose(pg_adv.numpy(), rho * (r + disc * vs[1:] - v[:-1]), atol=1e-12)


def test_gae_matches_sum_of_discounted_td_errors():
    """A_t = sum_{l>=0} (gamma lambda)^l delta_{t+l} with episode boundaries cutting the sum (Schulman et al. 2016, eq. 16),
    with the reference's done convention: dones[t] marks that obs[t] starts a new episode (cleanba_ppo.py:543-548)."""
    rng = np.random.default_rng(4)
    T, B = 9, 5
    rewards = rng.standard_normal((T, B)).astype(np.float32)
    values = rng.standard_normal((T, B)).astype(np.float32)
    dones = rng.random((T, B)) < 0.25
    next_value = rng.standard_normal(B).astype(np.float32)
    next_done = rng.random(B) < 0.25
    adv, ret = ppo.compute_gae(rewards, values, dones, next_value, next_done, 0.99, 0.95)
    v_ext = np.concatenate([values, next_value[None]], 0).astype(np.float64)
    d_ext = np.concatenate([dones, next_done[None]], 0)
    want = np.zeros((T, B))
    for t in range(T):
        for b in range(B):
            acc, w = 0.0, 1.0
            for l in range(t, T):
                nonterm = 0.0 if d_ext[l + 1, b] else 1.0
                delta = rewards[l, b] + 0.99 * v_ext[l + 1, b] * nonterm - v_ext[l, b]
                acc += w * delta
                w *= 0.99 * 0.95 * nonterm
                if w == 0.0:
                    break
            want[t, b] = acc
    np.testing.assert_allclose(adv, want, atol=2e-5)
    np.testing.assert_allclose(ret, want + values, atol=2e-5)


def test_logprob_entropy_match_torch_categorical():
    rng = np.random.default_rng(5)
    logits = torch.tensor(rng.standard_normal((64, 18)) * 3.0)
    actions = torch.tensor(rng.integers(0, 18, 64))
    dist = torch.distributions.Categorical(logits=logits)
    lp = torch.log_softmax(logits, -1).gather(1, actions[:, None]).squeeze(1)
    nl = logits - torch.logsumexp(logits, -1, keepdim=True)
    ent = -(nl * torch.softmax(nl, -1)).sum(-1)
    np.testing.assert_allclose(lp.numpy(), dist.log_prob(actions).numpy(), atol=1e-12)
    np.testing.assert_allclose(ent.numpy(), dist.entropy().numpy(), atol=1e-12)
    np.testing.assert_allclose(ppo.log_softmax(logits.numpy().astype(np.float32)), torch.log_softmax(logits.float(), -1).numpy(), atol=2e-6)


def test_gumbel_max_samples_from_softmax():
    logits = np.array([[2.0, 0.5, -1.0, 0.0, 1.0]], np.float32)
    n = 200_000
    u = tf.uniform(tf.PRNGKey(7), (n, 5))
    a = ppo.gumbel_argmax(np.repeat(logits, n, 0), u)
    counts = np.bincount(a, minlength=5).astype(np.float64)
    pr = np.exp(logits[0] - logits[0].max()); pr /= pr.sum()
    chi2 = float(((counts - n * pr) ** 2 / (n * pr)).sum())
    assert chi2 < 23.5, (chi2, counts / n, pr)       # chi-square(4 dof) 99.99th percentile
