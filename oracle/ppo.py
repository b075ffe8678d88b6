"""PPO hot path of cleanba/cleanba_ppo.py restated on CPU (numpy + torch autograd).

actor  : get_action_and_value           cleanba_ppo.py:245-261
learner: compute_gae / compute_gae_once cleanba_ppo.py:532-560
         advantage normalisation        cleanba_ppo.py:592-595
         update_epoch shuffle           cleanba_ppo.py:597-615
         get_logprob_entropy_value      cleanba_ppo.py:516-530
         ppo_loss                       cleanba_ppo.py:562-577
         update_minibatch / single_device_update / pmap+pmean   cleanba_ppo.py:579-660
PARITY UNPINNED vs JAX (see oracle/__init__.py); tolerances vs this oracle are stated in tests/.
"""
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import network as net
from . import optim, threefry

F32 = np.float32


# ----------------------------------------------------------------------------- actor
def gumbel_argmax(logits: np.ndarray, u: np.ndarray) -> np.ndarray:
    """action = argmax(logits - log(-log(u)), axis=1), first index on ties, int32 (cleanba_ppo.py:258)."""
    with np.errstate(divide="ignore"):
        g = logits.astype(F32) - np.log(-np.log(u.astype(F32)))
    return np.argmax(g, axis=1).astype(np.int32)


def log_softmax(logits: np.ndarray) -> np.ndarray:
    x = logits.astype(F32)
    m = x.max(axis=-1, keepdims=True)
    s = x - m
    return s - np.log(np.exp(s).sum(axis=-1, keepdims=True, dtype=F32))


def get_action_and_value(flat_params: np.ndarray, next_obs: np.ndarray, key: np.ndarray):
    """cleanba_ppo.py:245-261 -> (obs, action i32[N], logprob f32[N], value f32[N], new_key, logits)."""
    with torch.no_grad():
        logits, value, _ = net.forward(flat_params, next_obs)
    logits = logits.numpy()
    key, subkey = threefry.split(key)
    u = threefry.uniform(subkey, logits.shape)
    action = gumbel_argmax(logits, u)
    logprob = log_softmax(logits)[np.arange(action.shape[0]), action]
    return next_obs, action, logprob.astype(F32), value.numpy().astype(F32), key, logits


# ----------------------------------------------------------------------------- GAE
def compute_gae(rewards, values, dones, next_value, next_done, gamma=0.99, gae_lambda=0.95):
    """cleanba_ppo.py:532-560.  [T,B] inputs; reverse scan with separate fp32 mul/add (no FMA), in the
    operation order of compute_gae_once:
        delta = reward + gamma*nextvalues*nextnonterminal - curvalues
        adv   = delta + gamma*gae_lambda*nextnonterminal*adv
    """
    T, B = rewards.shape
    r = rewards.astype(F32)
    v = np.concatenate([values.astype(F32), next_value.astype(F32)[None]], 0)
    d = np.concatenate([dones.astype(F32), next_done.astype(F32)[None]], 0)
    g, gl = F32(gamma), F32(gamma * gae_lambda)  # python-float product, then cast (cleanba_ppo.py:538)
    adv = np.zeros(B, F32)
    out = np.zeros((T, B), F32)
    for t in range(T - 1, -1, -1):
        nn = F32(1.0) - d[t + 1]
        delta = r[t] + g * v[t + 1] * nn - v[t]
        adv = delta + gl * nn * adv
        out[t] = adv
    return out, out + values.astype(F32)


def normalize_advantages(adv: np.ndarray, num_minibatches: int) -> np.ndarray:
    """cleanba_ppo.py:592-595: per contiguous column group, population std, eps outside."""
    T, B = adv.shape
    a = adv.astype(F32).reshape(T, num_minibatches, -1)
    mean = a.mean((0, -1), keepdims=True, dtype=F32)
    std = a.std((0, -1), keepdims=True, dtype=F32)
    return ((a - mean) / (std + F32(1e-8))).reshape(T, B).astype(F32)


# ----------------------------------------------------------------------------- loss
def logprob_entropy_from_logits(logits: torch.Tensor, actions: torch.Tensor):
    """The part of get_logprob_entropy_value after the network (cleanba_ppo.py:524-528)."""
    logp_all = torch.log_softmax(logits, dim=-1)
    logprob = logp_all.gather(1, actions.long()[:, None]).squeeze(1)
    nl = logits - torch.logsumexp(logits, dim=-1, keepdim=True)
    nl = nl.clamp(min=torch.finfo(nl.dtype).min)
    p_log_p = nl * torch.softmax(nl, dim=-1)
    entropy = -p_log_p.sum(-1)
    return logprob, entropy


def logprob_entropy_value(p, obs_u8: torch.Tensor, actions: torch.Tensor):
    """get_logprob_entropy_value (cleanba_ppo.py:516-530)."""
    hidden = net.trunk_forward(p, obs_u8)
    logits, value = net.heads(p, hidden)
    logprob, entropy = logprob_entropy_from_logits(logits, actions)
    return logprob, entropy, value


def ppo_loss_from_heads(newlogprob, entropy, newvalue, behavior_logprobs, advantages, target_values,
                        clip_coef=0.1, ent_coef=0.01, vf_coef=0.5):
    """ppo_loss after the network (cleanba_ppo.py:564-577)."""
    logratio = newlogprob - behavior_logprobs
    ratio = torch.exp(logratio)
    approx_kl = ((ratio - 1) - logratio).mean().detach()
    pg_loss1 = -advantages * ratio
    pg_loss2 = -advantages * torch.clamp(ratio, 1 - clip_coef, 1 + clip_coef)
    pg_loss = torch.maximum(pg_loss1, pg_loss2).mean()
    v_loss = 0.5 * ((newvalue - target_values) ** 2).mean()
    entropy_loss = entropy.mean()
    loss = pg_loss - ent_coef * entropy_loss + v_loss * vf_coef
    return loss, (pg_loss, v_loss, entropy_loss, approx_kl)


def ppo_loss_and_grad(flat_params: np.ndarray, obs_u8, actions, behavior_logprobs, advantages, target_values,
                      clip_coef=0.1, ent_coef=0.01, vf_coef=0.5, dtype=torch.float32):
    """value_and_grad(ppo_loss) (cleanba_ppo.py:590,619-627) -> (5 scalars, flat grad)."""
    fp = torch.tensor(np.asarray(flat_params), dtype=dtype, requires_grad=True)
    p = net.unflatten(fp)
    lp, ent, val = logprob_entropy_value(p, torch.as_tensor(np.asarray(obs_u8)), torch.as_tensor(np.asarray(actions)))
    loss, (pg, vl, el, kl) = ppo_loss_from_heads(
        lp, ent, val,
        torch.as_tensor(np.asarray(behavior_logprobs)).to(dtype), torch.as_tensor(np.asarray(advantages)).to(dtype),
        torch.as_tensor(np.asarray(target_values)).to(dtype), clip_coef, ent_coef, vf_coef)
    loss.backward()
    stats = np.array([loss.item(), pg.item(), vl.item(), el.item(), kl.item()], np.float64)
    return stats, fp.grad.detach().numpy().copy()


def ppo_loss_and_grad_chunked(flat_params: np.ndarray, obs_u8, actions, behavior_logprobs, advantages, target_values,
                              clip_coef=0.1, ent_coef=0.01, vf_coef=0.5, dtype=torch.float32, chunk=256):
    """Same result as ppo_loss_and_grad for LARGE minibatches (config shape: 3840 frames) in bounded memory: every term of
    ppo_loss is a mean over the minibatch, so loss = sum_c (n_c / n) * loss(chunk c); gradients are accumulated chunk by chunk."""
    fp = torch.tensor(np.asarray(flat_params), dtype=dtype, requires_grad=True)
    n = len(actions)
    stats = np.zeros(5, np.float64)
    for lo in range(0, n, chunk):
        sl = slice(lo, min(n, lo + chunk))
        w = (sl.stop - sl.start) / n
        p = net.unflatten(fp)
        lp, ent, val = logprob_entropy_value(p, torch.as_tensor(np.asarray(obs_u8[sl])), torch.as_tensor(np.asarray(actions[sl])))
        loss, (pg, vl, el, kl) = ppo_loss_from_heads(
            lp, ent, val, torch.as_tensor(np.asarray(behavior_logprobs[sl])).to(dtype),
            torch.as_tensor(np.asarray(advantages[sl])).to(dtype), torch.as_tensor(np.asarray(target_values[sl])).to(dtype),
            clip_coef, ent_coef, vf_coef)
        (loss * w).backward()
        stats += w * np.array([loss.item(), pg.item(), vl.item(), el.item(), kl.item()], np.float64)
    return stats, fp.grad.detach().numpy().copy()


# ----------------------------------------------------------------------------- learner update
@dataclass
class PPOConfig:
    num_minibatches: int = 4
    update_epochs: int = 4
    gamma: float = 0.99
    gae_lambda: float = 0.95
    clip_coef: float = 0.1
    ent_coef: float = 0.01
    vf_coef: float = 0.5
    max_grad_norm: float = 0.5
    learning_rate: float = 2.5e-4
    anneal_lr: bool = True
    norm_adv: bool = True
    num_updates: int = 3255  # 50_000_000 // 15_360
    gradient_accumulation_steps: int = 1   # optax.MultiSteps(every_k_schedule) (cleanba_ppo.py:78,492-500,607)


@dataclass
class Shard:
    """One learner device's slice of an update: fields [T,Bl,...] + next_obs/next_done [Bl,...]
    (the hstack of the actor-thread payloads, cleanba_ppo.py:587-589)."""
    obs: np.ndarray
    dones: np.ndarray
    actions: np.ndarray
    logprobs: np.ndarray
    values: np.ndarray
    rewards: np.ndarray
    next_obs: np.ndarray
    next_done: np.ndarray


class PPOLearner:
    """single_device_update over L emulated devices with pmean'ed gradients (cleanba_ppo.py:579-660)."""

    def __init__(self, flat_params: np.ndarray, cfg: PPOConfig):
        self.cfg = cfg
        self.params = np.asarray(flat_params, F32).copy()
        self.opt = optim.Adam(self.params.size)

    def minibatch_indices(self, key: np.ndarray, n: int):
        """update_epoch's shuffle (cleanba_ppo.py:599-615): returns (new_key, idx[num_minibatches, mb])."""
        key, subkey = threefry.split(key)
        perm = threefry.permutation(subkey, n)
        return key, perm.reshape(self.cfg.num_minibatches * max(self.cfg.gradient_accumulation_steps, 1), -1)

    def prepare(self, shard: Shard):
        with torch.no_grad():
            _, next_value, _ = net.forward(self.params, shard.next_obs)
        adv, ret = compute_gae(shard.rewards, shard.values, shard.dones, next_value.numpy(), shard.next_done,
                               self.cfg.gamma, self.cfg.gae_lambda)
        if self.cfg.norm_adv:
            adv = normalize_advantages(adv, self.cfg.num_minibatches)
        return adv, ret

    def update(self, shards: Sequence[Shard], key: np.ndarray, record: Optional[list] = None):
        """-> (stats[5] averaged as cleanba_ppo.py:649-653, new_key).  All shards share `key` (quirk D.1/D.6)."""
        cfg = self.cfg
        prepared = [self.prepare(s) for s in shards]
        stats_all = []
        for _ in range(cfg.update_epochs):
            T, Bl = shards[0].rewards.shape
            key, idx = self.minibatch_indices(key, T * Bl)
            kacc = max(cfg.gradient_accumulation_steps, 1)
            for j in range(cfg.num_minibatches * kacc):
                grads, stats = [], []
                for s, (adv, ret) in zip(shards, prepared):
                    ii = idx[j]
                    st, g = ppo_loss_and_grad(
                        self.params, s.obs.reshape((-1,) + s.obs.shape[2:])[ii], s.actions.reshape(-1)[ii],
                        s.logprobs.reshape(-1)[ii], adv.reshape(-1)[ii], ret.reshape(-1)[ii],
                        cfg.clip_coef, cfg.ent_coef, cfg.vf_coef)
                    grads.append(g)
                    stats.append(st)
                g = np.mean(np.stack(grads), axis=0, dtype=F32)  # lax.pmean (cleanba_ppo.py:628)
                if getattr(self, "cross_allreduce", None) is not None:   # pmean also spans processes when --distributed
                    g = self.cross_allreduce(g)
                if kacc > 1:   # optax.MultiSteps (0.1.4): running mean of the mini-step gradients, inner update on the k-th
                    ms = j % kacc
                    self._acc = g.copy() if ms == 0 else (self._acc + (g - self._acc) / F32(ms + 1)).astype(F32)
                    stats_all.append(np.mean(np.stack(stats), axis=0))
                    if record is not None:
                        record.append(dict(mini_step=ms, raw_grad=g.copy(), stats=stats_all[-1].copy(), idx=idx[j].copy(), params_before=self.params.copy()))
                    if ms != kacc - 1:
                        continue
                    g = self._acc
                    stats_all.pop()
                    if record is not None:
                        record.pop()
                lr = optim.linear_schedule(self.opt.count, cfg.learning_rate, cfg.num_minibatches * cfg.update_epochs,
                                           cfg.num_updates, cfg.anneal_lr)
                if record is not None:   # the complete pre-step state: lets a test replay THIS step alone (no chained drift)
                    pre = dict(params_before=self.params.copy(), m_before=self.opt.m.copy(), v_before=self.opt.v.copy(),
                               count_before=int(self.opt.count), raw_grad=g.copy(), idx=idx[j].copy(),
                               shard_grads=[x.copy() for x in grads], shard_stats=[x.copy() for x in stats])
                    if kacc == 1:
                        # the same gradient in float64 on demand (shard index, or None = pmean): the tie-breaker when THIS fp32
                        # evaluation rounds a relu / max-pool gate the other way than the kernels under test (tests/_pin.py)
                        def grad64(si=None, p0=pre["params_before"], ii=idx[j].copy(), prep=prepared):
                            gs = [ppo_loss_and_grad(p0, s.obs.reshape((-1,) + s.obs.shape[2:])[ii], s.actions.reshape(-1)[ii],
                                                    s.logprobs.reshape(-1)[ii], adv.reshape(-1)[ii], ret.reshape(-1)[ii], cfg.clip_coef,
                                                    cfg.ent_coef, cfg.vf_coef, dtype=torch.float64)[1]
                                  for k, (s, (adv, ret)) in enumerate(zip(shards, prep)) if si is None or k == si]
                            return np.mean(np.stack(gs), axis=0)
                        pre["grad64"] = grad64
                g = optim.clip_by_global_norm(g, cfg.max_grad_norm)
                self.params = self.opt.step(self.params, g, lr)
                stats_all.append(np.mean(np.stack(stats), axis=0))
                if record is not None:
                    record.append(dict(grad=g.copy(), stats=stats_all[-1].copy(), lr=float(lr), params=self.params.copy(), **pre))
        return np.mean(np.stack(stats_all), axis=0), key
