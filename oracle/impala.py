"""IMPALA hot path of cleanba/cleanba_impala.py restated on CPU (numpy + torch autograd).

actor  : get_action                       cleanba_impala.py:287-301
learner: get_logits_and_value             cleanba_impala.py:547-555
         policy_gradient_loss             cleanba_impala.py:557-561   (rlax.policy_gradient_loss, summed)
         entropy_loss_fn                  cleanba_impala.py:563-567   (rlax.entropy_loss, summed)
         impala_loss                      cleanba_impala.py:569-597   (rlax.vtrace_td_error_and_advantage)
         single_device_update             cleanba_impala.py:599-639
rlax 0.1.5 (poetry.lock:1959) is not vendored: `vtrace` / `vtrace_td_error_and_advantage` restate its
published algorithm.  PARITY UNPINNED vs JAX/rlax; closed-form anchors in tests/.
"""
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import torch

from . import network as net
from . import optim, threefry
from .ppo import gumbel_argmax

F32 = np.float32


def get_action(flat_params: np.ndarray, next_obs: np.ndarray, key: np.ndarray):
    """cleanba_impala.py:287-301 -> (obs, action i32[N], logits f32[N,18], new_key)."""
    with torch.no_grad():
        logits, _, _ = net.forward(flat_params, next_obs)
    logits = logits.numpy()
    key, subkey = threefry.split(key)
    u = threefry.uniform(subkey, logits.shape)
    return next_obs, gumbel_argmax(logits, u), logits, key


def vtrace_errors(v_tm1, v_t, r_t, discount_t, rho_tm1, lambda_=1.0, clip_rho_threshold=1.0):
    """rlax.vtrace for one column set: all inputs [T,B] torch tensors; returns the *raw* errors
    (before the stop-gradient re-expression)."""
    c = torch.clamp(rho_tm1, max=1.0) * lambda_
    clipped = torch.clamp(rho_tm1, max=clip_rho_threshold)
    td = clipped * (r_t + discount_t * v_t - v_tm1)
    err = torch.zeros_like(v_t[0])
    out = [None] * v_t.shape[0]
    for i in reversed(range(v_t.shape[0])):
        err = td[i] + discount_t[i] * c[i] * err
        out[i] = err
    return torch.stack(out)


def vtrace_td_error_and_advantage(v_tm1, v_t, r_t, discount_t, rho_tm1, lambda_=1.0,
                                  clip_rho_threshold=1.0, clip_pg_rho_threshold=1.0):
    """rlax.vtrace_td_error_and_advantage (stop_target_gradients=True) vmapped over axis 1
    (cleanba_impala.py:585-587).  Returns (errors, pg_advantage, q_estimate)."""
    raw = vtrace_errors(v_tm1, v_t, r_t, discount_t, rho_tm1, lambda_, clip_rho_threshold)
    errors = (raw + v_tm1).detach() - v_tm1
    targets_tm1 = errors + v_tm1
    q_bootstrap = torch.cat([lambda_ * targets_tm1[1:] + (1 - lambda_) * v_tm1[1:], v_t[-1:]], 0)
    q_estimate = r_t + discount_t * q_bootstrap
    pg_adv = torch.clamp(rho_tm1, max=clip_pg_rho_threshold) * (q_estimate - v_tm1)
    return errors, pg_adv, q_estimate


def impala_loss_from_heads(policy_logits, newvalue, a, behaviour_logits, rewards, dones, firststeps,
                           gamma=0.99, vf_coef=0.5, ent_coef=0.01):
    """impala_loss after the network (cleanba_impala.py:570-597); inputs have T+1 rows."""
    dt = newvalue.dtype
    discounts = (1.0 - dones.to(dt)) * gamma
    mask = 1.0 - firststeps.to(dt)
    v_t = newvalue[1:]
    v_tm1 = newvalue[:-1]
    policy_logits = policy_logits[:-1]
    behaviour_logits = behaviour_logits[:-1]
    a = a[:-1].long()
    mask = mask[:-1]
    rewards = rewards[:-1].to(dt)
    discounts = discounts[:-1]

    logp = torch.log_softmax(policy_logits, -1)
    logp_a = logp.gather(-1, a[..., None]).squeeze(-1)
    logmu_a = torch.log_softmax(behaviour_logits.to(dt), -1).gather(-1, a[..., None]).squeeze(-1)
    rhos = torch.exp(logp_a - logmu_a)  # rlax.categorical_importance_sampling_ratios
    errors, pg_adv, _ = vtrace_td_error_and_advantage(v_tm1, v_t, rewards, discounts, rhos)
    pg_loss = torch.sum(-logp_a * pg_adv.detach() * mask)
    baseline_loss = 0.5 * torch.sum(errors ** 2 * mask)
    ent = -torch.nansum(torch.softmax(policy_logits, -1) * logp, dim=-1)
    ent_loss = torch.sum(-ent * mask)
    total = pg_loss + vf_coef * baseline_loss + ent_coef * ent_loss
    return total, (pg_loss, baseline_loss, ent_loss)


def impala_loss_and_grad(flat_params, obs_u8, a, behaviour_logits, rewards, dones, firststeps,
                         gamma=0.99, vf_coef=0.5, ent_coef=0.01, dtype=torch.float32):
    """value_and_grad(impala_loss) (cleanba_impala.py:606-618).  obs_u8 [T+1,B,4,84,84]."""
    fp = torch.tensor(np.asarray(flat_params), dtype=dtype, requires_grad=True)
    p = net.unflatten(fp)
    obs = torch.as_tensor(np.asarray(obs_u8))
    T1, B = obs.shape[:2]
    hidden = net.trunk_forward(p, obs.reshape((T1 * B,) + obs.shape[2:]))
    logits, value = net.heads(p, hidden)
    total, (pg, bl, el) = impala_loss_from_heads(
        logits.reshape(T1, B, -1), value.reshape(T1, B), torch.as_tensor(np.asarray(a)),
        torch.as_tensor(np.asarray(behaviour_logits)), torch.as_tensor(np.asarray(rewards)),
        torch.as_tensor(np.asarray(dones)), torch.as_tensor(np.asarray(firststeps)), gamma, vf_coef, ent_coef)
    total.backward()
    return np.array([total.item(), pg.item(), bl.item(), el.item()], np.float64), fp.grad.detach().numpy().copy()


@dataclass
class ImpalaConfig:
    num_minibatches: int = 4
    gamma: float = 0.99
    ent_coef: float = 0.01
    vf_coef: float = 0.5
    max_grad_norm: float = 40.0
    learning_rate: float = 6e-4
    anneal_lr: bool = True
    num_updates: int = 20833  # 50_000_000 // 2_400
    gradient_accumulation_steps: int = 1   # optax.MultiSteps(every_k_schedule) (cleanba_impala.py:76,532-540,626-633)


@dataclass
class Shard:
    """One learner device's slice: fields [T+1,Bl,...] (cleanba_impala.py:604)."""
    obs: np.ndarray
    dones: np.ndarray
    actions: np.ndarray
    logitss: np.ndarray
    rewards: np.ndarray
    firststeps: np.ndarray


class ImpalaLearner:
    """single_device_update over L emulated devices with pmean'ed gradients (cleanba_impala.py:599-645)."""

    def __init__(self, flat_params, cfg: ImpalaConfig):
        self.cfg = cfg
        self.params = np.asarray(flat_params, F32).copy()
        self.opt = optim.RMSPropPyTorchStyle(self.params.size)

    def update(self, shards: Sequence[Shard], record: Optional[list] = None):
        cfg = self.cfg
        stats_all = []
        Bl = shards[0].rewards.shape[1]
        kacc = max(cfg.gradient_accumulation_steps, 1)
        cols = np.split(np.arange(Bl), cfg.num_minibatches * kacc)  # contiguous column blocks (cleanba_impala.py:626-633)
        for j in range(cfg.num_minibatches * kacc):
            grads, stats = [], []
            for s in shards:
                c = cols[j]
                st, g = impala_loss_and_grad(self.params, s.obs[:, c], s.actions[:, c], s.logitss[:, c],
                                             s.rewards[:, c], s.dones[:, c], s.firststeps[:, c],
                                             cfg.gamma, cfg.vf_coef, cfg.ent_coef)
                grads.append(g)
                stats.append(st)
            g = np.mean(np.stack(grads), axis=0, dtype=F32)  # lax.pmean of summed-loss grads (quirk D.7)
            if getattr(self, "cross_allreduce", None) is not None:
                g = self.cross_allreduce(g)
            if kacc > 1:   # optax.MultiSteps (0.1.4): running mean of the mini-step gradients, inner update on the k-th
                ms = j % kacc
                self._acc = g.copy() if ms == 0 else (self._acc + (g - self._acc) / F32(ms + 1)).astype(F32)
                if ms != kacc - 1:
                    stats_all.append(np.mean(np.stack(stats), axis=0))
                    if record is not None:
                        record.append(dict(mini_step=ms, raw_grad=g.copy(), stats=stats_all[-1].copy(), cols=cols[j].copy(), params_before=self.params.copy()))
                    continue
                g = self._acc
            lr = optim.linear_schedule(self.opt.count, cfg.learning_rate, cfg.num_minibatches, cfg.num_updates, cfg.anneal_lr)
            if record is not None:   # the complete pre-step state: lets a test replay THIS step alone (no chained drift)
                pre = dict(params_before=self.params.copy(), nu_before=self.opt.nu.copy(), count_before=int(self.opt.count),
                           raw_grad=g.copy(), cols=cols[j].copy(), shard_grads=[x.copy() for x in grads],
                           shard_stats=[x.copy() for x in stats])
                if kacc == 1:
                    # the same gradient in float64 on demand (shard index, or None = pmean); see oracle/ppo.py
                    def grad64(si=None, p0=pre["params_before"], c=cols[j].copy()):
                        gs = [impala_loss_and_grad(p0, s.obs[:, c], s.actions[:, c], s.logitss[:, c], s.rewards[:, c], s.dones[:, c],
                                                   s.firststeps[:, c], cfg.gamma, cfg.vf_coef, cfg.ent_coef, dtype=torch.float64)[1]
                              for k, s in enumerate(shards) if si is None or k == si]
                        return np.mean(np.stack(gs), axis=0)
                    pre["grad64"] = grad64
            g = optim.clip_by_global_norm(g, cfg.max_grad_norm)
            self.params = self.opt.step(self.params, g, lr)
            stats_all.append(np.mean(np.stack(stats), axis=0))
            if record is not None:
                record.append(dict(grad=g.copy(), stats=stats_all[-1].copy(), lr=float(lr), params=self.params.copy(), **pre))
        return np.mean(np.stack(stats_all), axis=0)
