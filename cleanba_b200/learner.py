"""Learner-side host logic: one learner replica = one GPU = one `Context`; the per-update loop of the reference's
`single_device_update` is driven from Python over the C ABI, with the data-parallel gradient mean
(`jax.lax.pmean(grads, "local_devices")`, cleanba/cleanba_ppo.py:628, cleanba/cleanba_impala.py:619) as ONE allreduce
on the flat gradient buffer between the gradient call and the optimizer call.

PPO    : cleanba/cleanba_ppo.py:579-654      IMPALA : cleanba/cleanba_impala.py:599-639
"""
import os
from dataclasses import dataclass
from typing import Callable, Optional

import numpy as np
import torch

from .agent import Context, CB_ALGO_IMPALA, CB_ALGO_PPO, CB_CONV_TCGEN05, CB_MODEL_IMPALA_RESNET, CleanbaError

# allreduce hook: f(flat_grad_tensor) -> None, sums in place over all learner devices of all processes
AllReduce = Optional[Callable[[torch.Tensor], None]]


class _OverlappedExchange:
    """lax.pmean(grads) (cleanba_ppo.py:628) as TWO collectives on the flat buffer: the tail [dense | actor | critic] (91% of
    the bytes) is reduced on a side stream as soon as the backward pass has produced it -- under the whole conv backward -- and
    only the small conv-stage head after the gradient call.  Every element is reduced exactly once, so replicas stay bit-identical."""

    def __init__(self, ctx: Context, allreduce):
        self.allreduce = allreduce
        d = ctx.device
        with torch.cuda.device(d):
            self.stream = torch.cuda.Stream(d)
            self.event = torch.cuda.Event()
            self.event.record(torch.cuda.current_stream(d))        # materialises the cudaEvent_t handle
        self.split = ctx.set_grad_milestone(self.event)

    def __call__(self, grads: torch.Tensor):
        main = torch.cuda.current_stream(grads.device)
        self.stream.wait_event(self.event)                          # recorded inside the *_grad call that was just enqueued
        with torch.cuda.stream(self.stream):
            self.allreduce(grads[self.split:])
        self.allreduce(grads[:self.split])
        main.wait_stream(self.stream)


def _wrap_exchange(ctx: Context, allreduce: AllReduce, world_learners: int) -> AllReduce:
    if (allreduce is None or world_learners <= 1 or getattr(allreduce, "whole_buffer", False)
            or os.environ.get("CLEANBA_OVERLAP_EXCHANGE", "1") == "0"):
        return allreduce
    return _OverlappedExchange(ctx, allreduce)


def _graph_steps(ctx: Context):
    """The minibatch gradient step (~70 launches) replays as ONE captured CUDA graph (cb_graph_steps): the host-side cost of an
    update drops from ~50 ms of launches to a few hundred microseconds.  CLEANBA_GRAPH_STEPS=0 keeps plain launches."""
    if os.environ.get("CLEANBA_GRAPH_STEPS", "1") != "0":
        ctx.graph_steps(True)


def linear_schedule(count: int, base_lr: float, steps_per_update: int, num_updates: int, anneal: bool) -> float:
    """cleanba_ppo.py:475-479 / cleanba_impala.py:515-519, evaluated at the pre-increment optimizer count, in fp32."""
    if not anneal:
        return float(np.float32(base_lr))
    frac = 1.0 - (count // steps_per_update) / num_updates
    # the reference never steps past num_updates; a resumed run that does must not turn the step into gradient ascent
    return float(np.float32(base_lr * max(frac, 0.0)))


@dataclass
class PPOHyper:
    """Algorithm-specific arguments of cleanba_ppo.py:60-91 (defaults identical)."""
    learning_rate: float = 2.5e-4
    anneal_lr: bool = True
    gamma: float = 0.99
    gae_lambda: float = 0.95
    num_minibatches: int = 4
    update_epochs: int = 4
    norm_adv: bool = True
    clip_coef: float = 0.1
    ent_coef: float = 0.01
    vf_coef: float = 0.5
    max_grad_norm: float = 0.5
    num_updates: int = 3255
    gradient_accumulation_steps: int = 1    # optax.MultiSteps(every_k_schedule) (cleanba_ppo.py:78, 492-500)


class PPOLearner:
    def __init__(self, device, hyper: PPOHyper, T: int, Bl: int, world_learners: int = 1, allreduce: AllReduce = None,
                 conv_backend: int = CB_CONV_TCGEN05, num_actions: int = 18, model: int = CB_MODEL_IMPALA_RESNET):
        self.k = max(int(hyper.gradient_accumulation_steps), 1)
        if (T * Bl) % (hyper.num_minibatches * self.k):
            raise CleanbaError("T*Bl must be divisible by num_minibatches * gradient_accumulation_steps")
        if Bl % hyper.num_minibatches and hyper.norm_adv:
            raise CleanbaError("Bl must be divisible by num_minibatches (cleanba_ppo.py:416-418)")
        self.h, self.T, self.Bl = hyper, T, Bl
        self.mb = T * Bl // (hyper.num_minibatches * self.k)      # the shuffled batch is cut into num_minibatches * k mini-steps (:607)
        self.world_learners = world_learners
        self.ctx = Context(device, max_batch=max(self.mb, Bl), algo=CB_ALGO_PPO, train=True,
                           num_actions=num_actions, conv_backend=conv_backend, model=model)
        # with accumulation the exchanged buffer is complete only after the last mini-step's accumulate kernel: no overlap
        self.allreduce = _wrap_exchange(self.ctx, allreduce, world_learners) if self.k == 1 else allreduce
        _graph_steps(self.ctx)
        d = self.ctx.device
        self.grads = torch.zeros(self.ctx.num_params, dtype=torch.float32, device=d)
        self.acc = torch.zeros_like(self.grads) if self.k > 1 else None
        self.exchange_buffer = self.acc if self.k > 1 else self.grads     # what the (fused) gradient exchange reads
        self.stats = torch.zeros(hyper.update_epochs * hyper.num_minibatches * self.k, 5, dtype=torch.float32, device=d)
        self.opt_count = 0
        self.fused_step = None      # f(learner, grad_scale, lr, max_norm): gradient exchange fused into the optimizer step
        self.step_hook = None       # f(phase, k, learner), phase in "pre" | "grad" | "post" of minibatch step k: lets the parity
                                    # tests pin every step of an update to the oracle's recorded state (no chained drift)

    def update(self, obs, dones, actions, logprobs, values, rewards, next_obs, next_done, key) -> torch.Tensor:
        """single_device_update (cleanba_ppo.py:579-654).  Fields are [T,Bl,...] device tensors (the hstack of the actor
        payloads); `key` is the uint32[2] learner key (device, advanced in place).  Returns the 5 averaged scalars
        (loss, pg_loss, v_loss, entropy, approx_kl) as a device tensor (no host sync)."""
        h, c = self.h, self.ctx
        T, Bl = self.T, self.Bl
        _, next_value = c.policy_value(next_obs)                      # bootstrap value (cleanba_ppo.py:550-552)
        adv, ret = c.gae(rewards, values, dones, next_value, next_done, h.gamma, h.gae_lambda,
                         h.num_minibatches if h.norm_adv else 0)      # compute_gae + norm (cleanba_ppo.py:591-595)
        obs_f = obs.reshape(T * Bl, 4, 84, 84)
        act_f, lp_f, adv_f, ret_f = actions.reshape(-1), logprobs.reshape(-1), adv.reshape(-1), ret.reshape(-1)
        k = 0
        for _ in range(h.update_epochs):
            sub = c.split_key(key)                                    # key, subkey = split(key) (cleanba_ppo.py:599)
            perm = c.permutation(sub, T * Bl)                         # jax.random.permutation(subkey, .) (:606)
            for j in range(h.num_minibatches * self.k):
                idx = perm[j * self.mb:(j + 1) * self.mb]
                if self.step_hook is not None:
                    self.step_hook("pre", k, self)
                c.ppo_grad(obs_f, idx, self.mb, act_f, lp_f, adv_f, ret_f, h.clip_coef, h.ent_coef, h.vf_coef,
                           self.grads, self.stats[k])
                if self.step_hook is not None:
                    self.step_hook("grad", k, self)
                k += 1
                g = self.grads
                if self.k > 1:                                        # optax.MultiSteps: running mean over the k mini-steps
                    c.grad_accumulate(self.acc, self.grads, (j % self.k))
                    if (j + 1) % self.k:
                        continue
                    g = self.acc
                lr = linear_schedule(self.opt_count, h.learning_rate, h.num_minibatches * h.update_epochs,
                                     h.num_updates, h.anneal_lr)
                if self.fused_step is not None:                       # pmean + apply_gradients in one pass over peer memory
                    self.fused_step(self, 1.0 / self.world_learners, lr, h.max_grad_norm)
                else:
                    if self.allreduce is not None and self.world_learners > 1:
                        self.allreduce(g)                             # lax.pmean(grads) (cleanba_ppo.py:628)
                    c.optimizer_step(g, 1.0 / self.world_learners, lr, h.max_grad_norm)
                self.opt_count += 1
                if self.step_hook is not None:
                    self.step_hook("post", k - 1, self)
        return self.stats.mean(0)


@dataclass
class ImpalaHyper:
    """Algorithm-specific arguments of cleanba_impala.py:60-87 (defaults identical)."""
    learning_rate: float = 6e-4
    anneal_lr: bool = True
    gamma: float = 0.99
    num_minibatches: int = 4
    ent_coef: float = 0.01
    vf_coef: float = 0.5
    max_grad_norm: float = 40.0
    num_updates: int = 20833
    gradient_accumulation_steps: int = 1    # optax.MultiSteps(every_k_schedule) (cleanba_impala.py:76, 532-540, 626-633)


class ImpalaLearner:
    def __init__(self, device, hyper: ImpalaHyper, T1: int, Bl: int, world_learners: int = 1, allreduce: AllReduce = None,
                 conv_backend: int = CB_CONV_TCGEN05, num_actions: int = 18, model: int = CB_MODEL_IMPALA_RESNET):
        self.k = max(int(hyper.gradient_accumulation_steps), 1)
        if Bl % (hyper.num_minibatches * self.k):
            raise CleanbaError("Bl must be divisible by num_minibatches * gradient_accumulation_steps (cleanba_impala.py:456-458, 626-633)")
        self.h, self.T1, self.Bl = hyper, T1, Bl
        self.B = Bl // (hyper.num_minibatches * self.k)           # env columns per mini-step
        self.world_learners = world_learners
        self.ctx = Context(device, max_batch=T1 * self.B, algo=CB_ALGO_IMPALA, train=True, num_actions=num_actions,
                           conv_backend=conv_backend, model=model)
        self.allreduce = _wrap_exchange(self.ctx, allreduce, world_learners) if self.k == 1 else allreduce
        _graph_steps(self.ctx)
        d = self.ctx.device
        self.grads = torch.zeros(self.ctx.num_params, dtype=torch.float32, device=d)
        self.acc = torch.zeros_like(self.grads) if self.k > 1 else None
        self.exchange_buffer = self.acc if self.k > 1 else self.grads
        self.stats = torch.zeros(hyper.num_minibatches * self.k, 4, dtype=torch.float32, device=d)
        # contiguous env-column blocks, never shuffled (cleanba_impala.py:626-633): idx[j][t*B+b] = t*Bl + j*B + b
        t = torch.arange(T1, device=d, dtype=torch.int32)[:, None] * Bl
        self.fused_step = None
        self.step_hook = None       # see PPOLearner
        self.idx = [(t + (j * self.B + torch.arange(self.B, device=d, dtype=torch.int32))[None, :]).reshape(-1).contiguous()
                    for j in range(hyper.num_minibatches * self.k)]
        self.opt_count = 0

    def update(self, obs, dones, actions, logitss, rewards, firststeps) -> torch.Tensor:
        """single_device_update (cleanba_impala.py:599-639).  Fields are [T+1,Bl,...] device tensors.  Returns the 4
        averaged scalars (loss, pg_loss, v_loss, entropy_loss)."""
        h, c = self.h, self.ctx
        T1, Bl = self.T1, self.Bl
        obs_f = obs.reshape(T1 * Bl, 4, 84, 84)
        A = c.num_actions
        for j in range(h.num_minibatches * self.k):
            if self.step_hook is not None:
                self.step_hook("pre", j, self)
            c.impala_grad(obs_f, self.idx[j], T1, self.B, actions.reshape(-1), logitss.reshape(-1, A), rewards.reshape(-1),
                          dones.reshape(-1), firststeps.reshape(-1), h.gamma, h.vf_coef, h.ent_coef, self.grads, self.stats[j])
            if self.step_hook is not None:
                self.step_hook("grad", j, self)
            g = self.grads
            if self.k > 1:                                            # optax.MultiSteps: running mean over the k mini-steps
                c.grad_accumulate(self.acc, self.grads, j % self.k)
                if (j + 1) % self.k:
                    continue
                g = self.acc
            lr = linear_schedule(self.opt_count, h.learning_rate, h.num_minibatches, h.num_updates, h.anneal_lr)
            if self.fused_step is not None:                           # pmean + apply_gradients in one pass over peer memory
                self.fused_step(self, 1.0 / self.world_learners, lr, h.max_grad_norm)
            else:
                if self.allreduce is not None and self.world_learners > 1:
                    self.allreduce(g)                                 # lax.pmean(grads) (cleanba_impala.py:619)
                c.optimizer_step(g, 1.0 / self.world_learners, lr, h.max_grad_norm)
            self.opt_count += 1
            if self.step_hook is not None:
                self.step_hook("post", j, self)
        return self.stats.mean(0)
