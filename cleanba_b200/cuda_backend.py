"""Product backend of the Sebulba plumbing (cleanba_b200.sebulba): every hot-path call goes through the C ABI of
libcleanba_b200 on CUDA devices.  There is no CPU path here; constructing it without a GPU raises.

  actor -> learner payload (prepare_data + device_put_sharded, cleanba_ppo.py:276-278,357-363): each step's observation is
  copied from pinned host memory straight into row t of a pre-allocated [T,N,...] device buffer (no list / stack), the
  env axis is split into L contiguous slices and each slice is copied to its learner GPU as strided block copies on the
  actor's copy streams (one per learner, cb_memcpy_2d over NVLink); an event travels with the payload.
"""
import os
import threading
import time
from typing import List

import numpy as np
import torch

from . import agent as ag
from .learner import ImpalaHyper, ImpalaLearner, PPOHyper, PPOLearner
from .params import init_params
from .prng import first_key as _first_key


def _model_of(args) -> int:
    from .lib import MODELS
    return MODELS[getattr(args, "network", "impala_resnet")]


class _Storage:
    def __init__(self, actor, rows):
        d, N, A = actor.dev, actor.N, actor.ctx.num_actions
        self.rows = rows
        self.obs = torch.empty(rows, N, 4, 84, 84, dtype=torch.uint8, device=d)
        self.actions = torch.empty(rows, N, dtype=torch.int32, device=d)
        if actor.impala:
            self.logitss = torch.empty(rows, N, A, dtype=torch.float32, device=d)
        else:
            self.logprobs = torch.empty(rows, N, dtype=torch.float32, device=d)
            self.values = torch.empty(rows, N, dtype=torch.float32, device=d)
        self.host = {k: np.zeros((rows, N), dt) for k, dt in (("dones", bool), ("rewards", np.float32), ("firststeps", bool),
                                                              ("truncations", bool), ("terminations", np.int32), ("env_ids", np.int32))}
        self.impala = actor.impala
        self.stream = actor.stream

    def put_host(self, t, **fields):
        for k, v in fields.items():
            self.host[k][t] = v

    def take_carry(self):
        last = self.rows - 1
        c = {"obs": self.obs[last], "actions": self.actions[last], "logitss": self.logitss[last]}
        c["host"] = {k: v[last].copy() for k, v in self.host.items()}
        return c

    def put_carry(self, c):
        with torch.cuda.stream(self.stream):      # same stream as the actor steps that fill the other rows
            self.obs[0].copy_(c["obs"]); self.actions[0].copy_(c["actions"]); self.logitss[0].copy_(c["logitss"])
        for k, v in c["host"].items():
            self.host[k][0] = v


class CudaActor:
    def __init__(self, device_id, N, args, key):
        self.dev = torch.device("cuda", device_id)
        self.N = N
        self.impala = args.algo == "impala"
        self.ctx = ag.Context(self.dev, max_batch=N, algo=ag.CB_ALGO_IMPALA if self.impala else ag.CB_ALGO_PPO,
                              model=_model_of(args))
        self.stream = torch.cuda.Stream(self.dev)
        self.copy_streams = []                              # actor -> learner payload copies (overlap the next rollout)
        self._peers, self._payload_log, self._land = set(), [], {}
        self.key = ag.key_tensor(key, self.dev)
        self.act_host = torch.empty(N, dtype=torch.int32).pin_memory()
        self.staging = torch.empty(N, 4, 84, 84, dtype=torch.uint8).pin_memory()
        self.graphed = None     # the actor step is captured into a CUDA graph after the first parameter publish

    def new_storage(self, rows):
        # allocate on the actor's stream: the caching allocator recycles a block per allocation stream, and these buffers
        # are written on self.stream (the learner, which reads them on its own stream, calls record_stream on the payload)
        with torch.cuda.device(self.dev), torch.cuda.stream(self.stream):
            return _Storage(self, rows)

    def set_params(self, handle):
        snapshot, event = handle
        with torch.cuda.device(self.dev), torch.cuda.stream(self.stream):
            self.stream.wait_event(event)
            self.ctx.set_params(snapshot if snapshot.device == self.dev else snapshot.to(self.dev, non_blocking=True))
            self.stream.synchronize()      # the reference blocks on the new params too (cleanba_ppo.py:294-300)
        if self.graphed is None:
            with torch.cuda.device(self.dev):
                self.graphed = ag.RolloutActor(self.ctx, self.N, self.key, want_logits=self.impala, stream=self.stream)
                self._cur_storage = None

    def step(self, storage, t, obs_host):
        with torch.cuda.device(self.dev), torch.cuda.stream(self.stream):
            if isinstance(obs_host, np.ndarray):
                self.staging.numpy()[...] = obs_host
                obs_host = self.staging
            g = self.graphed
            if storage is not self._cur_storage:               # new rollout: point the device-side cursor at its storage
                if self.impala:
                    g.begin(storage.obs, storage.actions, logits=storage.logitss, first_row=t)
                else:
                    g.begin(storage.obs, storage.actions, storage.logprobs, storage.values, first_row=t)
                self._cur_storage = storage
            g.step(obs_host, t)                                # H2D of the frames into row t + one graph replay (~25 kernels),
            t0 = time.time()                                   # which writes this step's transition into row t
            self.act_host.copy_(storage.actions[t], non_blocking=True)
            self.stream.synchronize()
            return self.act_host.numpy().copy(), time.time() - t0

    def shard_to_learners(self, storage, next_obs, next_done, L, learner_devices=None):
        """prepare_data + jax.device_put_sharded (cleanba_ppo.py:276-278, 357-363): learner l receives env columns
        [l*N/L, (l+1)*N/L) of every field.  The column blocks go straight from the [T, N, ...] storage to the learner GPUs as
        strided block copies (cb_memcpy_2d: no contiguous temporary) on this actor's COPY stream, which only waits for the
        rollout's last step: the actor stream is free at once, so the next rollout overlaps the hand-off.  An event recorded
        on the copy stream travels with the payload."""
        N = self.N
        shards = []
        with torch.cuda.device(self.dev):
            done = torch.cuda.Event()
            done.record(self.stream)                      # the rollout (every step's writes into the storage rows)
            while len(self.copy_streams) < L:             # one copy stream per learner: the L block copies run concurrently
                self.copy_streams.append(torch.cuda.Stream(self.dev))
        nbytes, marks = 0, []
        for l in range(L):
            c = slice(l * N // L, (l + 1) * N // L)
            ld = self.learner_devices[l]
            cs = self.copy_streams[l]
            with torch.cuda.device(self.dev):
                cs.wait_event(done)
                t0 = torch.cuda.Event(enable_timing=True); t0.record(cs)
            local = L == 1 and ld == self.dev              # the only learner is this GPU: hand the storage itself over
            if ld != self.dev and ld not in self._peers:
                self.ctx.lib.cb_enable_peer_access(self.ctx.h, int(ld.index))    # DMA over NVLink instead of staging through the host
                self._peers.add(ld)                        # (a refusal only makes the copy slower: not an error)
            sh = {}
            fields = ["obs", "actions"] + (["logitss"] if self.impala else ["logprobs", "values"])
            for k in fields:
                src = getattr(storage, k)
                if local:
                    sh[k] = src
                    continue
                view = src[:, c]
                # allocate from a pool no compute stream frees into (this actor's copy stream / a landing stream on the learner
                # GPU): a recycled block is then only handed out once every stream that used it (record_stream) has drained,
                # so the unordered copy below can never overwrite memory another stream is still reading
                ls = cs if ld == self.dev else self._land.setdefault(ld, torch.cuda.Stream(ld))
                with torch.cuda.stream(ls):
                    dst = torch.empty(view.shape, dtype=view.dtype, device=ld)
                ag.copy_columns(dst, view, cs)
                src.record_stream(cs)                      # the storage block may not be recycled before the copy has run
                nbytes += dst.numel() * dst.element_size()
                sh[k] = dst
            for k in ("dones", "rewards", "firststeps"):
                sh[k] = torch.from_numpy(np.ascontiguousarray(storage.host[k][:, c])).to(ld, non_blocking=True)
            if next_obs is not None:
                no = next_obs if torch.is_tensor(next_obs) else torch.from_numpy(next_obs)
                sh["next_obs"] = no[c].to(ld, non_blocking=True)
                sh["next_done"] = torch.from_numpy(np.ascontiguousarray(next_done[c])).to(ld, non_blocking=True)
            with torch.cuda.device(self.dev):
                ev = torch.cuda.Event(enable_timing=True)
                ev.record(cs)
            sh["event"] = ev
            marks.append((t0, ev))
            shards.append(sh)
        if nbytes:
            self._payload_log.append((marks, nbytes))
        return shards

    def payload_bandwidth(self):
        """[(GB/s, bytes)] of the finished payload hand-offs (device-timed on the copy stream)."""
        out = []
        for marks, nb in self._payload_log:
            if all(t1.query() for _, t1 in marks):     # first start to last end over the per-learner copy streams
                t_first = marks[0][0]
                ms = max(t_first.elapsed_time(t1) for _, t1 in marks)
                out.append((nb / (ms * 1e-3) / 1e9, nb))
        return out


class CudaLearner:
    """multi_device_update over the local learner devices (cleanba_ppo.py:656-660) + parameter publish (:721-725)."""

    def __init__(self, args, key, allreduce):
        self.args = args
        self.impala = args.algo == "impala"
        self.devices = [torch.device("cuda", i) for i in args.learner_device_ids]
        L = len(self.devices)
        Bl = args.local_num_envs // L * args.num_actor_threads * len(args.actor_device_ids)
        world_learners = L * max(args.world_size, 1)
        self.cross = allreduce
        hooks = [self._make_hook(l) for l in range(L)]
        if self.impala:
            h = ImpalaHyper(learning_rate=args.learning_rate, anneal_lr=args.anneal_lr, gamma=args.gamma,
                            num_minibatches=args.num_minibatches, ent_coef=args.ent_coef, vf_coef=args.vf_coef,
                            max_grad_norm=args.max_grad_norm, num_updates=max(args.num_updates, 1),
                            gradient_accumulation_steps=args.gradient_accumulation_steps)
            self.learners = [ImpalaLearner(d, h, args.num_steps + 1, Bl, world_learners, hooks[l], model=_model_of(args))
                             for l, d in enumerate(self.devices)]
        else:
            h = PPOHyper(learning_rate=args.learning_rate, anneal_lr=args.anneal_lr, gamma=args.gamma, gae_lambda=args.gae_lambda,
                         num_minibatches=args.num_minibatches, update_epochs=args.update_epochs, norm_adv=args.norm_adv,
                         clip_coef=args.clip_coef, ent_coef=args.ent_coef, vf_coef=args.vf_coef, max_grad_norm=args.max_grad_norm,
                         num_updates=max(args.num_updates, 1), gradient_accumulation_steps=args.gradient_accumulation_steps)
            self.learners = [PPOLearner(d, h, args.num_steps, Bl, world_learners, hooks[l], model=_model_of(args))
                             for l, d in enumerate(self.devices)]
        params = init_params(args.seed, model=_model_of(args))
        for lr in self.learners:
            lr.ctx.set_params(params)
        self.keys = [ag.key_tensor(key, d) for d in self.devices]    # learner_keys = device_put_replicated(key) (cleanba_ppo.py:470)
        self.barrier = threading.Barrier(L, timeout=300) if L > 1 else None   # a failed replica thread breaks the barrier instead of hanging the others
        self.hyper = h
        # Several learner GPUs in ONE process and no cross-process exchange: the gradient mean is fused into the optimizer
        # kernels, which read every replica's flat gradient buffer from peer memory over NVLink (cb_optimizer_step_peers);
        # the replicas' streams are ordered with events, the host threads only rendezvous (no device synchronisation).
        # With `--distributed` on top (allreduce given), the exchange has two levels and still no device synchronisation:
        # replica 0 sums its process' buffers from peer memory (cb_reduce_peers), joins ONE NCCL allreduce on that sum, and
        # every replica applies the optimizer step on replica 0's reduced buffer (peer loads again).
        self.peer_fused = L > 1 and os.environ.get("CLEANBA_PEER_FUSED", "1") != "0"
        if self.peer_fused:
            for a in self.learners:
                for b in self.learners:
                    a.ctx.enable_peer_access(b.ctx)
            self.ev_done, self.ev_read, self.ev_sum = [None] * L, [None] * L, None
            if allreduce is not None:
                self.gsum = torch.zeros_like(self.learners[0].grads)
            for l, lr in enumerate(self.learners):
                lr.fused_step = self._make_fused(l) if allreduce is None else self._make_fused_cross(l)

    def _make_hook(self, l):
        """Gradient sum over all learner devices: local devices rendezvous on device 0, device 0 joins the cross-process
        allreduce (one NCCL allreduce on the flat buffer), the result is copied back."""
        def hook(g):
            L = len(self.devices)
            if L == 1:
                if self.cross is not None:
                    self.cross(g)
                return
            torch.cuda.current_stream(g.device).synchronize()
            self.barrier.wait()
            if l == 0:
                g0 = self.learners[0].exchange_buffer
                for k in range(1, L):
                    g0.add_(self.learners[k].exchange_buffer.to(g0.device))
                if self.cross is not None:
                    self.cross(g0)
                torch.cuda.current_stream(g0.device).synchronize()
            self.barrier.wait()
            if l != 0:
                g.copy_(self.learners[0].exchange_buffer)
                torch.cuda.current_stream(g.device).synchronize()
            self.barrier.wait()
        hook.whole_buffer = len(self.devices) > 1      # the multi-device fallback exchanges whole buffers: not splittable
        return hook

    def _make_fused(self, l):
        L = len(self.devices)

        def fused(lr, grad_scale, lrate, max_norm):
            st = torch.cuda.current_stream(self.devices[l])
            ev = torch.cuda.Event()
            ev.record(st)                                   # this replica's backward is complete
            self.ev_done[l] = ev
            self.barrier.wait()                             # host rendezvous only: every replica's event exists
            for k in range(L):
                if k != l:
                    st.wait_event(self.ev_done[k])
            lr.ctx.optimizer_step_peers([x.exchange_buffer for x in self.learners], grad_scale, lrate, max_norm)
            ev = torch.cuda.Event()
            ev.record(st)                                   # this replica has read every peer buffer
            self.ev_read[l] = ev
            self.barrier.wait()
            for k in range(L):
                if k != l:
                    st.wait_event(self.ev_read[k])          # the next backward may overwrite this replica's buffer
        return fused

    def _make_fused_cross(self, l):
        """pmean over ALL global learner devices (cleanba_ppo.py:628 under jax.distributed): in-process sum over peer memory
        -> one NCCL allreduce issued by replica 0 -> optimizer step of every replica on the reduced buffer."""
        L = len(self.devices)

        def fused(lr, grad_scale, lrate, max_norm):
            st = torch.cuda.current_stream(self.devices[l])
            ev = torch.cuda.Event()
            ev.record(st)                                   # this replica's backward is complete
            self.ev_done[l] = ev
            self.barrier.wait()                             # host rendezvous only: every replica's event exists
            if l == 0:
                for k in range(1, L):
                    st.wait_event(self.ev_done[k])
                lr.ctx.reduce_peers([x.exchange_buffer for x in self.learners], self.gsum)
                self.cross(self.gsum)                       # ONE NCCL allreduce per minibatch on the flat gradient buffer
                ev = torch.cuda.Event()
                ev.record(st)
                self.ev_sum = ev
            self.barrier.wait()
            if l != 0:
                st.wait_event(self.ev_sum)
            lr.ctx.optimizer_step_peers([self.gsum], grad_scale, lrate, max_norm)
            ev = torch.cuda.Event()
            ev.record(st)                                   # this replica has read the reduced buffer
            self.ev_read[l] = ev
            self.barrier.wait()
            if l == 0:
                for k in range(1, L):
                    st.wait_event(self.ev_read[k])          # the next reduction may overwrite gsum
        return fused

    def _update_one(self, l, payloads, out):
        lr, d = self.learners[l], self.devices[l]
        with torch.cuda.device(d):
            st = torch.cuda.current_stream(d)
            shards = [p[l] for p in payloads]
            for s in shards:
                st.wait_event(s["event"])
                for v in s.values():      # payload memory was allocated on the actor's stream: tell the caching allocator
                    if torch.is_tensor(v) and v.is_cuda:   # that this stream uses it too, so it is not recycled early
                        v.record_stream(st)
            # a same-device shard is a strided column view of the actor's storage (`.to(ld)` is a no-op there): make it dense
            cat = (lambda k: shards[0][k].contiguous()) if len(shards) == 1 else (lambda k: torch.cat([s[k] for s in shards], dim=1).contiguous())
            if self.impala:
                out[l] = lr.update(cat("obs"), cat("dones"), cat("actions"), cat("logitss"), cat("rewards"), cat("firststeps"))
            else:
                nobs = shards[0]["next_obs"].contiguous() if len(shards) == 1 else torch.cat([s["next_obs"] for s in shards]).contiguous()
                ndone = shards[0]["next_done"].contiguous() if len(shards) == 1 else torch.cat([s["next_done"] for s in shards]).contiguous()
                out[l] = lr.update(cat("obs"), cat("dones"), cat("actions"), cat("logprobs"), cat("values"), cat("rewards"),
                                   nobs, ndone, self.keys[l])

    def update(self, payloads):
        L = len(self.devices)
        out = [None] * L
        if L == 1:
            self._update_one(0, payloads, out)
        else:
            errs = []

            def run(l):
                try:
                    self._update_one(l, payloads, out)
                except BaseException as e:      # surface replica failures in the caller instead of losing them with the thread
                    errs.append(e)
                    self.barrier.abort()
            ths = [threading.Thread(target=run, args=(l,)) for l in range(L)]
            [t.start() for t in ths]
            [t.join() for t in ths]
            if errs:
                first = next((e for e in errs if not isinstance(e, threading.BrokenBarrierError)), errs[0])
                raise RuntimeError("a learner replica failed") from first
        if L == 1:
            return out[0]
        # loss scalars are pmean'ed over the learner devices (cleanba_ppo.py:649-653; cleanba_impala.py:635-638); the
        # reference then logs device [-1]'s copy of that mean (cleanba_ppo.py:745-749)
        d = self.devices[-1]
        with torch.cuda.device(d):
            st = torch.cuda.current_stream(d)
            for l in range(L - 1):
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(self.devices[l]))
                st.wait_event(ev)
            return torch.stack([o.to(d, non_blocking=True) for o in out]).mean(0)

    def params_for_actor(self, actor_device_id):
        lr = self.learners[0]
        with torch.cuda.device(lr.ctx.device):
            snap = lr.ctx.get_params()
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(lr.ctx.device))
        return snap, ev

    def flat_params(self):
        """Host copy of the (replicated) parameters: flax.jax_utils.unreplicate(agent_state).params (cleanba_ppo.py:756)."""
        return self.learners[0].ctx.get_params().cpu().numpy()

    def train_state(self):
        """Everything needed to resume: parameters, optimizer moments + count, learner keys."""
        lr = self.learners[0]
        m, v, count = lr.ctx.get_opt_state()
        return dict(params=self.flat_params(), m=m.cpu().numpy(), v=v.cpu().numpy(), count=count,
                    key=ag.key_numpy(self.keys[0]), opt_count=lr.opt_count)

    def load_train_state(self, st):
        for l, lr in enumerate(self.learners):
            lr.ctx.set_params(st["params"])
            lr.ctx.set_opt_state(st["m"], st["v"], st["count"])
            lr.opt_count = int(st["count"])
            self.keys[l] = ag.key_tensor(np.asarray(st["key"], np.uint32), self.devices[l])

    def stats_to_host(self, stats):
        s = stats.detach().cpu().numpy()
        names = ("loss", "pg_loss", "v_loss", "entropy_loss") + (() if self.impala else ("approx_kl",))
        return {k: float(v) for k, v in zip(names, s)}

    def current_lr(self):
        """`charts/learning_rate` (cleanba_ppo.py:737-739): the reference reads the learning rate out of the inject_hyperparams state,
        i.e. the rate the LAST optimizer step used (the schedule at the count before that step), not the next step's."""
        from .learner import linear_schedule
        lr = self.learners[0]
        spu = self.hyper.num_minibatches * (1 if self.impala else self.hyper.update_epochs)
        return linear_schedule(max(lr.opt_count - 1, 0), self.hyper.learning_rate, spu, self.hyper.num_updates, self.hyper.anneal_lr)


class CudaBackend:
    def __init__(self):
        if not torch.cuda.is_available():
            raise ag.CleanbaError("CudaBackend needs CUDA devices (sm_100a); there is no CPU fallback")
        self._learner_devices = None
        self.actors = []            # every CudaActor made (payload bandwidth log, probes)

    def first_key(self, seed):
        return _first_key(seed)

    def make_learner(self, args, key, allreduce):
        learner = CudaLearner(args, key, allreduce)
        self._learner_devices = learner.devices
        return learner

    def make_actor(self, device_id, N, args, key):
        a = CudaActor(device_id, N, args, key)
        a.learner_devices = self._learner_devices
        self.actors.append(a)
        return a
