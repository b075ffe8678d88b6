"""Host-side mirror of the reference's jitted hot-path callables, backed by libcleanba_b200 (CUDA, sm_100a).

  Actor.get_action_and_value  <-> cleanba/cleanba_ppo.py:245-261      Actor.get_action <-> cleanba/cleanba_impala.py:287-301
  Learner.compute_gae         <-> cleanba/cleanba_ppo.py:543-560 (+ :592-595 normalisation)
  Learner.ppo_update          <-> single_device_update, cleanba/cleanba_ppo.py:579-654
  Learner.impala_update       <-> single_device_update, cleanba/cleanba_impala.py:599-639
  publish_params              <-> cleanba/cleanba_ppo.py:721-725

PyTorch is plumbing only (device memory, streams, torch.distributed); all arithmetic runs in the C-ABI library.
"""
import ctypes
from typing import Callable, Optional, Sequence

import numpy as np
import torch

from . import lib as _lib
from .lib import (CB_ALGO_IMPALA, CB_ALGO_PPO, CB_CONV_SIMT, CB_CONV_TCGEN05, CB_MODEL_IMPALA_RESNET, CB_MODEL_NATURE_CNN, CleanbaError,
                  cb_config, check)

NUM_ACTIONS = 18


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _req(t: torch.Tensor, dtype, device, name):
    if t.dtype != dtype or t.device != device or not t.is_contiguous():
        raise CleanbaError(f"{name}: expected contiguous {dtype} on {device}, got {t.dtype} on {t.device} "
                           f"(contiguous={t.is_contiguous()})")
    return t


class Context:
    """One model replica on one GPU (an actor copy or a learner replica)."""

    def __init__(self, device, max_batch: int, algo: int = CB_ALGO_PPO, train: bool = False,
                 num_actions: int = NUM_ACTIONS, conv_backend: int = CB_CONV_TCGEN05, model: int = CB_MODEL_IMPALA_RESNET):
        if not torch.cuda.is_available():
            raise CleanbaError("cleanba_b200 needs a CUDA (sm_100a) device; there is no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise CleanbaError("cleanba_b200 contexts live on CUDA devices only")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.num_actions = num_actions
        self.algo = algo
        self.train = train
        self.max_batch = max_batch
        self.model = model
        cfg = cb_config(self.device.index, algo, max_batch, int(train), num_actions, conv_backend, model)
        h = ctypes.c_void_p()
        check(self.lib.cb_create(ctypes.byref(cfg), ctypes.byref(h)))
        self.h = h
        self.num_params = int(self.lib.cb_num_params_model(model, num_actions))
        self.hidden_width = int(self.lib.cb_hidden_width(h))

    def close(self):
        if getattr(self, "h", None):
            self.lib.cb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- parameters
    def set_params(self, flat):
        if isinstance(flat, np.ndarray):
            flat = torch.from_numpy(np.ascontiguousarray(flat, dtype=np.float32))
        flat = flat.to(torch.float32).contiguous()
        if flat.numel() != self.num_params:
            raise CleanbaError(f"expected {self.num_params} parameters, got {flat.numel()}")
        with torch.cuda.device(self.device):
            check(self.lib.cb_set_params(self.h, _ptr(flat), _stream(self.device)))
            if not flat.is_cuda:
                torch.cuda.current_stream(self.device).synchronize()

    def get_params(self) -> torch.Tensor:
        out = torch.empty(self.num_params, dtype=torch.float32, device=self.device)
        check(self.lib.cb_get_params(self.h, _ptr(out), _stream(self.device)))
        return out

    def params_view(self) -> torch.Tensor:
        """Zero-copy view of the master parameter vector (for NCCL broadcast); call refresh_weights() after writing."""
        ptr = self.lib.cb_params_ptr(self.h)
        iface = {"shape": (self.num_params,), "typestr": "<f4", "data": (ptr, False), "version": 2}
        holder = type("_CudaView", (), {"__cuda_array_interface__": iface})()
        return torch.as_tensor(holder, device=self.device)

    def refresh_weights(self):
        check(self.lib.cb_refresh_weights(self.h, _stream(self.device)))

    def publish_to(self, dst: "Context"):
        """learner -> actor parameter publish (cleanba_ppo.py:721-725), enqueued on dst's current stream."""
        check(self.lib.cb_publish_params(dst.h, self.h, _stream(dst.device)))

    def get_opt_state(self):
        m = torch.empty(self.num_params, dtype=torch.float32, device=self.device)
        v = torch.empty(self.num_params, dtype=torch.float32, device=self.device)
        cnt = ctypes.c_longlong()
        check(self.lib.cb_get_opt_state(self.h, _ptr(m), _ptr(v), ctypes.byref(cnt), _stream(self.device)))
        return m, v, cnt.value

    def set_opt_state(self, m, v, count: int):
        """Restore the optimizer moments and step count (resume; the reference has no equivalent)."""
        def dev(x):
            if isinstance(x, np.ndarray):
                x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
            x = x.to(self.device, torch.float32).contiguous()
            if x.numel() != self.num_params:
                raise CleanbaError(f"expected {self.num_params} optimizer-state elements, got {x.numel()}")
            return x
        m, v = dev(m), dev(v)
        with torch.cuda.device(self.device):
            check(self.lib.cb_set_opt_state(self.h, _ptr(m), _ptr(v), ctypes.c_longlong(int(count)), _stream(self.device)))
            torch.cuda.current_stream(self.device).synchronize()

    def profile(self, enable: bool):
        check(self.lib.cb_profile(self.h, int(enable)))

    def profile_report(self):
        import json
        buf = ctypes.create_string_buffer(1 << 16)
        check(self.lib.cb_profile_report(self.h, buf, len(buf)))
        return json.loads(buf.value.decode())

    def debug_tensor(self, name: str, shape) -> np.ndarray:
        out = np.empty(int(np.prod(shape)), np.float32)
        n = self.lib.cb_debug_tensor(self.h, name.encode(), out.ctypes.data_as(ctypes.c_void_p), out.size)
        if n < 0:
            raise CleanbaError(self.lib.cb_last_error().decode())
        return out[:n].reshape(shape)

    # ---- actor side
    def actor_step(self, obs: torch.Tensor, key: torch.Tensor, want_logprob_value=True, want_logits=False, out=None):
        """obs uint8 [n,4,84,84] (device); key uint32-as-int32 [2] (device, advanced in place).
        -> action int32[n], logprob f32[n]|None, value f32[n]|None, logits f32[n,A]|None
        `out` = (action, logprob, value, logits) pre-allocated contiguous device tensors (e.g. rows of the rollout
        storage) to write into instead of allocating."""
        d = self.device
        _req(obs, torch.uint8, d, "obs")
        n = obs.shape[0]
        if out is not None:
            action, logprob, value, logits = out
        else:
            action = torch.empty(n, dtype=torch.int32, device=d)
            logprob = torch.empty(n, dtype=torch.float32, device=d) if want_logprob_value else None
            value = torch.empty(n, dtype=torch.float32, device=d) if want_logprob_value else None
            logits = torch.empty(n, self.num_actions, dtype=torch.float32, device=d) if want_logits else None
        check(self.lib.cb_actor_step(self.h, _ptr(obs), n, _ptr(key), _ptr(action), _ptr(logprob), _ptr(value),
                                     _ptr(logits), _stream(d)))
        return action, logprob, value, logits

    def policy_value(self, obs: torch.Tensor, idx: Optional[torch.Tensor] = None, n: Optional[int] = None):
        d = self.device
        _req(obs, torch.uint8, d, "obs")
        n = int(idx.numel()) if idx is not None else (obs.shape[0] if n is None else n)
        logits = torch.empty(n, self.num_actions, dtype=torch.float32, device=d)
        value = torch.empty(n, dtype=torch.float32, device=d)
        check(self.lib.cb_policy_value(self.h, _ptr(obs), _ptr(idx), n, _ptr(logits), _ptr(value), _stream(d)))
        return logits, value

    # ---- learner pieces
    def gae(self, rewards, values, dones, next_value, next_done, gamma, gae_lambda, num_groups):
        d = self.device
        T, B = rewards.shape
        adv = torch.empty(T, B, dtype=torch.float32, device=d)
        ret = torch.empty(T, B, dtype=torch.float32, device=d)
        check(self.lib.cb_gae(self.h, _ptr(_req(rewards, torch.float32, d, "rewards")),
                              _ptr(_req(values, torch.float32, d, "values")), _ptr(_as_u8(dones, d, "dones")),
                              _ptr(_req(next_value, torch.float32, d, "next_value")),
                              _ptr(_as_u8(next_done, d, "next_done")), T, B, gamma, gae_lambda, num_groups,
                              _ptr(adv), _ptr(ret), _stream(d)))
        return adv, ret

    def split_key(self, key: torch.Tensor) -> torch.Tensor:
        sub = torch.empty(2, dtype=torch.int32, device=self.device)
        check(self.lib.cb_split_key(self.h, _ptr(key), _ptr(sub), _stream(self.device)))
        return sub

    def permutation(self, key: torch.Tensor, n: int) -> torch.Tensor:
        out = torch.empty(n, dtype=torch.int32, device=self.device)
        check(self.lib.cb_permutation(self.h, _ptr(key), n, _ptr(out), _stream(self.device)))
        return out

    def ppo_grad(self, obs, idx, mb, actions, logprobs, advantages, returns, clip_coef, ent_coef, vf_coef,
                 grads: torch.Tensor, stats: torch.Tensor):
        d = self.device
        check(self.lib.cb_ppo_grad(self.h, _ptr(_req(obs, torch.uint8, d, "obs")), _ptr(idx), mb,
                                   _ptr(_req(actions, torch.int32, d, "actions")),
                                   _ptr(_req(logprobs, torch.float32, d, "logprobs")),
                                   _ptr(_req(advantages, torch.float32, d, "advantages")),
                                   _ptr(_req(returns, torch.float32, d, "returns")), clip_coef, ent_coef, vf_coef,
                                   _ptr(grads), _ptr(stats), _stream(d)))

    def impala_grad(self, obs, idx, T1, B, actions, behaviour_logits, rewards, dones, firststeps, gamma, vf_coef,
                    ent_coef, grads, stats):
        d = self.device
        check(self.lib.cb_impala_grad(self.h, _ptr(_req(obs, torch.uint8, d, "obs")), _ptr(idx), T1, B,
                                      _ptr(_req(actions, torch.int32, d, "actions")),
                                      _ptr(_req(behaviour_logits, torch.float32, d, "behaviour_logits")),
                                      _ptr(_req(rewards, torch.float32, d, "rewards")), _ptr(_as_u8(dones, d, "dones")),
                                      _ptr(_as_u8(firststeps, d, "firststeps")), gamma, vf_coef, ent_coef,
                                      _ptr(grads), _ptr(stats), _stream(d)))

    def optimizer_step(self, grads, grad_scale, lr, max_norm, norm_out=None):
        check(self.lib.cb_optimizer_step(self.h, _ptr(grads), float(grad_scale), float(lr), float(max_norm),
                                         _ptr(norm_out), _stream(self.device)))


    def grad_accumulate(self, acc: torch.Tensor, grads: torch.Tensor, mini_step: int):
        """optax.MultiSteps running mean: acc += (grads - acc) / (mini_step + 1)."""
        check(self.lib.cb_grad_accumulate(self.h, _ptr(acc), _ptr(grads), int(mini_step), _stream(self.device)))

    def set_sm_budget(self, num_sms: int):
        """Size this context's persistent grids for an SM partition (see cleanba_b200.partition)."""
        check(self.lib.cb_set_sm_budget(self.h, int(num_sms)))

    def enable_peer_access(self, peer: "Context"):
        check(self.lib.cb_enable_peer_access(self.h, int(peer.device.index)))

    def optimizer_step_peers(self, grads_list, grad_scale, lr, max_norm, norm_out=None):
        """clip + Adam / RMSProp on the fixed-order SUM of the replicas' flat gradient buffers, read from peer memory inside
        the optimizer kernels (the fused form of pmean + apply_gradients, cleanba_ppo.py:628-629).  The caller orders the
        replicas' streams with events (see cuda_backend.CudaLearner)."""
        arr = (ctypes.c_void_p * len(grads_list))(*[g.data_ptr() for g in grads_list])
        check(self.lib.cb_optimizer_step_peers(self.h, arr, len(grads_list), float(grad_scale), float(lr), float(max_norm),
                                               _ptr(norm_out), _stream(self.device)))


    def set_grad_milestone(self, event: Optional[torch.cuda.Event]) -> int:
        """Have every *_grad call record `event` once the tail of the gradient vector (dense layer + heads) is final; returns
        the tail's first element offset.  See cb_set_grad_milestone."""
        off = ctypes.c_longlong()
        handle = None if event is None else ctypes.c_void_p(event.cuda_event)
        check(self.lib.cb_set_grad_milestone(self.h, handle, ctypes.byref(off)))
        return int(off.value)

    def set_actor_tail(self, cluster_size: int):
        """Opt in to the persistent ConvSequence 1+2 kernel (cluster_size 1 | 2; 0 = off); see cb_set_actor_tail."""
        check(self.lib.cb_set_actor_tail(self.h, int(cluster_size)))

    def graph_steps(self, enable: bool = True):
        """Replay ppo_grad / impala_grad as a captured CUDA graph (one small launch + one graph launch per call instead of ~70
        launches); see cb_graph_steps."""
        check(self.lib.cb_graph_steps(self.h, int(bool(enable))))

    @property
    def graph_replays(self) -> int:
        return int(self.lib.cb_graph_replays(self.h))

    def reduce_peers(self, grads_list, out: torch.Tensor):
        """out = fixed-order sum of the replicas' flat gradient buffers (peer memory); see cb_reduce_peers."""
        arr = (ctypes.c_void_p * len(grads_list))(*[g.data_ptr() for g in grads_list])
        check(self.lib.cb_reduce_peers(self.h, arr, len(grads_list), _ptr(out), _stream(self.device)))


def copy_columns(dst: torch.Tensor, src: torch.Tensor, stream: torch.cuda.Stream):
    """dst[r] <- src[r] for every row r of a [rows, ...] pair where each ROW is contiguous but the row pitch may differ (src is
    a column block `storage[:, c]` of a wider rollout storage): one cudaMemcpy2DAsync on `stream`, any device pair."""
    if dst.shape != src.shape or dst.dtype != src.dtype:
        raise CleanbaError(f"copy_columns: shape / dtype mismatch {tuple(dst.shape)} {dst.dtype} vs {tuple(src.shape)} {src.dtype}")
    rows = src.shape[0]
    if rows == 0 or src.numel() == 0:
        return
    if not (src[0].is_contiguous() and dst[0].is_contiguous()):
        raise CleanbaError("copy_columns: rows must be contiguous")
    es = src.element_size()
    width = src[0].numel() * es
    sp = src.stride(0) * es if rows > 1 else width
    dp = dst.stride(0) * es if rows > 1 else width
    check(_lib.load().cb_memcpy_2d(_ptr(dst), dp, _ptr(src), sp, width, rows, ctypes.c_void_p(stream.cuda_stream)))


import threading as _threading

_CAPTURE_LOCK = _threading.Lock()


class GraphedActor:
    """The actor step captured ONCE into a CUDA graph and replayed every environment step.

    At N = 60 the ~25 kernels of get_action_and_value are microseconds each, so launch overhead and inter-kernel gaps
    dominate an eager step; a replay costs one launch.  The graph reads a static observation buffer and the device-side
    PRNG key (advanced in place by every replay) and writes static output buffers; `step()` copies the frame in
    (host -> device or device -> device) on the actor's own stream and replays."""

    def __init__(self, ctx: Context, n: int, key: torch.Tensor, want_logits: bool = False, stream: Optional[torch.cuda.Stream] = None):
        d = ctx.device
        self.ctx, self.n, self.key = ctx, n, key
        self.stream = stream or torch.cuda.Stream(d)
        self.obs = torch.zeros(n, 4, 84, 84, dtype=torch.uint8, device=d)
        self.action = torch.zeros(n, dtype=torch.int32, device=d)
        self.logprob = None if want_logits else torch.zeros(n, dtype=torch.float32, device=d)
        self.value = None if want_logits else torch.zeros(n, dtype=torch.float32, device=d)
        self.logits = torch.zeros(n, ctx.num_actions, dtype=torch.float32, device=d) if want_logits else None
        out = (self.action, self.logprob, self.value, self.logits)
        key_backup = key.clone()
        with torch.cuda.device(d):
            self.stream.wait_stream(torch.cuda.current_stream(d))
            with torch.cuda.stream(self.stream):
                ctx.actor_step(self.obs, key, out=out)        # warm-up (sets function attributes, touches every buffer)
            self.stream.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with _CAPTURE_LOCK:      # one capture at a time; thread-local mode: other host threads keep using CUDA freely
                l0 = ctx.lib.cb_launch_count()
                with torch.cuda.graph(self.graph, stream=self.stream, capture_error_mode="thread_local"):
                    ctx.actor_step(self.obs, key, out=out)
            self.kernels_per_replay = int(ctx.lib.cb_launch_count() - l0)   # kernels of this library inside one replay
            self.replays = 0
            self.stream.synchronize()
        key.copy_(key_backup)                                 # warm-up and capture do not consume randomness

    def step(self, obs_src: torch.Tensor):
        """Enqueue one actor step on self.stream; results are in self.action / logprob / value / logits."""
        with torch.cuda.stream(self.stream):
            self.obs.copy_(obs_src, non_blocking=True)
            self.graph.replay()
            self.replays += 1


class RolloutActor:
    """The actor step captured ONCE into a CUDA graph that reads its frames from, and writes its outputs to, the current ROW
    of the rollout storage (cb_actor_step_cursor): per environment step the host copies the frames into their storage row and
    replays the graph -- the reference's list append + stack (`prepare_data`, cleanba_ppo.py:276-278, 342-356) without a staging
    buffer or per-field copies.  `begin()` points the device-side cursor at a rollout's storage views, `step()` runs row t."""

    def __init__(self, ctx: Context, n: int, key: torch.Tensor, want_logits: bool = False, stream: Optional[torch.cuda.Stream] = None):
        d = ctx.device
        self.ctx, self.n, self.key, self.want_logits = ctx, n, key, want_logits
        self.stream = stream or torch.cuda.Stream(d)
        self.cursor = torch.zeros(8, dtype=torch.int64, device=d)                 # struct cb_rollout_cursor (64 bytes)
        self._cursor_host = torch.zeros(8, dtype=torch.int64).pin_memory()
        self._cursor_event = None
        A = ctx.num_actions
        scratch = dict(obs=torch.zeros(1, n, 4, 84, 84, dtype=torch.uint8, device=d), action=torch.zeros(1, n, dtype=torch.int32, device=d),
                       logprob=None if want_logits else torch.zeros(1, n, dtype=torch.float32, device=d),
                       value=None if want_logits else torch.zeros(1, n, dtype=torch.float32, device=d),
                       logits=torch.zeros(1, n, A, dtype=torch.float32, device=d) if want_logits else None)
        self._scratch = scratch
        key_backup = key.clone()
        with torch.cuda.device(d):
            self.stream.wait_stream(torch.cuda.current_stream(d))
            self.begin(first_row=0, **scratch)
            with torch.cuda.stream(self.stream):
                self._launch()                                # warm-up (sets function attributes, touches every buffer)
            self.stream.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with _CAPTURE_LOCK:
                l0 = ctx.lib.cb_launch_count()
                with torch.cuda.graph(self.graph, stream=self.stream, capture_error_mode="thread_local"):
                    self._launch()
            self.kernels_per_replay = int(ctx.lib.cb_launch_count() - l0)
            self.replays = 0
            self.stream.synchronize()
        key.copy_(key_backup)                                 # warm-up and capture do not consume randomness
        self._obs = None

    def _launch(self):
        check(self.ctx.lib.cb_actor_step_cursor(self.ctx.h, _ptr(self.cursor), self.n, _ptr(self.key),
                                                ctypes.c_void_p(self.stream.cuda_stream)))

    def begin(self, obs, action, logprob=None, value=None, logits=None, first_row: int = 0):
        """Point the cursor at this rollout's storage: obs uint8 [rows, n, 4, 84, 84], action int32 [rows, n], logprob / value
        float32 [rows, n] (PPO) or logits float32 [rows, n, A] (IMPALA); views of a wider [rows, N_total, ...] storage are fine
        as long as every row of n entries is contiguous.  Enqueued on self.stream."""
        n, A, d = self.n, self.ctx.num_actions, self.ctx.device
        def chk(t, dtype, inner):
            if t is None:
                return 0
            if t.dtype != dtype or t.device != d or t.shape[1] != n or (n > 1 and t.stride(1) != inner) or not t[0].is_contiguous():
                raise CleanbaError(f"rollout storage view: expected {dtype} [rows, {n}, ...] on {d} with contiguous rows, got "
                                   f"{t.dtype} {tuple(t.shape)} strides {t.stride()} on {t.device}")
            return t.data_ptr()
        po = chk(obs, torch.uint8, 4 * 84 * 84)
        pa, pl, pv, pg = chk(action, torch.int32, 1), chk(logprob, torch.float32, 1), chk(value, torch.float32, 1), chk(logits, torch.float32, A)
        ors = action.stride(0)
        for t, mult in ((logprob, 1), (value, 1), (logits, A)):
            if t is not None and t.stride(0) != ors * mult:
                raise CleanbaError("rollout storage views must share one row stride (logits: row stride x num_actions)")
        if self._cursor_event is not None:
            self._cursor_event.synchronize()                  # the pinned mirror is free again
        h = self._cursor_host.numpy()
        h[:7] = [po, pa, pl, pv, pg, obs.stride(0), ors]
        h[7] = int(np.array([first_row - 1, 0], np.int32).view(np.int64)[0])
        with torch.cuda.stream(self.stream):
            self.cursor.copy_(self._cursor_host, non_blocking=True)
            self._cursor_event = torch.cuda.Event()
            self._cursor_event.record(self.stream)
        self._obs, self._next_row = obs, first_row

    def step(self, obs_src: torch.Tensor, t: int):
        """Enqueue the actor step of storage row t on self.stream: frames (host or device) -> row t, then one graph replay."""
        if t != self._next_row:
            raise CleanbaError(f"rollout rows must be stepped in order: expected row {self._next_row}, got {t}")
        if t >= self._obs.shape[0]:
            raise CleanbaError(f"row {t} is outside the rollout storage ({self._obs.shape[0]} rows): call begin() for the next rollout")
        with torch.cuda.stream(self.stream):
            self._obs[t].copy_(obs_src, non_blocking=True)
            self.graph.replay()
        self.replays += 1
        self._next_row += 1


def _as_u8(t: torch.Tensor, device, name):
    if t.dtype == torch.bool:
        t = t.view(torch.uint8)
    return _req(t, torch.uint8, device, name)


def key_tensor(key, device) -> torch.Tensor:
    """uint32[2] PRNG key as an int32 device tensor (bit pattern preserved)."""
    k = np.asarray(key, dtype=np.uint32).view(np.int32)
    return torch.from_numpy(k.copy()).to(device)


def key_numpy(key: torch.Tensor) -> np.ndarray:
    return key.detach().cpu().numpy().view(np.uint32).copy()
