"""Sebulba actor-learner plumbing (host side, Python threads + size-1 queues), mirroring the reference's `__main__` and
`rollout()`:

  Args + size derivation / asserts      cleanba/cleanba_ppo.py:34-118,410-430   (IMPALA: cleanba_impala.py:34-111,450-470)
  actor thread `rollout`                cleanba/cleanba_ppo.py:226-406          (IMPALA: cleanba_impala.py:268-447)
  params / rollout queues, thread spawn cleanba/cleanba_ppo.py:662-686
  learner loop                          cleanba/cleanba_ppo.py:691-751

The arithmetic is delegated to a *backend* object (injected): the product backend is `cleanba_b200.cuda_backend` (CUDA,
C ABI); tests drive the same plumbing with a CPU backend that lives under tests/ + oracle/.  This module never imports
the oracle and has no CPU fallback of its own.
"""
import os
import queue
import threading
import time
from collections import deque
from dataclasses import dataclass, field
from types import SimpleNamespace
from typing import Callable, List, Optional

import numpy as np

from .trace import NullTracer, Tracer


@dataclass
class Args:
    """Same fields, meaning and defaults as the reference `Args` (cleanba_ppo.py:34-118); `algo` selects the PPO or the
    IMPALA variant (the IMPALA script differs in the defaults overridden by `impala_defaults`)."""
    exp_name: str = "cleanba_ppo"
    seed: int = 1
    track: bool = False
    wandb_project_name: str = "cleanRL"
    wandb_entity: Optional[str] = None
    capture_video: bool = False
    save_model: bool = False
    upload_model: bool = False
    hf_entity: str = ""
    log_frequency: int = 10

    env_id: str = "Breakout-v5"
    total_timesteps: int = 50000000
    learning_rate: float = 2.5e-4
    local_num_envs: int = 64
    num_actor_threads: int = 2
    num_steps: int = 128
    anneal_lr: bool = True
    gamma: float = 0.99
    gae_lambda: float = 0.95
    num_minibatches: int = 4
    gradient_accumulation_steps: int = 1
    update_epochs: int = 4
    norm_adv: bool = True
    clip_coef: float = 0.1
    ent_coef: float = 0.01
    vf_coef: float = 0.5
    max_grad_norm: float = 0.5
    channels: List[int] = field(default_factory=lambda: [16, 32, 32])
    hiddens: List[int] = field(default_factory=lambda: [256])

    actor_device_ids: List[int] = field(default_factory=lambda: [0])
    learner_device_ids: List[int] = field(default_factory=lambda: [0])
    distributed: bool = False
    concurrency: bool = False

    # B200 build additions (not in the reference): which algorithm this Args drives and how many updates to run at most
    algo: str = "ppo"
    network: str = "impala_resnet"  # trunk: "impala_resnet" (cleanba_ppo.py:149-189) or "nature_cnn" (legacy_scripts/..._naturecnn.py:143-178)
    max_updates: int = 0          # 0 = run to total_timesteps
    synthetic_env: bool = True    # envpool is not installable here; frames come from cleanba_b200.envs.SyntheticAtari
    eval_max_steps: int = 27000   # step cap of one evaluation episode after --save-model (envpool's max_episode_steps)
    resume_from: str = ""         # train-state sidecar written by --save-model (parameters + optimizer state + keys)
    trace_path: str = ""          # write a Chrome / Perfetto timeline of the actor and learner threads here (cleanba_b200.trace)

    # runtime arguments to be filled in (cleanba_ppo.py:106-118)
    local_batch_size: int = 0
    local_minibatch_size: int = 0
    num_updates: int = 0
    world_size: int = 0
    local_rank: int = 0
    num_envs: int = 0
    batch_size: int = 0
    minibatch_size: int = 0


def impala_defaults(args: Args) -> Args:
    """Defaults of cleanba_impala.py:60-87 that differ from the PPO script."""
    args.exp_name = "cleanba_impala"
    args.algo = "impala"
    args.learning_rate = 6e-4
    args.num_steps = 20
    args.max_grad_norm = 40.0
    args.concurrency = True
    return args


def derive_sizes(args: Args, world_size: int = 1, local_rank: int = 0) -> Args:
    """cleanba_ppo.py:411-430: batch sizes, divisibility asserts, num_updates."""
    if args.network not in ("impala_resnet", "nature_cnn"):
        raise ValueError("--network must be impala_resnet or nature_cnn")
    if args.network == "impala_resnet" and (args.channels != [16, 32, 32] or args.hiddens != [256]):
        raise ValueError("libcleanba_b200 implements the reference's default IMPALA-ResNet (channels 16,32,32; hiddens 256) "
                         "and the legacy Nature-CNN (--network nature_cnn)")
    if args.gradient_accumulation_steps < 1:
        raise ValueError("gradient_accumulation_steps must be >= 1")
    args.local_batch_size = int(args.local_num_envs * args.num_steps * args.num_actor_threads * len(args.actor_device_ids))
    args.local_minibatch_size = int(args.local_batch_size // args.num_minibatches)
    assert args.local_num_envs % len(args.learner_device_ids) == 0, \
        "local_num_envs must be divisible by len(learner_device_ids)"
    assert int(args.local_num_envs / len(args.learner_device_ids)) * args.num_actor_threads % args.num_minibatches == 0, \
        "int(local_num_envs / len(learner_device_ids)) must be divisible by num_minibatches"
    args.world_size = world_size
    args.local_rank = local_rank
    args.num_envs = args.local_num_envs * world_size * args.num_actor_threads * len(args.actor_device_ids)
    args.batch_size = args.local_batch_size * world_size
    args.minibatch_size = args.local_minibatch_size * world_size
    args.num_updates = args.total_timesteps // (args.local_batch_size * world_size)
    return args


def distributed_env():
    """(world_size, rank, local_rank) from torchrun variables, or from the SLURM variables jax.distributed.initialize()
    auto-detects in the reference (cleanba_ppo.py:419-423, README.md:68-72)."""
    e = os.environ
    if "WORLD_SIZE" in e:
        return int(e["WORLD_SIZE"]), int(e.get("RANK", 0)), int(e.get("LOCAL_RANK", 0))
    if "SLURM_NTASKS" in e:
        return int(e["SLURM_NTASKS"]), int(e.get("SLURM_PROCID", 0)), int(e.get("SLURM_LOCALID", 0))
    return 1, 0, 0


class _NullWriter:
    def add_scalar(self, *a, **k):
        pass

    def add_text(self, *a, **k):
        pass

    def close(self):
        pass


def rollout(args, backend, make_env, rollout_queue, params_queue, writer, device_thread_id, actor_device_id, stop, key, errors):
    """Thread target: runs `_rollout` and hands any exception to the learner loop (in the reference a dead actor thread
    deadlocks the learner on Queue.get(), cleanba_ppo.py:708; here it is re-raised in the main thread)."""
    try:
        _rollout(args, backend, make_env, rollout_queue, params_queue, writer, device_thread_id, actor_device_id, stop, key)
    except BaseException as e:  # noqa: BLE001
        errors.append(e)
        stop.set()


def _rollout(args: Args, backend, make_env: Callable, rollout_queue: queue.Queue, params_queue: queue.Queue, writer,
             device_thread_id: int, actor_device_id: int, stop: threading.Event, key):
    """Actor thread (cleanba_ppo.py:226-406 / cleanba_impala.py:268-447)."""
    impala = args.algo == "impala"
    tracer = getattr(args, "_tracer", None) or NullTracer()
    tid = 1 + device_thread_id
    tracer.thread_name(tid, f"actor thread {device_thread_id} (device {actor_device_id})")
    envs = make_env(args.env_id, args.seed + args.local_rank + device_thread_id, args.local_num_envs)()
    len_actor_device_ids = len(args.actor_device_ids)
    N = args.local_num_envs
    # --resume-from: continue the update / step counters of the saved run (the learning-rate schedule and the logs depend on them)
    first_update, global_step = getattr(args, "_resume", (0, 0))
    start_step = global_step
    start_time = time.time()
    actor = backend.make_actor(actor_device_id, N, args, key)
    episode_returns = np.zeros((N,), dtype=np.float32)
    returned_episode_returns = np.zeros((N,), dtype=np.float32)
    episode_lengths = np.zeros((N,), dtype=np.float32)
    returned_episode_lengths = np.zeros((N,), dtype=np.float32)
    params_queue_get_time = deque(maxlen=10)
    rollout_time = deque(maxlen=10)
    rollout_queue_put_time = deque(maxlen=10)
    actor_policy_version = first_update
    if impala:
        envs.async_reset()
        next_obs = next_done = None
    else:
        next_obs = envs.reset()
        next_done = np.zeros(N, dtype=bool)
    carry = None   # IMPALA: last transition of the previous rollout (cleanba_impala.py:416)

    for update in range(first_update + 1, args.num_updates + 2):
        if stop.is_set():
            break
        update_time_start = time.time()
        env_recv_time = inference_time = storage_time = d2h_time = env_send_time = 0.0
        # NOTE: `update != 2` lets policy collection run concurrently with learning while keeping the actor's policy
        # exactly one version behind the learner's (cleanba_ppo.py:287-304)
        t0 = time.time()
        if not args.concurrency or update - first_update != 2:
            params = None
            with tracer.span("params_queue.get", tid, update=update):
                while not stop.is_set():
                    try:
                        params = params_queue.get(timeout=0.5)
                        break
                    except queue.Empty:
                        continue
                if params is not None:
                    actor.set_params(params)      # includes the block_until_ready of the reference
            if params is None:                    # shutdown (sentinel from train() or stop flag)
                break
            actor_policy_version += 1
        params_queue_get_time.append(time.time() - t0)
        rollout_time_start = time.time()
        rollout_span = tracer.span("rollout", tid, update=update, policy_version=actor_policy_version)
        rollout_span.__enter__()
        T = args.num_steps
        rows = T + 1 if impala else T
        storage = actor.new_storage(rows)
        row0 = 0
        if impala and carry is not None:
            storage.put_carry(carry)              # bootstrap step moved to the beginning of this update
            row0 = 1
        for t in range(row0, rows):
            if impala:
                t1 = time.time()
                next_obs, next_reward, next_done, info = envs.recv()
                env_recv_time += time.time() - t1
            cached_next_obs, cached_next_done = next_obs, next_done
            global_step += N * args.num_actor_threads * len_actor_device_ids * args.world_size
            t1 = time.time()
            cpu_action, t_d2h = actor.step(storage, t, cached_next_obs)     # get_action_and_value + np.array(action)
            inference_time += time.time() - t1 - t_d2h
            d2h_time += t_d2h
            t1 = time.time()
            if impala:
                envs.send(cpu_action, info["env_id"])
                reward_t, done_t = next_reward, next_done
            else:
                next_obs, next_reward, next_done, info = envs.step(cpu_action)
                reward_t, done_t = next_reward, cached_next_done
            env_send_time += time.time() - t1
            t1 = time.time()
            truncated = info["elapsed_step"] >= envs.spec.config.max_episode_steps
            storage.put_host(t, dones=done_t, env_ids=info["env_id"], rewards=reward_t, truncations=truncated,
                             terminations=info["terminated"], firststeps=info["elapsed_step"] == 0)
            env_id = info["env_id"]
            episode_returns[env_id] += info["reward"]
            returned_episode_returns[env_id] = np.where(info["terminated"] + truncated, episode_returns[env_id], returned_episode_returns[env_id])
            episode_returns[env_id] *= (1 - info["terminated"]) * (1 - truncated)
            episode_lengths[env_id] += 1
            returned_episode_lengths[env_id] = np.where(info["terminated"] + truncated, episode_lengths[env_id], returned_episode_lengths[env_id])
            episode_lengths[env_id] *= (1 - info["terminated"]) * (1 - truncated)
            storage_time += time.time() - t1
        rollout_time.append(time.time() - rollout_time_start)
        rollout_span.__exit__(None, None, None)
        avg_episodic_return = np.mean(returned_episode_returns)
        # prepare_data + device_put_sharded (cleanba_ppo.py:357-363): split the env axis over the learner devices
        with tracer.span("shard_to_learners", tid, update=update):
            sharded = actor.shard_to_learners(storage, None if impala else next_obs, None if impala else next_done,
                                              len(args.learner_device_ids))
        payload = (global_step, actor_policy_version, update, sharded, np.mean(params_queue_get_time), device_thread_id)
        t1 = time.time()
        with tracer.span("rollout_queue.put", tid, update=update):
            while not stop.is_set():
                try:
                    rollout_queue.put(payload, timeout=0.5)
                    break
                except queue.Full:
                    continue
        rollout_queue_put_time.append(time.time() - t1)
        if impala:
            carry = storage.take_carry()
        if update % args.log_frequency == 0:
            if device_thread_id == 0:
                print(f"global_step={global_step}, avg_episodic_return={avg_episodic_return}, rollout_time={np.mean(rollout_time)}")
                print("SPS:", int((global_step - start_step) / (time.time() - start_time)))
            writer.add_scalar("stats/rollout_time", np.mean(rollout_time), global_step)
            writer.add_scalar("charts/avg_episodic_return", avg_episodic_return, global_step)
            writer.add_scalar("charts/avg_episodic_length", np.mean(returned_episode_lengths), global_step)
            writer.add_scalar("stats/params_queue_get_time", np.mean(params_queue_get_time), global_step)
            writer.add_scalar("stats/env_recv_time", env_recv_time, global_step)
            writer.add_scalar("stats/inference_time", inference_time, global_step)
            writer.add_scalar("stats/storage_time", storage_time, global_step)
            writer.add_scalar("stats/d2h_time", d2h_time, global_step)
            writer.add_scalar("stats/env_send_time", env_send_time, global_step)
            writer.add_scalar("stats/rollout_queue_put_time", np.mean(rollout_queue_put_time), global_step)
            writer.add_scalar("charts/SPS", int((global_step - start_step) / (time.time() - start_time)), global_step)
            writer.add_scalar("charts/SPS_update", int(N * args.num_steps * len_actor_device_ids * args.num_actor_threads
                                                       * args.world_size / (time.time() - update_time_start)), global_step)


def train(args: Args, backend, make_env: Callable, writer=None, allreduce=None, on_update: Optional[Callable] = None):
    """The `__main__` of the reference after argument parsing (cleanba_ppo.py:465-751).  `backend` supplies the hot
    path; `allreduce(flat_grad)` sums gradients over the learner devices of other processes (None = single process).
    Returns a SimpleNamespace with the learner handle, the last stats and the measured SPS."""
    writer = writer or _NullWriter()
    tracer = Tracer(f"{args.exp_name} ({args.algo})") if getattr(args, "trace_path", "") else NullTracer()
    args._tracer = tracer
    tracer.thread_name(0, "learner (main thread)")
    key = backend.first_key(args.seed)            # key, network_key, actor_key, critic_key = split(PRNGKey(seed), 4)
    learner = backend.make_learner(args, key, allreduce)
    first_update = first_step = 0
    if getattr(args, "resume_from", ""):
        from .checkpoint import load_train_state
        st = load_train_state(args.resume_from)
        learner.load_train_state(st)
        first_update, first_step = int(st["learner_policy_version"]), int(st["global_step"])
        if first_update >= args.num_updates:
            raise ValueError(f"--resume-from: the saved run already finished {first_update} of {args.num_updates} updates "
                             "(raise --total-timesteps to continue it)")
        # the actor threads continue from the saved (advanced) key instead of replaying the first run's random stream
        key = np.asarray(st["key"], dtype=np.uint32).reshape(2).copy()
    args._resume = (first_update, first_step)
    params_queues, rollout_queues, threads = [], [], []
    stop = threading.Event()
    errors: list = []
    dummy_writer = _NullWriter()

    def get_payload(q):
        while True:
            try:
                return q.get(timeout=0.5)
            except queue.Empty:
                if errors:
                    raise RuntimeError("an actor thread failed") from errors[0]

    for d_idx, d_id in enumerate(args.actor_device_ids):
        device_params = learner.params_for_actor(d_id)
        for thread_id in range(args.num_actor_threads):
            params_queues.append(queue.Queue(maxsize=1))
            rollout_queues.append(queue.Queue(maxsize=1))
            params_queues[-1].put(device_params)
            th = threading.Thread(target=rollout, daemon=True, args=(
                args, backend, make_env, rollout_queues[-1], params_queues[-1],
                writer if d_idx == 0 and thread_id == 0 else dummy_writer,
                d_idx * args.num_actor_threads + thread_id, d_id, stop, key, errors))
            th.start()
            threads.append(th)

    rollout_queue_get_time = deque(maxlen=10)
    learner_policy_version = first_update
    global_step = first_step
    start = time.time()
    result = SimpleNamespace(learner=learner, stats=None, sps=0.0, updates=0, versions=[], update_seconds=[], queue_get_seconds=[],
                             update_done_at=[])
    try:
        while True:
            learner_policy_version += 1
            t0 = time.time()
            payloads = []
            with tracer.span("rollout_queue.get", 0, learner_policy_version=learner_policy_version):
                for d_idx, d_id in enumerate(args.actor_device_ids):
                    for thread_id in range(args.num_actor_threads):
                        (global_step, actor_policy_version, update, sharded, avg_params_queue_get_time,
                         device_thread_id) = get_payload(rollout_queues[d_idx * args.num_actor_threads + thread_id])
                        payloads.append(sharded)
            rollout_queue_get_time.append(time.time() - t0)
            result.queue_get_seconds.append(rollout_queue_get_time[-1])
            training_time_start = time.time()
            with tracer.span("multi_device_update", 0, learner_policy_version=learner_policy_version, actor_policy_version=actor_policy_version):
                stats = learner.update(payloads)      # multi_device_update (cleanba_ppo.py:714-720)
            with tracer.span("params_queue.put", 0, learner_policy_version=learner_policy_version):
                for d_idx, d_id in enumerate(args.actor_device_ids):
                    device_params = learner.params_for_actor(d_id)
                    for thread_id in range(args.num_actor_threads):
                        params_queues[d_idx * args.num_actor_threads + thread_id].put(device_params)
            result.stats, result.updates = stats, learner_policy_version
            result.update_seconds.append(time.time() - training_time_start)     # host time of multi_device_update + publish (enqueue)
            result.update_done_at.append((time.time(), global_step))
            result.versions.append((actor_policy_version, update, learner_policy_version))
            if on_update is not None:
                on_update(learner_policy_version, global_step, stats)
            if learner_policy_version % args.log_frequency == 0:
                s = learner.stats_to_host(stats)
                writer.add_scalar("stats/rollout_queue_get_time", np.mean(rollout_queue_get_time), global_step)
                writer.add_scalar("stats/rollout_params_queue_get_time_diff", np.mean(rollout_queue_get_time) - avg_params_queue_get_time, global_step)
                writer.add_scalar("stats/training_time", time.time() - training_time_start, global_step)
                writer.add_scalar("stats/rollout_queue_size", rollout_queues[-1].qsize(), global_step)
                writer.add_scalar("stats/params_queue_size", params_queues[-1].qsize(), global_step)
                print(global_step, f"actor_policy_version={actor_policy_version}, actor_update={update}, "
                      f"learner_policy_version={learner_policy_version}, training time: {time.time() - training_time_start}s")
                writer.add_scalar("charts/learning_rate", learner.current_lr(), global_step)
                writer.add_scalar("losses/value_loss", s["v_loss"], global_step)
                writer.add_scalar("losses/policy_loss", s["pg_loss"], global_step)
                writer.add_scalar("losses/entropy", s["entropy_loss"], global_step)
                if "approx_kl" in s:
                    writer.add_scalar("losses/approx_kl", s["approx_kl"], global_step)
                writer.add_scalar("losses/loss", s["loss"], global_step)
            if learner_policy_version >= args.num_updates or (args.max_updates and learner_policy_version - first_update >= args.max_updates):
                break
    finally:
        stop.set()
        for q in params_queues:   # unblock actor threads waiting for parameters
            try:
                q.put_nowait(None)
            except queue.Full:
                pass
        for th in threads:        # actor threads leave on `stop`; never let daemon threads die inside CUDA calls at exit
            th.join(timeout=10)
        args.__dict__.pop("_tracer", None)
        args.__dict__.pop("_resume", None)
        if errors:
            raise RuntimeError("an actor thread failed") from errors[0]
    result.sps = (global_step - first_step) / max(time.time() - start, 1e-9)
    result.global_step = global_step
    result.trace_path = tracer.save(args.trace_path) if getattr(args, "trace_path", "") else None
    return result
