#!/usr/bin/env python
"""Benchmark of the Sebulba hot path (BASELINE.json: env-steps/sec, Breakout-v5-shaped synthetic 84x84x4 frames).

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA, libcleanba_b200)
  python bench.py --impl reference --gpus N --steps K ...  the reference path's CPU restatement (oracle/) on host cores

One STEP = one full pass of the hot path over one batch = one PPO update cycle of config[1]
(`cleanba_ppo.py a0-l0-d1 --local-num-envs 60`): 128 rollout steps x 2 actor threads x 60 envs through the actor's
get_action_and_value, then single_device_update (GAE + 4 epochs x 4 shuffled minibatches of 3840: forward, loss,
backward, [allreduce], clip + Adam) and the parameter publish back to the actor = 15,360 env steps.
With N > 1 every rank runs that cycle on its own GPU (a0-l0 per process, `--distributed` d=N, weak scaling) and the
gradients are averaged with one NCCL allreduce on the flat gradient buffer per minibatch.

`value`  : frames already resident in HBM (a device pool larger than L2, cycled).
`e2e`    : same cycle through the public API with HOST frames: every actor step copies its [60,4,84,84] uint8 batch
           from pinned host memory and reads the int32 actions back (the reference's per-step sync, cleanba_ppo.py:313-317),
           every update reads the loss scalars back.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_ENVS, N_THREADS, N_MB, N_EPOCHS = 60, 2, 4, 4
FWD_FLOP = 108.46e6                     # trunk + heads forward per frame
# --workload: BASELINE.json configs[1] (PPO, the configuration the metric is quoted on; default) and configs[2] (IMPALA, V-trace)
WORKLOADS = {
    "ppo": dict(script="cleanba_ppo", T=128, ref_T=8, flop_per_env_step=FWD_FLOP * (1 + 3 * N_EPOCHS)),
    "impala": dict(script="cleanba_impala", T=20, ref_T=20, flop_per_env_step=FWD_FLOP * (1 + 3 * 21 / 20)),
}


def workload_name(wl, T):
    return (f"{WORKLOADS[wl]['script']} a0-l0-d1 Breakout-v5-shaped synthetic frames, local_num_envs={N_ENVS}, {N_THREADS} actor threads, "
            f"num_steps={T}")


DTYPE = "fp16x2 carrier (two fp16 tensor-core operand planes per tensor = 22 significant bits; fp32 accumulate, fp32 master weights and optimizer state)"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf=d.get("bf16_tflops_sustained", d["bf16_tflops"]), tf_burst=d["bf16_tflops"], src="measured")
    return dict(hbm=6650.0, tf=1400.0, tf_burst=1590.0, src="fallback")


class ClockSampler:
    """nvidia-smi clock / throttle-reason sampling during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.samples, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            try:
                sm.append(float(s[0])); mx.append(float(s[1]))
                for nme, v in zip(names, s[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ reference / CPU arm
def cpu_sample(T_s, workload="ppo", threads=N_THREADS, n_envs=N_ENVS, seed=1):
    """One bounded sample of the workload on the CPU oracle: a T_s-step rollout of `threads` x `n_envs` envs through
    get_action_and_value / get_action, then a full single_device_update on it.  Returns env steps processed."""
    from oracle import impala as oimpala, network as net, ppo as oppo, threefry as tf
    st = cpu_sample.state
    if st is None:
        rng = np.random.Generator(np.random.PCG64(seed))
        st = cpu_sample.state = dict(params=net.init_params(seed), rng=rng, key=tf.split(tf.PRNGKey(seed), 4)[0])
    rng, params = st["rng"], st["params"]
    Bl = threads * n_envs
    keys = [st["key"].copy() for _ in range(threads)]
    if workload == "impala":
        T1 = T_s + 1                                  # row 0 is the row carried over from the previous rollout
        obs = rng.integers(0, 256, (T1, Bl, 4, 84, 84), dtype=np.uint8)
        act = np.zeros((T1, Bl), np.int32); lg = np.zeros((T1, Bl, 18), np.float32)
        for t in range(1, T1):
            for th in range(threads):
                c = slice(th * n_envs, (th + 1) * n_envs)
                _, a, l, keys[th] = oimpala.get_action(params, obs[t, c], keys[th])
                act[t, c], lg[t, c] = a, l
        shard = oimpala.Shard(obs=obs, dones=rng.random((T1, Bl)) < 0.002, actions=act, logitss=lg,
                              rewards=rng.choice(np.array([-1, 0, 1], np.float32), size=(T1, Bl), p=[.05, .9, .05]),
                              firststeps=np.zeros((T1, Bl), bool))
        learner = oimpala.ImpalaLearner(params, oimpala.ImpalaConfig(num_minibatches=N_MB))
        learner.update([shard])
        st["params"] = learner.params
        return T_s * Bl
    obs = np.zeros((T_s, Bl, 4, 84, 84), np.uint8)
    act = np.zeros((T_s, Bl), np.int32); lp = np.zeros((T_s, Bl), np.float32); val = np.zeros((T_s, Bl), np.float32)
    for t in range(T_s):
        for th in range(threads):
            c = slice(th * n_envs, (th + 1) * n_envs)
            o = rng.integers(0, 256, (n_envs, 4, 84, 84), dtype=np.uint8)
            _, a, l, v, keys[th], _ = oppo.get_action_and_value(params, o, keys[th])
            obs[t, c], act[t, c], lp[t, c], val[t, c] = o, a, l, v
    shard = oppo.Shard(obs=obs, dones=rng.random((T_s, Bl)) < 0.002, actions=act, logprobs=lp, values=val,
                       rewards=rng.choice(np.array([-1, 0, 1], np.float32), size=(T_s, Bl), p=[.05, .9, .05]),
                       next_obs=rng.integers(0, 256, (Bl, 4, 84, 84), dtype=np.uint8), next_done=np.zeros(Bl, bool))
    learner = oppo.PPOLearner(params, oppo.PPOConfig(num_minibatches=N_MB, update_epochs=N_EPOCHS))
    learner.update([shard], st["key"])
    st["params"] = learner.params
    return T_s * Bl


cpu_sample.state = None


def run_cpu_arm(steps, warmup, workload="ppo", threads=None, T_s=None):
    """The CPU arm.  The sample is FIXED (no calibration, identical in every run): a `ref_T`-step rollout (PPO: 8 of the
    workload's 128 steps; IMPALA: the full 20) x 2 actor threads x 60 envs + one complete learner update on it per bench step."""
    import torch
    cores = threads or os.cpu_count() or 1
    torch.set_num_threads(cores)
    T_s = T_s or WORKLOADS[workload]["ref_T"]
    for _ in range(warmup):
        cpu_sample(T_s, workload)
    t0 = time.time(); n = 0
    for _ in range(steps):
        n += cpu_sample(T_s, workload)
    dt = time.time() - t0
    upd = f"one PPO update ({N_EPOCHS} epochs x {N_MB} minibatches)" if workload == "ppo" else f"one IMPALA update ({N_MB} minibatches of [{T_s + 1},{N_ENVS * N_THREADS // N_MB}])"
    return dict(value=n / dt, cores=cores, T_s=T_s, ms_per_step=1e3 * dt / max(steps, 1),
                sample=f"{T_s}-step rollout x {N_THREADS} actor threads x {N_ENVS} envs + {upd} "
                       f"= {T_s * N_THREADS * N_ENVS} env steps per bench step (fixed sample, no calibration)")


# ------------------------------------------------------------------------------------------------ our arm
class Cycle:
    """One actor device (2 logical actor threads, one CUDA-graphed step each) + one learner replica on one GPU: the a0-l0
    topology of BASELINE configs[1] (PPO) / configs[2] (IMPALA).  One step() = rollout of T steps + single_device_update +
    parameter publish."""

    def __init__(self, device, world, allreduce, workload="ppo", seed=1, network="impala_resnet"):
        import torch
        from cleanba_b200 import agent as ag
        from cleanba_b200.learner import ImpalaHyper, ImpalaLearner, PPOHyper, PPOLearner
        from cleanba_b200.params import init_params
        from cleanba_b200.prng import first_key
        from cleanba_b200.lib import MODELS
        model = MODELS[network]
        self.torch, self.ag = torch, ag
        self.dev = d = torch.device(device)
        self.impala = workload == "impala"
        self.T = T = WORKLOADS[workload]["T"]
        self.rows = rows = T + 1 if self.impala else T          # IMPALA: row 0 is carried over from the previous rollout
        Bl = self.Bl = N_ENVS * N_THREADS
        if self.impala:
            self.learner = ImpalaLearner(d, ImpalaHyper(), T1=rows, Bl=Bl, world_learners=world, allreduce=allreduce, model=model)
        else:
            self.learner = PPOLearner(d, PPOHyper(), T=T, Bl=Bl, world_learners=world, allreduce=allreduce, model=model)
        self.learner.ctx.set_params(init_params(seed, model=model))
        # one actor context + CUDA-graphed step per actor thread, each on its own stream (the reference runs
        # num_actor_threads = 2 Python threads per actor device, cleanba_ppo.py:670-686)
        self.actors, self.graphed = [], []
        for th in range(N_THREADS):
            a = ag.Context(d, max_batch=N_ENVS, train=False, algo=ag.CB_ALGO_IMPALA if self.impala else ag.CB_ALGO_PPO, model=model)
            self.learner.ctx.publish_to(a)
            torch.cuda.synchronize()
            key = ag.key_tensor(first_key(seed), d)
            self.actors.append(a)
            self.graphed.append(ag.RolloutActor(a, N_ENVS, key, want_logits=self.impala))   # writes straight into the storage rows
        self.obs = torch.zeros(rows, Bl, 4, 84, 84, dtype=torch.uint8, device=d)
        self.actions = torch.zeros(rows, Bl, dtype=torch.int32, device=d)
        if self.impala:
            self.logitss = torch.zeros(rows, Bl, 18, dtype=torch.float32, device=d)
            self.first = torch.zeros(rows, Bl, dtype=torch.bool, device=d)
        else:
            self.logprobs = torch.zeros(rows, Bl, dtype=torch.float32, device=d)
            self.values = torch.zeros(rows, Bl, dtype=torch.float32, device=d)
        g = torch.Generator(device="cpu"); g.manual_seed(seed)
        self.rew_pool = (torch.multinomial(torch.tensor([.05, .9, .05]), 8 * rows * Bl, True, generator=g).float() - 1).reshape(8, rows, Bl).to(d)
        self.done_pool = (torch.rand(8, rows, Bl, generator=g) < 1 / 500).to(d)
        self.next_done = torch.zeros(Bl, dtype=torch.bool, device=d)
        self.lkey = ag.key_tensor(first_key(seed), d)
        self.act_host = [torch.empty(N_ENVS, dtype=torch.int32).pin_memory() for _ in range(N_THREADS)]
        # frame pools: 256 distinct [60,4,84,84] batches = 433 MB (> 126 MB L2); host copy pinned for the e2e path
        rng = np.random.Generator(np.random.PCG64(seed))
        host = torch.empty(256, N_ENVS, 4, 84, 84, dtype=torch.uint8).pin_memory()
        host.numpy()[...] = rng.integers(0, 256, size=tuple(host.shape), dtype=np.uint8)
        self.host_pool = host
        self.dev_pool = host.to(d)
        self.cursor = 0
        self.h2d = self.d2h = 0

    def step(self, e2e: bool):
        torch = self.torch
        pool = self.host_pool if e2e else self.dev_pool
        main = torch.cuda.current_stream(self.dev)
        row0 = 1 if self.impala else 0
        for th, g in enumerate(self.graphed):
            g.stream.wait_stream(main)                 # new parameters (publish) are visible before the rollout starts
            c = slice(th * N_ENVS, (th + 1) * N_ENVS)
            if self.impala:                            # this thread's storage columns; the carried row first (cleanba_impala.py:327-329)
                with torch.cuda.stream(g.stream):
                    for buf in (self.obs, self.actions, self.logitss):
                        buf[0, c].copy_(buf[self.rows - 1, c], non_blocking=True)
                g.begin(self.obs[:, c], self.actions[:, c], logits=self.logitss[:, c], first_row=1)
            else:
                g.begin(self.obs[:, c], self.actions[:, c], self.logprobs[:, c], self.values[:, c])
        for t in range(row0, self.rows):
            for th, g in enumerate(self.graphed):
                c = slice(th * N_ENVS, (th + 1) * N_ENVS)
                if e2e and t > row0:
                    g.stream.synchronize()             # np.array(action) of this thread's previous step: the per-step sync of
                                                       # cleanba_ppo.py:317 (each actor thread waits for ITS OWN step only)
                g.step(pool[self.cursor % 256], t)     # frames -> storage row t (H2D from pinned memory when e2e) + graph replay:
                self.cursor += 1                       # the step reads row t and writes action / logprob / value into row t
                if e2e:
                    with torch.cuda.stream(g.stream):
                        self.act_host[th].copy_(self.actions[t, c], non_blocking=True)
                    self.h2d += N_ENVS * 4 * 84 * 84; self.d2h += N_ENVS * 4
        if e2e:
            for g in self.graphed:
                g.stream.synchronize()                 # actions of the last step
        for g in self.graphed:
            main.wait_stream(g.stream)
        k = (self.cursor // 256) % 8
        if self.impala:
            stats = self.learner.update(self.obs, self.done_pool[k], self.actions, self.logitss, self.rew_pool[k], self.first)
        else:
            stats = self.learner.update(self.obs, self.done_pool[k], self.actions, self.logprobs, self.values, self.rew_pool[k],
                                        self.obs[0], self.next_done, self.lkey)
        for a in self.actors:
            self.learner.ctx.publish_to(a)                             # params_queue.put(device_params) (cleanba_ppo.py:721-725)
        if e2e:
            _ = stats.cpu(); self.d2h += 4 * stats.numel()
        return stats


def run_our_arm(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    allreduce = (lambda g: dist.all_reduce(g)) if world > 1 else None
    from cleanba_b200 import lib
    wl = args.workload
    T_STEPS = WORKLOADS[wl]["T"]
    cyc = Cycle(f"cuda:{local_rank}", world, allreduce, workload=wl, network=args.network)
    if world > 1:   # identical initial parameters on every learner (the reference relies on equal seeds, cleanba_ppo.py:468)
        pv = cyc.learner.ctx.params_view(); dist.broadcast(pv, 0); cyc.learner.ctx.refresh_weights()
        for a in cyc.actors:
            cyc.learner.ctx.publish_to(a)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(e2e, steps, profile=False):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        if profile:
            cyc.learner.ctx.profile(True)
        l0 = lib.load().cb_launch_count()
        r0 = sum(g.replays for g in cyc.graphed)
        ev0.record()
        for _ in range(steps):
            cyc.step(e2e)
        ev1.record()
        barrier()
        # kernels of libcleanba_b200 in the timed region: eager launches + kernels inside the CUDA-graph replays
        launches = lib.load().cb_launch_count() - l0 + (sum(g.replays for g in cyc.graphed) - r0) * cyc.graphed[0].kernels_per_replay
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = t.item()
        return ms, launches

    for _ in range(max(args.warmup, 3)):
        cyc.step(False)
    cyc.step(True)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms, launches = timed(False, args.steps)                      # `value`: no instrumentation in the timed region
    ms_prof, _ = timed(False, args.steps, profile=True)           # same region again with per-kernel CUDA-event brackets
    rep = cyc.learner.ctx.profile_report()        # (the actor steps are CUDA-graph replays: no per-kernel brackets inside)
    cyc.learner.ctx.profile(False)
    cyc.h2d = cyc.d2h = 0
    ms_e2e, _ = timed(True, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    env_steps = T_STEPS * N_THREADS * N_ENVS * world
    value = env_steps * args.steps / (ms * 1e-3)
    e2e_value = env_steps * args.steps / (ms_e2e * 1e-3)
    if rank != 0:
        return
    peaks = load_peaks()
    agg = {}
    for r in rep:
        a = agg.setdefault(r["name"], dict(ms=0.0, calls=0, records=0, flops=0.0, bytes=0.0, abytes=0.0))
        for k in ("ms", "calls", "records", "flops", "bytes", "abytes"):
            a[k] += r[k]
    top = max(agg.items(), key=lambda kv: kv[1]["ms"])
    name, a = top
    # `achieved` uses the SURVEY 8(d) ALGORITHMIC bytes (every operand tensor of the operator once, unpadded fp32), not the
    # bytes of this library's own storage formats; those (`moved_bytes_per_launch`) and the ncu DRAM traffic are reported beside it.
    ai = a["flops"] / max(a["abytes"], 1.0)
    balance = peaks["tf"] * 1e12 / (peaks["hbm"] * 1e9)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(name)
    nrec = max(a["records"], 1)
    if a["flops"] > 0 and ai >= balance:
        roof = dict(bound="tensor", achieved=a["flops"] / (a["ms"] * 1e-3) / 1e12, peak=peaks["tf"], unit="TFLOP/s")
    else:
        roof = dict(bound="hbm", achieved=a["abytes"] / (a["ms"] * 1e-3) / 1e9, peak=peaks["hbm"], unit="GB/s")
    roof.update(frac=roof["achieved"] / roof["peak"], traffic=traffic, kernel=name, kernel_share_of_step=a["ms"] / ms_prof,
                algorithmic_bytes_per_launch=a["abytes"] / nrec, algorithmic_bytes_basis="fp32, each operand tensor once, unpadded (SURVEY 8d)",
                moved_bytes_per_launch=a["bytes"] / nrec, moved_gbs=a["bytes"] / (a["ms"] * 1e-3) / 1e9,
                traffic_over_algorithmic=(traffic / (a["abytes"] / nrec)) if traffic else None,
                ms_per_launch=a["ms"] / nrec, profiled_ms_per_step=ms_prof / args.steps,
                peak_source=peaks["src"], launches=a["calls"],
                algorithmic_tflops=a["flops"] / (a["ms"] * 1e-3) / 1e12 if a["flops"] else 0.0,
                tensor_frac=(a["flops"] / (a["ms"] * 1e-3) / 1e12 / peaks["tf"]) if a["flops"] else 0.0,
                kernels={k: round(v["ms"] / args.steps, 4) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])[:12]})
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        r = run_cpu_arm(4 if wl == "ppo" else 2, 0, workload=wl)
        cpu = dict(value=r["value"], unit="env-steps/s", cores=r["cores"], kind="port", sample=r["sample"] +
                   " (CPU restatement of the reference path in PyTorch-CPU fp32, not JAX: jax/flax/optax are not installable here)")
        # the reference pins XLA:CPU to ONE thread (XLA_FLAGS intra_op_parallelism_threads=1, cleanba_ppo.py:28): same sample, 1 thread
        r1 = run_cpu_arm(1, 0, workload=wl, threads=1, T_s=2)
        cpu["single_thread"] = dict(value=r1["value"], cores=1, sample=r1["sample"])
    out = {
        "metric": "env-steps/sec (Breakout-v5 84x84x4, synthetic frames)", "value": value, "unit": "env-steps/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": DTYPE,
        "data": "synthetic", "impl": "ours",
        "config": {"workload": workload_name(wl, T_STEPS) + ("" if args.network == "impala_resnet" else f", network={args.network}"),
                   "env_steps_per_bench_step": env_steps, "topology": f"a0-l0-d{world}",
                   "cache": "inputs larger than L2: 433 MB device frame pool cycled, 433 MB rollout storage per update",
                   "tensor_roof_frac_end_to_end": value * WORKLOADS[wl]["flop_per_env_step"] / (peaks["tf"] * 1e12 * world)},
        "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": cyc.h2d // args.steps,
                "d2h_bytes_per_step": cyc.d2h // args.steps, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches), "roofline": roof, "clocks": clocks,
    }
    if cpu is not None:
        out["cpu_baseline"] = cpu
    _emit(out)
    if world > 1:
        dist.destroy_process_group()


def _emit(obj):
    """The ONE JSON line on the real stdout (fd saved before libraries such as NCCL could write banners to it)."""
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)                 # anything else written to stdout (NCCL version banner, library prints) goes to stderr
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="ppo", choices=sorted(WORKLOADS), help="ppo = BASELINE configs[1] (default), impala = configs[2]")
    ap.add_argument("--network", default="impala_resnet", choices=["impala_resnet", "nature_cnn"],
                    help="trunk: the reference's IMPALA-ResNet (default, the configuration the metric is quoted on) or its legacy Nature-CNN")
    args = ap.parse_args()
    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return
        r = run_cpu_arm(args.steps, args.warmup, workload=args.workload)
        _emit({
            "impl": "reference", "metric": "env-steps/sec (Breakout-v5 84x84x4, synthetic frames)", "value": r["value"],
            "unit": "env-steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            # the ACTUAL rollout length of the timed sample (PPO: 8 of the workload's 128 steps, fixed; IMPALA: the full 20)
            "config": {"workload": workload_name(args.workload, r["T_s"]), "env_steps_per_bench_step": r["T_s"] * N_THREADS * N_ENVS,
                       "sample_of": workload_name(args.workload, WORKLOADS[args.workload]["T"])},
            "cpu_baseline": {"value": r["value"], "unit": "env-steps/s", "cores": r["cores"], "kind": "port",
                             "sample": r["sample"] + " (PyTorch-CPU restatement of the reference path; the JAX reference cannot be installed here)"},
            "e2e": {"value": r["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        return
    run_our_arm(args)


if __name__ == "__main__":
    main()
