"""Developer probe (GPU box): PPO config-2 cycle with the GPU spatially partitioned between actor and learner (green contexts)
and the rollout of update k+1 pipelined beside the learner step of update k, vs the serial cycle of bench.py."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import bench
from cleanba_b200 import agent as ag
from cleanba_b200.partition import SmPartition
from cleanba_b200.prng import first_key

actor_sms = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda:0")
cyc = bench.Cycle("cuda:0", 1, None)
for _ in range(2):
    cyc.step(False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    cyc.step(False)
e1.record(); torch.cuda.synchronize()
print(f"serial cycle (bench.py): {e0.elapsed_time(e1) / 3:.1f} ms")

use_part = actor_sms > 0
if use_part:
    part = SmPartition(dev, actor_sms)
    print(f"partition: actor {part.actor_sms} SMs, learner {part.learner_sms} SMs")
    s_learn = part.learner_stream()
    a_streams = [part.actor_stream() for _ in range(bench.N_THREADS)]
    cyc.learner.ctx.set_sm_budget(part.learner_sms)
else:
    s_learn = torch.cuda.Stream(dev)
    a_streams = [torch.cuda.Stream(dev, priority=-1) for _ in range(bench.N_THREADS)]
N, T, Bl = bench.N_ENVS, bench.WORKLOADS["ppo"]["T"], cyc.Bl
actors, graphed = [], []
for th in range(bench.N_THREADS):
    a = ag.Context(dev, max_batch=N, train=False)
    if use_part:
        a.set_sm_budget(part.actor_sms)
    cyc.learner.ctx.publish_to(a)
    torch.cuda.synchronize()
    actors.append(a)
    graphed.append(ag.GraphedActor(a, N, ag.key_tensor(first_key(1), dev), stream=a_streams[th]))
# two rollout storages: rollout k+1 is written while update k reads storage k
store = [dict(obs=torch.zeros(T, Bl, 4, 84, 84, dtype=torch.uint8, device=dev), actions=torch.zeros(T, Bl, dtype=torch.int32, device=dev),
              logprobs=torch.zeros(T, Bl, device=dev), values=torch.zeros(T, Bl, device=dev)) for _ in range(2)]
cursor = 0


def rollout(S, after_update_event):
    """enqueue: [wait params] publish -> 128 steps x 2 threads into storage S; returns the 'rollout done' events"""
    global cursor
    evs = []
    for th, g in enumerate(graphed):
        with torch.cuda.stream(g.stream):
            if after_update_event is not None:
                g.stream.wait_event(after_update_event)
            cyc.learner.ctx.publish_to(actors[th])
    for t in range(T):
        for th, g in enumerate(graphed):
            c = slice(th * N, (th + 1) * N)
            g.step(cyc.dev_pool[cursor % 256]); cursor += 1
            with torch.cuda.stream(g.stream):
                S["obs"][t, c].copy_(g.obs, non_blocking=True)
                S["actions"][t, c].copy_(g.action, non_blocking=True)
                S["logprobs"][t, c].copy_(g.logprob, non_blocking=True)
                S["values"][t, c].copy_(g.value, non_blocking=True)
    for g in graphed:
        ev = torch.cuda.Event(); ev.record(g.stream); evs.append(ev)
    return evs


def update(S, rollout_events):
    with torch.cuda.stream(s_learn):
        for ev in rollout_events:
            s_learn.wait_event(ev)
        cyc.learner.update(S["obs"], cyc.done_pool[0], S["actions"], S["logprobs"], S["values"], cyc.rew_pool[0], S["obs"][0],
                           cyc.next_done, cyc.lkey)
        ev = torch.cuda.Event(); ev.record(s_learn)
    return ev


def run(iters):
    upd_ev = None
    rv = rollout(store[0], None)
    for i in range(iters):
        nxt = rollout(store[(i + 1) % 2], upd_ev)      # rollout i+1 (parameters of update i-1) beside ...
        upd_ev = update(store[i % 2], rv)              # ... update i
        rv = nxt
        # the storage written by rollout i+2 is the one update i reads: order them
        for g in graphed:
            g.stream.wait_event(upd_ev)
    torch.cuda.synchronize()


run(2)
t0 = time.perf_counter()
K = 5
run(K)
ms = (time.perf_counter() - t0) * 1e3 / K
print(f"pipelined cycle ({'green contexts' if use_part else 'plain streams, actor priority -1'}): {ms:.1f} ms per update "
      f"-> {15360 / ms * 1e3:.0f} env-steps/s")
p = cyc.learner.ctx.get_params()
print("finite params:", bool(torch.isfinite(p).all()))
