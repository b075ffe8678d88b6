"""Developer probe (GPU box): one IMPALA cycle of config 3 (`cleanba_impala.py a0-l0-d1 --local-num-envs 60`, V-trace path):
21 rollout rows (20 new steps + the carried row) x 2 actor threads x 60 envs through get_action (CUDA-graph replays), then
single_device_update (4 contiguous column minibatches of [21,30] = 630 frames: forward, V-trace, backward, clip + RMSProp)
and the parameter publish = 2,400 env steps.  Prints env-steps/s and the per-kernel table of one update."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from cleanba_b200 import agent as ag
from cleanba_b200.learner import ImpalaHyper, ImpalaLearner
from cleanba_b200.params import init_params
from cleanba_b200.prng import first_key

N, TH, T = 60, 2, 20
Bl = N * TH
dev = torch.device("cuda:0")
L = ImpalaLearner(dev, ImpalaHyper(), T1=T + 1, Bl=Bl)
L.ctx.set_params(init_params(1))
actors, graphed = [], []
for th in range(TH):
    a = ag.Context(dev, max_batch=N, algo=ag.CB_ALGO_IMPALA)
    L.ctx.publish_to(a)
    torch.cuda.synchronize()
    actors.append(a)
    graphed.append(ag.GraphedActor(a, N, ag.key_tensor(first_key(1), dev), want_logits=True))
rng = np.random.Generator(np.random.PCG64(1))
pool = torch.from_numpy(rng.integers(0, 256, (64, N, 4, 84, 84), dtype=np.uint8)).to(dev)
obs = torch.zeros(T + 1, Bl, 4, 84, 84, dtype=torch.uint8, device=dev)
actions = torch.zeros(T + 1, Bl, dtype=torch.int32, device=dev)
logitss = torch.zeros(T + 1, Bl, 18, device=dev)
rewards = (torch.randint(0, 3, (T + 1, Bl), device=dev).float() - 1) * (torch.rand(T + 1, Bl, device=dev) < 0.1)
dones = torch.rand(T + 1, Bl, device=dev) < 1 / 500
first = torch.zeros(T + 1, Bl, dtype=torch.bool, device=dev)
cursor = 0


def cycle():
    global cursor
    main = torch.cuda.current_stream(dev)
    for g in graphed:
        g.stream.wait_stream(main)
    for t in range(1, T + 1):                       # row 0 is the row carried over from the previous rollout (cleanba_impala.py:327-329)
        for th, g in enumerate(graphed):
            c = slice(th * N, (th + 1) * N)
            g.step(pool[cursor % 64]); cursor += 1
            with torch.cuda.stream(g.stream):
                obs[t, c].copy_(g.obs, non_blocking=True)
                actions[t, c].copy_(g.action, non_blocking=True)
                logitss[t, c].copy_(g.logits, non_blocking=True)
    for g in graphed:
        main.wait_stream(g.stream)
    st = L.update(obs, dones, actions, logitss, rewards, first)
    for a in actors:
        L.ctx.publish_to(a)
    return st


for _ in range(3):
    cycle()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 20
e0.record()
for _ in range(K):
    st = cycle()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
print(f"IMPALA config 3: {ms:.3f} ms per cycle of {T * Bl} env steps -> {T * Bl / ms * 1e3:.0f} env-steps/s; stats {st.cpu().numpy()}")
t0 = time.perf_counter(); L.update(obs, dones, actions, logitss, rewards, first); torch.cuda.synchronize()
print(f"update alone: {(time.perf_counter() - t0) * 1e3:.3f} ms")
L.ctx.profile(True); L.update(obs, dones, actions, logitss, rewards, first); rep = L.ctx.profile_report(); L.ctx.profile(False)
tot = sum(r["ms"] for r in rep)
for r in sorted(rep, key=lambda r: -r["ms"])[:14]:
    print(f"  {r['name']:38s} calls={r['calls']:3d} ms={r['ms']:8.4f} ({100 * r['ms'] / tot:5.1f}%)")
print(f"  sum of kernels {tot:.3f} ms (4 minibatches)")
