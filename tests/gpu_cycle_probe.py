"""Developer probe (GPU box): where one bench cycle (bench.Cycle.step) spends its time -- rollout (256 actor steps on two
streams), the PPO update, the parameter publish -- as host time to enqueue vs device time to finish."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

cyc = bench.Cycle("cuda:0", 1, None)
for _ in range(2):
    cyc.step(False)
torch.cuda.synchronize()


def phase(fn, reps=1):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    return (t1 - t0) * 1e3 / reps, (t2 - t0) * 1e3 / reps


def rollout(nthreads=2, copies=True):
    main = torch.cuda.current_stream(cyc.dev)
    gs = cyc.graphed[:nthreads]
    for g in gs:
        g.stream.wait_stream(main)
    for t in range(bench.T_STEPS):
        for th, g in enumerate(gs):
            c = slice(th * bench.N_ENVS, (th + 1) * bench.N_ENVS)
            g.step(cyc.dev_pool[cyc.cursor % 256])
            cyc.cursor += 1
            if copies:
                with torch.cuda.stream(g.stream):
                    cyc.obs[t, c].copy_(g.obs, non_blocking=True)
                    cyc.actions[t, c].copy_(g.action, non_blocking=True)
                    cyc.logprobs[t, c].copy_(g.logprob, non_blocking=True)
                    cyc.values[t, c].copy_(g.value, non_blocking=True)
    for g in gs:
        main.wait_stream(g.stream)


def update():
    cyc.learner.update(cyc.obs, cyc.done_pool[0], cyc.actions, cyc.logprobs, cyc.values, cyc.rew_pool[0], cyc.obs[0],
                       cyc.next_done, cyc.lkey)


def publish():
    for a in cyc.actors:
        cyc.learner.ctx.publish_to(a)


def replay_only():
    g = cyc.graphed[0]
    with torch.cuda.stream(g.stream):
        for _ in range(256):
            g.graph.replay()


def gae_part():
    c = cyc.learner.ctx
    _, nv = c.policy_value(cyc.obs[0])
    c.gae(cyc.rew_pool[0], cyc.values, cyc.done_pool[0], nv, cyc.next_done, 0.99, 0.95, 4)
    for _ in range(4):
        sub = c.split_key(cyc.lkey)
        c.permutation(sub, bench.T_STEPS * cyc.Bl)


for name, fn in [("rollout 2 threads (256 steps + storage copies)", rollout),
                 ("rollout 2 threads, no storage copies", lambda: rollout(2, False)),
                 ("rollout 1 thread (128 steps)", lambda: rollout(1)),
                 ("256 graph replays on one stream", replay_only),
                 ("update (GAE + 16 minibatches)", update),
                 ("bootstrap + GAE + 4 permutations", gae_part),
                 ("publish x2", publish),
                 ("full cycle.step(False)", lambda: cyc.step(False)),
                 ("full cycle.step(True) [e2e]", lambda: cyc.step(True))]:
    h, d = phase(fn)
    h, d = phase(fn)
    print(f"{name:52s} host-enqueue {h:8.2f} ms   device-done {d:8.2f} ms")
