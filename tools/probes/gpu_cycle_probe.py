"""Developer probe (GPU box): where one bench cycle (bench.Cycle.step) spends its time -- rollout (256 actor steps on two
streams), the PPO update, the parameter publish -- as host time to enqueue vs device time to finish."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
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


def rollout(nthreads=2, e2e=False):
    main = torch.cuda.current_stream(cyc.dev)
    gs = cyc.graphed[:nthreads]
    pool = cyc.host_pool if e2e else cyc.dev_pool
    for th, g in enumerate(gs):
        g.stream.wait_stream(main)
        c = slice(th * bench.N_ENVS, (th + 1) * bench.N_ENVS)
        g.begin(cyc.obs[:, c], cyc.actions[:, c], cyc.logprobs[:, c], cyc.values[:, c])
    for t in range(bench.WORKLOADS["ppo"]["T"]):
        for th, g in enumerate(gs):
            c = slice(th * bench.N_ENVS, (th + 1) * bench.N_ENVS)
            g.step(pool[cyc.cursor % 256], t)
            cyc.cursor += 1
            if e2e:
                with torch.cuda.stream(g.stream):
                    cyc.act_host[th].copy_(cyc.actions[t, c], non_blocking=True)
        if e2e:
            for g in gs:
                g.stream.synchronize()
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
    g.begin(cyc.obs[:, :bench.N_ENVS], cyc.actions[:, :bench.N_ENVS], cyc.logprobs[:, :bench.N_ENVS], cyc.values[:, :bench.N_ENVS])
    with torch.cuda.stream(g.stream):
        for r in range(256):
            if r == 128:      # the cursor must stay inside the storage
                g.begin(cyc.obs[:, :bench.N_ENVS], cyc.actions[:, :bench.N_ENVS], cyc.logprobs[:, :bench.N_ENVS], cyc.values[:, :bench.N_ENVS])
            g.graph.replay()


def gae_part():
    c = cyc.learner.ctx
    _, nv = c.policy_value(cyc.obs[0])
    c.gae(cyc.rew_pool[0], cyc.values, cyc.done_pool[0], nv, cyc.next_done, 0.99, 0.95, 4)
    for _ in range(4):
        sub = c.split_key(cyc.lkey)
        c.permutation(sub, bench.WORKLOADS["ppo"]["T"] * cyc.Bl)


for name, fn in [("rollout 2 threads (256 steps into storage rows)", rollout),
                 ("rollout 2 threads, host frames + per-step sync", lambda: rollout(2, True)),
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
