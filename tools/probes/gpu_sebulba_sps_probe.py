"""Developer probe (GPU box): the real concurrent Sebulba pipeline (actor threads + queues + learner thread of
cleanba_b200.sebulba.train on the CUDA backend, synthetic Atari env) at config 2 / config 3 shapes; prints the steady-state
SPS (the reference's charts/SPS definition) between the 3rd and the last update."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from cleanba_b200.cuda_backend import CudaBackend
from cleanba_b200.sebulba import Args, derive_sizes, impala_defaults, train
from cleanba_b200.envs import SyntheticAtari

algo = sys.argv[1] if len(sys.argv) > 1 else "ppo"
updates = int(sys.argv[2]) if len(sys.argv) > 2 else 10


def make_env(env_id, seed, n):
    return lambda: SyntheticAtari(n, seed=seed, pool_batches=64)


a = Args(local_num_envs=60, num_actor_threads=2, total_timesteps=10 ** 9, log_frequency=10 ** 6, max_updates=updates)
if algo == "impala":
    a = impala_defaults(a)
a.concurrency = True
a = derive_sizes(a, 1)
marks = []
res = train(a, CudaBackend(), make_env, on_update=lambda v, gs, st: marks.append((time.perf_counter(), gs, float(st[0]))))
torch.cuda.synchronize()
t_end = time.perf_counter()
(t0, g0, _), (t1, g1, l1) = marks[2], marks[-1]
print(f"{algo}: {len(marks)} updates, steady-state SPS = {(g1 - g0) / (t1 - t0):.0f} env-steps/s "
      f"({(t1 - t0) / (len(marks) - 3) * 1e3:.1f} ms per update), last loss {l1:.4f}")
