"""Developer diagnostic: first two IMPALA payloads, CUDA backend vs oracle backend through the same plumbing."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from cleanba_b200.cuda_backend import CudaBackend
from cleanba_b200.sebulba import Args, derive_sizes, impala_defaults, train
from cleanba_b200.envs import SyntheticAtari
from oracle.backend import OracleBackend

def make_env(env_id, seed, n):
    return lambda: SyntheticAtari(n, seed=seed, pool_batches=8)

def args():
    a = impala_defaults(Args(local_num_envs=8, num_actor_threads=2, num_minibatches=2, total_timesteps=10**6, log_frequency=1000, max_updates=2))
    a.num_steps = 4; a.concurrency = False
    return derive_sizes(a, 1)

def record(backend, store):
    orig = backend.make_actor
    def mk(*a, **k):
        actor = orig(*a, **k)
        o2 = actor.shard_to_learners
        def sh(storage, *aa, **kk):
            out = o2(storage, *aa, **kk)
            f = lambda x: x.detach().cpu().numpy().copy() if torch.is_tensor(x) else np.array(x).copy()
            store.append({k: f(v) for k, v in out[0].items() if k != "event"})
            return out
        actor.shard_to_learners = sh
        return actor
    backend.make_actor = mk
    return backend

pc, po = [], []
sc, so = [], []
train(args(), record(CudaBackend(), pc), make_env, on_update=lambda v, g, st: sc.append(st.cpu().numpy()))
train(args(), record(OracleBackend(), po), make_env, on_update=lambda v, g, st: so.append(np.asarray(st)))
print("payloads", len(pc), len(po))
# payload order = (thread, update) interleaved by thread scheduling; match by content of host rewards
for i, c in enumerate(pc):
    best = min(range(len(po)), key=lambda j: np.abs(po[j]["obs"].astype(int) - c["obs"].astype(int)).sum())
    o = po[best]
    print(i, "->", best, "actions equal:", np.array_equal(c["actions"], o["actions"]), "n diff", int((c["actions"] != o["actions"]).sum()),
          "logits maxdiff", float(np.abs(c["logitss"] - o["logitss"]).max()), "rewards eq", np.array_equal(c["rewards"], o["rewards"]),
          "dones eq", np.array_equal(c["dones"], o["dones"]), "first eq", np.array_equal(c["firststeps"], o["firststeps"]))
print("stats cuda", sc); print("stats oracle", so)
