"""Developer probe (GPU box, >= 2 GPUs): the learner update of `--actor-device-ids 0 --learner-device-ids 0 1 [...]` in one
process at the config-4 shapes (local_num_envs 60, 128 steps, 2 actor threads): ms per multi_device_update with the
gradient exchange fused into the optimizer kernels over peer memory (default) and with the host-synchronised sum on
device 0 (CLEANBA_PEER_FUSED=0)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from cleanba_b200.cuda_backend import CudaBackend
from cleanba_b200.sebulba import Args, derive_sizes

L = min(torch.cuda.device_count(), int(sys.argv[1]) if len(sys.argv) > 1 else 2)
a = Args(local_num_envs=60, num_actor_threads=2, num_steps=128, actor_device_ids=[0], learner_device_ids=list(range(L)))
a = derive_sizes(a, 1)
be = CudaBackend()
learner = be.make_learner(a, be.first_key(1), None)
rng = np.random.default_rng(0)
Nl = 60 // L
T = 128


def payload():
    out = []
    for l in range(L):
        d = torch.device("cuda", l)
        with torch.cuda.device(d):
            s = dict(obs=torch.randint(0, 256, (T, Nl, 4, 84, 84), dtype=torch.uint8, device=d),
                     dones=torch.zeros(T, Nl, dtype=torch.bool, device=d), actions=torch.randint(0, 18, (T, Nl), dtype=torch.int32, device=d),
                     logprobs=torch.full((T, Nl), float(np.log(1 / 18)), device=d), values=torch.randn(T, Nl, device=d) * 0.1,
                     rewards=torch.zeros(T, Nl, device=d), next_obs=torch.randint(0, 256, (Nl, 4, 84, 84), dtype=torch.uint8, device=d),
                     next_done=torch.zeros(Nl, dtype=torch.bool, device=d))
            ev = torch.cuda.Event(); ev.record(torch.cuda.current_stream(d)); s["event"] = ev
        out.append(s)
    return out


pls = [payload(), payload()]           # two actor threads
for _ in range(2):
    learner.update(pls)
for d in learner.devices:
    torch.cuda.synchronize(d)
t0 = time.perf_counter()
K = 4
for _ in range(K):
    learner.update(pls)
for d in learner.devices:
    torch.cuda.synchronize(d)
ms = (time.perf_counter() - t0) * 1e3 / K
p = [lr.ctx.get_params().cpu() for lr in learner.learners]
same = all(torch.equal(p[0], x) for x in p[1:])
print(f"L={L} peer_fused={learner.peer_fused}: {ms:.2f} ms per multi_device_update ({T * 120 / ms * 1e3:.0f} env-steps/s learner-side), "
      f"replicas identical: {same}, finite: {bool(torch.isfinite(p[0]).all())}")
