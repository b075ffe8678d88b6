"""Developer probe (GPU box): per-tensor comparison of the actor's persistent tail kernel with the per-layer kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from oracle import network as net
from cleanba_b200 import agent as ag

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
params = net.init_params(1)
rng = np.random.default_rng(0)
obs = torch.from_numpy(rng.integers(0, 256, (n, 4, 84, 84), dtype=np.uint8)).cuda()
A = ag.Context("cuda:0", max_batch=n, train=False); L = ag.Context("cuda:0", max_batch=n, train=True)
for c in (A, L): c.set_params(params)
A.policy_value(obs); L.policy_value(obs); torch.cuda.synchronize()
H = {0: 42, 1: 21, 2: 11}; C = {0: 16, 1: 32, 2: 32}
for s in (0, 1, 2):
    for f in ("p", "pr", "a0", "b0", "b0r", "a1", "out"):
        shp = (n, H[s], H[s], C[s])
        a = A.debug_tensor(f"s{s}.{f}", shp); l = L.debug_tensor(f"s{s}.{f}", shp)
        d = np.abs(a - l)
        bad = np.argwhere(d > 0)
        print(f"s{s}.{f:4s} max|diff| {d.max():.3e}  mismatching {len(bad)}/{d.size}", ("first " + str(bad[:3].tolist()) + " last " + str(bad[-2:].tolist())) if len(bad) else "")
