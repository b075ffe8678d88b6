"""Developer probe (GPU box): ONE PPO minibatch gradient + optimizer step (mb from argv) inside cudaProfilerStart/Stop, for
`ncu --profile-from-start off`; also one actor step at n=60 when argv[2] == "actor"."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from oracle import network as net
from cleanba_b200 import agent as ag

mb = int(sys.argv[1]) if len(sys.argv) > 1 else 3840
what = sys.argv[2] if len(sys.argv) > 2 else "learner"
model = int(sys.argv[3]) if len(sys.argv) > 3 else 0          # 0 IMPALA-ResNet, 1 Nature-CNN
params = net.init_params(1, net.nature_param_spec() if model else None)
rng = np.random.default_rng(0)
obs = torch.from_numpy(rng.integers(0, 256, (mb, 4, 84, 84), dtype=np.uint8)).cuda()
cudart = torch.cuda.cudart()
if what == "learner":
    actions = torch.from_numpy(rng.integers(0, 18, mb).astype(np.int32)).cuda()
    oldlp = torch.full((mb,), float(np.log(1 / 18)), dtype=torch.float32).cuda()
    adv = torch.randn(mb, device="cuda"); ret = torch.randn(mb, device="cuda")
    idx = torch.from_numpy(rng.permutation(mb).astype(np.int32)).cuda()
    ctx = ag.Context("cuda:0", max_batch=mb, train=True, model=model)
    ctx.set_params(params)
    grads = torch.zeros(ctx.num_params, device="cuda"); stats = torch.zeros(5, device="cuda")
    # the once-per-update kernels too: bootstrap value is skipped, GAE scan + advantage normalisation at [T=128, Bl=120] and one
    # permutation of the 15,360 sample indices (cleanba_ppo.py:532-560, 592-595, 599-606)
    T, Bl = 128, 120
    rew = torch.randn(T, Bl, device="cuda"); val = torch.randn(T, Bl, device="cuda")
    dones = torch.rand(T, Bl, device="cuda") < 0.01
    nv = torch.randn(Bl, device="cuda"); nd = torch.zeros(Bl, dtype=torch.bool, device="cuda")
    lkey = ag.key_tensor(np.array([3, 4], np.uint32), ctx.device)
    def step():
        ctx.gae(rew, val, dones, nv, nd, 0.99, 0.95, 4)
        ctx.permutation(ctx.split_key(lkey), T * Bl)
        ctx.ppo_grad(obs, idx, mb, actions, oldlp, adv, ret, 0.1, 0.01, 0.5, grads, stats)
        ctx.optimizer_step(grads, 1.0, 2.5e-4, 0.5)
else:
    ctx = ag.Context("cuda:0", max_batch=mb, model=model)
    ctx.set_params(params)
    key = ag.key_tensor(np.array([1, 2], np.uint32), ctx.device)
    def step():
        ctx.actor_step(obs, key)
for _ in range(2): step()
torch.cuda.synchronize()
cudart.cudaProfilerStart()
step()
torch.cuda.synchronize()
cudart.cudaProfilerStop()
print("done", what, mb)
