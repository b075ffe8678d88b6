"""Developer probe (GPU box): per-kernel timing table of one PPO minibatch gradient + optimizer step and the actor step."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from oracle import network as net
from cleanba_b200 import agent as ag

mb = int(sys.argv[1]) if len(sys.argv) > 1 else 3840
backends = [int(b) for b in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0]
model = int(sys.argv[3]) if len(sys.argv) > 3 else 0          # 0 IMPALA-ResNet, 1 Nature-CNN
params = net.init_params(1, net.nature_param_spec() if model else None)
rng = np.random.default_rng(0)
N = mb
obs = torch.from_numpy(rng.integers(0, 256, (N, 4, 84, 84), dtype=np.uint8)).cuda()
actions = torch.from_numpy(rng.integers(0, 18, N).astype(np.int32)).cuda()
oldlp = torch.full((N,), float(np.log(1 / 18)), dtype=torch.float32).cuda()
adv = torch.randn(N, device="cuda"); ret = torch.randn(N, device="cuda")
idx = torch.from_numpy(rng.permutation(N).astype(np.int32)).cuda()
for backend in backends:
    ctx = ag.Context("cuda:0", max_batch=mb, train=True, conv_backend=backend, model=model)
    ctx.set_params(params)
    grads = torch.zeros(ctx.num_params, device="cuda"); stats = torch.zeros(5, device="cuda")
    def step():
        ctx.ppo_grad(obs, idx, mb, actions, oldlp, adv, ret, 0.1, 0.01, 0.5, grads, stats)
        ctx.optimizer_step(grads, 1.0, 2.5e-4, 0.5)
    for _ in range(2): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 5 if backend == 0 else 1
    e0.record()
    for _ in range(iters): step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"== model={'nature_cnn' if model else 'impala_resnet'} backend={'simt' if backend else 'tcgen05'} mb={mb}: {ms:.3f} ms per minibatch grad+opt step  ({mb/ms*1e3:.0f} samples/s)")
    ctx.profile(True); step(); rep = ctx.profile_report(); ctx.profile(False)
    tot = sum(r["ms"] for r in rep)
    for r in sorted(rep, key=lambda r: -r["ms"]):
        tf = r["flops"] / (r["ms"] * 1e-3) / 1e12 if r["ms"] > 0 else 0
        gb = r["bytes"] / (r["ms"] * 1e-3) / 1e9 if r["ms"] > 0 else 0
        print(f"  {r['name']:38s} calls={r['calls']:3d} ms={r['ms']:8.4f} ({100*r['ms']/tot:5.1f}%)  {tf:7.2f} TFLOP/s  {gb:8.1f} GB/s")
    print(f"  sum of kernels {tot:.3f} ms")
    ctx.close()
# actor step latency
for backend in backends:
    ctx = ag.Context("cuda:0", max_batch=60, conv_backend=backend, model=model)
    ctx.set_params(params)
    key = ag.key_tensor(np.array([1, 2], np.uint32), ctx.device)
    o = obs[:60].contiguous()
    for _ in range(5): ctx.actor_step(o, key)
    torch.cuda.synchronize()
    t = time.time()
    for _ in range(100): ctx.actor_step(o, key)
    torch.cuda.synchronize()
    print(f"== actor_step n=60 backend={backend}: {(time.time()-t)*10:.3f} ms per step (async launch, 100 steps)")
    ctx.profile(True); ctx.actor_step(o, key); rep = ctx.profile_report(); ctx.profile(False)
    for r in sorted(rep, key=lambda r: -r["ms"])[:8]:
        print(f"  {r['name']:38s} calls={r['calls']:3d} ms={r['ms']:8.4f}")
    ctx.close()
