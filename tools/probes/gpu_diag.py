"""Developer diagnostic (run on the GPU box): per-image, per-layer forward error table for several batch sizes."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from oracle import network as net
from cleanba_b200 import agent as ag

params = net.init_params(1)
p = net.unflatten(torch.tensor(params))
H = [84, 42, 21]; Ho = [42, 21, 11]; C = [16, 32, 32]
out = {}
for backend in (1, 0):
    for n in [int(a) for a in (sys.argv[1:] or ["1", "2", "3", "4", "5", "60"])]:
        rng = np.random.default_rng(100 + n)
        obs = rng.integers(0, 256, (n, 4, 84, 84), dtype=np.uint8)
        ctx = ag.Context("cuda:0", max_batch=max(n, 8), conv_backend=backend)
        ctx.set_params(params)
        logits, value = ctx.policy_value(torch.from_numpy(obs).to(ctx.device))
        torch.cuda.synchronize()
        rec = {}
        with torch.no_grad():
            hidden = net.trunk_forward(p, torch.from_numpy(obs), record=rec)
            ol, ov = net.heads(p, hidden)
        rows = []
        def per_image(got, want):
            got = got.reshape(n, -1).astype(np.float64); want = want.reshape(n, -1).astype(np.float64)
            return (np.abs(got - want).max(1) / np.abs(want).max()).tolist()
        for s in range(3):
            rows.append((f"s{s}.y", per_image(ctx.debug_tensor(f"s{s}.y", (n, H[s], H[s], C[s])), rec[f"s{s}.y"].permute(0, 2, 3, 1).numpy())))
            rows.append((f"s{s}.p", per_image(ctx.debug_tensor(f"s{s}.p", (n, Ho[s], Ho[s], C[s])), rec[f"s{s}.p"].permute(0, 2, 3, 1).numpy())))
            rows.append((f"s{s}.a0", per_image(ctx.debug_tensor(f"s{s}.a0", (n, Ho[s], Ho[s], C[s])), torch.relu(rec[f"s{s}.a0pre"]).permute(0, 2, 3, 1).numpy())))
            rows.append((f"s{s}.b0", per_image(ctx.debug_tensor(f"s{s}.b0", (n, Ho[s], Ho[s], C[s])), rec[f"s{s}.b0"].permute(0, 2, 3, 1).numpy())))
            rows.append((f"s{s}.a1", per_image(ctx.debug_tensor(f"s{s}.a1", (n, Ho[s], Ho[s], C[s])), torch.relu(rec[f"s{s}.a1pre"]).permute(0, 2, 3, 1).numpy())))
            w = rec[f"s{s}.b1"]; w = torch.relu(w) if s == 2 else w
            rows.append((f"s{s}.out", per_image(ctx.debug_tensor(f"s{s}.out", (n, Ho[s], Ho[s], C[s])), w.permute(0, 2, 3, 1).numpy())))
        rows.append(("hidden", per_image(ctx.debug_tensor("hidden", (n, 256)), hidden.numpy())))
        rows.append(("logits", per_image(logits.cpu().numpy(), ol.numpy())))
        rows.append(("value", per_image(value.cpu().numpy()[:, None], ov.numpy()[:, None])))
        print(f"== backend={'simt' if backend else 'tcgen05'} n={n}")
        for name, e in rows:
            worst = int(np.argmax(e))
            flag = " <<<" if max(e) > 1e-4 else ""
            tail = " ".join(f"{x:.1e}" for x in e[-3:])
            print(f"  {name:8s} max={max(e):.2e} (image {worst})  last3=[{tail}]{flag}")
        ctx.close()
