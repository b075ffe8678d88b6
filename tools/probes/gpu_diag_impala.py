"""Developer diagnostic: per-minibatch IMPALA update, CUDA learner vs oracle learner on identical shards."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from oracle import impala as oimpala, network as net
from cleanba_b200.learner import ImpalaHyper, ImpalaLearner
rng = np.random.default_rng(5)
T1, Bl = 5, 16
params = net.init_params(1)
sh = oimpala.Shard(obs=rng.integers(0, 256, (T1, Bl, 4, 84, 84), dtype=np.uint8), dones=rng.random((T1, Bl)) < 0.1,
                   actions=rng.integers(0, 18, (T1, Bl)).astype(np.int32), logitss=(rng.standard_normal((T1, Bl, 18)) * 0.01).astype(np.float32),
                   rewards=rng.choice([-1.0, 0.0, 1.0], size=(T1, Bl)).astype(np.float32), firststeps=rng.random((T1, Bl)) < 0.1)
ol = oimpala.ImpalaLearner(params, oimpala.ImpalaConfig(num_minibatches=2, num_updates=100)); rec = []
ol.update([sh], record=rec)
L = ImpalaLearner("cuda:0", ImpalaHyper(num_minibatches=2, num_updates=100), T1=T1, Bl=Bl)
L.ctx.set_params(params)
tt = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
st = L.update(tt(sh.obs), tt(sh.dones), tt(sh.actions), tt(sh.logitss), tt(sh.rewards), tt(sh.firststeps))
cs = L.stats.cpu().numpy()
for j in range(2):
    print("minibatch", j, "cuda", cs[j], "oracle", rec[j]["stats"], "lr", rec[j]["lr"])
p = L.ctx.get_params().cpu().numpy()
print("params maxabs diff", np.abs(p - ol.params).max(), "max param", np.abs(ol.params).max(), "step size", np.abs(ol.params - params).max())
# after first minibatch only
ol2 = oimpala.ImpalaLearner(params, oimpala.ImpalaConfig(num_minibatches=2, num_updates=100))
