"""GPU tests of the full Sebulba loop on the CUDA backend (actor threads + queues + learner), and of the data-parallel
gradient allreduce over NCCL when more than one GPU is visible."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _make_env(env_id, seed, num_envs):
    def thunk():
        from cleanba_b200.envs import SyntheticAtari
        return SyntheticAtari(num_envs, seed=seed, pool_batches=8)
    return thunk


def _run_pinned(args_fn, algo, L=1):
    """Runs the Sebulba loop twice on the same seeds / synthetic env / plumbing: first on the oracle backend (recording every
    minibatch step of every update), then on the CUDA backend with every learner replica's steps pinned to those records
    (tests/_pin.py).  Because each update ends with the oracle's exact parameters, the actors of the next rollout see the same
    weights as the oracle's actors: identical integer actions -> identical env trajectories -> the NEXT update's minibatches
    are the same data, so its first-step scalars can again be held to 1e-4 (a single flipped action would break that)."""
    from _pin import pin_hook
    from cleanba_b200.cuda_backend import CudaBackend
    from cleanba_b200.sebulba import train
    from oracle.backend import OracleBackend
    ob = OracleBackend()
    o_learner = [None]
    mk_o = ob.make_learner
    ob.make_learner = lambda *a, **k: o_learner.__setitem__(0, mk_o(*a, **k)) or o_learner[0]
    records, want = [], []

    def on_oracle(v, gs, st):
        records.append(list(o_learner[0].last_record))
        want.append(np.asarray(st, np.float64))

    ro = train(args_fn(), ob, _make_env, on_update=on_oracle)
    cb = CudaBackend()
    c_learner = [None]
    mk_c = cb.make_learner
    upd = [0]
    diag, got = [], []

    def make(*a, **k):
        cl = mk_c(*a, **k)
        c_learner[0] = cl
        shared = {}
        for l, lr in enumerate(cl.learners):
            lr.step_hook = pin_hook(lambda: records[upd[0]], diag, algo, shard=(l if L > 1 else None), nshards=L, shared=shared)
        return cl

    cb.make_learner = make

    def on_cuda(v, gs, st):
        got.append(st.detach().cpu().numpy().astype(np.float64))
        upd[0] += 1

    rc = train(args_fn(), cb, _make_env, on_update=on_cuda)
    assert rc.updates == ro.updates == len(records) and rc.global_step == ro.global_step
    for u in range(len(want)):
        # the update's averaged scalars (pmean'ed over the replicas when L > 1): 1e-4 relative to the size of the per-step
        # values they average (the policy loss of normalised advantages averages to ~0: its own magnitude is no scale)
        scale = np.maximum(np.abs(want[u][:4]), np.mean(np.abs(np.stack([r["stats"][:4] for r in records[u]])), axis=0))
        rel = np.abs(got[u][:4] - want[u][:4]) / np.maximum(scale, 1e-6)
        assert rel.max() < 1e-4, (u, got[u], want[u])
    return rc, ro, diag


@pytest.mark.parametrize("algo", ["ppo", "impala"])
def test_sebulba_loop_on_cuda_matches_cpu_plumbing(algo):
    """Actor threads + queues + learner on the CUDA backend against the same plumbing on the oracle backend, two updates, every
    minibatch step held to the single-step bars (loss scalars 1e-4, gradient 1e-3, post-step parameters 1% of an lr step)."""
    from cleanba_b200.sebulba import Args, derive_sizes, impala_defaults

    def args():
        a = Args(local_num_envs=8, num_actor_threads=2, num_steps=4, num_minibatches=2, update_epochs=1, total_timesteps=10 ** 6,
                 log_frequency=1000, max_updates=2)
        if algo == "impala":
            a = impala_defaults(a)
            a.num_steps = 4
        a.concurrency = False
        return derive_sizes(a, 1)

    rc, ro, diag = _run_pinned(args, algo)
    assert len(diag) == 2 * 2 * 2                       # 2 updates x 2 minibatch steps x (grad, post)
    p_cuda = rc.learner.learners[0].ctx.get_params().cpu().numpy()
    assert np.array_equal(p_cuda, ro.learner.learner.params)   # the pinned hand-over


def test_save_model_eval_and_resume(tmp_path, monkeypatch):
    """--save-model writes runs/{run}/{exp}.cleanrl_model in the reference's flax msgpack layout (cleanba_ppo.py:753-771),
    evaluates it (cleanrl_utils/evals/ppo_envpool_jax_eval.py), and the train-state sidecar resumes bit-exactly: a run of
    2 updates + a resumed run of 1 update equals a run of 3 updates when the rollouts are the same."""
    from cleanba_b200 import agent as ag, checkpoint as ck
    from cleanba_b200.cleanba_ppo import main
    from cleanba_b200.learner import PPOHyper, PPOLearner
    from cleanba_b200.params import init_params
    from cleanba_b200.sebulba import Args
    monkeypatch.chdir(tmp_path)
    a = Args(local_num_envs=8, num_actor_threads=1, num_steps=4, num_minibatches=2, update_epochs=1, total_timesteps=10 ** 6,
             log_frequency=1000, max_updates=2, save_model=True, eval_max_steps=5, seed=2)
    res = main(a)
    assert os.path.exists(res.model_path) and len(res.eval_returns) == 10
    saved_args, flat = ck.load_cleanrl_model(res.model_path)
    assert saved_args["seed"] == 2 and saved_args["local_num_envs"] == 8
    assert np.array_equal(flat, res.learner.flat_params())
    st = ck.load_train_state(res.model_path + ".train_state.npz")
    assert st["learner_policy_version"] == 2 and st["count"] == 4 and np.array_equal(st["params"], flat)   # 2 updates x 2 minibatches

    # resume determinism at the learner level (fixed synthetic rollouts): 3 updates == 2 updates + restore + 1 update
    T, Bl = 4, 8
    rng = np.random.default_rng(0)
    tt = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
    rolls = [dict(obs=tt(rng.integers(0, 256, (T, Bl, 4, 84, 84), dtype=np.uint8)), dones=tt(rng.random((T, Bl)) < 0.1),
                  actions=tt(rng.integers(0, 18, (T, Bl)).astype(np.int32)), logprobs=tt(np.full((T, Bl), np.log(1 / 18), np.float32)),
                  values=tt((rng.standard_normal((T, Bl)) * 0.1).astype(np.float32)),
                  rewards=tt(rng.choice([-1.0, 0.0, 1.0], size=(T, Bl)).astype(np.float32)),
                  next_obs=tt(rng.integers(0, 256, (Bl, 4, 84, 84), dtype=np.uint8)), next_done=tt(np.zeros(Bl, bool))) for _ in range(3)]

    def make():
        L = PPOLearner("cuda:0", PPOHyper(update_epochs=1, num_minibatches=2, num_updates=10), T=T, Bl=Bl)
        L.ctx.set_params(init_params(5))
        return L, ag.key_tensor(np.array([7, 9], np.uint32), L.ctx.device)

    def upd(L, key, r):
        return L.update(r["obs"], r["dones"], r["actions"], r["logprobs"], r["values"], r["rewards"], r["next_obs"], r["next_done"], key)

    A, ka = make()
    for r in rolls:
        upd(A, ka, r)
    want = A.ctx.get_params().cpu().numpy()
    B, kb = make()
    for r in rolls[:2]:
        upd(B, kb, r)
    m, v, count = B.ctx.get_opt_state()
    path = ck.save_train_state(str(tmp_path / "mid.npz"), B.ctx.get_params().cpu().numpy(), m.cpu().numpy(), v.cpu().numpy(), count,
                               ag.key_numpy(kb), 2)
    st = ck.load_train_state(path)
    C, _ = make()
    C.ctx.set_params(st["params"]); C.ctx.set_opt_state(st["m"], st["v"], st["count"]); C.opt_count = st["count"]
    kc = ag.key_tensor(st["key"], C.ctx.device)
    upd(C, kc, rolls[2])
    assert np.array_equal(C.ctx.get_params().cpu().numpy(), want)


WORKER = r"""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
rank = int(os.environ["RANK"]); torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
from cleanba_b200 import agent as ag
from cleanba_b200.learner import PPOHyper, PPOLearner
from cleanba_b200.params import init_params
from cleanba_b200.prng import first_key
world = dist.get_world_size()
T, Bl = 4, 8
L = PPOLearner(f"cuda:{{rank}}", PPOHyper(update_epochs=1, num_minibatches=2, num_updates=10), T=T, Bl=Bl, world_learners=world,
               allreduce=lambda g: dist.all_reduce(g))
L.ctx.set_params(init_params(1))
rng = np.random.default_rng(100 + rank)          # every replica sees different data
tt = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
shard = dict(obs=rng.integers(0, 256, (T, Bl, 4, 84, 84), dtype=np.uint8), dones=rng.random((T, Bl)) < 0.1,
             actions=rng.integers(0, 18, (T, Bl)).astype(np.int32), logprobs=np.full((T, Bl), np.log(1 / 18), np.float32),
             values=(rng.standard_normal((T, Bl)) * 0.1).astype(np.float32), rewards=rng.choice([-1.0, 0.0, 1.0], size=(T, Bl)).astype(np.float32),
             next_obs=rng.integers(0, 256, (Bl, 4, 84, 84), dtype=np.uint8), next_done=np.zeros(Bl, bool))
key = ag.key_tensor(first_key(1), L.ctx.device)
stats = L.update(tt(shard["obs"]), tt(shard["dones"]), tt(shard["actions"]), tt(shard["logprobs"]), tt(shard["values"]),
                 tt(shard["rewards"]), tt(shard["next_obs"]), tt(shard["next_done"]), key)
p = L.ctx.get_params()
gathered = [torch.zeros_like(p) for _ in range(world)]
dist.all_gather(gathered, p)
np.save({out!r} + f".shard{{rank}}.npy", shard, allow_pickle=True)
if rank == 0:
    assert all(torch.equal(gathered[0], g) for g in gathered), "learner replicas diverged after the allreduce"
    np.save({out!r}, gathered[0].cpu().numpy())
    print("NCCL_OK")
dist.destroy_process_group()
"""


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_nccl_allreduce_matches_shard_emulation(tmp_path):
    """2 learner replicas (one process per GPU) with ONE NCCL allreduce per minibatch on the flat gradient buffer: replicas
    stay bit-identical and agree with the oracle's 2-shard emulation (pmean of per-shard gradients, cleanba_ppo.py:628)."""
    from oracle import network as net, ppo as oppo, threefry as tf
    script = tmp_path / "worker.py"
    out = str(tmp_path / "params.npy")
    script.write_text(WORKER.format(root=ROOT, out=out))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29531", str(script)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "NCCL_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
    shards = [oppo.Shard(**np.load(out + f".shard{i}.npy", allow_pickle=True).item()) for i in range(2)]
    ol = oppo.PPOLearner(net.init_params(1), oppo.PPOConfig(update_epochs=1, num_minibatches=2, num_updates=10))
    ol.update(shards, tf.split(tf.PRNGKey(1), 4)[0])
    got = np.load(out)
    assert np.abs(got - ol.params).max() < 1e-3 * np.abs(ol.params).max()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("threads", [2, 1])
def test_in_process_two_learner_devices_match_shard_emulation(threads):
    """`--actor-device-ids 0 --learner-device-ids 0 1` in ONE process (the reference's pmap over local learner devices,
    cleanba_ppo.py:656-660): the actor splits every payload's env axis over the two learner GPUs (:278,357-363), the
    gradients are averaged once per minibatch (fused into the optimizer kernels over peer memory), the replicas stay
    bit-identical, and every replica's every minibatch step agrees with the oracle's 2-shard emulation at the single-step bars.
    threads = 1 covers the single-payload path where the learner on the actor's own GPU receives a strided column view."""
    from cleanba_b200.sebulba import Args, derive_sizes

    def args():
        a = Args(local_num_envs=8, num_actor_threads=threads, num_steps=4, num_minibatches=2, update_epochs=1, total_timesteps=10 ** 6,
                 log_frequency=1000, max_updates=2, actor_device_ids=[0], learner_device_ids=[0, 1])
        a.concurrency = False
        return derive_sizes(a, 1)

    rc, ro, diag = _run_pinned(args, "ppo", L=2)
    assert len(diag) == 2 * 2 * 2 * 2                   # 2 updates x 2 steps x (grad, post) x 2 replicas
    p0 = rc.learner.learners[0].ctx.get_params().cpu().numpy()
    p1 = rc.learner.learners[1].ctx.get_params().cpu().numpy()
    assert np.array_equal(p0, p1), "learner replicas diverged"


def test_sm_partition_actor_step_is_bit_identical():
    """cleanba_b200.partition.SmPartition (CUDA green contexts): an actor step launched into the 16-SM actor partition, with
    its grids sized by set_sm_budget, returns bit-identical actions / log-probs / values / key to the whole-GPU launch."""
    from cleanba_b200 import agent as ag
    from cleanba_b200.params import init_params
    from cleanba_b200.partition import SmPartition
    try:
        part = SmPartition("cuda:0", 16)
    except Exception as e:      # driver without green-context support
        pytest.skip(f"green contexts unavailable: {e}")
    assert part.actor_sms >= 8 and part.actor_sms + part.learner_sms <= torch.cuda.get_device_properties(0).multi_processor_count
    rng = np.random.default_rng(3)
    obs = torch.from_numpy(rng.integers(0, 256, (60, 4, 84, 84), dtype=np.uint8)).cuda()
    outs = []
    for partitioned in (False, True):
        ctx = ag.Context("cuda:0", max_batch=60)
        ctx.set_params(init_params(4))
        key = ag.key_tensor(np.array([5, 6], np.uint32), ctx.device)
        torch.cuda.synchronize()
        if partitioned:
            ctx.set_sm_budget(part.actor_sms)
            st = part.actor_stream()
            st.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(st):
                a, lp, v, _ = ctx.actor_step(obs, key)
            st.synchronize()
        else:
            a, lp, v, _ = ctx.actor_step(obs, key)
            torch.cuda.synchronize()
        outs.append((a.cpu().numpy(), lp.cpu().numpy(), v.cpu().numpy(), ag.key_numpy(key)))
        ctx.close()
    for x, y in zip(*outs):
        assert np.array_equal(x, y)


@pytest.mark.slow
@pytest.mark.parametrize("algo,network", [("ppo", "impala_resnet"), ("ppo", "nature_cnn"), ("impala", "impala_resnet")])
def test_the_system_learns_a_signal_task(algo, network):
    """End to end, free running (no pinning): actor threads + queues + PPO learner on the CUDA backend learn envs.SignalAtari
    (frame brightness encodes the rewarded action; a random policy earns 1/18 = 0.056 per step) to > 0.3 reward per step (5x the random policy) within
    120 updates, with finite losses throughout -- kernels, loss scale, optimizer, parameter publish and the plumbing all have to be
    right for that."""
    from cleanba_b200.cuda_backend import CudaBackend
    from cleanba_b200.envs import SignalAtari
    from cleanba_b200.sebulba import Args, derive_sizes, impala_defaults, train
    envs = []

    def make_env(env_id, seed, num_envs):
        def thunk():
            e = SignalAtari(num_envs, seed=seed)
            envs.append(e)
            return e
        return thunk

    a = Args(local_num_envs=32, num_actor_threads=2, num_steps=16, num_minibatches=2, update_epochs=2, total_timesteps=10 ** 7, log_frequency=10 ** 6,
             max_updates=120, learning_rate=1e-3, anneal_lr=False, ent_coef=0.0, gamma=0.0, gae_lambda=0.0, network=network)
    if algo == "impala":          # V-trace learner, RMSProp, async env interface, actors one policy version behind (concurrency)
        a = impala_defaults(a)
        # The reference pairs transition t -> t+1 with the reward that ARRIVED with obs[t], i.e. the reward of action t-1
        # (cleanba_impala.py:352,375-379,580-582; SURVEY appendix D.8, replicated).  With gamma = 0 action t would never see its own
        # reward; with gamma > 0 it does through the bootstrapped V-trace target gamma * vs[t+1].
        a.num_steps, a.learning_rate, a.max_updates, a.gamma = 16, 1e-3, 800, 0.9
    else:
        a.concurrency = False
    losses = []
    res = train(derive_sizes(a, 1), CudaBackend(), make_env, on_update=lambda v, gs, st: losses.append(st.detach().cpu().numpy()))
    assert res.updates == a.max_updates and np.isfinite(np.stack(losses)).all()
    got = float(np.mean([e.mean_reward for e in envs]))
    bar = 0.3 if algo == "ppo" else 0.15      # IMPALA's credit reaches the action only through the bootstrapped target (see above)
    assert got > bar, f"mean reward per step {got:.3f} after {a.max_updates} updates (random policy: 0.056)"
