"""The oracle -- and, under `-m gpu`, the CUDA path -- against tests/golden/reference_exec.npz: vectors produced by executing
the REFERENCE'S OWN function bodies (lifted from cleanba/cleanba_ppo.py and cleanba/cleanba_impala.py with `ast`) over stand-ins
for the third-party names they call (tests/golden/make_reference_exec.py says exactly what that pins and what it cannot).

Bars: integers, keys and the GAE recurrence bit-exact; fp32 quantities 1e-5 relative (reduction order); the fp64 loss values and
gradients 1e-9 / 1e-7 against the oracle evaluated in fp64."""
import dataclasses
import json
import os

import numpy as np
import pytest
import torch

from oracle import impala as oimpala, network as net, optim, ppo as oppo

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_exec.npz"))
META = json.loads(str(G["meta_json"]))


def frames(seed, shape):
    return np.random.default_rng(int(seed)).integers(0, 256, shape, dtype=np.uint8)


def relerr(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


# ------------------------------------------------------------------------------------------------ host logic (product side)
def test_args_defaults_match_the_reference_dataclass():
    """cleanba_b200.sebulba.Args field by field against the reference's `class Args` of both scripts."""
    from cleanba_b200.sebulba import Args, impala_defaults
    for ours, ref in ((Args(), META["ppo_args"]), (impala_defaults(Args()), META["impala_args"])):
        mine = {f.name: getattr(ours, f.name) for f in dataclasses.fields(ours)}
        skipped = {"exp_name", "global_learner_decices", "actor_devices", "learner_devices"}   # file name / jax device objects
        for k, v in ref.items():
            if k in skipped:
                continue
            assert k in mine, f"reference Args field {k} missing"
            assert mine[k] == v, f"Args.{k}: {mine[k]!r} != reference {v!r}"


def test_size_derivation_matches_the_reference_main_block():
    from cleanba_b200.sebulba import Args, derive_sizes, impala_defaults
    for case in META["sizes"]:
        a = Args() if case["script"] == "ppo" else impala_defaults(Args())
        for k, v in case["overrides"].items():
            setattr(a, k, v)
        derive_sizes(a, world_size=case["world_size"])
        for k, v in case["derived"].items():
            assert int(getattr(a, k)) == v, (case, k)


def test_learning_rate_schedules():
    """linear_schedule of both scripts (cleanba_ppo.py:475-479, cleanba_impala.py:515-519): oracle and product host code."""
    from cleanba_b200 import learner
    for name, per_update in (("ppo", 16), ("impala", 4)):
        nu = int(G[f"{name}_sched_num_updates"])
        base = 2.5e-4 if name == "ppo" else 6e-4
        for c, lr in zip(G[f"{name}_sched_counts"], G[f"{name}_sched_lr"]):
            assert abs(float(optim.linear_schedule(int(c), base, per_update, nu)) - lr) <= 1e-7 * base
            assert abs(float(learner.linear_schedule(int(c), base, per_update, nu, True)) - lr) <= 1e-7 * base


def test_benchmark_launcher_against_the_reference_tool(tmp_path, monkeypatch, capsys):
    """cleanba_b200.benchmark against cleanrl_utils/benchmark.py run as written (plain Python, so no stand-ins but `distutils`): the
    same argv gives the same command lines in the same order and the same rendered SLURM script."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import tiny_env
    from cleanba_b200 import benchmark as bm
    monkeypatch.chdir(tmp_path)
    monkeypatch.delenv("WANDB_TAGS", raising=False)
    (tmp_path / "template.slurm").write_text(tiny_env.SLURM_TEMPLATE)
    argv = [str(tmp_path / "template.slurm") if x == "<tiny_env.SLURM_TEMPLATE>" else x for x in json.loads(str(G["bench_tool_argv"]))]
    bm.main(argv)
    out = capsys.readouterr().out.splitlines()
    commands = [l for l in out[out.index("======= commands to run:") + 1:] if l.startswith("python -m")]
    assert commands == json.loads(str(G["bench_tool_commands"]))
    scripts = [f for f in os.listdir(tmp_path / "slurm") if f.endswith(".slurm")]
    assert len(scripts) == 1 and (tmp_path / "slurm" / scripts[0]).read_text() == str(G["bench_tool_slurm"])
    assert os.path.isdir(tmp_path / "slurm" / "logs")


# ------------------------------------------------------------------------------------------------ oracle vs reference lines
def test_actor_sampling_lines():
    flat = net.init_params(int(G["act_params_seed"]))
    obs = frames(G["act_obs_seed"], (6, 4, 84, 84))
    _, a, lp, v, key2, logits = oppo.get_action_and_value(flat, obs, G["act_key"])
    assert np.array_equal(a, G["act_action"]) and np.array_equal(key2, G["act_key_after"])
    assert relerr(lp, G["act_logprob"]) < 1e-6 and relerr(v, G["act_value"]) < 1e-6
    _, ai, li, key3 = oimpala.get_action(flat, obs, G["act_key"])
    assert np.array_equal(ai, G["act_impala_action"]) and np.array_equal(key3, G["act_impala_key_after"])
    assert relerr(li, G["act_impala_logits"]) < 1e-6


def test_gae_recurrence_is_bit_exact_and_normalisation_axes():
    adv, ret = oppo.compute_gae(G["gae_rewards"], G["gae_values"], G["gae_dones"], G["gae_next_value"], G["gae_next_done"], 0.99, 0.95)
    assert np.array_equal(adv, G["gae_adv"]), "operand order of compute_gae_once"
    assert np.array_equal(ret, G["gae_ret"])
    assert relerr(oppo.normalize_advantages(adv, 4), G["gae_norm4"]) < 1e-6


def test_ppo_loss_lines_value_and_gradient_fp64():
    lg = torch.tensor(G["ppo_logits"], requires_grad=True)
    vl = torch.tensor(G["ppo_value"], requires_grad=True)
    acts = torch.tensor(G["ppo_actions"])
    lp, ent = oppo.logprob_entropy_from_logits(lg, acts)
    np.testing.assert_allclose(lp.detach().numpy(), G["ppo_newlogprob"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(ent.detach().numpy(), G["ppo_entropy"], rtol=1e-12, atol=1e-12)
    loss, (pg, v, e, kl) = oppo.ppo_loss_from_heads(lp, ent, vl, torch.tensor(G["ppo_behavior_logprobs"]), torch.tensor(G["ppo_advantages"]),
                                                    torch.tensor(G["ppo_targets"]))
    np.testing.assert_allclose([loss.item(), pg.item(), v.item(), e.item(), kl.item()], G["ppo_scalars"], rtol=1e-12)
    dlg, dvl = torch.autograd.grad(loss, (lg, vl))
    np.testing.assert_allclose(dlg.numpy(), G["ppo_dlogits"], rtol=1e-9, atol=1e-14)
    np.testing.assert_allclose(dvl.numpy(), G["ppo_dvalue"], rtol=1e-9, atol=1e-14)


def test_impala_loss_lines_value_and_gradient_fp64():
    pl = torch.tensor(G["imp_policy_logits"], requires_grad=True)
    nv = torch.tensor(G["imp_values"], requires_grad=True)
    total, (pg, bl, el) = oimpala.impala_loss_from_heads(pl, nv, torch.tensor(G["imp_actions"]), torch.tensor(G["imp_behaviour_logits"]),
                                                         torch.tensor(G["imp_rewards"]), torch.tensor(G["imp_dones"]), torch.tensor(G["imp_firststeps"]))
    np.testing.assert_allclose([total.item(), pg.item(), bl.item(), el.item()], G["imp_scalars"], rtol=1e-12)
    dpl, dnv = torch.autograd.grad(total, (pl, nv))
    np.testing.assert_allclose(dpl.numpy(), G["imp_dlogits"], rtol=1e-9, atol=1e-13)
    np.testing.assert_allclose(dnv.numpy(), G["imp_dvalue"], rtol=1e-9, atol=1e-13)


def test_rmsprop_pytorch_style_transform():
    """scale_by_rms_pytorch_style (cleanba_impala.py:152-170) over three updates; the oracle applies p -= lr * u, so lr = 1, p = 0."""
    rms = optim.RMSPropPyTorchStyle(40, decay=0.99, eps=0.01)
    for g, u in zip(G["rms_grads"], G["rms_updates"]):
        p1 = rms.step(np.zeros(40, np.float32), g.astype(np.float32), 1.0)
        assert relerr(-p1, u) < 1e-6
    assert relerr(rms.nu, G["rms_nu"]) < 1e-6


def _ppo_shard():
    T, Bl, nmb, epochs, num_updates = (int(x) for x in G["upd_ppo_cfg"])
    cat = lambda k, ax=1: np.concatenate([G[f"upd_ppo_{k}0"], G[f"upd_ppo_{k}1"]], axis=ax)
    obs = np.concatenate([frames(s, (T, Bl // 2, 4, 84, 84)) for s in G["upd_ppo_obs_seeds"]], axis=1)
    nobs = np.concatenate([frames(s, (Bl // 2, 4, 84, 84)) for s in G["upd_ppo_next_obs_seeds"]], axis=0)
    shard = oppo.Shard(obs=obs, dones=cat("dones"), actions=cat("actions"), logprobs=cat("logprobs"), values=cat("values"), rewards=cat("rewards"),
                       next_obs=nobs, next_done=cat("next_done", 0))
    cfg = oppo.PPOConfig(num_minibatches=nmb, update_epochs=epochs, num_updates=num_updates)
    return shard, cfg


def test_whole_ppo_single_device_update():
    """The reference's single_device_update executed end to end (GAE -> normalisation -> 2 epochs x 4 shuffled minibatches ->
    clip + Adam with its own schedule) against oracle.PPOLearner.update on the same payloads: same shuffles, same scan order, same
    averaging of the scalars, same parameters afterwards."""
    torch.set_num_threads(1)
    shard, cfg = _ppo_shard()
    flat = net.init_params(int(G["upd_ppo_params_seed"]))
    learner = oppo.PPOLearner(flat, cfg)
    stats, key2 = learner.update([shard], G["upd_ppo_key"])
    assert np.array_equal(key2, G["upd_ppo_key_after"]) and learner.opt.count == int(G["upd_ppo_opt_count"])
    np.testing.assert_allclose(stats, G["upd_ppo_scalars"], rtol=2e-5)
    d = learner.params.astype(np.float64) - flat.astype(np.float64)
    assert abs(np.sqrt((d * d).sum()) - float(G["upd_ppo_step_l2"])) < 1e-4 * float(G["upd_ppo_step_l2"])
    # Adam turns rounding noise of near-zero gradient elements into O(lr) differences: compare the sample in units of the step size
    lr = 2.5e-4
    diff = np.abs(learner.params[::211] - G["upd_ppo_params_after_every211"])
    assert np.quantile(diff, 0.999) < 0.05 * lr and diff.max() < 8 * 2 * lr


def test_whole_impala_single_device_update():
    torch.set_num_threads(1)
    T1, Bl, nmb, num_updates = (int(x) for x in G["upd_imp_cfg"])
    cat = lambda k: np.concatenate([G[f"upd_imp_{k}0"], G[f"upd_imp_{k}1"]], axis=1)
    obs = np.concatenate([frames(s, (T1, Bl // 2, 4, 84, 84)) for s in G["upd_imp_obs_seeds"]], axis=1)
    shard = oimpala.Shard(obs=obs, dones=cat("dones"), actions=cat("actions"), logitss=cat("logitss"), rewards=cat("rewards"), firststeps=cat("firststeps"))
    flat = net.init_params(int(G["upd_imp_params_seed"]))
    learner = oimpala.ImpalaLearner(flat, oimpala.ImpalaConfig(num_minibatches=nmb, num_updates=num_updates))
    stats = learner.update([shard])
    assert learner.opt.count == int(G["upd_imp_opt_count"])
    np.testing.assert_allclose(stats, G["upd_imp_scalars"], rtol=2e-5)
    d = learner.params.astype(np.float64) - flat.astype(np.float64)
    assert abs(np.sqrt((d * d).sum()) - float(G["upd_imp_step_l2"])) < 1e-4 * float(G["upd_imp_step_l2"])
    assert relerr(learner.params[::211], G["upd_imp_params_after_every211"]) < 1e-5


def test_whole_ppo_update_with_gradient_accumulation():
    """gradient_accumulation_steps = 2 (cleanba_ppo.py:78, 492-500, 607): the reference's reshape of the shuffled batch into
    num_minibatches * 2 mini-steps and its scan over them, with optax.MultiSteps restated in the stand-in (running mean of the
    mini-step gradients, inner chain and schedule count advance on every second mini-step)."""
    torch.set_num_threads(1)
    shard, cfg = _ppo_shard()
    cfg.gradient_accumulation_steps = 2
    flat = net.init_params(int(G["upd_ppo_params_seed"]))
    learner = oppo.PPOLearner(flat, cfg)
    stats, key2 = learner.update([shard], G["upd_ppo_key"])
    assert np.array_equal(key2, G["upd_ppok2_key_after"]) and learner.opt.count == int(G["upd_ppok2_opt_count"])
    np.testing.assert_allclose(stats, G["upd_ppok2_scalars"], rtol=2e-5)
    d = learner.params.astype(np.float64) - flat.astype(np.float64)
    assert abs(np.sqrt((d * d).sum()) - float(G["upd_ppok2_step_l2"])) < 1e-4 * float(G["upd_ppok2_step_l2"])
    diff = np.abs(learner.params[::211] - G["upd_ppok2_params_after_every211"])
    assert np.quantile(diff, 0.999) < 0.05 * 2.5e-4 and diff.max() < 8 * 2 * 2.5e-4


def _two_device_shards(prefix, make):
    """Device l holds env columns [2l, 2l + 2) of each actor thread's payload (prepare_data's split of the env axis)."""
    return [make(slice(2 * l, 2 * l + 2)) for l in range(2)]


def test_whole_ppo_update_on_two_learner_devices():
    """multi_device_update (cleanba_ppo.py:656-660) with the reference's update function running on two emulated devices (one thread
    each, `lax.pmean` = a rendezvous): local advantage normalisation, the same shuffle key on both devices, pmean'ed gradients
    and scalars -- against the oracle's 2-shard emulation (what the multi-GPU tests compare the CUDA learners with)."""
    torch.set_num_threads(1)
    T, Bl, nmb, epochs, num_updates = (int(x) for x in G["upd_ppo_cfg"])
    obs = [frames(s, (T, Bl // 2, 4, 84, 84)) for s in G["upd_ppo_obs_seeds"]]
    nobs = [frames(s, (Bl // 2, 4, 84, 84)) for s in G["upd_ppo_next_obs_seeds"]]

    def make(c):
        cat = lambda k: np.concatenate([G[f"upd_ppo_{k}0"][:, c], G[f"upd_ppo_{k}1"][:, c]], axis=1)
        return oppo.Shard(obs=np.concatenate([o[:, c] for o in obs], axis=1), dones=cat("dones"), actions=cat("actions"), logprobs=cat("logprobs"),
                          values=cat("values"), rewards=cat("rewards"), next_obs=np.concatenate([o[c] for o in nobs], axis=0),
                          next_done=np.concatenate([G["upd_ppo_next_done0"][c], G["upd_ppo_next_done1"][c]]))
    flat = net.init_params(int(G["upd_ppo_params_seed"]))
    learner = oppo.PPOLearner(flat, oppo.PPOConfig(num_minibatches=nmb, update_epochs=epochs, num_updates=num_updates))
    stats, key2 = learner.update(_two_device_shards("upd_ppo", make), G["upd_ppo_key"])
    assert np.array_equal(key2, G["upd_ppo2_key_after"]) and learner.opt.count == int(G["upd_ppo2_opt_count"])
    np.testing.assert_allclose(stats, G["upd_ppo2_scalars"], rtol=2e-5)
    d = learner.params.astype(np.float64) - flat.astype(np.float64)
    assert abs(np.sqrt((d * d).sum()) - float(G["upd_ppo2_step_l2"])) < 1e-4 * float(G["upd_ppo2_step_l2"])
    diff = np.abs(learner.params[::211] - G["upd_ppo2_params_after_every211"])
    assert np.quantile(diff, 0.999) < 0.05 * 2.5e-4 and diff.max() < 8 * 2 * 2.5e-4


def test_whole_impala_update_on_two_learner_devices():
    torch.set_num_threads(1)
    T1, Bl, nmb, num_updates = (int(x) for x in G["upd_imp_cfg"])
    obs = [frames(s, (T1, Bl // 2, 4, 84, 84)) for s in G["upd_imp_obs_seeds"]]

    def make(c):
        cat = lambda k: np.concatenate([G[f"upd_imp_{k}0"][:, c], G[f"upd_imp_{k}1"][:, c]], axis=1)
        return oimpala.Shard(obs=np.concatenate([o[:, c] for o in obs], axis=1), dones=cat("dones"), actions=cat("actions"), logitss=cat("logitss"),
                             rewards=cat("rewards"), firststeps=cat("firststeps"))
    flat = net.init_params(int(G["upd_imp_params_seed"]))
    learner = oimpala.ImpalaLearner(flat, oimpala.ImpalaConfig(num_minibatches=nmb, num_updates=num_updates))
    stats = learner.update(_two_device_shards("upd_imp", make))
    assert learner.opt.count == int(G["upd_imp2_opt_count"])
    np.testing.assert_allclose(stats, G["upd_imp2_scalars"], rtol=2e-5)
    d = learner.params.astype(np.float64) - flat.astype(np.float64)
    assert abs(np.sqrt((d * d).sum()) - float(G["upd_imp2_step_l2"])) < 1e-4 * float(G["upd_imp2_step_l2"])
    assert relerr(learner.params[::211], G["upd_imp2_params_after_every211"]) < 1e-5


# ppoconc: the PPO script with --concurrency, three updates; *l2: two learner devices (replicate / device_put_sharded / pmap / unreplicate)
@pytest.mark.parametrize("algo", ["ppo", "impala", "ppoconc", "ppol2", "impalal2"])
def test_whole_program_against_the_reference_main_block(algo):
    """cleanba_b200.sebulba.train (the product's host program: actor threads, size-1 queues, learner loop; here over the CPU oracle
    backend) against the reference's whole `if __name__ == "__main__":` block executed as written -- its own rollout() in real threads,
    its own learner loop -- for two updates on the same deterministic env: the order and names of every scalar the learner and the
    first actor thread log, the loss scalars and the logged learning rate of both updates, learner / actor policy versions (the
    actor one behind under IMPALA's concurrency), update index, global_step, optimizer count and the final parameters."""
    import contextlib
    import io
    import sys
    torch.set_num_threads(1)
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import tiny_env
    from cleanba_b200 import sebulba
    from oracle.backend import OracleBackend
    a = sebulba.impala_defaults(sebulba.Args()) if algo.startswith("impala") else sebulba.Args()
    for k, v in json.loads(str(G[f"main_{algo}_overrides"])).items():
        setattr(a, k, v)
    sebulba.derive_sizes(a)
    assert a.num_updates == int(G[f"main_{algo}_cfg"][4])
    learner_sc, actor_sc = [], []

    class W:
        def add_scalar(self, name, value, step):
            import threading
            (learner_sc if threading.current_thread() is threading.main_thread() else actor_sc).append((name, float(value), int(step)))

        def add_text(self, *args, **kw):
            pass

        def close(self):
            pass
    with contextlib.redirect_stdout(io.StringIO()):
        res = sebulba.train(a, OracleBackend(), tiny_env.make_env, writer=W())
    assert [n for n, _, _ in learner_sc] == json.loads(str(G[f"main_{algo}_learner_scalar_names"]))
    assert [n for n, _, _ in actor_sc] == json.loads(str(G[f"main_{algo}_actor_scalar_names"]))
    keep = ("charts/learning_rate", "losses/value_loss", "losses/policy_loss", "losses/entropy", "losses/approx_kl", "losses/loss")
    mine = np.array([[v, s] for n, v, s in learner_sc if n in keep])
    np.testing.assert_allclose(mine, G[f"main_{algo}_learner_scalars"], rtol=1e-4, atol=1e-6)      # atol: policy losses that cancel to ~1e-5
    ret = np.array([[v, s] for n, v, s in actor_sc if n in ("charts/avg_episodic_return", "charts/avg_episodic_length")])
    np.testing.assert_allclose(ret, G[f"main_{algo}_actor_returns"], rtol=1e-6)
    lpv, apv, upd, gs = (int(x) for x in G[f"main_{algo}_versions"])
    assert res.updates == lpv and res.versions[-1] == (apv, upd, lpv) and res.global_step == gs
    inner = res.learner.learner
    assert inner.opt.count == int(G[f"main_{algo}_opt_count"])
    flat0 = net.init_params(int(G[f"main_{algo}_cfg"][3]))
    d = inner.params.astype(np.float64) - flat0.astype(np.float64)
    assert abs(np.sqrt((d * d).sum()) - float(G[f"main_{algo}_step_l2"])) < 1e-3 * float(G[f"main_{algo}_step_l2"])
    assert np.abs(inner.params[::211] - G[f"main_{algo}_params_after_every211"]).max() < 2e-5


# ------------------------------------------------------------------------------------------------ CUDA path vs reference lines
@pytest.mark.gpu
def test_cuda_actor_and_gae_against_reference_lines():
    """cb_actor_step and cb_gae through the C ABI against the vectors directly (not through the oracle)."""
    from cleanba_b200 import agent as ag
    dev = torch.device("cuda:0")
    flat = net.init_params(int(G["act_params_seed"]))
    obs = torch.from_numpy(frames(G["act_obs_seed"], (6, 4, 84, 84))).to(dev)
    for algo, want in ((ag.CB_ALGO_PPO, "ppo"), (ag.CB_ALGO_IMPALA, "impala")):
        ctx = ag.Context(dev, max_batch=8, train=False, algo=algo)
        ctx.set_params(flat)
        key = ag.key_tensor(G["act_key"], dev)
        action, logprob, value, logits = ctx.actor_step(obs, key, want_logprob_value=(want == "ppo"), want_logits=(want == "impala"))
        torch.cuda.synchronize()
        if want == "ppo":
            assert np.array_equal(action.cpu().numpy(), G["act_action"]) and np.array_equal(ag.key_numpy(key), G["act_key_after"])
            assert relerr(logprob.cpu().numpy(), G["act_logprob"]) < 1e-4 and relerr(value.cpu().numpy(), G["act_value"]) < 1e-4
        else:
            assert np.array_equal(action.cpu().numpy(), G["act_impala_action"]) and np.array_equal(ag.key_numpy(key), G["act_impala_key_after"])
            assert relerr(logits.cpu().numpy(), G["act_impala_logits"]) < 1e-4
        ctx.close()
    ctx = ag.Context(dev, max_batch=4)
    tt = lambda k: torch.from_numpy(G[k]).to(dev)
    adv, ret = ctx.gae(tt("gae_rewards"), tt("gae_values"), tt("gae_dones"), tt("gae_next_value"), tt("gae_next_done"), 0.99, 0.95, 0)
    assert np.array_equal(adv.cpu().numpy(), G["gae_adv"]) and np.array_equal(ret.cpu().numpy(), G["gae_ret"])
    adv_n, _ = ctx.gae(tt("gae_rewards"), tt("gae_values"), tt("gae_dones"), tt("gae_next_value"), tt("gae_next_done"), 0.99, 0.95, 4)
    assert relerr(adv_n.cpu().numpy(), G["gae_norm4"]) < 1e-5
    ctx.close()


@pytest.mark.gpu
def test_cuda_ppo_update_first_step_against_reference_lines():
    """learner.PPOLearner.update on the payloads of the reference-executed update: the shuffle (key after the update) is bit-exact and
    the averaged scalars agree at the free-running bar of a 2-epoch update on 8-frame minibatches (1e-2: relu / pool gate flips are
    amplified by Adam on tiny minibatches; the single-step bars are held by the pinned tests)."""
    from cleanba_b200 import agent as ag
    from cleanba_b200.learner import PPOHyper, PPOLearner
    dev = torch.device("cuda:0")
    shard, cfg = _ppo_shard()
    T, Bl = shard.rewards.shape
    hyper = PPOHyper(num_minibatches=cfg.num_minibatches, update_epochs=cfg.update_epochs, num_updates=cfg.num_updates)
    learner = PPOLearner(dev, hyper, T=T, Bl=Bl)
    learner.ctx.set_params(net.init_params(int(G["upd_ppo_params_seed"])))
    key = ag.key_tensor(G["upd_ppo_key"], dev)
    tt = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    stats = learner.update(tt(shard.obs), tt(shard.dones), tt(shard.actions), tt(shard.logprobs), tt(shard.values), tt(shard.rewards),
                           tt(shard.next_obs), tt(shard.next_done), key)
    torch.cuda.synchronize()
    assert np.array_equal(ag.key_numpy(key), G["upd_ppo_key_after"])
    got = stats.cpu().numpy().astype(np.float64)
    np.testing.assert_allclose(got[:4], G["upd_ppo_scalars"][:4], rtol=1e-2, atol=2e-3)


@pytest.mark.skipif(not os.path.exists("/root/reference/cleanba/cleanba_ppo.py"), reason="the reference tree only exists in the build container")
def test_fixture_regenerates_from_the_reference_sources():
    """The committed .npz is what tests/golden/make_reference_exec.py produces from /root/reference today."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_reference_exec", os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "make_reference_exec.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    try:
        fresh = mod.build()
    finally:
        torch.set_default_dtype(torch.float32)
    assert set(fresh) == set(G.files)
    for k in G.files:
        a, b = np.asarray(fresh[k]), G[k]
        if a.dtype.kind == "f":
            np.testing.assert_allclose(a, b, rtol=1e-6, atol=1e-9, err_msg=k)
        else:
            assert np.array_equal(a, b), k


@pytest.mark.parametrize("algo", ["ppo", "impala"])
def test_rollout_thread_against_the_reference_rollout(algo):
    """cleanba_b200.sebulba._rollout (the product's actor thread, here over the CPU oracle backend) against the reference's whole
    rollout() function executed on the same deterministic env for three updates with two learner devices: every payload field of
    every learner shard (frames as checksums), next_obs / next_done, global_step, policy version (one behind under IMPALA's
    concurrency), update index, thread id, the scalar names in order and the episodic-return scalars."""
    import queue
    import sys
    import threading
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import tiny_env
    from cleanba_b200 import sebulba
    from oracle.backend import OracleBackend
    N, T, L, updates = (int(x) for x in G[f"ro_{algo}_cfg"])
    a = sebulba.Args() if algo == "ppo" else sebulba.impala_defaults(sebulba.Args())
    a.local_num_envs, a.num_steps, a.num_actor_threads, a.learner_device_ids, a.log_frequency, a.seed = N, T, 2, [0, 1], 1, 3
    a.num_updates, a.world_size, a.local_rank = updates - 1, 1, 0
    scalars = []

    class W:
        def add_scalar(self, name, value, step):
            scalars.append((name, float(value), int(step)))
    pq, rq = queue.Queue(), queue.Queue()
    for s in G[f"ro_{algo}_param_seeds"]:
        pq.put(net.init_params(int(s)))
    sebulba._rollout(a, OracleBackend(), tiny_env.make_env, rq, pq, W(), 1, 0, threading.Event(), G[f"ro_{algo}_key"])
    for u in range(updates):
        gs, ver, upd, sharded, _, dtid = rq.get_nowait()
        assert [gs, ver, upd, dtid] == G[f"ro_{algo}_u{u}_meta"].tolist()
        for l in range(L):
            sh = sharded[l]
            fields = ("obs", "dones", "actions", "rewards", "firststeps") + (("logprobs", "values") if algo == "ppo" else ("logitss",))
            for f in fields:
                want = G[f"ro_{algo}_u{u}_l{l}_{f}"]
                got = np.asarray(sh[f])
                if f == "obs":
                    got = got.reshape(got.shape[0], got.shape[1], -1).astype(np.int64).sum(-1)
                if got.dtype.kind == "f":
                    assert got.shape == want.shape and relerr(got, want) < 1e-6, (u, l, f)
                else:
                    assert np.array_equal(got, want), (u, l, f)
            if algo == "ppo":
                assert np.asarray(sh["next_obs"]).astype(np.int64).sum() == int(G[f"ro_ppo_u{u}_l{l}_next_obs_sum"])
                assert np.array_equal(np.asarray(sh["next_done"]), G[f"ro_ppo_u{u}_l{l}_next_done"])
    assert rq.empty()
    assert [n for n, _, _ in scalars] == json.loads(str(G[f"ro_{algo}_scalar_names"]))
    keep = ("charts/avg_episodic_return", "charts/avg_episodic_length")
    np.testing.assert_allclose(np.array([[v, s] for n, v, s in scalars if n in keep]), G[f"ro_{algo}_scalars"], rtol=1e-6)
