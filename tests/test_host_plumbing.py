"""CPU tests of the host side: C-ABI symbol table, Args size derivation, Sebulba plumbing (queues, policy-version lag,
payload sharding) driven by the oracle backend, and the world_size-2 data-parallel path over gloo."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cabi_library_exports_every_declared_symbol():
    from cleanba_b200 import build, lib
    build.build()
    header = open(os.path.join(ROOT, "include", "cleanba_b200.h")).read()
    declared = set(re.findall(r"\b(cb_[a-z_0-9]+)\s*\(", header))
    declared -= {"cb_ctx", "cb_config", "cb_stream"}
    assert declared, "no declarations found"
    dll = ctypes.CDLL(lib.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(dll, s)]
    assert not missing, f"symbols declared in include/cleanba_b200.h but not exported: {missing}"
    assert declared == set(lib.SIGNATURES), (declared ^ set(lib.SIGNATURES))
    # metadata calls work without a GPU
    assert lib.load().cb_num_params(18) == 1_094_115 and len(lib.leaves()) == 36


def test_product_path_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from cleanba_b200 import agent
    with pytest.raises(agent.CleanbaError):
        agent.Context("cuda:0", max_batch=4)
    from cleanba_b200.cuda_backend import CudaBackend
    with pytest.raises(agent.CleanbaError):
        CudaBackend()


def test_product_params_and_keys_match_oracle():
    from cleanba_b200 import params, prng
    from oracle import network as net, threefry as tf
    assert np.array_equal(params.init_params(3), net.init_params(3))
    assert [(n, s) for n, _, s in params.leaves()] == net.param_spec()
    assert prng.first_key(1).tolist() == tf.split(tf.PRNGKey(1), 4)[0].tolist()
    assert prng.split(prng.first_key(5), 3).tolist() == tf.split(tf.split(tf.PRNGKey(5), 4)[0], 3).tolist()


def test_size_derivation_matches_reference_configs():
    from cleanba_b200.sebulba import Args, derive_sizes, impala_defaults
    a = derive_sizes(Args(local_num_envs=60), 1)              # config 2: a0-l0-d1
    assert (a.local_batch_size, a.local_minibatch_size, a.num_updates, a.batch_size) == (15360, 3840, 3255, 15360)
    a = derive_sizes(Args(local_num_envs=60, learner_device_ids=[1, 2, 3]), 2)   # config 5: a0-l1,2,3-d2
    assert (a.batch_size, a.num_envs, a.num_updates) == (30720, 240, 1627)
    a = derive_sizes(Args(local_num_envs=8), 1)               # config 1
    assert (a.local_batch_size, a.local_minibatch_size) == (2048, 512)
    a = derive_sizes(impala_defaults(Args(local_num_envs=60)), 1)   # config 3
    assert (a.local_batch_size, a.num_updates, a.concurrency, a.num_steps) == (2400, 20833, True, 20)
    with pytest.raises(AssertionError):
        derive_sizes(Args(local_num_envs=60, learner_device_ids=[0, 1, 2, 3, 4, 5, 6]), 1)   # 60 % 7
    with pytest.raises(AssertionError):
        derive_sizes(Args(local_num_envs=6, num_actor_threads=1, learner_device_ids=[0, 1]), 1)   # 3 % 4


def _tiny_args(algo="ppo", **kw):
    from cleanba_b200.sebulba import Args, derive_sizes, impala_defaults
    a = Args(local_num_envs=4, num_actor_threads=2, num_steps=3, num_minibatches=2, update_epochs=1, total_timesteps=10 ** 6,
             log_frequency=1000, max_updates=3, **kw)
    if algo == "impala":
        a = impala_defaults(a)
        a.num_steps = 3
    return derive_sizes(a, 1)


def _make_env(env_id, seed, num_envs):
    def thunk():
        from cleanba_b200.envs import SyntheticAtari
        return SyntheticAtari(num_envs, seed=seed, pool_batches=8, pin=False)
    return thunk


@pytest.mark.parametrize("algo,concurrency", [("ppo", False), ("ppo", True), ("impala", True)])
def test_plumbing_runs_and_keeps_policy_lag(algo, concurrency):
    """README pseudocode of the reference: with --concurrency the actor's policy is exactly one version behind from the
    second rollout on (`update != 2` rule, cleanba_ppo.py:287-304); without it the versions are equal."""
    from cleanba_b200.sebulba import train
    from oracle.backend import OracleBackend
    args = _tiny_args(algo)
    args.concurrency = concurrency
    res = train(args, OracleBackend(), _make_env)
    assert res.updates == 3 and np.isfinite(np.asarray(res.stats, dtype=np.float64)).all()
    for actor_version, actor_update, learner_version in res.versions:
        assert actor_update == learner_version
        assert actor_version == (learner_version if not concurrency or learner_version == 1 else learner_version - 1)
    assert res.global_step == 3 * args.local_num_envs * args.num_actor_threads * args.num_steps * (1 if algo == "ppo" else 1) \
        + (args.local_num_envs * args.num_actor_threads if algo == "impala" else 0)


def test_plumbing_is_deterministic():
    from cleanba_b200.sebulba import train
    from oracle.backend import OracleBackend
    r1 = train(_tiny_args(), OracleBackend(), _make_env)
    r2 = train(_tiny_args(), OracleBackend(), _make_env)
    assert np.array_equal(r1.learner.learner.params, r2.learner.learner.params)
    np.testing.assert_array_equal(np.asarray(r1.stats), np.asarray(r2.stats))


def test_two_learner_devices_shard_the_env_axis():
    """a0-l0,1: payloads are split along the env axis and each learner sees [T, N/L * threads] (cleanba_ppo.py:278,587)."""
    from cleanba_b200.sebulba import train
    from oracle.backend import OracleBackend
    seen = []
    args = _tiny_args(learner_device_ids=[0, 1])
    backend = OracleBackend()
    res = train(args, backend, _make_env, on_update=lambda v, gs, st: seen.append(v))
    assert seen == [1, 2, 3]
    assert res.learner.L == 2


WORKER = r"""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
from cleanba_b200.sebulba import Args, derive_sizes, train
from cleanba_b200.envs import SyntheticAtari
from oracle.backend import OracleBackend
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
def allreduce(g):   # jax.lax.pmean over the learner devices of all processes
    t = torch.from_numpy(np.ascontiguousarray(g)); dist.all_reduce(t); return (t / world).numpy()
a = Args(local_num_envs=4, num_actor_threads=1, num_steps=2, num_minibatches=2, update_epochs=1, total_timesteps=10**6,
         log_frequency=1000, max_updates=2, distributed=True)
derive_sizes(a, world, rank)
def make_env(env_id, seed, n):
    return lambda: SyntheticAtari(n, seed=seed, pool_batches=8, pin=False)
res = train(a, OracleBackend(), make_env, allreduce=allreduce)
p = torch.from_numpy(res.learner.learner.params.copy())
gathered = [torch.zeros_like(p) for _ in range(world)]
dist.all_gather(gathered, p)
if rank == 0:
    assert all(torch.equal(gathered[0], g) for g in gathered), "replicas diverged"
    np.save({out!r}, gathered[0].numpy())
    print("OK", a.num_updates, a.batch_size, a.num_envs)
dist.destroy_process_group()
"""


def test_world_size_2_gloo_replicas_stay_identical(tmp_path):
    """Two processes (one learner each) with an allreduce-mean on the flat gradient: parameters stay bit-identical
    across replicas, and the global sizes follow cleanba_ppo.py:425-430."""
    script = tmp_path / "worker.py"
    out = str(tmp_path / "params.npy")
    script.write_text(WORKER.format(root=ROOT, out=out))
    env = dict(os.environ, OMP_NUM_THREADS="2")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29517", str(script)], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "OK 62500 16 8" in r.stdout, r.stdout[-500:]
    assert np.isfinite(np.load(out)).all()


def test_trace_export_records_actor_and_learner_phases(tmp_path):
    """--trace-path: the Sebulba loop writes a Chrome / Perfetto timeline with one track per actor thread and one for the
    learner; spans carry the policy versions, so the one-version lag of the concurrent mode is visible (cleanba_ppo.py:287-304)."""
    import json
    from cleanba_b200.envs import SyntheticAtari
    from cleanba_b200.sebulba import Args, derive_sizes, train
    from oracle.backend import OracleBackend
    path = str(tmp_path / "trace.json")
    a = Args(local_num_envs=4, num_actor_threads=2, num_steps=2, num_minibatches=2, update_epochs=1, total_timesteps=10 ** 6,
             log_frequency=1000, max_updates=3, trace_path=path)
    a.concurrency = True
    a = derive_sizes(a, 1)
    res = train(a, OracleBackend(), lambda env_id, seed, n: (lambda: SyntheticAtari(n, seed=seed, pool_batches=2)))
    assert res.trace_path == path and not hasattr(a, "_tracer")
    doc = json.load(open(path))
    evs = [e for e in doc["traceEvents"] if e["ph"] == "X"]
    names = {(e["tid"], e["name"]) for e in evs}
    for tid in (1, 2):
        assert (tid, "rollout") in names and (tid, "rollout_queue.put") in names and (tid, "params_queue.get") in names
    assert (0, "multi_device_update") in names and (0, "rollout_queue.get") in names and (0, "params_queue.put") in names
    upd = [e for e in evs if e["name"] == "multi_device_update"]
    assert len(upd) == 3 and all(e["dur"] > 0 for e in upd)
    # concurrency: the second rollout of an actor thread starts before the first learner update has finished
    r2 = min(e["ts"] for e in evs if e["name"] == "rollout" and e["args"]["update"] == 2)
    assert r2 < upd[0]["ts"] + upd[0]["dur"]
    tracks = {e["tid"]: e["args"]["name"] for e in doc["traceEvents"] if e["ph"] == "M" and e["name"] == "thread_name"}
    assert tracks[0].startswith("learner") and tracks[1].startswith("actor thread 0")


def test_header_is_plain_c_and_a_c_program_can_drive_the_abi(tmp_path):
    """include/cleanba_b200.h compiles as C99 and examples/abi_smoke.c (dlopen, no Python, no torch types) reaches the library:
    without a GPU cb_create fails with a message from cb_last_error() (exit status 2); on a B200 it runs one actor step (0)."""
    import shutil
    import subprocess
    import torch
    from cleanba_b200 import build as b
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    so = b.build()
    exe = str(tmp_path / "abi_smoke")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(root, "include"), os.path.join(root, "examples", "abi_smoke.c"),
                        "-o", exe, "-ldl"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe, so], capture_output=True, text=True, timeout=300)
    assert "parameters: 1094115 floats" in r.stdout
    if torch.cuda.is_available():
        assert r.returncode == 0 and "actions:" in r.stdout, r.stdout + r.stderr
    else:
        assert r.returncode == 2 and "cb_create failed" in r.stdout and "cuda" in r.stdout.lower(), r.stdout + r.stderr


def test_benchmark_fanout_launcher(tmp_path, monkeypatch, capsys):
    """cleanba_b200.benchmark (the reference's cleanrl_utils/benchmark.py:12-137): seed-major expansion, local workers, SLURM script."""
    from cleanba_b200 import benchmark as bm
    monkeypatch.chdir(tmp_path)
    monkeypatch.delenv("WANDB_TAGS", raising=False)
    cmds = bm.expand("run", ["A", "B"], 2, 5)
    assert cmds == ["run --env-id A --seed 5", "run --env-id B --seed 5", "run --env-id A --seed 6", "run --env-id B --seed 6"]
    marker = tmp_path / "out"
    marker.mkdir()
    script = tmp_path / "job.py"
    script.write_text("import sys, pathlib\na = sys.argv\npathlib.Path(r'%s', a[a.index('--env-id') + 1] + '_' + a[a.index('--seed') + 1]).write_text('ok')\n" % marker)
    res = bm.main(["--env-ids", "X", "Y", "--num-seeds", "2", "--workers", "2", "--auto-tag", "False", "--command", f"{sys.executable} {script}"])
    assert len(res) == 4 and all(rc == 0 for rc in res.values())
    assert sorted(p.name for p in marker.iterdir()) == ["X_1", "X_2", "Y_1", "Y_2"]
    tpl = tmp_path / "tpl.slurm"
    tpl.write_text("#SBATCH --array={{array}}\n{{nodes}}\nenvs={{env_ids}} seeds={{seeds}} n={{len_seeds}} cpg={{cpus_per_gpu}} g={{gpus_per_task}} t={{ntasks}}\n{{command}}\n")
    bm.main(["--env-ids", "X", "--num-seeds", "3", "--workers", "0", "--auto-tag", "False", "--command", "cmd", "--slurm-template-path", str(tpl),
             "--slurm-total-cpus", "10", "--slurm-gpus-per-task", "2", "--slurm-ntasks", "2", "--slurm-nodes", "1"])
    out = next((tmp_path / "slurm").glob("*.slurm")).read_text()
    assert "--array=0-2%0" in out and "#SBATCH --nodes=1" in out and "envs=(X) seeds=(1 2 3) n=3 cpg=3 g=2 t=2" in out and out.strip().endswith("cmd")
    with pytest.raises(SystemExit):
        bm.main(["--env-ids", "X", "--num-seeds", "1", "--workers", "1", "--auto-tag", "False", "--command", f"{sys.executable} -c \"raise SystemExit(3)\""])
