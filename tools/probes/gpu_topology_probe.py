"""Developer probe (GPU box, 4 or 8 GPUs): BASELINE configs[3] / configs[4] through the product entry point.

  config 4:  python tools/probes/gpu_topology_probe.py --updates 12
             = cleanba_ppo.py --actor-device-ids 0 --learner-device-ids 1 2 3 --local-num-envs 60      (one process, 4 GPUs)
  config 5:  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
                 tools/probes/gpu_topology_probe.py --distributed --updates 12
             = cleanba_ppo.py --distributed --actor-device-ids 0 --learner-device-ids 1 2 3            (two processes, 8 GPUs)

Prints one JSON line per process: steady-state env-steps/s (whole job), learner update time, rollout time, the actor -> learner
payload bandwidth over NVLink (device-timed on the actors' copy streams) and whether all learner replicas hold identical
parameters at the end (bit-exact, across processes too)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--distributed", action="store_true")
    ap.add_argument("--updates", type=int, default=12)
    ap.add_argument("--algo", default="ppo")
    ap.add_argument("--actor", type=int, nargs="+", default=[0])
    ap.add_argument("--learners", type=int, nargs="+", default=[1, 2, 3])
    ap.add_argument("--num-envs", type=int, default=60)
    a = ap.parse_args()
    from cleanba_b200 import cleanba_ppo, cuda_backend
    from cleanba_b200.sebulba import Args, impala_defaults
    args = Args(local_num_envs=a.num_envs, actor_device_ids=list(a.actor), learner_device_ids=list(a.learners), distributed=a.distributed,
                max_updates=a.updates, log_frequency=4, total_timesteps=50_000_000)
    if a.algo == "impala":
        args = impala_defaults(args)
    made = []
    scalars = {}

    class Writer:                       # keeps the last value of every scalar the rollout / learner loops log
        def add_scalar(self, k, v, step):
            scalars[k] = float(v)

        def add_text(self, *a, **k):
            pass

        def close(self):
            pass
    import torch.utils.tensorboard as tb
    tb.SummaryWriter = lambda *a, **k: Writer()
    args.log_frequency = 4
    orig = cuda_backend.CudaBackend
    class Probe(orig):
        def __init__(self):
            super().__init__(); made.append(self)
    cuda_backend.CudaBackend = Probe
    t0 = time.time()
    res = cleanba_ppo.main(args)
    wall = time.time() - t0
    for d in range(torch.cuda.device_count()):
        torch.cuda.synchronize(d)
    be = made[0]
    # steady state: from the end of update 3 to the end of the last update
    (ta, sa), (tb, sb) = res.update_done_at[2], res.update_done_at[-1]
    bw = [x for act in be.actors for x in act.payload_bandwidth()]
    learners = res.learner.learners
    same = all(torch.equal(learners[0].ctx.get_params().cpu(), l.ctx.get_params().cpu()) for l in learners[1:])
    free_b, total_b = torch.cuda.mem_get_info(learners[0].ctx.device)
    final_stats = res.learner.stats_to_host(res.stats)
    out = dict(final_stats={k: round(v, 6) for k, v in final_stats.items()}, stats_finite=bool(all(np.isfinite(v) for v in final_stats.values())),
               learner_gpu_used_gb=round((total_b - free_b) / 1e9, 2),
               rank=int(os.environ.get("RANK", 0)), world=int(os.environ.get("WORLD_SIZE", 1)), algo=a.algo,
               topology=f"a{','.join(map(str, a.actor))}-l{','.join(map(str, a.learners))}-d{int(os.environ.get('WORLD_SIZE', 1))}",
               updates=res.updates, global_step=res.global_step, wall_s=round(wall, 2),
               steady_env_steps_per_s=round((sb - sa) / (tb - ta), 1),
               update_ms_mean=round(1e3 * float(np.mean(res.update_seconds[3:])), 2), update_ms_min=round(1e3 * float(np.min(res.update_seconds[3:])), 2),
               queue_get_ms_mean=round(1e3 * float(np.mean(res.queue_get_seconds[3:])), 2),
               payload_gbs_median=round(float(np.median([b for b, _ in bw])), 1) if bw else None,
               payload_mb_per_handoff=round(bw[0][1] / 1e6, 1) if bw else None, payload_handoffs=len(bw),
               replicas_identical_in_process=bool(same),
               actor_thread0_ms_per_update={k.split("/")[1]: round(1e3 * v, 2) for k, v in scalars.items()
                                            if k.startswith("stats/") and k.endswith("_time")})
    if a.distributed:
        import torch.distributed as dist
        # cleanba_ppo.main destroyed the process group; a fresh gloo group compares the replicas across processes
        dist.init_process_group("gloo")
        mine = learners[0].ctx.get_params().cpu()
        allp = [torch.zeros_like(mine) for _ in range(dist.get_world_size())]
        dist.all_gather(allp, mine)
        out["replicas_identical_across_processes"] = bool(all(torch.equal(allp[0], p) for p in allp))
        dist.destroy_process_group()
    print("TOPOLOGY_PROBE " + json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
