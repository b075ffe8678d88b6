"""`python -m cleanba_b200.cleanba_ppo ...` -- drop-in for `python cleanba/cleanba_ppo.py ...` (same flags:
--actor-device-ids / --learner-device-ids / --distributed / --local-num-envs / --num-actor-threads / --concurrency ...),
with the hot path on libcleanba_b200.  One process drives the listed local GPUs; `--distributed` joins the processes
started by torchrun (or SLURM, as jax.distributed does in the reference) into one data-parallel learner group."""
import time
import uuid

import tyro

from .sebulba import Args, derive_sizes, distributed_env, impala_defaults, train


def make_env(env_id, seed, num_envs):
    """cleanba_ppo.py:126-146.  envpool is not installable in this image: the synthetic Atari-shaped env stands in."""
    def thunk():
        from .envs import SyntheticAtari
        envs = SyntheticAtari(num_envs, seed=seed, pool_batches=64)
        envs.num_envs = num_envs
        return envs
    return thunk


def main(args: Args):
    import torch
    import torch.distributed as dist
    from .cuda_backend import CudaBackend
    world, rank, local_rank = (1, 0, 0)
    allreduce = None
    if args.distributed:
        world, rank, local_rank = distributed_env()
        ndev = len(args.learner_device_ids) + len(set(args.actor_device_ids) - set(args.learner_device_ids))
        base = local_rank * ndev       # local_device_ids=range(len(learner)+len(actor)) per process (cleanba_ppo.py:420-422)
        args.actor_device_ids = [base + d for d in args.actor_device_ids]
        args.learner_device_ids = [base + d for d in args.learner_device_ids]
        torch.cuda.set_device(args.learner_device_ids[0])
        dist.init_process_group("nccl", device_id=torch.device("cuda", args.learner_device_ids[0]))
        allreduce = lambda g: dist.all_reduce(g)
    derive_sizes(args, world, rank)
    run_name = f"{args.env_id}__{args.exp_name}__{args.seed}__{uuid.uuid4()}"
    writer = None
    if rank == 0:
        try:
            from torch.utils.tensorboard import SummaryWriter
            writer = SummaryWriter(f"runs/{run_name}")
            writer.add_text("hyperparameters", "|param|value|\n|-|-|\n%s" % ("\n".join([f"|{k}|{v}|" for k, v in vars(args).items()])))
        except Exception:
            writer = None
    backend = CudaBackend()
    t0 = time.time()
    res = train(args, backend, make_env, writer=writer, allreduce=allreduce)
    if rank == 0:
        print(f"done: {res.updates} updates, global_step={res.global_step}, SPS={int(res.sps)}, wall={time.time() - t0:.1f}s")
    if writer is not None:
        writer.close()
    if args.distributed:
        dist.destroy_process_group()
    return res


if __name__ == "__main__":
    main(tyro.cli(Args))
