"""`python -m cleanba_b200.cleanba_ppo ...` -- drop-in for `python cleanba/cleanba_ppo.py ...` (same flags:
--actor-device-ids / --learner-device-ids / --distributed / --local-num-envs / --num-actor-threads / --concurrency ...),
with the hot path on libcleanba_b200.  One process drives the listed local GPUs; `--distributed` joins the processes
started by torchrun (or SLURM, as jax.distributed does in the reference) into one data-parallel learner group."""
import sys
import time
import uuid

import tyro

from .sebulba import Args, derive_sizes, distributed_env, impala_defaults, train


ATARI_MAX_FRAMES = int(108000 / 4)    # cleanba_ppo.py:120-123
USE_SYNTHETIC_ENV = True              # set from Args.synthetic_env by main(); envpool is not installable in this image


def make_env(env_id, seed, num_envs):
    """cleanba_ppo.py:126-146.  With `--no-synthetic-env` and envpool importable this is the reference's envpool Atari
    vector env (same keyword arguments); otherwise the synthetic Atari-shaped env with the same call surface
    (reset/step, async_reset/recv/send, the info dict keys the rollout loop reads) stands in."""
    def thunk():
        if not USE_SYNTHETIC_ENV:
            try:
                import envpool
            except ImportError as e:
                raise RuntimeError("--no-synthetic-env needs envpool (not installable in this image)") from e
            envs = envpool.make(env_id, env_type="gym", num_envs=num_envs, episodic_life=False, repeat_action_probability=0.25,
                                noop_max=1, full_action_space=True, max_episode_steps=ATARI_MAX_FRAMES, reward_clip=True, seed=seed)
            envs.num_envs = num_envs
            envs.single_action_space = envs.action_space
            envs.single_observation_space = envs.observation_space
            envs.is_vector_env = True
            return envs
        from .envs import SyntheticAtari
        envs = SyntheticAtari(num_envs, seed=seed, pool_batches=64)
        envs.num_envs = num_envs
        return envs
    return thunk


def main(args: Args):
    global USE_SYNTHETIC_ENV
    USE_SYNTHETIC_ENV = bool(args.synthetic_env)
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
    if args.track and rank == 0:
        # cleanba_ppo.py:447-458: wandb mirrors the TensorBoard scalars (sync_tensorboard) of rank 0
        try:
            import wandb
        except ImportError as e:
            raise RuntimeError("--track needs the wandb package (not installed in this image); run without --track to log to "
                               "TensorBoard only") from e
        wandb.init(project=args.wandb_project_name, entity=args.wandb_entity, sync_tensorboard=True,
                   config={k: v for k, v in vars(args).items() if not k.startswith("_")}, name=run_name, monitor_gym=True, save_code=True)
    if rank == 0:
        try:
            from torch.utils.tensorboard import SummaryWriter
            writer = SummaryWriter(f"runs/{run_name}")
            writer.add_text("hyperparameters", "|param|value|\n|-|-|\n%s" % ("\n".join([f"|{k}|{v}|" for k, v in vars(args).items() if not k.startswith("_")])))
        except Exception as e:  # noqa: BLE001
            print(f"warning: TensorBoard logging disabled ({type(e).__name__}: {e})", file=sys.stderr)
            writer = None
        for flag in ("capture_video", "upload_model"):
            if getattr(args, flag):
                print(f"warning: --{flag.replace('_', '-')} is accepted for CLI compatibility but not implemented "
                      "(video capture / HF upload are outside the hot path, DESIGN.md section 7)", file=sys.stderr)
    backend = CudaBackend()
    t0 = time.time()
    res = train(args, backend, make_env, writer=writer, allreduce=allreduce)
    if rank == 0:
        print(f"done: {res.updates} updates, global_step={res.global_step}, SPS={int(res.sps)}, wall={time.time() - t0:.1f}s")
    if args.save_model and rank == 0:
        # cleanba_ppo.py:753-783: write runs/{run_name}/{exp_name}.cleanrl_model (flax msgpack of [vars(args), [network,
        # actor, critic]]), then evaluate 10 episodes and log them.  A sidecar with the optimizer state allows --resume-from.
        from .checkpoint import save_cleanrl_model, save_train_state
        from .evals import evaluate
        model_path = f"runs/{run_name}/{args.exp_name}.cleanrl_model"
        save_cleanrl_model(model_path, args, res.learner.flat_params())
        st = res.learner.train_state()
        save_train_state(model_path + ".train_state.npz", st["params"], st["m"], st["v"], st["count"], st["key"],
                         res.updates, res.global_step)
        print(f"model saved to {model_path}")
        episodic_returns = evaluate(model_path, make_env, args.env_id, eval_episodes=10, run_name=f"{run_name}-eval",
                                    device=f"cuda:{args.learner_device_ids[0]}", max_episode_steps=args.eval_max_steps)
        for idx, episodic_return in enumerate(episodic_returns):
            if writer is not None:
                writer.add_scalar("eval/episodic_return", episodic_return, idx)
        res.model_path, res.eval_returns = model_path, episodic_returns
    if writer is not None:
        writer.close()
    if args.distributed:
        dist.destroy_process_group()
    return res


if __name__ == "__main__":
    main(tyro.cli(Args))
