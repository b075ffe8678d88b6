"""Seed x environment fan-out launcher with the CLI of the reference's `cleanrl_utils/benchmark.py:12-137` (the tool behind
`benchmark.sh`): expands `--command` over `--env-ids` and `--num-seeds` seeds, runs the resulting command lines on `--workers`
local worker threads, or renders them into a SLURM array script from `--slurm-template-path` (same `{{placeholders}}`) and
submits it with sbatch.  Example (BASELINE config 2 over three seeds, two at a time):

    python -m cleanba_b200.benchmark --env-ids Breakout-v5 --num-seeds 3 --workers 2 \
        --command "python -m cleanba_b200.cleanba_ppo --local-num-envs 60 --actor-device-ids 0 --learner-device-ids 0"

Differences from the reference tool: no GitHub lookup for the auto tag (there is no network here: the tag is `git describe`
and the commit only), worker failures are collected and reported at the end instead of dying inside a pool thread."""
import argparse
import math
import os
import shlex
import subprocess
import uuid
from concurrent.futures import ThreadPoolExecutor
from typing import Dict, List, Optional


def _flag(x) -> bool:
    return str(x).lower() in ("1", "true", "t", "yes", "y", "on")


def parse_args(argv=None):
    p = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    p.add_argument("--env-ids", nargs="+", default=["Breakout-v5"], help="environment ids to fan out over")
    p.add_argument("--command", type=str, default="python -m cleanba_b200.cleanba_ppo", help="the command to run")
    p.add_argument("--num-seeds", type=int, default=3, help="number of random seeds")
    p.add_argument("--start-seed", type=int, default=1, help="first seed")
    p.add_argument("--workers", type=int, default=0, help="local worker threads (0 = only print the commands)")
    p.add_argument("--auto-tag", type=_flag, default=True, nargs="?", const=True, help="export WANDB_TAGS from git describe / commit")
    p.add_argument("--slurm-template-path", type=str, default=None, help="SLURM template with {{array}}, {{env_ids}}, {{seeds}}, ...")
    p.add_argument("--slurm-gpus-per-task", type=int, default=1)
    p.add_argument("--slurm-total-cpus", type=int, default=50)
    p.add_argument("--slurm-ntasks", type=int, default=1)
    p.add_argument("--slurm-nodes", type=int, default=None)
    return p.parse_args(argv)


def expand(command: str, env_ids: List[str], num_seeds: int, start_seed: int) -> List[str]:
    """Seed-major order, as the reference: all environments of seed s before seed s + 1."""
    return [f"{command} --env-id {env} --seed {start_seed + s}" for s in range(num_seeds) for env in env_ids]


def git_tag() -> str:
    try:
        tag = subprocess.check_output(["git", "describe", "--tags"], stderr=subprocess.DEVNULL).decode().strip()
    except Exception:
        return ""
    try:
        commit = subprocess.check_output(["git", "rev-parse", "--short", "HEAD"], stderr=subprocess.DEVNULL).decode().strip()
        return f"{tag},{commit}"
    except Exception:
        return tag


def run_local(commands: List[str], workers: int) -> Dict[str, int]:
    def one(cmd):
        print(f"running {cmd}", flush=True)
        return cmd, subprocess.call(shlex.split(cmd))
    with ThreadPoolExecutor(max_workers=workers, thread_name_prefix="cleanba-benchmark-worker-") as ex:
        return dict(ex.map(one, commands))


def render_slurm(template: str, args, ncommands: int) -> str:
    gpus = args.slurm_gpus_per_task * args.slurm_ntasks
    fields = {
        "array": f"0-{ncommands - 1}%{args.workers}",
        "env_ids": "(" + " ".join(args.env_ids) + ")",
        "seeds": "(" + " ".join(str(args.start_seed + s) for s in range(args.num_seeds)) + ")",
        "len_seeds": str(args.num_seeds),
        "command": args.command,
        "gpus_per_task": str(args.slurm_gpus_per_task),
        "cpus_per_gpu": str(math.ceil(args.slurm_total_cpus / gpus)),
        "ntasks": str(args.slurm_ntasks),
        "nodes": f"#SBATCH --nodes={args.slurm_nodes}" if args.slurm_nodes is not None else "",
    }
    for k, v in fields.items():
        template = template.replace("{{" + k + "}}", v)
    return template


def main(argv=None) -> Optional[Dict[str, int]]:
    args = parse_args(argv)
    if args.auto_tag:
        if "WANDB_TAGS" in os.environ:
            raise ValueError("WANDB_TAGS is already set: unset it or pass --auto-tag False")
        tag = git_tag()
        if tag:
            os.environ["WANDB_TAGS"] = tag
    commands = expand(args.command, args.env_ids, args.num_seeds, args.start_seed)
    print("======= commands to run:")
    for c in commands:
        print(c)
    results = None
    if args.workers > 0 and args.slurm_template_path is None:
        results = run_local(commands, args.workers)
        failed = {c: rc for c, rc in results.items() if rc != 0}
        if failed:
            raise SystemExit("benchmark: %d of %d runs failed: %s" % (len(failed), len(commands), failed))
    elif args.slurm_template_path is None:
        print("not running the experiments because --workers is 0; only the commands were printed")
    if args.slurm_template_path is not None:
        os.makedirs(os.path.join("slurm", "logs"), exist_ok=True)
        with open(args.slurm_template_path) as f:
            script = render_slurm(f.read(), args, len(commands))
        path = os.path.join("slurm", f"{uuid.uuid4()}.slurm")
        with open(path, "w") as f:
            f.write(script)
        print(f"======= slurm script saved to {path}")
        if args.workers > 0:
            subprocess.check_call(["sbatch", path])
    return results


if __name__ == "__main__":
    main()
