"""A small deterministic vector env with the envpool call surface the reference's rollout() uses (gym-style reset/step for the
PPO script, async_reset/recv/send for the IMPALA script).  Episodes end often (termination 25 %, truncation after 4 steps) so a
3-step rollout exercises the done / truncation / first-step / episodic-return bookkeeping.  Shared by the fixture generator
(tests/golden/make_reference_exec.py: the REFERENCE's rollout runs on it) and tests/test_reference_exec.py (the product's)."""
from types import SimpleNamespace

import numpy as np


class TinyEnv:
    MAX_STEPS = 4

    def __init__(self, num_envs, seed):
        self.num_envs = num_envs
        self.rng = np.random.default_rng(1000 + int(seed))
        self.elapsed = np.zeros(num_envs, np.int32)
        self.needs_reset = np.zeros(num_envs, bool)
        self.spec = SimpleNamespace(config=SimpleNamespace(max_episode_steps=self.MAX_STEPS))
        self.single_action_space = SimpleNamespace(n=18)
        self.single_observation_space = SimpleNamespace(shape=(4, 84, 84), dtype=np.uint8, sample=lambda: np.zeros((4, 84, 84), np.uint8))
        self.pending = None

    def _obs(self):
        return self.rng.integers(0, 256, (self.num_envs, 4, 84, 84), dtype=np.uint8)

    def _transition(self, action):
        n = self.num_envs
        assert np.asarray(action).shape == (n,)
        reward = self.rng.choice(np.array([-1.0, 0.0, 1.0, 2.0], np.float32), size=n) + np.asarray(action, np.float32) * 0.0
        terminated = self.rng.random(n) < 0.25
        self.elapsed += 1
        self.elapsed[self.needs_reset] = 0
        reward[self.needs_reset] = 0.0
        terminated[self.needs_reset] = False
        truncated = self.elapsed >= self.MAX_STEPS
        done = terminated | truncated
        self.needs_reset = done.copy()
        info = {"env_id": np.arange(n, dtype=np.int32), "elapsed_step": self.elapsed.copy(), "terminated": terminated.astype(np.int32),
                "reward": reward.copy(), "TimeLimit.truncated": truncated.copy()}
        return reward, done, info

    def reset(self):
        self.elapsed[:] = 0
        return self._obs()

    def step(self, action):
        reward, done, info = self._transition(action)
        return self._obs(), reward, done, info

    def async_reset(self):
        n = self.num_envs
        self.elapsed[:] = 0
        self.pending = (np.zeros(n, np.float32), np.zeros(n, bool),
                        {"env_id": np.arange(n, dtype=np.int32), "elapsed_step": np.zeros(n, np.int32), "terminated": np.zeros(n, np.int32),
                         "reward": np.zeros(n, np.float32)})

    def recv(self):
        reward, done, info = self.pending
        return self._obs(), reward, done, info

    def send(self, action, env_id=None):
        self.pending = self._transition(action)

    def close(self):
        pass


def make_env(env_id, seed, num_envs):
    return lambda: TinyEnv(num_envs, seed)


# A SLURM template with every placeholder the reference's cleanrl_utils/benchmark.py substitutes (our own text, not the reference's file)
SLURM_TEMPLATE = """#!/bin/bash
#SBATCH --gpus-per-task={{gpus_per_task}}
#SBATCH --cpus-per-gpu={{cpus_per_gpu}}
#SBATCH --ntasks={{ntasks}}
#SBATCH --array={{array}}
{{nodes}}
envs={{env_ids}}
seeds={{seeds}}
srun {{command}} --env-id ${envs[$SLURM_ARRAY_TASK_ID / {{len_seeds}}]} --seed ${seeds[$SLURM_ARRAY_TASK_ID % {{len_seeds}}]}
"""
