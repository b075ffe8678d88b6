"""Synthetic Atari-shaped vector environment with the envpool call surface the reference uses
(cleanba/cleanba_ppo.py:126-146,273,321-340; async form cleanba/cleanba_impala.py:308,352,365).

SURVEY.md section 8(d): frames are uint8 [N,4,84,84] drawn i.i.d. uniform{0..255} from numpy PCG64(seed) into a pinned
pool of `pool_batches` distinct batches that is cycled; rewards in {-1,0,+1} with P = {.05,.9,.05} (reward_clip=True);
terminated ~ Bernoulli(1/500); elapsed_step resets on termination; truncation at 27000 steps; env_id = arange(N).
"""
from types import SimpleNamespace

import numpy as np
import torch

ATARI_MAX_FRAMES = int(108000 / 4)   # cleanba_ppo.py:121-123


class SyntheticAtari:
    def __init__(self, num_envs: int, seed: int = 1, pool_batches: int = 256, pin: bool = True,
                 num_actions: int = 18):
        self.num_envs = num_envs
        self.rng = np.random.Generator(np.random.PCG64(seed))
        pool = torch.empty((pool_batches, num_envs, 4, 84, 84), dtype=torch.uint8)
        if pin and torch.cuda.is_available():
            pool = pool.pin_memory()
        pool.numpy()[...] = self.rng.integers(0, 256, size=pool.shape, dtype=np.uint8)
        self.pool = pool
        self.t = 0
        self.elapsed = np.zeros(num_envs, np.int32)
        self.spec = SimpleNamespace(config=SimpleNamespace(max_episode_steps=ATARI_MAX_FRAMES))
        self.single_action_space = SimpleNamespace(n=num_actions)
        self.action_space = self.single_action_space
        self.single_observation_space = SimpleNamespace(shape=(4, 84, 84), dtype=np.uint8)
        self.observation_space = self.single_observation_space
        self.is_vector_env = True
        self._pending = None
        self._needs_reset = np.zeros(num_envs, bool)

    def _obs(self) -> torch.Tensor:
        o = self.pool[self.t % self.pool.shape[0]]
        self.t += 1
        return o

    def _transition(self):
        """One env step.  As in envpool, the step after a terminal step is the auto-reset step: it returns the first
        observation of the new episode with elapsed_step == 0, zero reward and done == False."""
        n = self.num_envs
        resetting = self._needs_reset
        reward = self.rng.choice(np.array([-1.0, 0.0, 1.0], np.float32), size=n, p=[0.05, 0.9, 0.05])
        terminated = self.rng.random(n) < (1.0 / 500.0)
        self.elapsed += 1
        self.elapsed[resetting] = 0
        reward[resetting] = 0.0
        terminated[resetting] = False
        truncated = self.elapsed >= ATARI_MAX_FRAMES
        done = terminated | truncated
        self._needs_reset = done.copy()
        info = {"env_id": np.arange(n, dtype=np.int32), "elapsed_step": self.elapsed.copy(),
                "terminated": terminated.astype(np.int32), "reward": reward.copy(),
                "TimeLimit.truncated": truncated.copy()}
        return reward, done, info

    # gym-style (PPO, cleanba_ppo.py:273,321)
    def reset(self):
        self.elapsed[:] = 0
        return self._obs()

    def step(self, action):
        reward, done, info = self._transition()
        return self._obs(), reward, done, info

    # envpool async style (IMPALA, cleanba_impala.py:308,352,365)
    def async_reset(self):
        self.elapsed[:] = 0
        n = self.num_envs
        self._pending = (np.zeros(n, np.float32), np.zeros(n, bool),
                         {"env_id": np.arange(n, dtype=np.int32), "elapsed_step": np.zeros(n, np.int32),
                          "terminated": np.zeros(n, np.int32), "reward": np.zeros(n, np.float32)})

    def recv(self):
        reward, done, info = self._pending
        return self._obs(), reward, done, info

    def send(self, action, env_id=None):
        self._pending = self._transition()

    def close(self):
        pass
