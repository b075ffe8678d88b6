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


class SignalAtari(SyntheticAtari):
    """A learnable stand-in with the same call surface: every env shows a frame whose brightness encodes a target action
    (18 levels + pixel noise) and pays reward 1 when the agent picks it, else 0; a new target is drawn every step and episodes
    last `horizon` steps.  Random policy: 1/18 per step.  Used by the end-to-end "does the whole actor-learner system learn" test
    (the i.i.d.-noise SyntheticAtari has nothing to learn)."""

    def __init__(self, num_envs: int, seed: int = 1, horizon: int = 64, num_actions: int = 18, **kw):
        super().__init__(num_envs, seed=seed, pool_batches=1, pin=kw.get("pin", True), num_actions=num_actions)
        self.horizon, self.num_actions = horizon, num_actions
        self.target = self.rng.integers(0, num_actions, num_envs)
        self.frame = torch.empty((num_envs, 4, 84, 84), dtype=torch.uint8)
        if torch.cuda.is_available() and kw.get("pin", True):
            self.frame = self.frame.pin_memory()
        self.mean_reward = 0.0

    def _obs(self) -> torch.Tensor:
        level = (self.target * (255 // self.num_actions) + 7).astype(np.int16)[:, None, None, None]
        noise = self.rng.integers(-6, 7, size=(self.num_envs, 4, 84, 84), dtype=np.int16)
        self.frame.numpy()[...] = np.clip(level + noise, 0, 255).astype(np.uint8)
        return self.frame

    def _transition_signal(self, action):
        n = self.num_envs
        reward = (np.asarray(action).reshape(-1)[:n] == self.target).astype(np.float32)
        self.mean_reward = 0.98 * self.mean_reward + 0.02 * float(reward.mean())
        self.elapsed += 1
        terminated = self.elapsed >= self.horizon
        self.elapsed[terminated] = 0
        self.target = self.rng.integers(0, self.num_actions, n)
        info = {"env_id": np.arange(n, dtype=np.int32), "elapsed_step": self.elapsed.copy(), "terminated": terminated.astype(np.int32),
                "reward": reward.copy(), "TimeLimit.truncated": np.zeros(n, bool)}
        return reward, terminated.copy(), info

    def step(self, action):
        reward, done, info = self._transition_signal(action)
        return self._obs(), reward, done, info

    def send(self, action, env_id=None):
        self._pending = self._transition_signal(action)
