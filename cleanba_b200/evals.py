"""Evaluation loop of a saved `.cleanrl_model` (cleanrl_utils/evals/ppo_envpool_jax_eval.py:13-82), with the policy forward
on the CUDA actor path.  Same call shape: `evaluate(model_path, make_env, env_id, eval_episodes, run_name, ...)`; same
semantics: one environment, `key = PRNGKey(seed)` split four ways first, actions sampled with the Gumbel-max trick
(never arg-max), an episode ends when `terminated` or `TimeLimit.truncated` is set, the return is the sum of
`infos["reward"]`.  Video capture (cv2 / moviepy in the reference) is out of scope."""
from typing import Callable, List

import numpy as np
import torch

from . import agent as ag
from .checkpoint import load_cleanrl_model
from .prng import first_key


def evaluate(model_path: str, make_env: Callable, env_id: str, eval_episodes: int, run_name: str = "", Model=None,
             capture_video: bool = False, seed: int = 1, device="cuda:0", max_episode_steps: int = 27000) -> List[float]:
    envs = make_env(env_id, seed, num_envs=1)()
    envs.reset()                                          # the reference resets once before the episode loop too (:24)
    _, flat = load_cleanrl_model(model_path)
    from .checkpoint import _model_of_size
    ctx = ag.Context(device, max_batch=1, train=False, model=_model_of_size(flat.size, ag.NUM_ACTIONS))
    ctx.set_params(flat)
    key = ag.key_tensor(first_key(seed), ctx.device)      # key, *_ = jax.random.split(PRNGKey(seed), 4)
    limit = getattr(getattr(getattr(envs, "spec", None), "config", None), "max_episode_steps", max_episode_steps)
    episodic_returns: List[float] = []
    for _ in range(eval_episodes):
        episodic_return = 0.0
        next_obs = envs.reset()
        for _ in range(limit):
            obs = torch.as_tensor(np.asarray(next_obs)).to(ctx.device, non_blocking=True)
            action, _, _ = ctx.actor_step(obs, key)[:3]
            step = envs.step(action.cpu().numpy())
            next_obs, infos = step[0], step[-1]
            episodic_return += float(infos["reward"][0])
            if int(np.sum(infos["terminated"])) + int(np.sum(infos["TimeLimit.truncated"])) >= 1:
                break
        print(f"eval_episode={len(episodic_returns)}, episodic_return={episodic_return}")
        episodic_returns.append(episodic_return)
    ctx.close()
    return episodic_returns
