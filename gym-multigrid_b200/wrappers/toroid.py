"""Batched twin of the reference's `ToroidObservation` (gym_multigrid/wrappers/toroid.py:6-68): wraps a
CollectVecEnv and replaces the observation by per-agent, agent-centred, wrap-around one-hot planes
(float32 [N, A, W, H, num_ball_types + num_agents]) computed by one CUDA kernel from the state planes."""
from __future__ import annotations

import numpy as np

from ..spaces import Box


class ToroidObservation:
    def __init__(self, env):
        self.env = env
        self.depth = env.num_ball_types + env.num_agents   # toroid.py:24
        self.single_observation_space = Box(-np.inf, np.inf, (env.width, env.height, self.depth), np.float32)  # :25-27
        self.observation_space = Box(-np.inf, np.inf, (env.num_envs, env.num_agents, env.width, env.height, self.depth), np.float32)
        self.single_action_space, self.action_space = env.single_action_space, env.action_space
        self.num_envs = env.num_envs

    def __getattr__(self, name):
        return getattr(self.env, name)

    def observation(self, obs=None):
        return self.env.toroid_obs()

    def reset(self, **kwargs):
        _, info = self.env.reset(**kwargs)
        return self.observation(), info

    def step(self, actions):
        _, rewards, terminated, truncated, info = self.env.step(actions)
        return self.observation(), rewards, terminated, truncated, info
