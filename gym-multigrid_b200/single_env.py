"""Reference-style single-env adaptors for the map families: the reference's class names, constructor kwargs and
return types (`MazeSingleAgentEnv` maze.py:26-377, `CtFMvNEnv` ctf.py:657-1433, `Ctf1v1Env` ctf.py:50-654) over ONE env
of the batched CUDA classes - so that code written against the reference (its tests, `scripts/main_mvn_ctf_rl.py`) runs
unchanged apart from the import.  Everything is computed by the same kernels as the vector envs; `render()` returns the
rgb_array frame whatever `render_mode` says (there is no `human` window).  Red agents: the built-in RwPolicy on the device,
or the reference's `enemy_policies` objects (`policy/ctf/heuristic.py` has this package's own) deciding on the host."""
from __future__ import annotations

import numpy as np
import torch

from .actions import CtfActions, MazeActions  # noqa: F401
from .map_env import Ctf1v1VecEnv, CtfVecEnv, MazeVecEnv
from .policy.ctf.heuristic import RwPolicy  # noqa: F401  (re-exported beside the env classes, as gym_multigrid.envs users expect)
from .spaces import Discrete, MultiDiscrete


class _SingleMapEnv:
    actions_set = MazeActions

    def _wrap(self, vec, observation_option):
        self.vec, self.observation_option = vec, observation_option
        self.max_steps, self.width, self.height = vec.max_steps, vec.width, vec.height
        self.observation_space = vec.single_observation_space
        self._seed = None

    @property
    def step_count(self):
        return int(self.vec.step_count[0])

    @property
    def agent_positions(self):
        return self.vec.agent_pos[0].cpu().numpy()

    def _info(self):
        return {k: float(v[0]) for k, v in self.vec.get_info().items()}

    def render(self, close=False, highlight=False, tile_size=32):
        """MultiGridEnv.render (multigrid.py:546-606) as ndarray (H * tile_size, W * tile_size, 3) uint8."""
        if close or highlight:
            return None
        frames = self.vec.render(tile_size=tile_size)
        return None if frames is None else frames[0].cpu().numpy()

    def close(self):
        self.vec.close()


class MazeSingleAgentEnv(_SingleMapEnv):
    """maze.py:31-40 kwargs; `reset(seed=None) -> (obs, info)`, `step(action) -> (obs float64 (W,H), reward float, bool, bool, info)`."""

    def __init__(self, map_path, max_steps=100, flag_reward=1.0, obstacle_penalty_ratio=0.0, step_penalty_ratio=0.01,
                 observation_option="map", render_mode="rgb_array", device="cuda:0", seed=0):
        if observation_option != "map":
            raise NotImplementedError('the adaptor returns observation_option="map" (maze.py:245-260)')
        self._wrap(MazeVecEnv(1, map_path, max_steps=max_steps, flag_reward=flag_reward, obstacle_penalty_ratio=obstacle_penalty_ratio,
                              step_penalty_ratio=step_penalty_ratio, device=device, seed=seed, autoreset=False, reference_dtypes=True),
                   observation_option)
        self.action_space = Discrete(5)

    def reset(self, seed=None):
        obs, _ = self.vec.reset()
        return obs[0].cpu().numpy(), self._info()

    def step(self, action):
        a = torch.as_tensor([int(action)], dtype=torch.int8, device=self.vec.device)
        obs, rew, term, trunc, _ = self.vec.step(a)
        if self.vec.status() & 8:
            raise ValueError(f"Invalid action: {action}")       # maze.py:286
        return obs[0].cpu().numpy(), float(rew[0]), bool(term[0]), bool(trunc[0]), self._info()


class CtFMvNEnv(_SingleMapEnv):
    """ctf.py:662-679 kwargs (red agents follow RwPolicy, drawn on the device); `step(blue_actions)` returns the scalar team
    reward.  observation_option: "map" (int64 (H,W)), "flattened" (int64 vector) or "positional" (dict of int64 arrays)."""
    _vec_cls = CtfVecEnv

    def __init__(self, map_path, num_blue_agents=2, num_red_agents=2, enemy_policies=None, battle_range=1, randomness=0.75,
                 flag_reward=1, battle_reward_ratio=0.25, obstacle_penalty_ratio=0, step_penalty_ratio=0.01, max_steps=100,
                 observation_option="positional", observation_scaling=1, render_mode="rgb_array", device="cuda:0", seed=0,
                 carry_agent_flags=False):
        if observation_option not in ("map", "flattened", "positional"):
            raise ValueError(f"Invalid observation_option: {observation_option}")
        kw = dict(battle_range=battle_range, randomness=randomness, flag_reward=flag_reward, battle_reward_ratio=battle_reward_ratio,
                  obstacle_penalty_ratio=obstacle_penalty_ratio, step_penalty_ratio=step_penalty_ratio, max_steps=max_steps,
                  device=device, seed=seed, autoreset=False, reference_dtypes=True)
        if self._vec_cls is CtfVecEnv:
            # carry_agent_flags=True: this object behaves like ONE reference instance across reset() calls - defeated agents stay
            # defeated, as the reference's reset never clears Agent.terminated / collided; default: a fresh one per episode
            kw.update(num_blue_agents=num_blue_agents, num_red_agents=num_red_agents, carry_agent_flags=carry_agent_flags)
        self._wrap(self._vec_cls(1, map_path, **kw), observation_option)
        self.num_blue_agents, self.num_red_agents = self.vec.num_blue, self.vec.num_red
        self.action_space = MultiDiscrete([5] * self.vec.num_blue) if self._vec_cls is CtfVecEnv else Discrete(5)
        self._set_enemy_policies(enemy_policies, seed)

    def _set_enemy_policies(self, enemy_policies, seed):
        """ctf.py:775-826: one policy for every red agent or a list of `num_red_agents` policies - any object with
        `act(observation_dict, curr_pos) -> int` (the reference's CtfPolicy interface).  None / RwPolicy entries only = the built-in
        device opponent.  Otherwise every red action comes from the host (`CtfVecEnv.set_enemy_policies`): `act` is called with
        the positional observation dict and the agent's position before each step, exactly where the reference calls it
        (ctf.py:1297-1301), and the actions reach the kernel through `set_red_actions`; `random_generator`, `field_map` and
        `action_set` attributes are filled in as the reference's constructor does (the generator is the env's `np_random`)."""
        self.np_random = np.random.default_rng(seed)
        self._policies = self.vec.set_enemy_policies(enemy_policies, random_generator=self.np_random)

    def _obs(self, map_obs):
        if self.observation_option == "map":
            return map_obs[0].cpu().numpy()
        if self.observation_option == "flattened":
            return self.vec.flattened_obs()[0].cpu().numpy()
        d = {k: v[0].cpu().numpy() for k, v in self.vec.positional_obs().items()}
        if "is_red_agent_defeated" in d:      # Ctf1v1Env returns a plain int there (ctf.py:395)
            d["is_red_agent_defeated"] = int(d["is_red_agent_defeated"][0])
        return d

    def reset(self, *, seed=None, options=None):
        if seed is not None:      # gymnasium.Env.reset(seed=): a fresh np_random (tests/test_ctf.py:37-48); policies keep the generator they were given
            self.np_random = np.random.default_rng(seed)
        obs, _ = self.vec.reset(seed=seed)
        return self._obs(obs), self._info()

    def step(self, blue_actions):
        a = torch.as_tensor(np.round(np.asarray(blue_actions, dtype=np.float64)).astype(np.int8).reshape(1, -1), device=self.vec.device)
        obs, rew, term, trunc, _ = self.vec.step(a)
        if self.vec.status() & 8:
            raise ValueError(f"Invalid action: {blue_actions}")  # ctf.py:1200-1201
        return self._obs(obs), float(rew[0]), bool(term[0]), bool(trunc[0]), self._info()


class Ctf1v1Env(CtFMvNEnv):
    """ctf.py:55-70 kwargs; one blue agent (scalar Discrete(5) action) against one RwPolicy red agent."""
    _vec_cls = Ctf1v1VecEnv

    def __init__(self, map_path, enemy_policy=None, battle_range=1.0, randomness=0.75, flag_reward=1.0, battle_reward_ratio=0.25,
                 obstacle_penalty_ratio=0.0, step_penalty_ratio=0.01, max_steps=100, observation_option="positional",
                 observation_scaling=1, render_mode="rgb_array", device="cuda:0", seed=0):
        super().__init__(map_path, enemy_policies=enemy_policy, battle_range=battle_range, randomness=randomness, flag_reward=flag_reward,
                         battle_reward_ratio=battle_reward_ratio, obstacle_penalty_ratio=obstacle_penalty_ratio,
                         step_penalty_ratio=step_penalty_ratio, max_steps=max_steps, observation_option=observation_option,
                         observation_scaling=observation_scaling, render_mode=render_mode, device=device, seed=seed)
