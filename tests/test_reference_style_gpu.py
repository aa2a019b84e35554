"""The scenarios of the reference's own tests (tests/test_collect.py, tests/test_ctf.py, tests/test_maze.py) exercised
through this package's drop-in surface: the same ids, class names and constructor kwargs, an episode loop with sampled
actions until terminated / truncated - a user only changes the import.  Rendering is out of scope.  Each scenario also
asserts the reference's return types, and one golden episode per family is replayed through the single-env adaptors."""
import os

import numpy as np
import pytest

from replay import load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def maps(tmp_path_factory):
    """tests/assets/board.txt and board_maze.txt of the reference, re-created from the field maps stored in the fixtures
    (load_text_map transposes, utils/map.py:37)."""
    d = tmp_path_factory.mktemp("assets")
    out = {}
    for name, stem in (("board.txt", "ctf_2v2"), ("board_maze.txt", "maze_board13")):
        p = os.path.join(d, name)
        np.savetxt(p, load_golden(stem)["field_map"].T, fmt="%d")
        out[name] = p
    return out


@pytest.mark.parametrize("env_id", ["gym_multigrid:multigrid-collect-v0"])
def test_collect_game(env_id, cuda_device):
    """tests/test_collect.py:9-22"""
    import gym_multigrid_b200 as gym
    env = gym.make(env_id)
    obs, info = env.reset()
    assert obs.shape == (10, 10, 3) and obs.dtype == np.uint8
    steps = 0
    while True:
        actions = [env.action_space.sample() for a in env.agents]
        obs, reward, terminated, truncated, info = env.step(actions)
        steps += 1
        if terminated or truncated:
            assert env.step_count == steps <= 100 and 0 <= env.collected_balls <= 15
            break
    assert reward.shape == (2,) and reward.dtype == np.float64 and sum(info.values()) == env.collected_balls


def test_ctf(maps, cuda_device):
    """tests/test_ctf.py:20-35 (Ctf1v1Env, flattened observations, random actions until the episode ends)"""
    from gym_multigrid_b200 import Ctf1v1Env
    env = Ctf1v1Env(map_path=maps["board.txt"], render_mode="human", observation_option="flattened")
    obs, _ = env.reset()
    env.render()
    n = 0
    while True:
        action = np.random.choice(list(env.actions_set))
        obs, reward, terminated, truncated, info = env.step(action)
        n += 1
        if terminated or truncated:
            break
    assert obs.dtype == np.int64 and obs.ndim == 1 and isinstance(reward, float) and n <= 100
    assert set(info) == {"d_ba_ra", "d_ba_bf", "d_ba_rf", "d_ra_bf", "d_ra_rf", "d_bf_rf", "d_ba_bb", "d_ba_rb", "d_ra_bb", "d_ra_rb", "d_ba_ob"}


def test_ctf_mvn(maps, cuda_device):
    """tests/test_ctf.py:52-71 (CtFMvNEnv 2v2, action_space.sample())"""
    from gym_multigrid_b200 import CtFMvNEnv
    env = CtFMvNEnv(num_blue_agents=2, num_red_agents=2, map_path=maps["board.txt"], render_mode="human", observation_option="flattened")
    obs, _ = env.reset()
    while True:
        action = env.action_space.sample()
        obs, reward, terminated, truncated, info = env.step(action)
        if terminated or truncated:
            break
    assert terminated or truncated
    with pytest.raises(ValueError):
        env.step([7, 0])            # ctf.py:1200-1201


def test_maze(maps, cuda_device):
    """tests/test_maze.py:6-21"""
    from gym_multigrid_b200 import MazeSingleAgentEnv
    env = MazeSingleAgentEnv(map_path=maps["board_maze.txt"], render_mode="human", max_steps=200, step_penalty_ratio=0)
    obs, _ = env.reset()
    assert obs.dtype == np.float64 and obs.shape == (13, 13)
    n = 0
    while True:
        action = np.random.choice(list(env.actions_set))
        obs, reward, terminated, truncated, info = env.step(action)
        n += 1
        if terminated or truncated:
            break
    assert n <= 200 and reward in (0.0, 1.0) and set(info) == {"d_a_f", "d_a_ob"}


def test_single_env_adaptors_replay_reference_episodes(maps, cuda_device):
    """One recorded reference episode per class through the adaptors: observations, rewards, flags and infos with the
    reference's dtypes, bit for bit."""
    from gym_multigrid_b200 import CtFMvNEnv, MazeSingleAgentEnv
    g = load_golden("maze_board13")
    ep = int(np.argmax(g["length"]))
    env = MazeSingleAgentEnv(map_path=maps["board_maze.txt"])
    env.vec.set_trace(start_index=g["start_index"][ep:ep + 1])
    obs, info = env.reset()
    assert np.array_equal(obs, g["init_obs"][ep].astype(np.float64)) and list(info.values()) == list(g["init_info"][ep])
    for t in range(int(g["length"][ep])):
        obs, rew, term, trunc, info = env.step(int(g["actions"][ep, t]))
        assert np.array_equal(obs, g["obs"][ep, t].astype(np.float64)) and rew == g["reward"][ep, t]
        assert (term, trunc) == (bool(g["terminated"][ep, t]), bool(g["truncated"][ep, t]))
        assert list(info.values()) == list(g["info"][ep, t])
    env.close()
    g = load_golden("ctf_2v2")
    ep = int(np.argmax(g["n_battles"].sum(axis=1)))
    env = CtFMvNEnv(map_path=maps["board.txt"], observation_option="map")
    env.vec.set_trace(blue_place=g["blue_place"][ep:ep + 1], red_place=g["red_place"][ep:ep + 1])
    obs, info = env.reset()
    assert obs.dtype == np.int64 and np.array_equal(obs, g["init_obs"][ep]) and list(info.values()) == list(g["init_info"][ep])
    for t in range(int(g["length"][ep])):
        env.vec.set_trace(red_actions=g["red_actions"][ep:ep + 1, t], order=g["order"][ep:ep + 1, t], blue_win=g["blue_win"][ep:ep + 1, t])
        obs, rew, term, trunc, info = env.step(g["actions"][ep, t])
        assert np.array_equal(obs, g["obs"][ep, t]) and rew == g["reward"][ep, t]
        assert (term, trunc) == (bool(g["terminated"][ep, t]), bool(g["truncated"][ep, t]))
        assert list(info.values()) == list(g["info"][ep, t])
    env.close()
