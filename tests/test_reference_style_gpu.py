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


class _ChasePolicy:
    """A caller-supplied `enemy_policies` entry with the reference's CtfPolicy interface (policy/ctf/heuristic.py:18-37, 75-177):
    steps greedily towards the nearest blue agent read from the observation dict; `random_generator` / `field_map` /
    `action_set` are filled in by the env as ctf.py:785-826 does."""
    name = "chase"

    def __init__(self, field_map=None):
        self.field_map, self.random_generator, self.action_set, self.calls = field_map, None, None, 0

    def act(self, observation, curr_pos):
        assert set(observation) == {"blue_agent", "red_agent", "blue_flag", "red_flag", "blue_territory", "red_territory", "obstacle", "terminated_agents"}
        self.calls += 1
        blue = observation["blue_agent"].reshape(-1, 2)
        t = blue[np.argmin(np.abs(blue - np.array(curr_pos)).sum(1))]
        d = t - np.array(curr_pos)
        if d[0] != 0:
            return int(self.action_set.up if d[0] > 0 else self.action_set.down)      # up = (+1, 0), down = (-1, 0)
        if d[1] != 0:
            return int(self.action_set.right if d[1] > 0 else self.action_set.left)   # right = (0, +1), left = (0, -1)
        return int(self.action_set.stay)


def test_enemy_policies_objects(maps, cuda_device):
    """tests/test_ctf.py:97-125 (`enemy_policies=[FightPolicy(), RwPolicy()]`): host-side policy objects drive the red agents
    through the adaptor; the first red agent closes in on the blue agents, frames are rendered every step as the test does."""
    from gym_multigrid_b200 import CtFMvNEnv, RwPolicy
    chase = _ChasePolicy()
    env = CtFMvNEnv(num_blue_agents=2, num_red_agents=2, map_path=maps["board.txt"], render_mode="human", observation_option="flattened",
                    enemy_policies=[chase, RwPolicy()])
    assert chase.random_generator is env.np_random and chase.field_map is not None and chase.action_set is env.actions_set
    obs, _ = env.reset()
    frames = [env.render()]
    nb = 2
    dist = lambda: np.abs(env.agent_positions[:nb] - env.agent_positions[nb]).sum(1).min()  # noqa: E731
    d0 = dist()
    for _ in range(3):
        obs, reward, terminated, truncated, info = env.step([0, 0])            # blue stays: the chaser must get closer
        frames.append(env.render())
        if terminated or truncated:
            break
    assert chase.calls >= 1 and (dist() < d0 or terminated)
    assert all(f.shape == frames[0].shape and f.dtype == np.uint8 for f in frames)
    env.close()
    with pytest.raises(AssertionError):
        CtFMvNEnv(map_path=maps["board.txt"], enemy_policies=[chase])           # ctf.py:779: one policy per red agent
    one = CtFMvNEnv(map_path=maps["board.txt"], enemy_policies=_ChasePolicy())  # a single policy is shared by every red agent (ctf.py:776-777)
    one.reset(); one.step([1, 2])
    assert one._policies[0] is one._policies[1] and one._policies[0].calls == 2
    one.close()


def test_ctf_random_seeding(maps, cuda_device):
    """tests/test_ctf.py:37-48: `reset(seed=1)` twice gives the same np_random stream - and, here, the same episode."""
    from gym_multigrid_b200 import Ctf1v1Env, CtFMvNEnv
    env = Ctf1v1Env(map_path=maps["board.txt"], render_mode="human", observation_option="flattened")
    env.reset(seed=1)
    array1 = env.np_random.random(10)
    env.reset(seed=1)
    array2 = env.np_random.random(10)
    np.testing.assert_allclose(array1, array2)
    env.close()
    env = CtFMvNEnv(map_path=maps["board.txt"], observation_option="flattened")
    runs = []
    for seed in (5, 5, 6):
        obs, _ = env.reset(seed=seed)
        traj = [obs]
        for t in range(12):
            obs, rew, term, trunc, _ = env.step([t % 5, (t + 2) % 5])
            traj.append(obs)
            if term or trunc:
                break
        runs.append(np.stack(traj))
    assert np.array_equal(runs[0], runs[1]) and not (runs[0].shape == runs[2].shape and np.array_equal(runs[0], runs[2]))
    env.close()


_MOVES = {0: (0, 0), 1: (0, -1), 2: (-1, 0), 3: (0, 1), 4: (1, 0)}       # core/agent.py:54-67, ctf.py:1189-1199


def _spy(policy):
    """Record every decision of a policy object without changing it."""
    policy.decisions = []
    inner = policy.act

    def act(observation, curr_pos):
        a = inner(observation, curr_pos)
        policy.decisions.append((tuple(int(v) for v in curr_pos), int(a)))
        return a
    policy.act = act
    return policy


def _policy_episode(env, red_policies):
    """The loop of tests/test_ctf.py:110-121 (sampled blue actions, a frame per step) with one extra check per step: every red
    agent either stood still (stay, blocked, defeated) or made exactly the move its policy decided on, from the cell the policy
    was shown - i.e. the host-side decisions are the ones the CUDA step executed."""
    nb = env.num_blue_agents
    obs, _ = env.reset()
    frames = [env.render()]
    moved = steps = 0
    while True:
        before = env.agent_positions.astype(np.int64)
        marks = [len(p.decisions) for p in red_policies]
        obs, reward, terminated, truncated, info = env.step(env.action_space.sample())
        frames.append(env.render())
        steps += 1
        after = env.agent_positions.astype(np.int64)
        seen = {}
        for k, p in enumerate(red_policies):
            shown, a = p.decisions[marks[k] + seen.get(id(p), 0)]         # a shared policy object decides for its agents in order
            seen[id(p)] = seen.get(id(p), 0) + 1
            assert shown == tuple(before[nb + k].tolist())
            delta = tuple((after[nb + k] - before[nb + k]).tolist())
            assert delta in ((0, 0), _MOVES[a]), (steps, k, a, delta)
            moved += delta != (0, 0)
        if terminated or truncated:
            break
    assert steps <= env.max_steps and moved > 0 and isinstance(reward, float)
    assert all(f.shape == frames[0].shape and f.dtype == np.uint8 for f in frames)
    return steps


def test_fight_policy(maps, cuda_device):
    """tests/test_ctf.py:97-125: `enemy_policies=[FightPolicy(), RwPolicy()]` - the env fills in the map and its generator."""
    from gym_multigrid_b200 import CtFMvNEnv
    from gym_multigrid_b200.policy.ctf.heuristic import FightPolicy, RwPolicy
    fight, rw = _spy(FightPolicy()), _spy(RwPolicy())
    env = CtFMvNEnv(num_blue_agents=2, num_red_agents=2, map_path=maps["board.txt"], render_mode="human", observation_option="flattened",
                    enemy_policies=[fight, rw])
    assert fight.field_map is not None and fight.random_generator is env.np_random and rw.random_generator is env.np_random
    _policy_episode(env, [fight, rw])
    env.close()


def test_capture_policy(maps, cuda_device):
    """tests/test_ctf.py:127-155"""
    from gym_multigrid_b200 import CtFMvNEnv
    from gym_multigrid_b200.map_env import load_text_map
    from gym_multigrid_b200.policy.ctf.heuristic import CapturePolicy, RwPolicy
    field_map = load_text_map(maps["board.txt"])
    capture, rw = _spy(CapturePolicy(field_map)), _spy(RwPolicy())
    env = CtFMvNEnv(num_blue_agents=2, num_red_agents=2, map_path=maps["board.txt"], render_mode="human", observation_option="flattened",
                    enemy_policies=[capture, rw])
    assert capture.field_map is field_map                                      # a map given by the caller is kept (ctf.py:796-799)
    _policy_episode(env, [capture, rw])
    env.close()


@pytest.mark.parametrize("cls_name", ["PatrolPolicy", "PatrolFightPolicy"])
def test_patrol_policies(cls_name, maps, cuda_device):
    """tests/test_ctf.py:157-215: one policy object shared by both red agents."""
    from gym_multigrid_b200 import CtFMvNEnv
    from gym_multigrid_b200.map_env import load_text_map
    from gym_multigrid_b200.policy.ctf import heuristic
    policy = _spy(getattr(heuristic, cls_name)(load_text_map(maps["board.txt"])))
    assert len(policy.border) > 0
    env = CtFMvNEnv(num_blue_agents=2, num_red_agents=2, map_path=maps["board.txt"], render_mode="human", observation_option="flattened",
                    enemy_policies=policy)
    steps = _policy_episode(env, [policy, policy])
    assert len(policy.decisions) == 2 * steps
    env.close()
