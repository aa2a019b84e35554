"""GPU parity suite for the base-class MultiGridEnv.step with DefaultWorld (csrc/generic_kernels.cu) through the C ABI:
bit-exact replay of golden traces recorded from reference classes, and CUDA vs the C oracle with same-step autoreset,
final observations, Philox agent orders (shard invariance) and the error bit for actions the reference raises on."""
import numpy as np
import pytest
import torch

import oracle as oc
from replay import load_golden

pytestmark = pytest.mark.gpu
GENERIC = ["generic_9x9_a3", "generic_12x12_a5", "generic_7x7_a1"]


def _np(t):
    return t.cpu().numpy()


def _tile(x, k):
    return np.concatenate([x] * k)


@pytest.mark.parametrize("stem", GENERIC)
def test_generic_replay_bit_exact(stem, cuda_device):
    import gym_multigrid_b200 as mg
    g = load_golden(stem)
    E, T, A = g["actions"].shape
    S, k = int(g["meta_size"]), 5          # E * k is not a multiple of the 32-env tile for any fixture
    env = mg.make_generic_vec(E * k, S, num_agents=A, max_steps=int(g["meta_max_steps"]), autoreset=False)
    env.set_layout(_tile(g["init_obs"][:, 0], k), _tile(g["init_pos"], k))
    obs, _ = env.reset()
    assert np.array_equal(_np(obs), _tile(g["init_obs"], k)), "reset observation"
    ident = np.arange(A, dtype=np.uint8)[None]
    for t in range(T):
        lv = g["length"] > t
        live = _tile(lv, k)
        env.set_order(_tile(np.where(lv[:, None], g["order"][:, t], ident), k))
        act = _tile(np.where(lv[:, None], g["actions"][:, t], 0), k).astype(np.int8)
        obs, rew, term, trunc, _ = env.step(torch.as_tensor(act, device=cuda_device))
        assert np.array_equal(_np(obs)[live], _tile(g["obs"][:, t], k)[live]), f"step {t}: obs"
        assert np.array_equal(_np(rew)[live], _tile(g["rewards"][:, t], k)[live]), f"step {t}: rewards (float64 bit-exact)"
        assert np.array_equal(_np(term)[live], _tile(g["terminated"][:, t], k)[live])
        assert np.array_equal(_np(trunc)[live], _tile(g["truncated"][:, t], k)[live])
        assert np.array_equal(_np(env.agent_pos)[live], _tile(g["pos"][:, t], k)[live])
    assert env.status() == 0
    env.close()


def test_generic_autoreset_matches_oracle(cuda_device):
    import gym_multigrid_b200 as mg
    g = load_golden("generic_9x9_a3")
    E, _, A = g["actions"].shape
    S, n, max_steps = 9, 24 * 11 + 7, 25
    idx = np.arange(n) % E
    init_obs, init_pos = g["init_obs"][idx, 0], g["init_pos"][idx]
    env = mg.make_generic_vec(n, S, num_agents=A, max_steps=max_steps, autoreset=True)
    env.enable_final_observation()
    env.set_layout(init_obs, init_pos)
    env.reset()
    o = oc.GenericOracle(n, S, S, A, max_steps)
    o.set_state_from_obs(init_obs, init_pos)
    init = (o.gcell.copy(), o.gstate.copy(), o.pos.copy())
    rng = np.random.default_rng(3)
    n_term = 0
    for t in range(80):
        order = np.stack([rng.permutation(A) for _ in range(n)]).astype(np.uint8)
        act = rng.choice(4, size=(n, A), p=[0.1, 0.2, 0.2, 0.5]).astype(np.int8)
        env.set_order(order)
        obs, rew, term, trunc, info = env.step(torch.as_tensor(act, device=cuda_device))
        oobs, orew, oterm, otrunc = o.step(act, order)
        d = oterm | otrunc
        assert np.array_equal(_np(rew), orew) and np.array_equal(_np(term), oterm) and np.array_equal(_np(trunc), otrunc), f"step {t}"
        assert np.array_equal(_np(info["final_observation"])[d], oobs[d]), f"step {t}: final observation"
        o.gcell[d], o.gstate[d], o.pos[d], o.step_count[d] = init[0][d], init[1][d], init[2][d], 0     # gymnasium 0.29 same-step autoreset
        assert np.array_equal(_np(obs), o.encode()), f"step {t}: obs"
        assert np.array_equal(_np(env.agent_pos), o.pos) and np.array_equal(_np(env.step_count), o.step_count)
        n_term += int(oterm.sum())
    assert n_term > 0 and int(env.episode_count.min()) >= 3 and env.status() == 0
    env.close()


def test_generic_philox_orders_are_shard_invariant_and_bad_actions_flagged(cuda_device):
    import gym_multigrid_b200 as mg
    g = load_golden("generic_12x12_a5")
    A, S, n = 5, 12, 200
    idx = np.arange(n) % g["actions"].shape[0]
    acts = np.random.default_rng(0).integers(0, 4, size=(30, n, A)).astype(np.int8)

    def run(base, m):
        env = mg.make_generic_vec(m, S, num_agents=A, max_steps=20, seed=7, env_id_base=base)
        env.set_layout(g["init_obs"][idx[base:base + m], 0], g["init_pos"][idx[base:base + m]])
        env.reset()
        out = []
        for t in range(30):
            obs, rew, *_ = env.step(torch.as_tensor(acts[t, base:base + m], device=cuda_device))
            out.append((_np(obs).copy(), _np(rew).copy()))
        assert env.status() == 0
        env.close()
        return out

    full, lo, hi = run(0, n), run(0, 77), run(77, n - 77)
    for t in range(30):
        assert np.array_equal(full[t][0], np.concatenate([lo[t][0], hi[t][0]]))
        assert np.array_equal(full[t][1], np.concatenate([lo[t][1], hi[t][1]]))
    assert any(not np.array_equal(full[t][0], full[0][0]) for t in range(1, 30))
    env = mg.make_generic_vec(4, S, num_agents=A, max_steps=20)
    env.set_layout(g["init_obs"][0, 0], g["init_pos"][0])
    env.reset()
    env.step(torch.full((4, A), 6, dtype=torch.int8, device=cuda_device))      # toggle: the reference raises (multigrid.py:447)
    assert env.status() & 8
    env.close()


@pytest.mark.parametrize("stem", ["partial6_9x9_a3", "partial6_12x12_a5"])
def test_generic_partial_views_match_reference(stem, cuda_device):
    """encode_dim-6 gen_obs: the reference's gen_obs_grid + encode_for_agents outputs on recorded DefaultWorld states."""
    import gym_multigrid_b200 as mg
    g = load_golden(stem)
    S, A = int(g["meta_size"]), int(g["meta_num_agents"])
    for V in (3, 5, 7):
        for st in (False, True):
            sel = np.where((g["V"] == V) & (g["see_through"] == st))[0]
            if len(sel) == 0:
                continue
            env = mg.make_generic_vec(len(sel), S, num_agents=A, autoreset=False)
            env.set_state_from_obs(g["obs6"][sel], g["pos"][sel])
            assert np.array_equal(_np(env.gen_obs(V, st)), g["views"][sel][:, :, :V, :V]), f"V={V} see_through={st}"
            env.close()


@pytest.mark.parametrize("V", [1, 2, 4, 9, 15])
def test_generic_partial_views_match_oracle_any_size(V, cuda_device):
    import gym_multigrid_b200 as mg
    g = load_golden("partial6_12x12_a5")
    n, A, S = 700, 5, 12
    idx = np.arange(n) % len(g["V"])
    env = mg.make_generic_vec(n, S, num_agents=A, autoreset=False)
    env.set_state_from_obs(g["obs6"][idx], g["pos"][idx])
    o = oc.GenericOracle(n, S, S, A, 100)
    o.set_state_from_obs(g["obs6"][idx], g["pos"][idx])
    dirs = np.random.default_rng(V).integers(0, 4, size=(n, A)).astype(np.uint8)
    for st in (False, True):
        assert np.array_equal(_np(env.gen_obs(V, st)), o.partial_views(V, st))
        assert np.array_equal(_np(env.gen_obs(V, st, dirs=dirs)), o.partial_views(V, st, dirs=dirs))
    env.close()
