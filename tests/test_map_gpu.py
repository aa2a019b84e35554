"""GPU parity suite for the static-map families (Maze, CtF MvN) through the C ABI.
 * bit-exact replay of golden traces recorded from the reference MazeSingleAgentEnv / CtFMvNEnv;
 * CUDA vs the C oracle in Philox mode (same counter-based RNG), incl. autoreset + final_observation;
 * reference dtypes (float64 / int64 observations), host-buffer path, invariants at scale."""
import numpy as np
import pytest
import torch

import oracle as oc
from replay import load_golden

pytestmark = pytest.mark.gpu
MAZE = ["maze_board13", "maze_board13_penalty", "maze_gen64", "maze_gen64_penalty"]
CTF = ["ctf_2v2", "ctf_3v4", "ctf_2v2_penalty", "ctf_1v1", "ctf_3v4_penalty_battles"]   # the last: collision penalty AND battles in the same episodes


def _np(t):
    return t.cpu().numpy()


def _tile(x, k):
    return np.concatenate([x] * k)


@pytest.mark.parametrize("stem", MAZE)
@pytest.mark.parametrize("ref_dtypes", [False, True])
def test_maze_replay_bit_exact(stem, ref_dtypes, cuda_device):
    import gym_multigrid_b200 as mg
    g = load_golden(stem)
    E, T = g["actions"].shape
    k = 7
    env = mg.make_maze_vec(E * k, g["field_map"], obstacle_penalty_ratio=float(g["meta_obstacle_penalty_ratio"]),
                           autoreset=False, reference_dtypes=ref_dtypes)
    env.set_trace(start_index=_tile(g["start_index"], k))
    obs, _ = env.reset()
    assert obs.dtype == (torch.float64 if ref_dtypes else torch.uint8)
    assert np.array_equal(_np(obs), _tile(g["init_obs"], k))
    info = env.get_info()
    assert list(info) == ["d_a_f", "d_a_ob"]
    assert np.array_equal(np.stack([_np(v) for v in info.values()], 1), _tile(g["init_info"], k))
    env.with_info = True
    for t in range(T):
        live = _tile(g["length"] > t, k)
        act = _tile(np.where(g["length"] > t, g["actions"][:, t], 0), k).astype(np.int8)
        obs, rew, term, trunc, info = env.step(torch.as_tensor(act, device=cuda_device))
        assert np.array_equal(np.stack([_np(info[q]) for q in ("d_a_f", "d_a_ob")], 1)[live], _tile(g["info"][:, t], k)[live]), f"step {t}: info"
        assert np.array_equal(_np(obs)[live], _tile(g["obs"][:, t], k)[live]), f"step {t}: obs"
        assert np.array_equal(_np(rew)[live], _tile(g["reward"][:, t], k)[live]), f"step {t}: reward (float64 bit-exact)"
        assert np.array_equal(_np(term)[live], _tile(g["terminated"][:, t], k)[live])
        assert np.array_equal(_np(trunc)[live], _tile(g["truncated"][:, t], k)[live])
        assert np.array_equal(_np(env.agent_pos)[live, 0], _tile(g["pos"][:, t], k)[live])
        assert np.array_equal(_np(env.agent_dir)[live, 0], _tile(g["dir"][:, t], k)[live])
    assert env.status() == 0
    env.close()


@pytest.mark.parametrize("stem", CTF)
def test_ctf_replay_bit_exact(stem, cuda_device):
    import gym_multigrid_b200 as mg
    g = load_golden(stem)
    E, T, nb = g["actions"].shape
    nr = int(g["meta_num_red"])
    k = 9
    env = mg.make_ctf_vec(E * k, g["field_map"], num_blue_agents=nb, num_red_agents=nr,
                          obstacle_penalty_ratio=float(g["meta_obstacle_penalty_ratio"]), autoreset=False,
                          reference_dtypes=(stem == "ctf_3v4"))
    env.set_trace(blue_place=_tile(g["blue_place"], k), red_place=_tile(g["red_place"], k))
    obs, _ = env.reset()
    assert np.array_equal(_np(obs), _tile(g["init_obs"], k)) and np.array_equal(_np(env.agent_pos), _tile(g["init_pos"], k))
    assert np.array_equal(np.stack([_np(v) for v in env.get_info().values()], 1), _tile(g["init_info"], k)), "reset info (ctf.py:1165-1182)"
    ident = np.arange(nb + nr, dtype=np.uint8)[None]
    for t in range(T):
        lv = g["length"] > t
        live = _tile(lv, k)
        tr = env.set_trace(red_actions=_tile(g["red_actions"][:, t], k), order=_tile(np.where(lv[:, None], g["order"][:, t], ident), k),
                           blue_win=_tile(g["blue_win"][:, t], k))
        act = _tile(np.where(lv[:, None], g["actions"][:, t], 0), k).astype(np.int8)
        obs, rew, term, trunc, _ = env.step(torch.as_tensor(act, device=cuda_device))
        assert np.array_equal(_np(obs)[live], _tile(g["obs"][:, t], k)[live]), f"step {t}: obs"
        assert np.array_equal(_np(rew)[live], _tile(g["reward"][:, t], k)[live]), f"step {t}: reward (float64 bit-exact)"
        assert np.array_equal(_np(term)[live], _tile(g["terminated"][:, t], k)[live])
        assert np.array_equal(_np(trunc)[live], _tile(g["truncated"][:, t], k)[live])
        assert np.array_equal(_np(env.agent_pos)[live], _tile(g["pos"][:, t], k)[live])
        assert np.array_equal(_np(env.agent_dir)[live], _tile(g["dir"][:, t], k)[live])
        assert np.array_equal(_np(env.agent_terminated)[live], _tile(g["dead"][:, t], k)[live].astype(bool))
        assert np.array_equal(_np(tr["battles_used"])[live], _tile(g["n_battles"][:, t], k)[live])
        assert np.array_equal(np.stack([_np(v) for v in env.get_info().values()], 1)[live], _tile(g["info"][:, t], k)[live]), f"step {t}: info"
        gs = env.game_stats()
        assert np.array_equal(np.stack([_np(gs["blue_flag_captured"]), _np(gs["red_flag_captured"])], 1)[live], _tile(g["stats_flags"][:, t], k)[live].astype(bool))
        assert np.array_equal(np.concatenate([_np(gs["blue_agent_defeated"]), _np(gs["red_agent_defeated"])], 1)[live],
                              _tile(g["stats_defeated"][:, t], k)[live].astype(bool)), f"step {t}: game_stats"
    assert env.status() == 0
    env.close()


@pytest.mark.parametrize("stem", ["ctf_2v2_carry", "ctf_3v4_penalty_carry"])
def test_ctf_flags_survive_reset_like_the_reference(stem, cuda_device):
    """Consecutive episodes of ONE reference env instance per slot (`carry_agent_flags`): masked resets keep the agents'
    terminated / collided flags, every session walks through its episodes at its own pace, k copies of every session."""
    import gym_multigrid_b200 as mg
    from replay import ctf_session_schedule
    g = load_golden(stem)
    nb, nr = int(g["meta_num_blue"]), int(g["meta_num_red"])
    S, k = int(g["meta_sessions"]), 5
    env = mg.make_ctf_vec(S * k, g["field_map"], num_blue_agents=nb, num_red_agents=nr, max_steps=int(g["meta_max_steps"]),
                          obstacle_penalty_ratio=float(g["meta_obstacle_penalty_ratio"]), autoreset=False, carry_agent_flags=True)
    ident = np.arange(nb + nr, dtype=np.uint8)[None]
    steps = 0
    for ev in ctf_session_schedule(g):
        if ev[0] == "reset":
            _, mask, ep = ev
            m = _tile(mask, k)
            env.set_trace(blue_place=_tile(g["blue_place"][ep], k), red_place=_tile(g["red_place"][ep], k))
            obs, _ = env.reset(mask=torch.as_tensor(m, device=cuda_device))
            assert np.array_equal(_np(obs)[m], _tile(g["init_obs"][ep], k)[m]) and np.array_equal(_np(env.agent_pos)[m], _tile(g["init_pos"][ep], k)[m])
            assert np.array_equal(_np(env.agent_terminated)[m], _tile(g["init_dead"][ep], k)[m].astype(bool)), "terminated flags the episode starts with"
            assert np.array_equal(((_np(env.agent_flags) >> 1) & 1)[m], _tile(g["init_collided"][ep], k)[m])
            continue
        _, lv, ep, t = ev
        live = _tile(lv, k)
        tr = env.set_trace(red_actions=_tile(np.where(lv[:, None], g["red_actions"][ep, t], 0).astype(np.int8), k),
                           order=_tile(np.where(lv[:, None], g["order"][ep, t], ident).astype(np.uint8), k), blue_win=_tile(g["blue_win"][ep, t], k))
        act = _tile(np.where(lv[:, None], g["actions"][ep, t], 0), k).astype(np.int8)
        obs, rew, term, trunc, _ = env.step(torch.as_tensor(act, device=cuda_device))
        assert np.array_equal(_np(obs)[live], _tile(g["obs"][ep, t], k)[live]), "obs"
        assert np.array_equal(_np(rew)[live], _tile(g["reward"][ep, t], k)[live]), "reward (float64 bit-exact)"
        assert np.array_equal(_np(term)[live], _tile(g["terminated"][ep, t], k)[live]) and np.array_equal(_np(trunc)[live], _tile(g["truncated"][ep, t], k)[live])
        assert np.array_equal(_np(env.agent_pos)[live], _tile(g["pos"][ep, t], k)[live]) and np.array_equal(_np(env.agent_dir)[live], _tile(g["dir"][ep, t], k)[live])
        assert np.array_equal(_np(env.agent_terminated)[live], _tile(g["dead"][ep, t], k)[live].astype(bool))
        assert np.array_equal(((_np(env.agent_flags) >> 1) & 1)[live], _tile(g["collided"][ep, t], k)[live])
        assert np.array_equal(_np(tr["battles_used"])[live], _tile(g["n_battles"][ep, t], k)[live])
        steps += int(lv.sum())
    assert steps == int(g["length"].sum()) and env.status() == 0
    env.close()


@pytest.mark.parametrize("carry", [False, True])
def test_ctf_autoreset_with_carried_flags_matches_oracle(carry, cuda_device):
    """Philox mode with same-step autoreset, short episodes: the flags of defeated agents survive (or not) the autoreset exactly as in the oracle."""
    import gym_multigrid_b200 as mg
    g = load_golden("ctf_2v2")
    n, nb, nr = 4099, 3, 2
    env = mg.make_ctf_vec(n, g["field_map"], num_blue_agents=nb, num_red_agents=nr, max_steps=9, seed=21, carry_agent_flags=carry, obstacle_penalty_ratio=0.25)
    o = oc.CtfOracle(g["field_map"], n, nb, nr, max_steps=9, carry_agent_flags=carry, obstacle_penalty_ratio=0.25)
    r = oc.map_rng(mode=1, seed=21)
    obs, _ = env.reset()
    assert np.array_equal(_np(obs), o.reset(r))
    rng = np.random.default_rng(4)
    started_dead = 0
    for t in range(60):
        act = rng.integers(1, 5, size=(n, nb)).astype(np.int8)
        obs, rew, term, trunc, _ = env.step(torch.as_tensor(act, device=cuda_device))
        oo, orew, oterm, otrunc = o.step(act, r, autoreset=True)
        assert np.array_equal(_np(obs), oo) and np.array_equal(_np(rew), orew), f"step {t}"
        assert np.array_equal(_np(term), oterm) and np.array_equal(_np(trunc), otrunc)
        assert np.array_equal(_np(env.agent_flags) & 3, o.flags & 3), f"step {t}: flags"
        started_dead += int(((o.flags & 1).any(1) & (o.step_count == 0)).sum())
    assert (started_dead > 0) == carry and env.status() == 0
    env.close()


@pytest.mark.parametrize("stem,n,pen", [("maze_board13", 1000, 0.0), ("maze_board13", 333, 0.5), ("maze_gen64", 300, 0.5)])
def test_maze_philox_matches_oracle(stem, n, pen, cuda_device):
    import gym_multigrid_b200 as mg
    g = load_golden(stem)
    env = mg.make_maze_vec(n, g["field_map"], obstacle_penalty_ratio=pen, max_steps=30, seed=99, env_id_base=5)
    env.enable_final_observation()
    o = oc.MazeOracle(g["field_map"], n, obstacle_penalty_ratio=pen, max_steps=30)
    r = oc.map_rng(mode=1, seed=99, env_id_base=5)
    obs, _ = env.reset()
    assert np.array_equal(_np(obs), o.reset(r))
    rng = np.random.default_rng(1)
    for t in range(100):
        act = rng.integers(0, 5, size=n).astype(np.int8)
        obs, rew, term, trunc, info = env.step(torch.as_tensor(act, device=cuda_device))
        oo, orew, oterm, otrunc, ofin = o.step(act, r, autoreset=True, want_final_obs=True)
        assert np.array_equal(_np(obs), oo), f"step {t}: obs"
        assert np.array_equal(_np(rew), orew) and np.array_equal(_np(term), oterm) and np.array_equal(_np(trunc), otrunc)
        d = oterm | otrunc
        assert np.array_equal(_np(info["final_observation"])[d], ofin[d])
        assert np.array_equal(_np(env.agent_pos), o.pos) and np.array_equal(_np(env.step_count), o.step_count)
    assert env.status() == 0 and int(env.episode_count.min()) >= 3
    env.close()


@pytest.mark.parametrize("nb,nr,n,pen", [(2, 2, 2000, 0.0), (3, 4, 515, 0.0), (2, 2, 300, 0.5), (8, 8, 200, 0.0),
                                         (8, 8, 260, 0.5), (1, 3, 300, 0.25), (7, 9, 150, 0.0), (5, 2, 129, 0.5)])
def test_ctf_philox_matches_oracle(nb, nr, n, pen, cuda_device):
    import gym_multigrid_b200 as mg
    g = load_golden("ctf_2v2")
    env = mg.make_ctf_vec(n, g["field_map"], num_blue_agents=nb, num_red_agents=nr, obstacle_penalty_ratio=pen, max_steps=40,
                          seed=7, env_id_base=3)
    env.enable_final_observation()
    o = oc.CtfOracle(g["field_map"], n, nb, nr, obstacle_penalty_ratio=pen, max_steps=40)
    r = oc.map_rng(mode=1, seed=7, env_id_base=3)
    obs, _ = env.reset()
    assert np.array_equal(_np(obs), o.reset(r))
    rng = np.random.default_rng(2)
    battles = 0
    for t in range(130):
        act = rng.integers(0, 5, size=(n, nb)).astype(np.int8)
        obs, rew, term, trunc, info = env.step(torch.as_tensor(act, device=cuda_device))
        oo, orew, oterm, otrunc, ofin = o.step(act, r, autoreset=True, want_final_obs=True)
        assert np.array_equal(_np(obs), oo), f"step {t}: obs"
        assert np.array_equal(_np(rew), orew), f"step {t}: reward"
        assert np.array_equal(_np(term), oterm) and np.array_equal(_np(trunc), otrunc)
        d = oterm | otrunc
        assert np.array_equal(_np(info["final_observation"])[d], ofin[d])
        assert np.array_equal(_np(env.agent_pos), o.pos) and np.array_equal(_np(env.agent_flags), o.flags)
        if t % 8 == 0:
            assert np.array_equal(np.stack([_np(v) for v in env.get_info().values()], 1), o.info()), f"step {t}: info"
        assert np.array_equal(_np(env.agent_dir), o.dir)
        battles += int((np.abs(orew + 0.01 * nb) > 0.2).sum())
    assert env.status() == 0 and battles > 0
    env.close()


@pytest.mark.parametrize("family", ["ctf", "maze"])
def test_large_batch_kernel_variant_matches_oracle(family, cuda_device):
    """Launches of >= 262 144 envs take the 8-CTAs-per-SM register allocation of map_kernel: same results, ragged last tile."""
    import gym_multigrid_b200 as mg
    n = 262144 + 77
    if family == "ctf":
        fm = load_golden("ctf_2v2")["field_map"]
        env = mg.make_ctf_vec(n, fm, max_steps=12, seed=4, env_id_base=1)
        o = oc.CtfOracle(fm, n, 2, 2, max_steps=12)
    else:
        fm = load_golden("maze_board13")["field_map"]
        env = mg.make_maze_vec(n, fm, max_steps=12, seed=4, env_id_base=1)
        o = oc.MazeOracle(fm, n, max_steps=12)
    r = oc.map_rng(mode=1, seed=4, env_id_base=1)
    obs, _ = env.reset()
    assert np.array_equal(_np(obs), o.reset(r))
    gen = torch.Generator(device=cuda_device).manual_seed(5)
    for t in range(16):
        act = torch.randint(0, 5, (n, 2) if family == "ctf" else (n,), generator=gen, device=cuda_device, dtype=torch.int8)
        obs, rew, term, trunc, _ = env.step(act)
        oo, orew, oterm, otrunc = o.step(_np(act), r, autoreset=True)
        assert np.array_equal(_np(obs), oo) and np.array_equal(_np(rew), orew), f"step {t}"
        assert np.array_equal(_np(term), oterm) and np.array_equal(_np(trunc), otrunc)
    assert np.array_equal(_np(env.agent_pos), o.pos) and env.status() == 0 and int(env.episode_count.min()) >= 2
    env.close()


def test_ctf_other_observation_options_and_host_path(cuda_device):
    import gym_multigrid_b200 as mg
    g = load_golden("ctf_2v2")
    n = 700
    a, b = (mg.make_ctf_vec(n, g["field_map"], seed=4) for _ in range(2))
    a.reset(); b.reset()
    rng = np.random.default_rng(0)
    for t in range(30):
        act = rng.integers(0, 5, size=(n, 2)).astype(np.int8)
        x = a.step(torch.as_tensor(act, device=cuda_device))
        y = b.step(act.astype(np.float32) + 0.2)   # numpy in/out through mg_step_host; rounded like ctf.py:1303-1304
        for u, v in zip(x[:4], y[:4]):
            assert np.array_equal(_np(u), v)
    d = a.positional_obs()
    assert d["blue_agent"].shape == (n, 4) and d["blue_territory"].shape == (n, 2 * 48) and d["terminated_agents"].shape == (n, 4)
    f = a.flattened_obs()
    assert f.shape == (n, 216) and f.dtype == torch.int64   # ctf.py:940-948
    # the map observation is consistent with the positions
    obs = _np(a._obs); pos = _np(a.agent_pos); dead = _np(a.agent_terminated)
    for e in range(0, n, 97):
        for i in range(4):
            want = 6 if dead[e, i] else (2 if i < 2 else 3)
            assert obs[e, pos[e, i, 1], pos[e, i, 0]] == want
    a.close(); b.close()


def test_maze_invariants_at_scale(cuda_device):
    import gym_multigrid_b200 as mg
    g = load_golden("maze_gen64")
    n = 16384
    env = mg.make_maze_vec(n, g["field_map"], seed=1)
    fm = torch.as_tensor(g["field_map"], device=cuda_device)
    obs, _ = env.reset()
    gen = torch.Generator(device=cuda_device).manual_seed(0)
    for t in range(1, 121):
        act = torch.randint(0, 5, (n,), generator=gen, device=cuda_device, dtype=torch.int8)
        obs, rew, term, trunc, _ = env.step(act)
        assert bool(((obs == 1).sum(dim=(1, 2)) == 1).all()), "exactly one agent cell"
        assert bool(((obs != 1) <= (obs == fm[None])).all()), "everything else is the static map"
        p = env.agent_pos[:, 0].long()
        assert bool((fm[p[:, 0], p[:, 1]] != 3).all()), "penalty 0: never on an obstacle"
        assert bool((env.step_count[term | trunc] == 0).all()), "finished envs were reset in the same step"
        assert bool((env.step_count[~(term | trunc)] > 0).all())
    assert env.status() == 0
    env.close()


def test_map_create_rejects_bad_configs(cuda_device):
    import gym_multigrid_b200 as mg
    with pytest.raises(ValueError):
        mg.make_maze_vec(4, np.zeros((5, 6)))
    with pytest.raises(ValueError):
        mg.make_maze_vec(4, np.full((5, 5), 3))
    with pytest.raises(ValueError):
        mg.make_ctf_vec(4, np.zeros((6, 6)))            # no flags
    with pytest.raises(ValueError):
        mg.make_ctf_vec(4, load_golden("ctf_2v2")["field_map"], num_blue_agents=12, num_red_agents=12)
    # Ctf1v1Env with a collision penalty: the reference raises (ctf.py:639); the C ABI refuses it too, not only the Python class
    with pytest.raises(ValueError, match="obstacle_penalty"):
        mg.make_ctf_vec(4, load_golden("ctf_2v2")["field_map"], num_blue_agents=1, num_red_agents=1, variant_1v1=True, obstacle_penalty_ratio=0.5)
    e = mg.make_ctf_vec(4, load_golden("ctf_2v2")["field_map"])
    with pytest.raises(ValueError, match="mask"):
        e.reset(seed=3, mask=np.array([1, 0, 0, 1]))
    e.close()


def test_ctf1v1_replay_and_philox(cuda_device):
    """Ctf1v1Env (ctf.py:50-654): golden replay, then Philox mode against the oracle."""
    import gym_multigrid_b200 as mg
    g = load_golden("ctf1v1")
    E, T, _ = g["actions"].shape
    k = 5
    env = mg.make_ctf1v1_vec(E * k, g["field_map"], autoreset=False)
    env.set_trace(blue_place=_tile(g["blue_place"], k), red_place=_tile(g["red_place"], k))
    obs, _ = env.reset()
    assert np.array_equal(_np(obs), _tile(g["init_obs"], k))
    for t in range(T):
        lv = g["length"] > t
        live = _tile(lv, k)
        tr = env.set_trace(red_actions=_tile(g["red_actions"][:, t], k), blue_win=_tile(g["blue_win"][:, t], k))
        act = _tile(np.where(lv[:, None], g["actions"][:, t], 0), k).astype(np.int8)
        obs, rew, term, trunc, _ = env.step(torch.as_tensor(act, device=cuda_device))
        assert np.array_equal(_np(obs)[live], _tile(g["obs"][:, t], k)[live]), f"step {t}: obs"
        assert np.array_equal(_np(rew)[live], _tile(g["reward"][:, t], k)[live]), f"step {t}: reward"
        assert np.array_equal(_np(term)[live], _tile(g["terminated"][:, t], k)[live])
        assert np.array_equal(_np(trunc)[live], _tile(g["truncated"][:, t], k)[live])
        assert np.array_equal(_np(tr["battles_used"])[live], _tile(g["n_battles"][:, t], k)[live])
        assert np.array_equal(np.stack([_np(v) for v in env.get_info().values()], 1)[live], _tile(g["info"][:, t], k)[live]), f"step {t}: info"
        gs = env.game_stats()
        assert np.array_equal(np.stack([_np(gs["blue_flag_captured"]), _np(gs["red_flag_captured"])], 1)[live], _tile(g["stats_flags"][:, t], k)[live].astype(bool))
        assert np.array_equal(np.concatenate([_np(gs["blue_agent_defeated"]), _np(gs["red_agent_defeated"])], 1)[live],
                              _tile(g["stats_defeated"][:, t], k)[live].astype(bool)), f"step {t}: game_stats"
    assert env.status() == 0
    env.close()
    n = 3000
    env = mg.make_ctf1v1_vec(n, g["field_map"], max_steps=30, seed=21)
    o = oc.CtfOracle(g["field_map"], n, 1, 1, max_steps=30, variant_1v1=True)
    r = oc.map_rng(mode=1, seed=21)
    obs, _ = env.reset()
    assert np.array_equal(_np(obs), o.reset(r))
    rng = np.random.default_rng(0)
    for t in range(80):
        act = rng.integers(0, 5, size=(n, 1)).astype(np.int8)
        obs, rew, term, trunc, _ = env.step(torch.as_tensor(act, device=cuda_device))
        oo, orew, oterm, otrunc = o.step(act, r, autoreset=True)
        assert np.array_equal(_np(obs), oo) and np.array_equal(_np(rew), orew)
        assert np.array_equal(_np(term), oterm) and np.array_equal(_np(trunc), otrunc)
    with pytest.raises(ValueError):
        mg.make_ctf1v1_vec(4, g["field_map"], obstacle_penalty_ratio=0.5)
    env.close()


def test_full_size_invariants_ctf_and_maze(cuda_device):
    """BASELINE configs 3 / 4 at full size (1 M CtF envs, 131 072 Maze envs on a 64x64 map): size-independent properties."""
    import gym_multigrid_b200 as mg
    fm = load_golden("ctf_2v2")["field_map"]
    n = 1 << 20
    env = mg.make_ctf_vec(n, fm, seed=12)
    obs, _ = env.reset()
    gen = torch.Generator(device=cuda_device).manual_seed(9)
    static = torch.as_tensor(fm.T.copy(), device=cuda_device)      # obs[y][x] = map[x][y]
    steps_before = env.step_count.clone()
    for t in range(12):
        obs, rew, term, trunc, _ = env.step(torch.randint(0, 5, (n, 2), generator=gen, device=cuda_device, dtype=torch.int8))
        agents = (obs == 2) | (obs == 3)
        dead = env.agent_terminated
        # every living agent is drawn exactly once unless a later agent hides it (positions are distinct for living agents)
        assert int(agents.sum(dim=(1, 2)).max()) <= 4 and bool((agents.sum(dim=(1, 2)) == (~dead).sum(dim=1)).all())
        assert bool(((obs == static) | agents | (obs == 6)).all()), "cells are the static map, an agent, or a defeated agent"
        assert bool((rew >= -(1.0 + 2 * 0.25 + 0.02) - 1e-12).all()) and bool((rew <= 2 * 1.0 + 4 * 0.25).all())
        done = term | trunc
        assert bool((env.step_count[done] == 0).all()) and bool((env.step_count[~done] == steps_before[~done] + 1).all())
        steps_before = env.step_count.clone()
    assert env.status() == 0 and int(env.episode_count.max()) >= 2
    env.close()

    g = load_golden("maze_gen64")
    fm = torch.as_tensor(g["field_map"], device=cuda_device)
    n = 131072
    env = mg.make_maze_vec(n, g["field_map"], seed=3)
    obs, _ = env.reset()
    for t in range(6):
        obs, rew, term, trunc, _ = env.step(torch.randint(0, 5, (n,), generator=gen, device=cuda_device, dtype=torch.int8))
        diff = obs != fm
        assert bool((diff.sum(dim=(1, 2)) == 1).all()), "the observation is the map plus exactly one agent cell"
        pos = env.agent_pos[:, 0].long()
        assert bool((obs[torch.arange(n, device=cuda_device), pos[:, 0], pos[:, 1]] == 1).all())
        assert bool((fm[pos[:, 0], pos[:, 1]] != 3).all()), "obstacles are never entered with penalty 0"
        assert bool(((rew == -0.01) | (rew == 1.0 - 0.01) | term).all())
    assert env.status() == 0
    env.close()


def test_ctf_external_enemy_policy(cuda_device):
    """`enemy_policies` other than RwPolicy: red actions supplied by the caller (here: a scripted 'always move up / left'
    opponent computed from the observation on the device), shuffle and battles from the Philox stream; CUDA == oracle."""
    import gym_multigrid_b200 as mg
    g = load_golden("ctf_2v2")
    n, nb, nr = 1500, 2, 2
    env = mg.make_ctf_vec(n, g["field_map"], num_blue_agents=nb, num_red_agents=nr, max_steps=30, seed=9)
    o = oc.CtfOracle(g["field_map"], n, nb, nr, max_steps=30)
    obs, _ = env.reset()
    assert np.array_equal(_np(obs), o.reset(oc.map_rng(mode=1, seed=9)))
    red = env.set_red_actions(torch.zeros((n, nr), dtype=torch.int8, device=cuda_device))
    gen = torch.Generator(device=cuda_device).manual_seed(4)
    r = None
    for t in range(70):
        red.copy_(torch.where(env.agent_pos[:, nb:, 0] > 4, 2, 1).to(torch.int8))      # policy: head for x <= 4, then walk left
        act = torch.randint(0, 5, (n, nb), generator=gen, device=cuda_device, dtype=torch.int8)
        r = oc.map_rng(mode=1, seed=9, red_actions=_np(red))
        obs, rew, term, trunc, _ = env.step(act)
        oo, orew, oterm, otrunc = o.step(_np(act), r, autoreset=True)
        assert np.array_equal(_np(obs), oo) and np.array_equal(_np(rew), orew), f"step {t}"
        assert np.array_equal(_np(term), oterm) and np.array_equal(_np(trunc), otrunc)
    env.set_red_actions(None)
    obs, rew, *_ = env.step(torch.zeros((n, nb), dtype=torch.int8, device=cuda_device))
    oo, orew, *_ = o.step(np.zeros((n, nb), np.int8), oc.map_rng(mode=1, seed=9), autoreset=True)
    assert np.array_equal(_np(obs), oo) and env.status() == 0
    env.close()


def test_ctf_vec_enemy_policies(cuda_device):
    """`CtfVecEnv.set_enemy_policies`: the reference's heuristic opponents (policy/ctf/heuristic.py) deciding on the host for a whole
    batch.  Twin policies with an equally seeded generator, fed the same positional observations in the same order (env-major, red
    agents in index order), must decide the same actions; the CUDA step with those decisions == the oracle stepped with the twins'."""
    import gym_multigrid_b200 as mg
    from gym_multigrid_b200.policy.ctf.heuristic import FightPolicy, PatrolFightPolicy
    g = load_golden("ctf_2v2")
    fm = g["field_map"].astype(np.float64)
    n, nb, nr = 37, 2, 2
    env = mg.make_ctf_vec(n, g["field_map"], num_blue_agents=nb, num_red_agents=nr, max_steps=25, seed=5)
    o = oc.CtfOracle(g["field_map"], n, nb, nr, max_steps=25)
    mine = env.set_enemy_policies([FightPolicy(randomness=0.8), PatrolFightPolicy(fm, randomness=0.8)], random_generator=np.random.default_rng(11))
    assert mine[0].field_map is not None and mine[0].random_generator is mine[1].random_generator
    twin_gen = np.random.default_rng(11)
    twins = [FightPolicy(fm, random_generator=twin_gen, randomness=0.8), PatrolFightPolicy(fm, random_generator=twin_gen, randomness=0.8)]
    obs, _ = env.reset()
    assert np.array_equal(_np(obs), o.reset(oc.map_rng(mode=1, seed=5)))
    gen = torch.Generator(device=cuda_device).manual_seed(6)
    for t in range(60):
        d = {k: _np(v) for k, v in env.positional_obs().items()}
        pos = _np(env.agent_pos).astype(np.int64)
        want = np.array([[int(twins[k].act({key: v[e] for key, v in d.items()}, tuple(pos[e, nb + k].tolist()))) for k in range(nr)]
                         for e in range(n)], np.int8)
        act = torch.randint(0, 5, (n, nb), generator=gen, device=cuda_device, dtype=torch.int8)
        obs, rew, term, trunc, _ = env.step(act)
        assert np.array_equal(env._red_host, want), f"step {t}"
        oo, orew, oterm, otrunc = o.step(_np(act), oc.map_rng(mode=1, seed=5, red_actions=want), autoreset=True)
        assert np.array_equal(_np(obs), oo) and np.array_equal(_np(rew), orew), f"step {t}"
        assert np.array_equal(_np(term), oterm) and np.array_equal(_np(trunc), otrunc)
    assert len(set(env._red_host.reshape(-1).tolist())) > 1
    assert env.set_enemy_policies(None) is None                      # back to the device opponent
    obs, rew, *_ = env.step(torch.zeros((n, nb), dtype=torch.int8, device=cuda_device))
    oo, orew, *_ = o.step(np.zeros((n, nb), np.int8), oc.map_rng(mode=1, seed=5), autoreset=True)
    assert np.array_equal(_np(obs), oo) and env.status() == 0
    with pytest.raises(AssertionError):
        env.set_enemy_policies([FightPolicy()])                      # ctf.py:779: one policy per red agent
    env.close()


@pytest.mark.parametrize("stem", ["ctf_2v2_flat", "ctf_3v4_flat"])
def test_ctf_flattened_obs_matches_reference(stem, cuda_device):
    """observation_option="flattened" (ctf.py:1084-1104; the option scripts/main_mvn_ctf_rl.py trains on) recorded from the
    reference, replayed through the CUDA step (trace mode) + mg_ctf_flat_obs, tiled over ragged tiles."""
    import gym_multigrid_b200 as mg
    g = load_golden(stem)
    E, T, nb = g["actions"].shape
    nr = int(g["meta_num_red"])
    k = 7
    env = mg.make_ctf_vec(E * k, g["field_map"], num_blue_agents=nb, num_red_agents=nr, autoreset=False)
    env.set_trace(blue_place=_tile(g["blue_place"], k), red_place=_tile(g["red_place"], k))
    env.reset()
    f = env.flattened_obs()
    assert f.dtype == torch.int64 and np.array_equal(_np(f), _tile(g["init_obs"], k))
    ident = np.arange(nb + nr, dtype=np.uint8)[None]
    for t in range(T):
        lv = g["length"] > t
        live = _tile(lv, k)
        env.set_trace(red_actions=_tile(g["red_actions"][:, t], k), order=_tile(np.where(lv[:, None], g["order"][:, t], ident), k),
                      blue_win=_tile(g["blue_win"][:, t], k))
        env.step(torch.as_tensor(_tile(np.where(lv[:, None], g["actions"][:, t], 0), k).astype(np.int8), device=cuda_device))
        assert np.array_equal(_np(env.flattened_obs())[live], _tile(g["obs"][:, t], k)[live]), f"step {t}"
    d = env.positional_obs()
    assert list(d) == ["blue_agent", "red_agent", "blue_flag", "red_flag", "blue_territory", "red_territory", "obstacle", "terminated_agents"]
    assert np.array_equal(_np(d["blue_agent"]).reshape(-1, nb, 2), _np(env.agent_pos)[:, :nb])
    assert np.array_equal(_np(d["terminated_agents"]).astype(bool), _np(env.agent_terminated))
    assert env.status() == 0
    env.close()


def test_ctf_flattened_obs_vs_oracle_philox(cuda_device):
    """Philox mode with autoreset, a ragged env count and an output pointer that is only 8-byte aligned (plain-store path)."""
    import gym_multigrid_b200 as mg
    fm = load_golden("ctf_2v2")["field_map"]
    n = 1003
    env = mg.make_ctf_vec(n, fm, max_steps=25, seed=31)
    o = oc.CtfOracle(fm, n, 2, 2, max_steps=25)
    env.reset(); o.reset(oc.map_rng(mode=1, seed=31))
    gen = torch.Generator(device=cuda_device).manual_seed(3)
    buf = torch.zeros(n * 216 + 1, dtype=torch.int64, device=cuda_device)
    for t in range(40):
        act = torch.randint(0, 5, (n, 2), generator=gen, device=cuda_device, dtype=torch.int8)
        env.step(act); o.step(_np(act), oc.map_rng(mode=1, seed=31), autoreset=True)
        want = o.flattened()
        assert np.array_equal(_np(env.flattened_obs()), want), f"step {t}"
        if t % 10 == 0:
            assert np.array_equal(_np(env.flattened_obs(out=buf[1:].view(n, 216))), want)
    env.close()


def test_ctf_observation_option_in_step_and_1v1_layout(cuda_device):
    """`observation_option="flattened" / "positional"` as constructor option: reset / step return that observation (the map
    observation is not written); Ctf1v1Env's layout ends with is_red_agent_defeated alone.  Reference episodes replayed."""
    import gym_multigrid_b200 as mg
    g = load_golden("ctf1v1_flat")
    E, T, _ = g["actions"].shape
    env = mg.make_ctf1v1_vec(E, g["field_map"], autoreset=False, observation_option="flattened", reference_dtypes=True)
    assert env.single_observation_space.shape == (g["obs"].shape[-1],)
    env.set_trace(blue_place=g["blue_place"], red_place=g["red_place"])
    obs, _ = env.reset()
    assert obs.dtype == torch.int64 and np.array_equal(_np(obs), g["init_obs"])
    for t in range(T):
        live = g["length"] > t
        env.set_trace(red_actions=g["red_actions"][:, t], blue_win=g["blue_win"][:, t])
        obs, rew, term, trunc, _ = env.step(torch.as_tensor(np.where(live[:, None], g["actions"][:, t], 0).astype(np.int8), device=cuda_device))
        assert np.array_equal(_np(obs)[live], g["obs"][live, t]), f"step {t}"
    assert list(env.positional_obs())[-1] == "is_red_agent_defeated" and env.status() == 0
    env.close()
    g = load_golden("ctf_2v2_flat")
    n = 500
    a = mg.make_ctf_vec(n, g["field_map"], seed=8, observation_option="positional")
    b = mg.make_ctf_vec(n, g["field_map"], seed=8)
    da, _ = a.reset(); b.reset()
    gen = torch.Generator(device=cuda_device).manual_seed(1)
    for t in range(30):
        act = torch.randint(0, 5, (n, 2), generator=gen, device=cuda_device, dtype=torch.int8)
        da, ra, ta, ua, _ = a.step(act)
        ob, rb, tb, ub, _ = b.step(act)
        assert torch.equal(ra, rb) and torch.equal(ta, tb) and torch.equal(ua, ub)
        f8 = torch.cat(list(da.values()), 1)                      # default dtypes: the compact uint8 form
        assert isinstance(da, dict) and f8.dtype == torch.uint8 and torch.equal(f8.to(torch.int64), b.flattened_obs())
        assert torch.equal(b.flattened_obs(dtype=torch.uint8), f8)
    with pytest.raises(ValueError):
        mg.make_ctf_vec(4, g["field_map"], observation_option="pixels")
    a.close(); b.close()


def test_maze_positional_obs(cuda_device):
    """maze.py:224-231: agent position + the static cell lists in np.where order."""
    import gym_multigrid_b200 as mg
    g = load_golden("maze_board13")
    env = mg.make_maze_vec(33, g["field_map"], seed=2)
    env.reset()
    env.step(torch.ones(33, dtype=torch.int8, device=cuda_device))
    d = env.positional_obs()
    assert list(d) == ["agent", "background", "flag", "obstacle"] and all(v.dtype == torch.int64 for v in d.values())
    assert np.array_equal(_np(d["agent"]), _np(env.agent_pos)[:, 0])
    for key, code in (("background", 0), ("flag", 2), ("obstacle", 3)):
        want = np.array(list(zip(*np.where(g["field_map"] == code))), np.int64).reshape(-1)
        assert np.array_equal(_np(d[key]), np.broadcast_to(want, (33, want.size)))
    env.close()


def test_map_envs_step_async_wait(cuda_device):
    """step_async / step_wait on the map families: two CtF batches and two Maze batches in flight == the blocking device path."""
    import gym_multigrid_b200 as mg
    g = load_golden("ctf_2v2")
    mz = load_golden("maze_board13")
    n = 900
    pairs = [([mg.make_ctf_vec(n, g["field_map"], seed=s) for s in (1, 2)], [mg.make_ctf_vec(n, g["field_map"], seed=s) for s in (1, 2)], (n, 2)),
             ([mg.make_maze_vec(n, mz["field_map"], seed=s) for s in (3, 4)], [mg.make_maze_vec(n, mz["field_map"], seed=s) for s in (3, 4)], (n,))]
    rng = np.random.default_rng(5)
    for ref, pip, shape in pairs:
        for e in ref + pip:
            e.reset()
        acts = [[rng.integers(0, 5, size=shape).astype(np.int8) for _ in range(2)] for _ in range(25)]
        for b in range(2):
            pip[b].step_async(acts[0][b])
        for t in range(25):
            for b in range(2):
                want = ref[b].step(torch.as_tensor(acts[t][b], device=cuda_device))
                got = pip[b].step_wait()
                for x, y in zip(want[:4], got[:4]):
                    assert np.array_equal(_np(x), y), f"step {t} batch {b}"
                if t + 1 < 25:
                    pip[b].step_async(acts[t + 1][b])
        for e in ref + pip:
            assert e.status() == 0
            e.close()
