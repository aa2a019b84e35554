"""CPU suite: the C oracle for Maze and CtF (oracle/mg_oracle_map.c) against golden traces recorded
from the unmodified reference MazeSingleAgentEnv / CtFMvNEnv (oracle/gen_golden.py)."""
import numpy as np
import pytest

import oracle as oc
from replay import load_golden

MAZE = ["maze_board13", "maze_board13_penalty", "maze_gen64", "maze_gen64_penalty"]
CTF = ["ctf_2v2", "ctf_3v4", "ctf_2v2_penalty", "ctf_1v1", "ctf_3v4_penalty_battles"]   # the last: collision penalty AND battles in the same episodes


@pytest.mark.parametrize("stem", MAZE)
def test_maze_matches_reference(stem):
    g = load_golden(stem)
    E, T = g["actions"].shape
    o = oc.MazeOracle(g["field_map"], E, obstacle_penalty_ratio=float(g["meta_obstacle_penalty_ratio"]))
    obs = o.reset(oc.map_rng(mode=0, start_index=g["start_index"]))
    assert np.array_equal(obs, g["init_obs"]) and np.array_equal(o.info(), g["init_info"])
    checked = 0
    for t in range(T):
        live = g["length"] > t
        obs, rew, term, trunc = o.step(np.where(live, g["actions"][:, t], 0), oc.map_rng(mode=0))
        assert np.array_equal(o.info()[live], g["info"][live, t]), f"step {t}: _get_info (float64, bit-exact)"
        assert np.array_equal(obs[live], g["obs"][live, t]), f"step {t}: obs"
        assert np.array_equal(rew[live], g["reward"][live, t]), f"step {t}: reward (float64, bit-exact)"
        assert np.array_equal(term[live], g["terminated"][live, t]) and np.array_equal(trunc[live], g["truncated"][live, t])
        assert np.array_equal(o.pos[live, 0], g["pos"][live, t]) and np.array_equal(o.dir[live, 0], g["dir"][live, t])
        checked += int(live.sum())
    assert checked == int(g["length"].sum()) and o.status.value == 0


@pytest.mark.parametrize("stem", CTF)
def test_ctf_matches_reference(stem):
    g = load_golden(stem)
    E, T, nb = g["actions"].shape
    nr = int(g["meta_num_red"])
    o = oc.CtfOracle(g["field_map"], E, nb, nr, obstacle_penalty_ratio=float(g["meta_obstacle_penalty_ratio"]))
    obs = o.reset(oc.map_rng(mode=0, blue_place=g["blue_place"], red_place=g["red_place"]))
    assert np.array_equal(obs, g["init_obs"]) and np.array_equal(o.pos, g["init_pos"]) and np.array_equal(o.dir, g["init_dir"])
    assert np.array_equal(o.info(), g["init_info"])
    ident = np.arange(nb + nr, dtype=np.uint8)[None]
    for t in range(T):
        live = g["length"] > t
        used = np.zeros(E, np.int32)
        r = oc.map_rng(mode=0, red_actions=g["red_actions"][:, t], order=np.where(live[:, None], g["order"][:, t], ident),
                       blue_win=g["blue_win"][:, t], battles_used=used)
        obs, rew, term, trunc = o.step(np.where(live[:, None], g["actions"][:, t], 0), r)
        assert np.array_equal(obs[live], g["obs"][live, t]), f"step {t}: obs"
        assert np.array_equal(rew[live], g["reward"][live, t]), f"step {t}: reward (float64, bit-exact)"
        assert np.array_equal(term[live], g["terminated"][live, t]) and np.array_equal(trunc[live], g["truncated"][live, t])
        assert np.array_equal(o.pos[live], g["pos"][live, t]) and np.array_equal(o.dir[live], g["dir"][live, t])
        assert np.array_equal((o.flags & 1)[live], g["dead"][live, t]) and np.array_equal(used[live], g["n_battles"][live, t])
        assert np.array_equal(o.info()[live], g["info"][live, t]), f"step {t}: _get_info (float64, bit-exact)"
        gf, gd = o.game_stats()
        assert np.array_equal(gf[live], g["stats_flags"][live, t]) and np.array_equal(gd[live], g["stats_defeated"][live, t]), f"step {t}: game_stats"
    assert o.status.value == 0


@pytest.mark.parametrize("stem", ["ctf_2v2_carry", "ctf_3v4_penalty_carry"])
def test_ctf_flags_survive_reset_like_the_reference(stem):
    """Several episodes of ONE reference env instance: Agent.terminated / collided are never cleared by reset(), so an
    agent defeated in episode k starts episode k + 1 defeated.  `carry_agent_flags` reproduces that; every session steps at its own pace."""
    from replay import ctf_session_schedule
    g = load_golden(stem)
    nb, nr = int(g["meta_num_blue"]), int(g["meta_num_red"])
    S = int(g["meta_sessions"])
    assert g["init_dead"].any(1).sum() > S, "the fixture must hold episodes that START with a defeated agent"
    o = oc.CtfOracle(g["field_map"], S, nb, nr, obstacle_penalty_ratio=float(g["meta_obstacle_penalty_ratio"]),
                     max_steps=int(g["meta_max_steps"]), carry_agent_flags=True)
    ident = np.arange(nb + nr, dtype=np.uint8)[None]
    steps = 0
    for ev in ctf_session_schedule(g):
        if ev[0] == "reset":
            _, mask, ep = ev
            obs = o.reset(oc.map_rng(mode=0, blue_place=g["blue_place"][ep], red_place=g["red_place"][ep]), mask=mask)
            assert np.array_equal(obs[mask], g["init_obs"][ep][mask]) and np.array_equal(o.pos[mask], g["init_pos"][ep][mask])
            assert np.array_equal((o.flags & 1)[mask], g["init_dead"][ep][mask]), "terminated flags the episode starts with"
            assert np.array_equal(((o.flags >> 1) & 1)[mask], g["init_collided"][ep][mask])
            assert np.array_equal(o.info()[mask], g["init_info"][ep][mask])
            continue
        _, live, ep, t = ev
        used = np.zeros(S, np.int32)
        r = oc.map_rng(mode=0, red_actions=np.where(live[:, None], g["red_actions"][ep, t], 0).astype(np.int8),
                       order=np.where(live[:, None], g["order"][ep, t], ident).astype(np.uint8), blue_win=g["blue_win"][ep, t], battles_used=used)
        obs, rew, term, trunc = o.step(np.where(live[:, None], g["actions"][ep, t], 0), r)
        assert np.array_equal(obs[live], g["obs"][ep, t][live]), "obs"
        assert np.array_equal(rew[live], g["reward"][ep, t][live]), "reward (float64, bit-exact)"
        assert np.array_equal(term[live], g["terminated"][ep, t][live]) and np.array_equal(trunc[live], g["truncated"][ep, t][live])
        assert np.array_equal(o.pos[live], g["pos"][ep, t][live]) and np.array_equal(o.dir[live], g["dir"][ep, t][live])
        assert np.array_equal((o.flags & 1)[live], g["dead"][ep, t][live]) and np.array_equal(((o.flags >> 1) & 1)[live], g["collided"][ep, t][live])
        assert np.array_equal(used[live], g["n_battles"][ep, t][live])
        assert np.array_equal(o.info()[live], g["info"][ep, t][live])
        gf, gd = o.game_stats()
        assert np.array_equal(gf[live], g["stats_flags"][ep, t][live]) and np.array_equal(gd[live], g["stats_defeated"][ep, t][live])
        steps += int(live.sum())
    assert steps == int(g["length"].sum()) and o.status.value == 0
    # and without the flag the second episode of some session already differs: the fixture is not satisfiable by fresh instances
    assert g["init_dead"].reshape(S, -1)[:, 1:].any()


def test_map_philox_mode_shard_invariant():
    g = load_golden("ctf_3v4")
    acts = np.random.default_rng(0).integers(0, 5, size=(40, 48, 3)).astype(np.int8)

    def run(base, n):
        o = oc.CtfOracle(g["field_map"], n, 3, 4)
        r = oc.map_rng(mode=1, seed=5, env_id_base=base)
        o.reset(r)
        out = []
        for t in range(40):
            out.append(o.step(acts[t, base:base + n], r, autoreset=True)[:2])
        return out

    full, lo, hi = run(0, 48), run(0, 20), run(20, 28)
    for t in range(40):
        assert np.array_equal(full[t][0], np.concatenate([lo[t][0], hi[t][0]]))
        assert np.array_equal(full[t][1], np.concatenate([lo[t][1], hi[t][1]]))


def test_ctf1v1_matches_reference():
    g = load_golden("ctf1v1")
    E, T, _ = g["actions"].shape
    o = oc.CtfOracle(g["field_map"], E, 1, 1, variant_1v1=True)
    obs = o.reset(oc.map_rng(mode=0, blue_place=g["blue_place"], red_place=g["red_place"]))
    assert np.array_equal(obs, g["init_obs"]) and np.array_equal(o.pos, g["init_pos"])
    for t in range(T):
        live = g["length"] > t
        used = np.zeros(E, np.int32)
        r = oc.map_rng(mode=0, red_actions=g["red_actions"][:, t], blue_win=g["blue_win"][:, t], battles_used=used)
        obs, rew, term, trunc = o.step(np.where(live[:, None], g["actions"][:, t], 0), r)
        assert np.array_equal(obs[live], g["obs"][live, t]) and np.array_equal(rew[live], g["reward"][live, t])
        assert np.array_equal(term[live], g["terminated"][live, t]) and np.array_equal(trunc[live], g["truncated"][live, t])
        assert np.array_equal(o.pos[live], g["pos"][live, t]) and np.array_equal((o.flags & 1)[live], g["dead"][live, t])
        assert np.array_equal(used[live], g["n_battles"][live, t])
        assert np.array_equal(o.info()[live], g["info"][live, t])
        gf, gd = o.game_stats()
        assert np.array_equal(gf[live], g["stats_flags"][live, t]) and np.array_equal(gd[live], g["stats_defeated"][live, t])


@pytest.mark.parametrize("stem", ["ctf_2v2_flat", "ctf_3v4_flat"])
def test_ctf_flattened_obs_matches_reference(stem):
    """observation_option="flattened" (ctf.py:1084-1104), the option the reference's RL script uses: every step of every episode."""
    g = load_golden(stem)
    E, T, nb = g["actions"].shape
    nr = int(g["meta_num_red"])
    o = oc.CtfOracle(g["field_map"], E, nb, nr)
    o.reset(oc.map_rng(mode=0, blue_place=g["blue_place"], red_place=g["red_place"]))
    f = o.flattened()
    assert f.dtype == np.int64 and np.array_equal(f, g["init_obs"])
    ident = np.arange(nb + nr, dtype=np.uint8)[None]
    for t in range(T):
        live = g["length"] > t
        o.step(np.where(live[:, None], g["actions"][:, t], 0),
               oc.map_rng(mode=0, red_actions=g["red_actions"][:, t], order=np.where(live[:, None], g["order"][:, t], ident), blue_win=g["blue_win"][:, t]))
        assert np.array_equal(o.flattened()[live], g["obs"][live, t]), f"step {t}"
    assert o.status.value == 0


def test_ctf1v1_flattened_obs_matches_reference():
    """Ctf1v1Env's flattened vector (ctf.py:359-371): the tail is the single is_red_agent_defeated flag."""
    g = load_golden("ctf1v1_flat")
    E, T, _ = g["actions"].shape
    o = oc.CtfOracle(g["field_map"], E, 1, 1, variant_1v1=True)
    o.reset(oc.map_rng(mode=0, blue_place=g["blue_place"], red_place=g["red_place"]))
    assert np.array_equal(o.flattened(), g["init_obs"])
    for t in range(T):
        live = g["length"] > t
        o.step(np.where(live[:, None], g["actions"][:, t], 0), oc.map_rng(mode=0, red_actions=g["red_actions"][:, t], blue_win=g["blue_win"][:, t]))
        assert np.array_equal(o.flattened()[live], g["obs"][live, t]), f"step {t}"
    assert o.status.value == 0
