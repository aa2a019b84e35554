"""CPU suite: the C oracle (oracle/mg_oracle.c) against golden traces recorded from the
unmodified reference (oracle/gen_golden.py).  This is what pins the oracle."""
import numpy as np
import pytest

import oracle as oc
from replay import COLLECT_FIXTURES, collect_kwargs, expected, load_golden, step_inputs


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10
    assert oc.philox4x32_10([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert oc.philox4x32_10([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert oc.philox4x32_10([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == [
        0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_encode3_table():
    cells = np.arange(256, dtype=np.uint8)
    obs = oc.encode3(cells)
    assert np.array_equal(obs[:, 0], cells & 3)
    assert np.array_equal(obs[:, 1], (cells >> 2) & 15)
    # STATE is the agent's dir; a ball (type 2) always encodes STATE 0 (object.py:58-74) - its bit 6 is the internal
    # "placed by _respawn" mark (include/multigrid_b200.h, MG_PLANE_GRID)
    assert np.array_equal(obs[:, 2], np.where((cells & 3) == 2, 0, cells >> 6))
    # the codes the Collect world produces
    assert obs[0].tolist() == [0, 0, 0] and obs[1 | 7 << 2].tolist() == [1, 7, 0]
    assert obs[3 | 5 << 2 | 3 << 6].tolist() == [3, 5, 3] and obs[2 | 2 << 2].tolist() == [2, 2, 0]


@pytest.mark.parametrize("stem", sorted(COLLECT_FIXTURES))
def test_collect_reset_matches_reference(stem):
    g = load_golden(stem)
    E = len(g["length"])
    o = oc.CollectOracle(oc.make_collect_cfg(**collect_kwargs(g, stem)), E)
    tr = oc.TraceRng(draws=g["reset_draws"], n_draws=g["n_reset_draws"])
    obs = o.reset(tr)
    assert o.status.value == 0
    assert np.array_equal(tr.draws_used, g["n_reset_draws"])
    assert np.array_equal(obs, g["init_obs"])
    assert np.array_equal(o.agent_pos, g["init_pos"])


@pytest.mark.parametrize("stem", sorted(COLLECT_FIXTURES))
def test_collect_step_matches_reference(stem):
    g = load_golden(stem)
    E, T, A = g["actions"].shape
    o = oc.CollectOracle(oc.make_collect_cfg(**collect_kwargs(g, stem)), E)
    o.set_state_from_obs(g["init_obs"], g["init_pos"])
    checked = 0
    for t in range(T):
        act, order, draws, n_draws, live = step_inputs(g, t)
        tr = oc.TraceRng(order=order, draws=draws, n_draws=n_draws)
        obs, rew, term, trunc = o.step(act, tr)
        x = expected(g, t)
        assert np.array_equal(tr.draws_used[live], n_draws[live]), f"step {t}: draws consumed"
        assert np.array_equal(obs[live], x["obs"][live]), f"step {t}: obs"
        assert np.array_equal(rew[live], x["rewards"][live]), f"step {t}: rewards"
        assert np.array_equal(term[live], x["terminated"][live]), f"step {t}: terminated"
        assert np.array_equal(trunc[live], x["truncated"][live]), f"step {t}: truncated"
        assert np.array_equal(o.agent_pos[live], x["pos"][live]), f"step {t}: agent positions"
        assert np.array_equal(o.collected[live], x["collected"][live]), f"step {t}: collected"
        ni = o.info.shape[1]
        assert np.array_equal(o.info[live], x["info"][live][:, :ni]), f"step {t}: info counters"
        checked += int(live.sum())
    assert o.status.value == 0
    assert checked == int(g["length"].sum())


def test_ball_loss_quirk_is_in_the_fixture():
    """The respawn-onto-the-vacated-cell quirk must be exercised by the golden data:
    a pickup step whose recorded respawn draw equals the cell the agent then enters."""
    g = load_golden("collect_respawn_clustered")
    hits = 0
    E, T, A = g["actions"].shape
    for e in range(E):
        for t in range(int(g["length"][e])):
            n = int(g["n_draws"][e, t])
            if n == 0:
                continue
            balls_before = (g["obs"][e, t - 1, :, :, 0] == 2).sum() if t else (g["init_obs"][e, :, :, 0] == 2).sum()
            balls_after = (g["obs"][e, t, :, :, 0] == 2).sum()
            hits += int(balls_after < balls_before)
    assert hits > 0


def test_philox_mode_is_deterministic_and_shard_invariant():
    g = load_golden("collect_respawn_clustered")
    kw = collect_kwargs(g, "collect_respawn_clustered")
    rng = np.random.default_rng(0)
    acts = rng.integers(0, 4, size=(20, 64, 2)).astype(np.int8)

    def run(base, n):
        o = oc.CollectOracle(oc.make_collect_cfg(**kw), n)
        r = oc.PhiloxRng(seed=123, env_id_base=base)
        o.reset(r)
        outs = []
        for t in range(20):
            obs, rew, term, trunc = o.step(acts[t, base:base + n], r, autoreset=True)
            outs.append((obs.copy(), rew.copy()))
        return outs

    full = run(0, 64)
    lo, hi = run(0, 32), run(32, 32)
    for t in range(20):
        assert np.array_equal(full[t][0], np.concatenate([lo[t][0], hi[t][0]]))
        assert np.array_equal(full[t][1], np.concatenate([lo[t][1], hi[t][1]]))


def test_oracle_openmp_threads_agree():
    g = load_golden("collect_respawn")
    kw = collect_kwargs(g, "collect_respawn")
    rng = np.random.default_rng(1)
    acts = rng.integers(0, 4, size=(30, 512, 2)).astype(np.int8)
    res = []
    for nt in (1, 4):
        o = oc.CollectOracle(oc.make_collect_cfg(**kw), 512, nthreads=nt)
        r = oc.PhiloxRng(seed=7)
        o.reset(r)
        for t in range(30):
            obs, rew, term, trunc = o.step(acts[t], r, autoreset=True)
        res.append((obs, rew, o.grid.copy(), o.rng_ctr.copy()))
    for a, b in zip(res[0], res[1]):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("stem", ["partial_rooms", "partial_clustered", "partial_quadrants15"])
def test_partial_view_matches_reference(stem):
    """MultiGridEnv.gen_obs_grid + encode_for_agents of the reference vs the oracle's direct-index form,
    all four directions, V in {3,5,7}, with and without see_through_walls (incl. two agents on one cell)."""
    g = load_golden(stem)
    W = g["grid_obs"].shape[1]
    seen = set()
    for i in range(len(g["V"])):
        V, st = int(g["V"][i]), bool(g["see_through"][i])
        out = oc.partial_view3(oc.pack_obs(g["grid_obs"][i]).reshape(1, -1), g["pos"][i][None], W, W, V, st, dirs=g["dirs"][i][None])
        assert np.array_equal(out[0], g["views"][i][:, :V, :V]), f"state {i}"
        seen |= {(V, st, int(d)) for d in g["dirs"][i]}
    assert len(seen) >= 20


@pytest.mark.parametrize("stem", ["toroid_clustered", "toroid_rooms", "toroid_single"])
def test_toroid_matches_reference(stem):
    g = load_golden(stem)
    W = g["grid_obs"].shape[1]
    out = oc.toroid(oc.pack_obs(g["grid_obs"]).reshape(len(g["pos"]), -1), g["pos"], W, 3)
    assert out.dtype == np.float32 and np.array_equal(out, g["toroid"])
