"""CPU suite: the C oracle of the base-class MultiGridEnv.step with DefaultWorld (oracle/mg_oracle_generic.c) against
golden traces recorded from reference classes only (oracle/ref_harness.py::make_generic_env, oracle/gen_golden.py)."""
import numpy as np
import pytest

import oracle as oc
from replay import load_golden

GENERIC = ["generic_9x9_a3", "generic_12x12_a5", "generic_7x7_a1"]


@pytest.mark.parametrize("stem", GENERIC)
def test_generic_step_matches_reference(stem):
    g = load_golden(stem)
    E, T, A = g["actions"].shape
    S = int(g["meta_size"])
    o = oc.GenericOracle(E, S, S, A, int(g["meta_max_steps"]))
    o.set_state_from_obs(g["init_obs"][:, 0], g["init_pos"])
    assert np.array_equal(o.encode(), g["init_obs"]), "reset observation (encode_for_agents, encode_dim 6)"
    ident = np.arange(A, dtype=np.uint8)[None]
    checked = 0
    for t in range(T):
        live = g["length"] > t
        act = np.where(live[:, None], g["actions"][:, t], 0)
        obs, rew, term, trunc = o.step(act, np.where(live[:, None], g["order"][:, t], ident))
        assert np.array_equal(obs[live], g["obs"][live, t]), f"step {t}: obs"
        assert np.array_equal(rew[live], g["rewards"][live, t]), f"step {t}: rewards (float64, bit-exact)"
        assert np.array_equal(term[live], g["terminated"][live, t]) and np.array_equal(trunc[live], g["truncated"][live, t])
        assert np.array_equal(o.pos[live], g["pos"][live, t])
        checked += int(live.sum())
    assert checked == int(g["length"].sum()) and o.status.value == 0
    assert g["terminated"].any() and (g["rewards"] > 0).any()     # the goal branch and `_reward` are exercised


def test_generic_bad_action_sets_status():
    o = oc.GenericOracle(1, 5, 5, 1, 10)
    o.gcell[0, 2 * 5 + 2] = 10
    o.pos[0, 0] = (2, 2)
    o.step(np.array([[5]], np.int8), np.array([[0]], np.uint8))      # toggle: the reference raises (multigrid.py:447)
    assert o.status.value & 8      # OC_ERR_BAD_ACTION


@pytest.mark.parametrize("stem", ["partial6_9x9_a3", "partial6_12x12_a5"])
def test_generic_partial_views_match_reference(stem):
    """encode_dim-6 partial observations (gen_obs_grid + encode_for_agents) on recorded DefaultWorld states."""
    g = load_golden(stem)
    S, A = int(g["meta_size"]), int(g["meta_num_agents"])
    checked = 0
    for V in (3, 5, 7):
        for st in (False, True):
            sel = np.where((g["V"] == V) & (g["see_through"] == st))[0]
            if len(sel) == 0:
                continue
            o = oc.GenericOracle(len(sel), S, S, A, 100)
            o.set_state_from_obs(g["obs6"][sel], g["pos"][sel])
            assert np.array_equal(o.partial_views(V, st), g["views"][sel][:, :, :V, :V]), f"V={V} see_through={st}"
            checked += len(sel)
    assert checked == len(g["V"])
