"""Scripted CtF opponents (`gym_multigrid_b200.policy.ctf`) against the unmodified reference classes
(policy/ctf/heuristic.py, policy/ctf/utils.py): the recorded decisions of `tests/golden/ctf_policies.npz`
(`oracle/gen_golden.py policies`) and, where /root/reference exists, a live differential run on fresh random inputs.
Host-side code: no GPU involved (the policies are callers of the CUDA step)."""
import numpy as np
import pytest

from replay import POLICY_NAMES, load_golden, policy_maps, policy_observation

from gym_multigrid_b200.policy.ctf import heuristic as H
from gym_multigrid_b200.policy.ctf.utils import a_star, closest_area_pos, manhattan_distance, position_in_positions


@pytest.fixture(scope="module")
def golden():
    return load_golden("ctf_policies")


@pytest.fixture(scope="module")
def maps():
    return policy_maps()


@pytest.mark.parametrize("mname", ["board", "wide", "walls"])
def test_a_star_routes_are_the_reference_routes(golden, maps, mname):
    """Cell for cell - tie-breaking included; `walls` holds the blocking value 8 and pairs without any route."""
    g, fm = golden, maps[mname]
    cells, k = g[f"astar_{mname}_cells"], 0
    none = 0
    for s, e, n in zip(g[f"astar_{mname}_start"], g[f"astar_{mname}_end"], g[f"astar_{mname}_len"]):
        path = a_star(tuple(s), tuple(e), fm)
        assert np.array_equal(np.array(path, np.int64).reshape(-1, 2), cells[k:k + n]), (mname, s, e)
        if n:
            assert path[0] == tuple(s) and path[-1] == tuple(e) and len(path) >= manhattan_distance(s, e) + 1
        none += n == 0
        k += n
    assert k == len(cells) and (none > 0) == (mname == "walls")


def test_a_star_ignores_ctf_obstacles(maps):
    """policy/ctf/utils.py:73 blocks on the value 8 only; CtF obstacles are 6 (core/world.py:66-79) and are walked through."""
    fm = maps["board"]
    assert (fm == 6).sum() == 4 and fm[4, 4] == 6 and fm[5, 4] == 6
    path = a_star((3, 4), (6, 4), fm)
    assert path == [(3, 4), (4, 4), (5, 4), (6, 4)]
    assert a_star((2, 2), (2, 2), fm) == [(2, 2)]
    assert a_star((0, 0), (99, 0), fm) == []                     # a target outside the map: no route


def _replay(golden, maps, mname, pname, ego, **extra):
    stem = f"{mname}_{pname}_{ego}"
    gen = np.random.Generator(np.random.PCG64(int(golden[f"{stem}_seed"])))
    pol = getattr(H, pname)(field_map=maps[mname], random_generator=gen, ego_agent=ego, randomness=float(golden[f"{stem}_randomness"]), **extra)
    got = [int(pol.act(policy_observation(maps[mname], b, r), tuple(c)))
           for c, b, r in zip(golden[f"{stem}_curr"], golden[f"{stem}_blue"], golden[f"{stem}_red"])]
    return pol, gen, got


@pytest.mark.parametrize("ego", ["red", "blue"])
@pytest.mark.parametrize("pname", POLICY_NAMES)
@pytest.mark.parametrize("mname", ["board", "wide"])
def test_policy_decisions_and_rng_stream(golden, maps, mname, pname, ego):
    """320 recorded decisions per policy x team x map from ONE seeded generator: equal actions need equal targets, equal routes
    and the same draws in the same order; `tail` checks where the stream stands afterwards."""
    stem = f"{mname}_{pname}_{ego}"
    pol, gen, got = _replay(golden, maps, mname, pname, ego)
    want = golden[f"{stem}_action"].tolist()
    assert got == want, next(i for i, (a, b) in enumerate(zip(got, want)) if a != b)
    assert int(gen.integers(0, 2 ** 31)) == int(golden[f"{stem}_tail"])
    if f"{stem}_border" in golden:
        assert np.array_equal(np.array(pol.border, np.int64).reshape(-1, 2), golden[f"{stem}_border"])


def test_interface_mirrors_the_reference():
    """Names, defaults and attributes user code touches (heuristic.py:51-67, 84-106, 189-214, 284-319; ctf.py:785-826)."""
    from gym_multigrid_b200.actions import CtfActions
    from gym_multigrid_b200.world import CtfWorld
    assert [H.RwPolicy().name, H.FightPolicy().name, H.CapturePolicy().name, H.PatrolPolicy().name, H.PatrolFightPolicy().name] == \
        ["rw", "fight", "capture", "patrol", "patrol_fight"]
    p = H.FightPolicy()
    assert p.field_map is None and p.action_set is CtfActions and p.randomness == 0.75 and p.ego_agent == "red"
    assert isinstance(p.random_generator, np.random.Generator)
    assert issubclass(H.PatrolFightPolicy, H.PatrolPolicy) and issubclass(H.PatrolPolicy, H.DestinationPolicy)
    assert issubclass(H.DestinationPolicy, H.CtfPolicy) and issubclass(H.RwPolicy, H.CtfPolicy)
    assert CtfWorld.OBJECT_TO_IDX["obstacle"] == 6 and [a.name for a in CtfActions] == ["stay", "left", "down", "right", "up"]
    assert 0 <= int(H.RwPolicy(random_generator=np.random.default_rng(0)).act()) < 5
    with pytest.raises(NotImplementedError):
        H.CtfPolicy().act({}, (0, 0))
    # a patrol policy built without a map has an empty border for good and fails on its first decision (heuristic.py:319, utils/map.py:61)
    q = H.PatrolPolicy()
    assert q.border == [] and q.obstacle == []
    with pytest.raises(ValueError):
        q.get_target({}, (1, 1))
    assert closest_area_pos((0, 0), [(3, 0), (0, 3), (1, 1), (1, 1)]) == (1, 1)
    assert closest_area_pos((0, 0), [(0, 3), (3, 0)]) == (0, 3)                   # first of equals
    assert position_in_positions((1, 2), [(0, 0), (1, 2)]) and not position_in_positions((2, 1), [(0, 0), (1, 2)])


def test_route_cache_follows_the_map(maps):
    """The env installs `field_map` after construction (ctf.py:796-799); memoised routes belong to one map object."""
    fm = maps["walls"].copy()
    pol = H.CapturePolicy(field_map=fm, random_generator=np.random.default_rng(0), randomness=1.0)
    obs = {"blue_flag": np.array([2, 3])}
    first = pol._next_cell((3, 0), (2, 3))
    assert first == (3, 1) == tuple(a_star((3, 0), (2, 3), fm)[1]) and pol._next_cell((3, 0), (2, 3)) is first   # around the 8s
    assert int(pol.act(obs, (3, 0))) == 3                                        # right = (0, +1)
    open_map = np.zeros_like(fm)
    pol.field_map = open_map
    assert pol._next_cell((3, 0), (2, 3)) == tuple(a_star((3, 0), (2, 3), open_map)[1]) == (2, 0)
    assert int(pol.act(obs, (3, 0))) == 2                                        # down = (-1, 0)


def _reference_heuristic():
    import ref_harness as rh
    if not rh.reference_available():
        pytest.skip("needs /root/reference (build container only)")
    rh.import_reference()
    from gym_multigrid.policy.ctf import heuristic, utils
    return heuristic, utils


def test_live_differential_against_the_reference(maps):
    """Fresh random maps, team sizes and positions through both implementations with equally seeded generators."""
    RH, RU = _reference_heuristic()
    rng = np.random.default_rng(5)
    for trial in range(6):
        rows, cols = int(rng.integers(6, 15)), int(rng.integers(6, 15))
        fm = np.where(np.arange(rows)[:, None] + rng.integers(-1, 2, size=(rows, cols)) < rows // 2, 1.0, 0.0)
        fm[rng.random(fm.shape) < 0.1] = 6.0
        fm[0, 0], fm[rows - 1, cols - 1] = 5.0, 4.0
        blocked = fm.copy()
        blocked[rng.random(fm.shape) < 0.2] = 8.0
        for _ in range(60):
            s, e = (int(rng.integers(0, rows)), int(rng.integers(0, cols))), (int(rng.integers(0, rows)), int(rng.integers(0, cols)))
            assert [tuple(int(v) for v in p) for p in RU.a_star(s, e, blocked)] == a_star(s, e, blocked)
        for pname in POLICY_NAMES:
            for ego in ("red", "blue"):
                ga, gb = (np.random.Generator(np.random.PCG64(100 + trial)) for _ in range(2))
                kw = dict(field_map=fm, ego_agent=ego, randomness=float(rng.choice([0.75, 0.3, 1.0])))
                try:
                    ref = getattr(RH, pname)(random_generator=ga, **kw)
                except Exception as exc:      # noqa: BLE001 - whatever the reference raises, ours must raise too
                    with pytest.raises(type(exc)):
                        getattr(H, pname)(random_generator=gb, **kw)
                    continue
                ours = getattr(H, pname)(random_generator=gb, **kw)
                nb, nr = int(rng.integers(1, 5)), int(rng.integers(1, 5))
                for k in range(80):
                    blue = np.stack([rng.integers(0, rows, nb), rng.integers(0, cols, nb)], 1)
                    red = np.stack([rng.integers(0, rows, nr), rng.integers(0, cols, nr)], 1)
                    border = getattr(ref, "border", [])
                    cur = tuple(int(v) for v in border[rng.integers(0, len(border))]) if len(border) and k % 2 else \
                        (int(rng.integers(0, rows)), int(rng.integers(0, cols)))
                    obs = policy_observation(fm, blue, red)
                    try:
                        want = int(ref.act(obs, cur))
                    except Exception as exc:  # noqa: BLE001
                        with pytest.raises(type(exc)):
                            ours.act(obs, cur)
                        break
                    assert int(ours.act(obs, cur)) == want, (trial, pname, ego, k)
                assert ga.integers(0, 2 ** 31) == gb.integers(0, 2 ** 31)


def _random_ctf_map(size, seed):
    """A square CtF map: ragged frontier between the territories, scattered obstacles, one flag each."""
    rng = np.random.default_rng(seed)
    fm = np.where(np.arange(size)[:, None] + rng.integers(-2, 3, size=(size, size)) < size // 2, 1.0, 0.0)
    fm[rng.random(fm.shape) < 0.07] = 6.0
    fm[1, 1], fm[size - 2, size - 2] = 5.0, 4.0
    return fm


@pytest.mark.parametrize("size,nb,nr,seed", [(12, 2, 3, 1), (9, 4, 2, 2), (16, 1, 1, 3), (7, 3, 5, 4)])
def test_oracle_rule_on_random_maps(size, nb, nr, seed):
    """Same comparison on generated maps, team sizes from 1v1 to 3v5 and every assignment of the four policies."""
    import oracle as oc
    from gym_multigrid_b200.policy.ctf.device import build_tables
    fm = _random_ctf_map(size, seed)
    names = [("FightPolicy", "CapturePolicy", "PatrolPolicy", "PatrolFightPolicy")[(k + seed) % 4] for k in range(nr)]
    host = [getattr(H, name)(fm, randomness=1.0, random_generator=np.random.default_rng(0)) for name in names]
    tables = build_tables(host, fm)
    n = 48
    o = oc.CtfOracle(fm.astype(np.uint8), n, nb, nr, max_steps=20)
    o.reset(oc.map_rng(mode=1, seed=seed))
    rng = np.random.default_rng(seed)
    episode = np.zeros(n, np.int32)
    compared = 0
    for t in range(45):
        red = o.policy_actions(tables, seed, episode)
        blue, reds = o.pos[:, :nb].astype(np.int64), o.pos[:, nb:].astype(np.int64)
        for e in range(n):
            obs = policy_observation(fm, blue[e], reds[e])
            intruder = any(fm[tuple(b)] in (1, 5) for b in blue[e].tolist())
            for k, p in enumerate(host):
                cur = tuple(reds[e, k].tolist())
                patrols = names[k] == "PatrolPolicy" or (names[k] == "PatrolFightPolicy" and not intruder)
                if patrols and cur in p._border_cells:
                    continue
                assert int(p.act(obs, cur)) == int(red[e, k]), (t, e, k, names[k], cur)
                compared += 1
        _, _, term, trunc = o.step(rng.integers(0, 5, (n, nb)).astype(np.int8), oc.map_rng(mode=1, seed=seed, red_actions=red), autoreset=True)
        episode += (term | trunc)
    assert compared > 1000


def test_oracle_rule_of_the_device_policies_matches_the_host_policies():
    """The C restatement of csrc/policy_kernels.cu's decision rule (`oc_ctf_policy_actions`, the checker of
    tests/test_policy_device_gpu.py) against the host policies above, on CtF states stepped by the oracle: with randomness 1
    every decision that involves no patrol draw must be the host policy's; with randomness 0 every action is a uniform draw."""
    import oracle as oc
    from gym_multigrid_b200.policy.ctf.device import build_tables
    g = load_golden("ctf_3v4")
    fm = g["field_map"].astype(np.float64)
    nb, nr, n = 3, 4, 64
    names = ("FightPolicy", "CapturePolicy", "PatrolPolicy", "PatrolFightPolicy")
    host = [getattr(H, name)(fm, randomness=1.0, random_generator=np.random.default_rng(0)) for name in names]
    tables = build_tables(host, fm)
    o = oc.CtfOracle(g["field_map"], n, nb, nr, max_steps=25)
    o.reset(oc.map_rng(mode=1, seed=4))
    rng = np.random.default_rng(1)
    episode = np.zeros(n, np.int32)
    compared = drawn = 0
    for t in range(60):
        red = o.policy_actions(tables, 4, episode)
        for e in range(n):
            obs = policy_observation(fm, o.pos[e, :nb].astype(np.int64), o.pos[e, nb:].astype(np.int64))
            intruder = any(fm[tuple(b)] in (1, 5) for b in o.pos[e, :nb].astype(np.int64).tolist())
            for k, p in enumerate(host):
                cur = tuple(o.pos[e, nb + k].astype(np.int64).tolist())
                if (k == 2 or (k == 3 and not intruder)) and cur in p._border_cells:
                    drawn += 1
                    assert tables["first_move"][cur[0] * 10 + cur[1]].max() <= 4
                    continue
                assert int(p.act(obs, cur)) == int(red[e, k]), (t, e, k, cur)
                compared += 1
        _, _, term, trunc = o.step(rng.integers(0, 5, (n, nb)).astype(np.int8), oc.map_rng(mode=1, seed=4, red_actions=red), autoreset=True)
        episode += (term | trunc)
    assert compared > 10000 and drawn > 50
    tables0 = build_tables([getattr(H, name)(fm, randomness=0.0) for name in names], fm)
    acts = np.concatenate([o.policy_actions(tables0, s, episode).reshape(-1) for s in range(40)])
    assert np.bincount(acts, minlength=5).min() > 0.15 * len(acts)


def test_native_first_move_table_matches_the_python_a_star(maps):
    """`mg_astar_first_moves` (csrc/astar_host.cu, host-only C++) against one `utils.a_star` run per (start, target) pair - the
    Python A* being the one pinned to the reference's routes above: non-square maps, blocking cells, pairs without a route."""
    from gym_multigrid_b200.policy.ctf.device import first_move_table, first_move_table_py
    rng = np.random.default_rng(12)
    cases = dict(maps)
    for i in range(5):
        r, c = int(rng.integers(2, 9)), int(rng.integers(2, 9))
        m = np.zeros((r, c))
        m[rng.random((r, c)) < 0.3] = 8.0
        cases[f"random{i}"] = m
    for name, fm in cases.items():
        native, py = first_move_table(fm), first_move_table_py(fm)
        assert native.shape == py.shape == (fm.size, fm.size) and np.array_equal(native, py), name
    with pytest.raises(ValueError):
        first_move_table(np.zeros((300, 300)))


def test_device_tables(maps):
    """`policy/ctf/device.build_tables` (what mg_set_red_policies uploads): the first-move table restates `DestinationPolicy.act`'s
    route step for every (cell, target) pair, the patrol tables restate `PatrolPolicy.get_target`; inputs with no device form
    are refused."""
    from gym_multigrid_b200.policy.ctf.device import build_tables, first_move_table
    fm = maps["board"]
    pols = [H.FightPolicy(fm, randomness=0.3), H.PatrolFightPolicy(fm), None, H.CapturePolicy(fm, randomness=1.0)]
    t = build_tables(pols, fm)
    assert t["kind"].tolist() == [1, 4, 0, 2] and t["randomness"].tolist() == [0.3, 0.75, 0.0, 1.0]
    fmv = t["first_move"]
    assert fmv.shape == (100, 100) and fmv.dtype == np.uint8 and np.array_equal(fmv, first_move_table(fm))
    probe = H.CapturePolicy(fm, randomness=1.0, random_generator=np.random.default_rng(0))
    rng = np.random.default_rng(3)
    for _ in range(300):
        s, g = (int(rng.integers(0, 10)), int(rng.integers(0, 10))), (int(rng.integers(0, 10)), int(rng.integers(0, 10)))
        assert int(probe.act({"blue_flag": np.array(g)}, s)) == fmv[s[0] * 10 + s[1], g[0] * 10 + g[1]]
    patrol = pols[1]
    assert sorted(set(np.flatnonzero(t["on_border"]).tolist())) == sorted({x * 10 + y for x, y in patrol._border_cells})
    assert t["along_border"].tolist() == [int(x) * 10 + int(y) for x, y in patrol._along_border]
    for c in range(100):
        x, y = closest_area_pos(divmod(c, 10), patrol.border)
        assert t["patrol_goal"][c] == x * 10 + y
    no_patrol = build_tables([H.FightPolicy(fm), H.RwPolicy()], fm)
    assert len(no_patrol["along_border"]) == 0 and not no_patrol["on_border"].any()

    class Mine(H.FightPolicy):
        pass
    with pytest.raises(TypeError):
        build_tables([Mine(fm)], fm)
    with pytest.raises(ValueError):
        build_tables([H.PatrolPolicy()], fm)                                  # no map at construction: empty border
    with pytest.raises(ValueError):
        build_tables([H.FightPolicy(fm, ego_agent="blue")], fm)
    with pytest.raises(ValueError):
        build_tables([H.FightPolicy(maps["wide"][:10, :10].copy())], fm)      # another map
    with pytest.raises(ValueError):
        build_tables([H.PatrolPolicy(fm), H.PatrolPolicy(fm, ego_agent="red", world=type("W", (), {"OBJECT_TO_IDX": dict(red_territory=0, blue_territory=1, obstacle=6)}))], fm)
    with pytest.raises(ValueError):
        build_tables([H.FightPolicy()], np.zeros((33, 33)))                   # > 1024 cells


@pytest.mark.parametrize("names", [("FightPolicy", "CapturePolicy"), ("PatrolPolicy", "PatrolFightPolicy"), ("FightPolicy", "RwPolicy"),
                                   ("PatrolFightPolicy", "PatrolFightPolicy")])
def test_policies_drop_into_the_unmodified_reference_env(names):
    """The other direction of the drop-in: this package's policy objects handed to the REFERENCE `CtFMvNEnv(enemy_policies=...)`
    (tests/test_ctf.py:97-215) reproduce, step for step, the episodes the reference's own policies produce - observations, rewards,
    flags - with the policies' draws interleaved with the env's own (shuffle, battles) as ctf.py:821-826 arranges."""
    import os
    RH, _ = _reference_heuristic()
    import ref_harness as rh
    from gym_multigrid.envs.ctf import CtFMvNEnv
    map_path = os.path.join(rh.REFERENCE_ROOT, "tests", "assets", "board.txt")
    fm = np.loadtxt(map_path).T
    for seed in (7, 8, 9):
        runs = []
        for mod in (RH, H):
            pols = [mod.RwPolicy() if n == "RwPolicy" else getattr(mod, n)(fm.copy()) for n in names]
            env = CtFMvNEnv(map_path, num_blue_agents=2, num_red_agents=2, enemy_policies=pols, observation_option="map")
            shared = np.random.Generator(np.random.PCG64(40 + seed))     # the construction-time np_random is unseeded: pin it
            env._np_random = shared
            for agent in env.agents[2:]:
                assert agent.policy.action_set is env.actions_set
                agent.policy.random_generator = shared
            obs, _ = env.reset(seed=seed)
            arng, traj, rews, flags = np.random.default_rng(seed), [np.asarray(obs)], [], []
            while True:
                obs, rew, term, trunc, info = env.step([int(v) for v in arng.integers(0, 5, 2)])
                traj.append(np.asarray(obs)); rews.append(rew); flags.append((term, trunc))
                if term or trunc:
                    break
            runs.append((np.stack(traj), rews, flags, shared.integers(0, 2 ** 31)))
        assert runs[0][0].shape == runs[1][0].shape and np.array_equal(runs[0][0], runs[1][0]), (names, seed)
        assert runs[0][1] == runs[1][1] and runs[0][2] == runs[1][2] and runs[0][3] == runs[1][3] and len(runs[0][1]) >= 10
