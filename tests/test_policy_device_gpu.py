"""Scripted CtF opponents decided on the device (`CtfVecEnv.set_enemy_policies(..., device=True)` -> mg_set_red_policies /
mg_red_policy_actions, csrc/policy_kernels.cu):
 * the kernel's actions == the C oracle's restatement of the rule (same Philox blocks), step after step with autoreset, and the
   CUDA step driven by them == the oracle's step driven by the oracle's;
 * with randomness 1 (always follow the route) the kernel's actions == the decisions of the host policies - which are pinned
   to the unmodified reference classes (tests/test_policies.py) - for every env and red agent whose decision involves no draw."""
import numpy as np
import pytest
import torch

import oracle as oc
from replay import load_golden

pytestmark = pytest.mark.gpu


def _np(t):
    return t.cpu().numpy()


def _policies(names, fm, randomness):
    from gym_multigrid_b200.policy.ctf import heuristic as H
    out = []
    for name, r in zip(names, randomness):
        out.append(H.RwPolicy() if name == "RwPolicy" else getattr(H, name)(fm, randomness=r))
    return out


@pytest.mark.parametrize("stem,names,randomness", [
    ("ctf_2v2", ("FightPolicy", "PatrolFightPolicy"), (0.75, 0.6)),
    ("ctf_2v2", ("CapturePolicy", "PatrolPolicy"), (0.9, 0.5)),
    ("ctf_2v2", ("RwPolicy", "FightPolicy"), (0.0, 0.0)),
    ("ctf_3v4", ("FightPolicy", "CapturePolicy", "PatrolPolicy", "PatrolFightPolicy"), (0.75, 1.0, 0.75, 0.25)),
])
@pytest.mark.parametrize("fused", [True, False])   # decided inside the step kernel (mg_set_red_policy_fusion) / by an explicit call before it
def test_device_policies_match_oracle(stem, names, randomness, fused, cuda_device):
    import gym_multigrid_b200 as mg
    g = load_golden(stem)
    fm = g["field_map"].astype(np.float64)
    nb, nr = int(g["meta_num_blue"]), int(g["meta_num_red"])
    n, seed = 777, 5
    env = mg.make_ctf_vec(n, g["field_map"], num_blue_agents=nb, num_red_agents=nr, max_steps=20, seed=seed)
    o = oc.CtfOracle(g["field_map"], n, nb, nr, max_steps=20)
    env.set_enemy_policies(_policies(names, fm, randomness), device=True, fused=fused)
    assert env._fused_policies == fused
    tables = env._policy_tables
    assert tables["kind"].tolist() == [dict(RwPolicy=0, FightPolicy=1, CapturePolicy=2, PatrolPolicy=3, PatrolFightPolicy=4)[k] for k in names]
    obs, _ = env.reset()
    assert np.array_equal(_np(obs), o.reset(oc.map_rng(mode=1, seed=seed)))
    gen = torch.Generator(device=cuda_device).manual_seed(8)
    seen = set()
    for t in range(50):
        want = o.policy_actions(tables, seed, _np(env.episode_count))
        act = torch.randint(0, 5, (n, nb), generator=gen, device=cuda_device, dtype=torch.int8)
        obs, rew, term, trunc, _ = env.step(act)
        got = _np(env._red_buf)
        assert np.array_equal(got, want), f"step {t}: {np.argwhere(got != want)[:5].tolist()}"
        oo, orew, oterm, otrunc = o.step(_np(act), oc.map_rng(mode=1, seed=seed, red_actions=want), autoreset=True)
        assert np.array_equal(_np(obs), oo) and np.array_equal(_np(rew), orew), f"step {t}"
        assert np.array_equal(_np(term), oterm) and np.array_equal(_np(trunc), otrunc)
        seen.update(got.reshape(-1).tolist())
    assert seen == {0, 1, 2, 3, 4} and env.status() == 0
    assert int(_np(env.episode_count).max()) >= 2          # autoresets happened: the episode word of the counter was exercised
    env.set_enemy_policies(None)                            # back to the built-in RwPolicy
    obs, *_ = env.step(torch.zeros((n, nb), dtype=torch.int8, device=cuda_device))
    oo, *_ = o.step(np.zeros((n, nb), np.int8), oc.map_rng(mode=1, seed=seed), autoreset=True)
    assert np.array_equal(_np(obs), oo)
    env.close()


def test_fused_policy_step_equals_two_launches_at_full_size(cuda_device):
    """The 64-register variant of the fused kernel (batches of >= 262 144 envs) against policy kernel + step kernel."""
    import gym_multigrid_b200 as mg
    g = load_golden("ctf_2v2")
    fm = g["field_map"].astype(np.float64)
    n = 262144 + 77
    envs = []
    for fused in (True, False):
        e = mg.make_ctf_vec(n, g["field_map"], max_steps=12, seed=3)
        e.set_enemy_policies(_policies(("PatrolFightPolicy", "FightPolicy"), fm, (0.7, 0.85)), device=True, fused=fused)
        envs.append(e)
    o0, o1 = envs[0].reset()[0], envs[1].reset()[0]
    assert torch.equal(o0, o1)
    gen = torch.Generator(device=cuda_device).manual_seed(2)
    for t in range(16):
        act = torch.randint(0, 5, (n, 2), generator=gen, device=cuda_device, dtype=torch.int8)
        a, b = envs[0].step(act), envs[1].step(act)
        assert torch.equal(envs[0]._red_buf, envs[1]._red_buf), f"step {t}"
        for x, y in zip(a[:4], b[:4]):
            assert torch.equal(x, y), f"step {t}"
    assert torch.equal(envs[0].state, envs[1].state) and envs[0].status() == 0


def test_set_red_actions_takes_the_red_team_over_from_device_policies(cuda_device):
    """`set_red_actions(buffer)` after `set_enemy_policies(device=True)`: the caller's buffer drives the red team (fused or not)."""
    import gym_multigrid_b200 as mg
    g = load_golden("ctf_2v2")
    fm = g["field_map"].astype(np.float64)
    n, seed = 513, 12
    for fused in (True, False):
        env = mg.make_ctf_vec(n, g["field_map"], max_steps=15, seed=seed)
        o = oc.CtfOracle(g["field_map"], n, 2, 2, max_steps=15)
        env.set_enemy_policies(_policies(("FightPolicy", "CapturePolicy"), fm, (1.0, 1.0)), device=True, fused=fused)
        red = env.set_red_actions(torch.zeros((n, 2), dtype=torch.int8, device=cuda_device))
        assert not env._device_policies and not env._fused_policies
        assert np.array_equal(_np(env.reset()[0]), o.reset(oc.map_rng(mode=1, seed=seed)))
        rng = np.random.default_rng(3)
        for t in range(20):
            act, ra = rng.integers(0, 5, size=(n, 2)).astype(np.int8), rng.integers(0, 5, size=(n, 2)).astype(np.int8)
            red.copy_(torch.as_tensor(ra))
            obs, rew, term, trunc, _ = env.step(torch.as_tensor(act, device=cuda_device))
            oo, orew, oterm, otrunc = o.step(act, oc.map_rng(mode=1, seed=seed, red_actions=ra), autoreset=True)
            assert np.array_equal(_np(obs), oo) and np.array_equal(_np(rew), orew), f"step {t}"
        env.close()


def test_device_decisions_are_the_host_policies_decisions(cuda_device):
    """randomness = 1: `choice([True, False], p=[1, 0])` always follows the route, so a decision without a patrol draw is a pure
    function of the state - the device must return what the (reference-pinned) host policy returns."""
    import gym_multigrid_b200 as mg
    g = load_golden("ctf_3v4")
    fm = g["field_map"].astype(np.float64)
    nb, nr, n = 3, 4, 300
    names = ("FightPolicy", "CapturePolicy", "PatrolPolicy", "PatrolFightPolicy")
    env = mg.make_ctf_vec(n, g["field_map"], num_blue_agents=nb, num_red_agents=nr, max_steps=30, seed=2)
    env.set_enemy_policies(_policies(names, fm, (1.0,) * 4), device=True)
    host = _policies(names, fm, (1.0,) * 4)
    for p in host:
        p.random_generator = np.random.default_rng(0)
    along = set(env._policy_tables["along_border"].tolist())
    env.reset()
    gen = torch.Generator(device=cuda_device).manual_seed(3)
    compared = drawn = 0
    for t in range(40):
        d = {k: _np(v) for k, v in env.positional_obs().items()}
        pos = _np(env.agent_pos).astype(np.int64)
        env.step(torch.randint(0, 5, (n, nb), generator=gen, device=cuda_device, dtype=torch.int8))
        got = _np(env._red_buf)
        for e in range(n):
            obs = {key: v[e] for key, v in d.items()}
            for k, p in enumerate(host):
                cur = tuple(pos[e, nb + k].tolist())
                patrolling = k == 2 or (k == 3 and not any(fm[tuple(b)] in (1, 5) for b in pos[e, :nb].tolist()))
                if patrolling and cur in p._border_cells:
                    drawn += 1                      # the target is a draw: numpy's on the host, Philox's on the device
                    continue
                assert int(p.act(obs, cur)) == int(got[e, k]), (t, e, k, cur)
                compared += 1
    assert compared > 30000 and drawn > 100 and along
    env.close()


def test_device_policies_reject_what_they_cannot_express(cuda_device):
    import gym_multigrid_b200 as mg
    from gym_multigrid_b200.policy.ctf import heuristic as H
    g = load_golden("ctf_2v2")
    fm = g["field_map"].astype(np.float64)
    env = mg.make_ctf_vec(4, g["field_map"], num_blue_agents=2, num_red_agents=2)

    class Mine(H.FightPolicy):
        pass
    with pytest.raises(TypeError):
        env.set_enemy_policies([Mine(fm), H.RwPolicy()], device=True)
    with pytest.raises(ValueError):
        env.set_enemy_policies([H.PatrolPolicy(), H.RwPolicy()], device=True)          # built without a map: empty border
    with pytest.raises(ValueError):
        env.set_enemy_policies([H.FightPolicy(fm, ego_agent="blue"), H.RwPolicy()], device=True)
    with pytest.raises(ValueError):
        env.set_enemy_policies([H.FightPolicy(fm.T.copy()), H.RwPolicy()], device=True)
    assert env.set_enemy_policies([H.FightPolicy(fm), None], device=True) is not None
    env.reset()
    env.step(torch.zeros((4, 2), dtype=torch.int8, device=cuda_device))
    assert env.status() == 0
    env.close()


def test_step_async_runs_the_opponents_like_step(cuda_device):
    """step_async / step_wait on a CtF batch with opponents (device-decided, and host policies): the red team must decide before
    every step exactly as in `step` - same observations, rewards and flags as the blocking device path; non-"map" observation
    options are refused before the state is touched."""
    import gym_multigrid_b200 as mg
    g = load_golden("ctf_2v2")
    fm = g["field_map"].astype(np.float64)
    n = 600
    for device_side in (True, False):
        m = n if device_side else 12
        ref = mg.make_ctf_vec(m, g["field_map"], seed=4, max_steps=25)
        pip = mg.make_ctf_vec(m, g["field_map"], seed=4, max_steps=25)
        for e in (ref, pip):
            e.set_enemy_policies(_policies(("FightPolicy", "PatrolFightPolicy"), fm, (0.75, 0.6)), device=device_side,
                                 random_generator=np.random.default_rng(3))
            e.reset()
        rng = np.random.default_rng(1)
        moved = 0
        for t in range(40):
            act = rng.integers(0, 5, size=(m, 2)).astype(np.int8)
            want = ref.step(torch.as_tensor(act, device=cuda_device))
            pip.step_async(act)
            got = pip.step_wait()
            for x, y in zip(want[:4], got[:4]):
                assert np.array_equal(_np(x), y), f"device={device_side} step {t}"
            assert np.array_equal(_np(ref._red_buf), _np(pip._red_buf))
            moved += int((_np(pip._red_buf) != 0).sum())
        assert moved > 0, "the opponents never acted"
        ref.close(); pip.close()
    flat = mg.make_ctf_vec(8, g["field_map"], observation_option="flattened")
    flat.reset()
    before = flat.state.clone()
    with pytest.raises(NotImplementedError):
        flat.step_async(np.zeros((8, 2), np.int8))
    with pytest.raises(NotImplementedError):
        flat.step(np.zeros((8, 2), np.int8))
    assert torch.equal(flat.state, before), "a refused host step must not advance the state"
    flat.close()


class _TappedGenerator:
    """A numpy Generator that logs what the policy drew: ("follow", bool) from choice([True, False], p=...), ("cell", (x, y))
    from choice(border cells), ("action", int) from integers(0, n)."""

    def __init__(self, seed):
        self.g, self.log = np.random.Generator(np.random.PCG64(seed)), []

    def choice(self, a, p=None, **kw):
        out = self.g.choice(a, p=p, **kw)
        if p is not None:
            self.log.append(("follow", bool(out)))
        else:
            self.log.append(("cell", tuple(int(v) for v in np.asarray(out).reshape(-1))))
        return out

    def integers(self, lo, hi=None, **kw):
        out = self.g.integers(lo, hi, **kw)
        self.log.append(("action", int(out)))
        return out

    def random(self, *a, **kw):
        raise AssertionError("the reference's policies draw through choice / integers only")


@pytest.mark.parametrize("pname", ["FightPolicy", "CapturePolicy", "PatrolPolicy", "PatrolFightPolicy"])
def test_device_decisions_replay_the_reference_draws(pname, cuda_device):
    """The device policy kernel against the REFERENCE's recorded decisions with randomness < 1 (tests/golden/ctf_policies.npz:
    320 decisions per policy on the 10x10 board, recorded from the unmodified reference classes with one seeded generator).
    The host policy - pinned to the reference decision for decision and draw for draw (tests/test_policies.py) - replays the run
    with a tapped generator; its draws (patrol cell, follow-or-not, uniform action) are handed to the kernel through
    mg_set_policy_trace; the kernel's action must be the recorded one for every decision."""
    import gym_multigrid_b200 as mg
    from gym_multigrid_b200.policy.ctf import heuristic as H
    from replay import policy_maps, policy_observation
    g = load_golden("ctf_policies")
    fm = policy_maps()["board"]
    stem = f"board_{pname}_red"
    curr, blue, red, want = g[f"{stem}_curr"], g[f"{stem}_blue"], g[f"{stem}_red"], g[f"{stem}_action"]
    randomness = float(g[f"{stem}_randomness"])
    assert 0.0 < randomness < 1.0
    tap = _TappedGenerator(int(g[f"{stem}_seed"]))
    pol = getattr(H, pname)(field_map=fm, random_generator=tap, ego_agent="red", randomness=randomness)
    n, S = len(want), fm.shape[0]
    patrol, follow, action = np.zeros((n, 2), np.uint16), np.zeros((n, 2), np.uint8), np.zeros((n, 2), np.int8)
    kinds = set()
    for e in range(n):
        tap.log.clear()
        got = int(pol.act(policy_observation(fm, blue[e], red[e]), tuple(curr[e])))
        assert got == int(want[e]), "host policy != reference (tests/test_policies.py covers this)"
        for what, v in tap.log:
            kinds.add(what)
            if what == "follow":
                follow[e, 0] = v
            elif what == "cell":
                patrol[e, 0] = v[0] * S + v[1]
            else:
                action[e, 0] = v
    assert "follow" in kinds and "action" in kinds and (("cell" in kinds) == pname.startswith("Patrol"))
    assert 0 < follow[:, 0].sum() < n, "the recording must hold followed and random decisions"
    nb = blue.shape[1]
    env = mg.make_ctf_vec(n, g_map(fm), num_blue_agents=nb, num_red_agents=2, autoreset=False)
    env.reset()
    ag = env._agents                                     # [n, agents, (x, y, dir, flags)]
    ag[:, :nb, 0:2] = torch.as_tensor(blue.astype(np.uint8), device=cuda_device)
    ag[:, nb, 0:2] = torch.as_tensor(curr.astype(np.uint8), device=cuda_device)       # the deciding agent = red agent 0
    ag[:, nb + 1, 0:2] = torch.as_tensor(red[:, 1].astype(np.uint8), device=cuda_device)
    ag[:, :, 3] = 0
    env.set_enemy_policies([getattr(H, pname)(fm, randomness=randomness), H.RwPolicy()], device=True)
    env.set_policy_trace(patrol, follow, action)
    import ctypes as C
    env._check(env._lib.mg_red_policy_actions(env._h, C.c_void_p(env.state.data_ptr()), C.c_void_p(env._red_buf.data_ptr()), env._stream()))
    got = _np(env._red_buf)[:, 0]
    # The fixture's observation dict lists the red territory WITHOUT the red flag cell (replay.policy_observation: the cells with
    # code 1), while the env hands its policies the territory plus the flag (ctf.py:765-769) - which is what the kernel tests.  The
    # two differ only when a blue agent stands on the red flag: those decisions (a handful) are left out of the comparison.
    flag = np.argwhere(fm == 5)[0]
    on_flag = (blue == flag[None, None, :]).all(axis=2).any(axis=1)
    assert on_flag.sum() < n // 10
    bad = np.nonzero((got != want) & ~on_flag)[0]
    assert len(bad) == 0, f"{pname}: device decision != reference at {bad[:8].tolist()}: {got[bad[:8]].tolist()} vs {want[bad[:8]].tolist()}"
    if pname != "PatrolFightPolicy":
        assert np.array_equal(got, want)        # the other policies never look at the territory
    env.set_policy_trace()
    env.close()


def g_map(fm):
    return fm.astype(np.uint8)
