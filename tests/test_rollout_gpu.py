"""mg_rollout (T steps per launch, state in shared memory, collect_rollout_kernels.cu) against the oracle and against T calls of
the per-step kernel: bit-exact, autoresets and final observations included.  mg_step launches the same warp-tile kernel with
T = 1 by default (every test of test_collect_gpu.py runs through it); the CTA-tile kernel of collect_kernels.cu stays selectable
(MG_STEP_IMPL=tile, also the fallback for grids too large for a warp's shared-memory slice) and is replayed here against the
reference's golden traces, with and without its early observation store."""
import os

import numpy as np
import pytest
import torch

import oracle as oc
from replay import COLLECT_FIXTURES, collect_kwargs, expected, load_golden, step_inputs

pytestmark = pytest.mark.gpu

ENV_ID = "multigrid-collect-respawn-clustered-v0"


def _np(t):
    return t.cpu().numpy()


def _oracle(n, seed, env_id_base=0, **kw):
    import gym_multigrid_b200 as mg
    s = mg.spec(ENV_ID)
    k = dict(layout="quadrants_respawn", time_limit=s.max_episode_steps, **s.kwargs)
    k.update(kw)
    o = oc.CollectOracle(oc.make_collect_cfg(**k), n)
    r = oc.PhiloxRng(seed=seed, env_id_base=env_id_base)
    return o, r


@pytest.mark.parametrize("n,T", [(64, 1), (4096, 60), (1000, 53), (33, 7), (2080, 128)])
def test_rollout_matches_oracle_and_step(n, T, cuda_device):
    """Given actions [T, N, A]: rollout == oracle stepped T times == T mg_step calls (obs, rewards, flags, final obs, state)."""
    import gym_multigrid_b200 as mg
    env = mg.make_vec(ENV_ID, n, device=cuda_device, seed=3, env_id_base=77)
    ref = mg.make_vec(ENV_ID, n, device=cuda_device, seed=3, env_id_base=77)
    ref.enable_final_observation()
    o, r = _oracle(n, 3, 77)
    env.reset(); ref.reset(); o.reset(r)
    act = np.random.default_rng(n + T).integers(0, 4, size=(T, n, 2)).astype(np.int8)
    obs, rew, term, trunc, info = env.rollout(torch.as_tensor(act, device=cuda_device), final_observation=True)
    assert obs.shape == (T, n, 10, 10, 3) and rew.shape == (T, n, 2)
    obs, rew, term, trunc, fin = _np(obs), _np(rew), _np(term), _np(trunc), _np(info["final_observation"])
    for t in range(T):
        oobs, orew, oterm, otrunc, ofin = o.step(act[t], r, autoreset=True, want_final_obs=True)
        assert np.array_equal(obs[t], oobs), f"obs, step {t}"
        assert np.array_equal(rew[t], orew) and np.array_equal(term[t], oterm) and np.array_equal(trunc[t], otrunc), f"step {t}"
        d = oterm | otrunc
        if d.any():
            assert np.array_equal(fin[t][d], ofin[d]), f"final obs, step {t}"
        s = ref.step(torch.as_tensor(act[t], device=cuda_device))
        assert np.array_equal(_np(s[0]), obs[t]) and np.array_equal(_np(s[1]), rew[t])
    assert torch.equal(env.state, ref.state), "state planes after the rollout differ from T separate steps"
    assert np.array_equal(_np(env.grid), o.grid) and np.array_equal(_np(env.agent_pos), o.agent_pos)
    assert np.array_equal(_np(env.pickups).reshape(n, -1), o.info)
    assert env.status() == 0
    env.close(); ref.close()


def test_rollout_without_observations_and_in_chunks(cuda_device):
    """obs=False (rewards only) gives the same rewards / flags / state; two rollouts of T/2 equal one of T."""
    import gym_multigrid_b200 as mg
    n, T = 3000, 40
    a, b, c = (mg.make_vec(ENV_ID, n, device=cuda_device, seed=5) for _ in range(3))
    for e in (a, b, c):
        e.reset()
    act = torch.randint(0, 4, (T, n, 2), device=cuda_device, dtype=torch.int8)
    oa = a.rollout(act)
    ob = b.rollout(act, obs=False)
    assert ob[0] is None
    for x, y in zip(oa[1:4], ob[1:4]):
        assert torch.equal(x, y)
    assert torch.equal(a.state, b.state)
    first = [t.clone() for t in c.rollout(act[:T // 2])[:4]]
    second = c.rollout(act[T // 2:])
    for k in range(4):
        assert torch.equal(torch.cat([first[k], second[k]]), oa[k])
    assert torch.equal(a.state, c.state)
    for e in (a, b, c):
        e.close()


def test_rollout_uniform_policy_on_device(cuda_device):
    """actions=None: the policy's actions are uniform over the 4 moves, a function of (seed, global env id, step, episode) only
    (shard-invariant), and the rollout equals the oracle stepped with exactly those actions."""
    import gym_multigrid_b200 as mg
    n, T = 8192, 64
    env = mg.make_vec(ENV_ID, n, device=cuda_device, seed=11)
    o, r = _oracle(n, 11)
    env.reset(); o.reset(r)
    obs, rew, term, trunc, info = env.rollout(steps=T)
    act = _np(info["actions"])
    assert act.shape == (T, n, 2) and act.min() == 0 and act.max() == 3
    freq = np.bincount(act.reshape(-1), minlength=4) / act.size
    assert np.all(np.abs(freq - 0.25) < 0.005), freq
    assert not np.array_equal(act[0], act[1]) and not np.array_equal(act[:, 0], act[:, 1])
    for t in range(T):
        oobs, orew, oterm, otrunc = o.step(act[t], r, autoreset=True)
        assert np.array_equal(_np(obs[t]), oobs) and np.array_equal(_np(rew[t]), orew), f"step {t}"
        assert np.array_equal(_np(term[t]), oterm) and np.array_equal(_np(trunc[t]), otrunc)
    # two shards of the same global env range draw the same actions
    lo = mg.make_vec(ENV_ID, n // 2, device=cuda_device, seed=11)
    hi = mg.make_vec(ENV_ID, n // 2, device=cuda_device, seed=11, env_id_base=n // 2)
    lo.reset(); hi.reset()
    a_lo, a_hi = _np(lo.rollout(steps=T)[4]["actions"]), _np(hi.rollout(steps=T)[4]["actions"])
    assert np.array_equal(np.concatenate([a_lo, a_hi], axis=1), act)
    for e in (env, lo, hi):
        e.close()


@pytest.mark.parametrize("early", ["0", "1"])
@pytest.mark.parametrize("stem", ["collect_respawn_clustered", "collect_rooms", "collect_quadrants15", "collect_single", "collect_rooms_respawn"])
def test_cta_tile_kernel_replays_reference_traces(stem, early, cuda_device, monkeypatch):
    """The CTA-tile kernel as the step implementation (MG_STEP_IMPL=tile, trace mode): the reference's own trajectories, on 64-env
    tiles with (MG_EARLY_OBS=1) and without the early observation store and its in-place patches."""
    from gym_multigrid_b200.vector_env import CollectVecEnv
    monkeypatch.setenv("MG_STEP_IMPL", "tile")
    monkeypatch.setenv("MG_EARLY_OBS", early)
    monkeypatch.setenv("MG_TILE", "0" if early == "1" else "9")
    g = load_golden(stem)
    E, T, A = g["actions"].shape
    tile = 3
    kw = collect_kwargs(g, stem)
    tl = kw.pop("time_limit")
    env = CollectVecEnv(E * tile, max_episode_steps=tl or None, autoreset=False, **kw)
    env.set_state_from_obs(np.concatenate([g["init_obs"]] * tile), np.concatenate([g["init_pos"]] * tile))
    for t in range(T):
        act, order, draws, n_draws, live = step_inputs(g, t, tile)
        tr = env.set_trace(order=order, draws=draws, n_draws=n_draws)
        obs, rew, term, trunc, info = env.step(torch.as_tensor(act, device=cuda_device))
        x = expected(g, t, tile)
        assert np.array_equal(_np(obs)[live], x["obs"][live]), f"step {t}"
        assert np.array_equal(_np(rew)[live], x["rewards"][live])
        assert np.array_equal(_np(tr["draws_used"])[live], n_draws[live])
    assert env.status() == 0
    env.close()


def test_early_store_patches_with_a_pickup_in_every_env(cuda_device, monkeypatch):
    """Every env of every (full, 64-env) tile picks a ball up at every step - each agent is surrounded by balls - so the early
    observation store's in-place patches of the slab already in global memory (bulk store, wait, byte stores by other threads
    to the same lines) run for all 128 agents of a tile at once; compared with the oracle on the same state, Philox mode."""
    import gym_multigrid_b200 as mg
    monkeypatch.setenv("MG_STEP_IMPL", "tile")
    monkeypatch.setenv("MG_EARLY_OBS", "1")
    monkeypatch.setenv("MG_TILE", "0")
    n = 64 * 37
    env = mg.make_vec(ENV_ID, n, device=cuda_device, seed=2, max_episode_steps=None, max_steps=10**6)
    o, r = _oracle(n, 2, time_limit=0, max_steps=10**6)
    env.reset(); o.reset(r)
    rng = np.random.default_rng(0)
    picked = 0.0
    for t in range(40):
        # surround both agents with balls (colours 0..2), then move them: whatever they do, they pick one up
        grid, pos = o.grid.reshape(n, 10, 10).copy(), o.agent_pos
        grid[(grid & 3) == 2] = 0          # drop the old balls first: the grid must never fill up (respawn samples until it finds a free cell)
        for a in range(2):
            for dx, dy in ((0, -1), (1, 0), (0, 1), (-1, 0)):
                x, y = pos[:, a, 0].astype(int) + dx, pos[:, a, 1].astype(int) + dy
                ok = (x >= 1) & (x <= 8) & (y >= 1) & (y <= 8)
                e = np.nonzero(ok)[0]
                free = grid[e, x[e], y[e]] == 0
                grid[e[free], x[e[free]], y[e[free]]] = 2 | (int(rng.integers(0, 3)) << 2)
        o.grid[:] = grid.reshape(n, 100)
        env.grid.copy_(torch.as_tensor(o.grid, device=cuda_device))
        act = rng.integers(0, 4, size=(n, 2)).astype(np.int8)
        obs, rew, term, trunc, _ = env.step(torch.as_tensor(act, device=cuda_device))
        oobs, orew, oterm, otrunc = o.step(act, r, autoreset=True)
        assert np.array_equal(_np(obs), oobs), f"step {t}"
        assert np.array_equal(_np(rew), orew)
        picked += orew.sum()
    assert picked > 0.8 * 40 * n, "the construction must force pickups almost everywhere"
    assert env.status() == 0
    env.close()


@pytest.mark.parametrize("transport", ["delta", "packed"])
def test_cta_tile_kernel_host_transports(transport, cuda_device, monkeypatch):
    """MG_STEP_IMPL=tile behind mg_step_host: delta records and reset rows written by the CTA-tile kernel."""
    import gym_multigrid_b200 as mg
    n = 2051
    dev = mg.make_vec(ENV_ID, n, device=cuda_device, seed=9)
    monkeypatch.setenv("MG_STEP_IMPL", "tile")
    host = mg.make_vec(ENV_ID, n, device=cuda_device, seed=9, host_transport=transport)
    dev.reset(); host.reset()
    rng = np.random.default_rng(3)
    for t in range(110):
        act = rng.integers(0, 4, size=(n, 2)).astype(np.int8)
        a = dev.step(torch.as_tensor(act, device=cuda_device))
        b = host.step(act)
        for x, y in zip(a[:4], b[:4]):
            assert np.array_equal(_np(x), y), f"step {t}"
    assert host.status() == 0
    dev.close(); host.close()


def test_grids_too_large_for_a_warp_slice_fall_back_to_the_cta_tile_kernel(cuda_device):
    """40x40 grids: 32 envs x 4 x 1600 bytes per warp do not fit two warps' shared-memory slices, so `mg_step` launches the
    CTA-tile kernel (16-env tiles) instead - same results as the oracle - and `mg_rollout` says so instead of launching."""
    from gym_multigrid_b200.vector_env import CollectVecEnv
    kw = dict(size=40, agents_index=[3, 5], balls_index=[0, 1, 2], balls_reward=[1, 1, 1], num_balls=[30], respawn=True, layout="even_dist", max_steps=20)
    n = 300
    env = CollectVecEnv(n, seed=4, **kw)
    o = oc.CollectOracle(oc.make_collect_cfg(size=40, agents_index=[3, 5], balls_index=[0, 1, 2], balls_reward=[1, 1, 1], num_balls=30, respawn=True,
                                             layout="even_dist", max_steps=20, time_limit=0), n)
    r = oc.PhiloxRng(seed=4)
    assert np.array_equal(_np(env.reset()[0]), o.reset(r))
    rng = np.random.default_rng(2)
    for t in range(45):
        act = rng.integers(0, 4, size=(n, 2)).astype(np.int8)
        obs, rew, term, trunc, _ = env.step(torch.as_tensor(act, device=cuda_device))
        oobs, orew, oterm, otrunc = o.step(act, r, autoreset=True)
        assert np.array_equal(_np(obs), oobs) and np.array_equal(_np(rew), orew), f"step {t}"
        assert np.array_equal(_np(term), oterm) and np.array_equal(_np(trunc), otrunc)
    with pytest.raises(RuntimeError, match="too large"):
        env.rollout(steps=4)
    assert env.status() == 0
    env.close()
