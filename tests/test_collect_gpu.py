"""GPU parity suite (run on the B200 box: `pytest -m gpu`).  Everything goes through the C ABI
(include/multigrid_b200.h) via gym_multigrid_b200; the oracle is only the checker.

 * bit-exact replay of the golden traces recorded from the reference (tests/golden/*.npz),
   tiled across thousands of envs so that every tile shape (full, ragged last tile) is hit;
 * CUDA vs the C oracle in Philox mode on seeded inputs (same counter-based RNG on both sides);
 * size-independent properties at BASELINE.json's full sizes (65 536 envs).
"""
import numpy as np
import pytest
import torch

import oracle as oc
from replay import COLLECT_FIXTURES, collect_kwargs, expected, load_golden, step_inputs

pytestmark = pytest.mark.gpu


def _make(g, stem, n, **extra):
    import gym_multigrid_b200 as mg
    from gym_multigrid_b200.vector_env import CollectVecEnv
    kw = collect_kwargs(g, stem)
    tl = kw.pop("time_limit")
    kw.update(extra)
    return CollectVecEnv(n, max_episode_steps=tl or None, **kw)


def _np(t):
    return t.cpu().numpy()


@pytest.mark.parametrize("stem", sorted(COLLECT_FIXTURES))
def test_reset_replay_bit_exact(stem, cuda_device):
    g = load_golden(stem)
    E = len(g["length"])
    tile = 3
    env = _make(g, stem, E * tile, autoreset=False)
    t = env.set_trace(reset_draws=np.concatenate([g["reset_draws"]] * tile), n_reset_draws=np.concatenate([g["n_reset_draws"]] * tile))
    obs, _ = env.reset()
    assert env.status() == 0
    assert np.array_equal(_np(obs), np.concatenate([g["init_obs"]] * tile))
    assert np.array_equal(_np(env.agent_pos), np.concatenate([g["init_pos"]] * tile))
    assert np.array_equal(_np(t["reset_draws_used"]), np.concatenate([g["n_reset_draws"]] * tile))
    assert int(env.step_count.abs().sum()) == 0 and int(env.collected_balls.abs().sum()) == 0
    env.close()


@pytest.mark.parametrize("stem", sorted(COLLECT_FIXTURES))
def test_step_replay_bit_exact(stem, cuda_device):
    """The reference's own trajectories, replayed through the CUDA step kernel."""
    g = load_golden(stem)
    E, T, A = g["actions"].shape
    tile = 5 if stem == "collect_respawn_clustered" else 11   # 1280 / 66..132 envs: full + ragged tiles
    env = _make(g, stem, E * tile, autoreset=False)
    env.set_state_from_obs(np.concatenate([g["init_obs"]] * tile), np.concatenate([g["init_pos"]] * tile))
    for t in range(T):
        act, order, draws, n_draws, live = step_inputs(g, t, tile)
        tr = env.set_trace(order=order, draws=draws, n_draws=n_draws)
        obs, rew, term, trunc, info = env.step(torch.as_tensor(act, device=cuda_device))
        x = expected(g, t, tile)
        assert np.array_equal(_np(tr["draws_used"])[live], n_draws[live]), f"step {t}: draws consumed"
        assert np.array_equal(_np(obs)[live], x["obs"][live]), f"step {t}: obs"
        assert np.array_equal(_np(rew)[live], x["rewards"][live]), f"step {t}: rewards"
        assert np.array_equal(_np(term)[live], x["terminated"][live]), f"step {t}: terminated"
        assert np.array_equal(_np(trunc)[live], x["truncated"][live]), f"step {t}: truncated"
        assert np.array_equal(_np(env.agent_pos)[live], x["pos"][live]), f"step {t}: agent positions"
        assert np.array_equal(_np(env.collected_balls)[live], x["collected"][live]), f"step {t}: collected"
        ni = env.num_agents * env.num_ball_types
        assert np.array_equal(_np(env.pickups).reshape(len(live), -1)[live], x["info"][live][:, :ni]), f"step {t}: info"
    assert env.status() == 0
    env.close()


def test_replay_at_65536_envs(cuda_device):
    """BASELINE config 2 at full size: 256 reference traces x 50 steps tiled over 65 536 envs."""
    stem = "collect_respawn_clustered"
    g = load_golden(stem)
    E, T, A = g["actions"].shape
    tile = 65536 // E
    env = _make(g, stem, E * tile, autoreset=False)
    env.set_state_from_obs(np.concatenate([g["init_obs"]] * tile), np.concatenate([g["init_pos"]] * tile))
    for t in range(T):
        act, order, draws, n_draws, live = step_inputs(g, t, tile)
        env.set_trace(order=order, draws=draws, n_draws=n_draws)
        obs, rew, term, trunc, _ = env.step(torch.as_tensor(act, device=cuda_device))
        if t % 7 == 0 or t == T - 1:
            x = expected(g, t, tile)
            assert np.array_equal(_np(obs), x["obs"]) and np.array_equal(_np(rew), x["rewards"])
            assert np.array_equal(_np(trunc), x["truncated"]) and np.array_equal(_np(term), x["terminated"])
    assert env.status() == 0
    env.close()


@pytest.mark.parametrize("stem,n", [("collect_respawn_clustered", 4099), ("collect_respawn", 1000), ("collect_rooms_respawn", 777),
                                    ("collect_quadrants15", 300), ("collect_single", 129), ("collect_even", 64)])
def test_philox_mode_matches_oracle(stem, n, cuda_device):
    """Production RNG mode: same Philox streams on both sides -> CUDA == oracle, incl. autoreset."""
    g = load_golden(stem)
    kw = collect_kwargs(g, stem)
    env = _make(g, stem, n, autoreset=True, seed=2024, env_id_base=17)
    env.enable_final_observation()
    o = oc.CollectOracle(oc.make_collect_cfg(**kw), n)
    r = oc.PhiloxRng(seed=2024, env_id_base=17)
    obs, _ = env.reset()
    assert np.array_equal(_np(obs), o.reset(r))
    rng = np.random.default_rng(5)
    steps = 120 if kw["time_limit"] else 210
    for t in range(steps):
        act = rng.integers(-1, 5, size=(n, env.num_agents)).astype(np.int8)
        obs, rew, term, trunc, info = env.step(torch.as_tensor(act, device=cuda_device))
        oobs, orew, oterm, otrunc, ofin = o.step(act, r, autoreset=True, want_final_obs=True)
        assert np.array_equal(_np(obs), oobs), f"step {t}: obs"
        assert np.array_equal(_np(rew), orew), f"step {t}: rewards"
        assert np.array_equal(_np(term), oterm) and np.array_equal(_np(trunc), otrunc), f"step {t}: flags"
        done = oterm | otrunc
        assert np.array_equal(_np(info["_final_observation"]), done)
        assert np.array_equal(_np(info["final_observation"])[done], ofin[done]), f"step {t}: final obs"
        assert np.array_equal(_np(env.grid), o.grid), f"step {t}: grids"
        assert np.array_equal(_np(env.agent_pos), o.agent_pos)
        assert np.array_equal(_np(env.step_count), o.step_count) and np.array_equal(_np(env.collected_balls), o.collected)
        assert np.array_equal(_np(env.pickups).reshape(n, -1), o.info)
    assert env.status() == 0
    assert int(env.episode_count.min()) >= 2   # every env went through the on-device autoreset
    env.close()


def test_encode_matches_oracle_all_codes(cuda_device):
    """Grid.encode kernel on every possible packed cell value, ragged size."""
    g = load_golden("collect_respawn_clustered")
    n = 1031
    env = _make(g, "collect_respawn_clustered", n, autoreset=False)
    cells = torch.randint(0, 256, (n, 100), dtype=torch.uint8, device=cuda_device)
    cells[0, :] = torch.arange(100, dtype=torch.uint8)
    cells[1, :] = torch.arange(156, 256, dtype=torch.uint8)
    env.grid.copy_(cells)
    out = torch.zeros((n, 10, 10, 3), dtype=torch.uint8, device=cuda_device)
    env.encode(out)
    assert np.array_equal(_np(out).reshape(n, 100, 3), oc.encode3(_np(cells)))
    # unaligned output pointer -> plain-store path must give the same bytes
    buf = torch.zeros(n * 300 + 1, dtype=torch.uint8, device=cuda_device)
    env.encode(buf[1:].view(n, 10, 10, 3))
    assert np.array_equal(_np(buf[1:]).reshape(n, 100, 3), oc.encode3(_np(cells)))
    env.close()


@pytest.mark.parametrize("transport", ["full", "packed", "delta"])
def test_step_host_matches_device_path(transport, cuda_device):
    """numpy in -> numpy out through mg_step_host equals the device path, whatever crosses PCIe for the observation:
    the expanded array, the packed plane, or per-env delta records patched into the host mirror (autoresets included)."""
    g = load_golden("collect_respawn_clustered")
    envs = [_make(g, "collect_respawn_clustered", 2051, autoreset=True, seed=9, host_transport=transport) for _ in range(2)]
    for e in envs:
        e.reset()
    rng = np.random.default_rng(3)
    for t in range(110):     # two lockstep TimeLimit autoresets inside
        act = rng.integers(0, 4, size=(2051, 2)).astype(np.int8)
        a = envs[0].step(torch.as_tensor(act, device=cuda_device))
        b = envs[1].step(act)   # numpy in -> numpy out through mg_step_host
        assert isinstance(b[0], np.ndarray)
        for x, y in zip(a[:4], b[:4]):
            assert np.array_equal(_np(x), y), f"step {t}"
        if t == 30:     # a device-path step in between invalidates the delta transport's mirror: the next host step refreshes it
            for e in envs:
                e.step(torch.as_tensor(act, device=cuda_device))
        if t == 60:     # ... and so does writing the state
            st = envs[0].get_state()
            envs[1].set_state(st)
    for e in envs:
        assert e.status() == 0
        e.close()


# (kwargs, n): staggered terminations (no respawn: episodes end when the last ball is collected), a wide-record grid
# (W*H > 256 -> 16-bit cell indices), three agents, rooms ghosts
DELTA_CASES = [
    (dict(size=8, agents_index=[1, 2], balls_index=[0, 1], balls_reward=[1, 2], num_balls=[4], respawn=False, layout="even_dist", max_steps=30), 1500),
    (dict(size=17, agents_index=[3, 5], balls_index=[0, 1, 2], balls_reward=[1, 1, 1], num_balls=[15], respawn=True, layout="quadrants_respawn", max_steps=40), 700),
    (dict(size=10, agents_index=[3, 5, 1], balls_index=[0, 1, 2], balls_reward=[1, 1, 1], num_balls=[15], respawn=True, layout="even_dist", max_steps=25), 999),
    (dict(size=11, agents_index=[3, 5], balls_index=[0, 1, 2], balls_reward=[1, 1, 1], num_balls=[15], respawn=False, layout="rooms", max_steps=35), 1111),
]


@pytest.mark.parametrize("kw,n", DELTA_CASES)
@pytest.mark.parametrize("transport", ["packed", "delta"])
def test_compact_transports_with_final_observation(kw, n, transport, cuda_device):
    """Compact transports vs the device path incl. `final_observation` (the terminal rows come from the host mirror in delta
    mode), on envs that finish at different steps, with greedy-ish actions so that episodes really terminate."""
    from gym_multigrid_b200.vector_env import CollectVecEnv
    dev = CollectVecEnv(n, seed=21, autoreset=True, **kw)
    host = CollectVecEnv(n, seed=21, autoreset=True, host_transport=transport, **kw)
    dev.enable_final_observation()
    dev.reset(); host.reset()
    A = len(kw["agents_index"])
    W = kw["size"]
    hb = None
    rng = np.random.default_rng(5)
    import ctypes as C
    from gym_multigrid_b200 import _lib
    fin = torch.zeros((n, W, W, 3), dtype=torch.uint8).pin_memory()
    finished = 0
    for t in range(90):
        act = rng.integers(0, 4, size=(n, A)).astype(np.int8)
        want = dev.step(torch.as_tensor(act, device=cuda_device))
        io = host._host_io(act)
        io.final_obs = fin.data_ptr()
        host._check(host._lib.mg_step_host(host._h, C.c_void_p(host.state.data_ptr()), C.byref(io), host._stream()))
        got = host._host_result()
        for x, y in zip(want[:4], got[:4]):
            assert np.array_equal(_np(x), y), f"step {t}"
        done = _np(want[2] | want[3])
        finished += int(done.sum())
        if done.any():
            assert np.array_equal(_np(want[4]["final_observation"])[done], fin.numpy()[done]), f"final_observation, step {t}"
    assert finished > n, "the case must exercise autoresets"
    assert dev.status() == 0 and host.status() == 0
    dev.close(); host.close()


@pytest.mark.parametrize("kw,n", DELTA_CASES[:2])
@pytest.mark.parametrize("transport", ["packed", "delta"])
def test_async_decode_three_batches_with_autoreset_rows(kw, n, transport, cuda_device):
    """mg_step_host_async / _wait with the decode running on the host pool while the caller enqueues the other batches (three in
    flight): staggered terminations, so most steps carry reset rows (the count-sized second copy joins the delta pass) and terminal
    observations taken from the mirror; the steady-state steps go out as one captured graph launch.  == the device path."""
    import ctypes as C
    from gym_multigrid_b200.vector_env import CollectVecEnv
    B = 3
    devs = [CollectVecEnv(n, seed=30 + b, autoreset=True, **kw) for b in range(B)]
    hosts = [CollectVecEnv(n, seed=30 + b, autoreset=True, host_transport=transport, **kw) for b in range(B)]
    W, A = kw["size"], len(kw["agents_index"])
    fins = [torch.zeros((n, W, W, 3), dtype=torch.uint8).pin_memory() for _ in range(B)]
    for d, h in zip(devs, hosts):
        d.enable_final_observation(); d.reset(); h.reset()
    rng = np.random.default_rng(6)
    T = 70
    acts = rng.integers(0, 4, size=(T, B, n, A)).astype(np.int8)

    def enqueue(b, t):
        io = hosts[b]._host_io(acts[t, b])
        io.final_obs = fins[b].data_ptr()
        hosts[b].step_async(acts[t, b])
    for b in range(B):
        enqueue(b, 0)
    finished = 0
    for t in range(T):
        for b in range(B):
            want = devs[b].step(torch.as_tensor(acts[t, b], device=cuda_device))
            got = hosts[b].step_wait()
            for x, y in zip(want[:4], got[:4]):
                assert np.array_equal(_np(x), y), f"step {t} batch {b}"
            done = _np(want[2] | want[3])
            finished += int(done.sum())
            if done.any():
                assert np.array_equal(_np(want[4]["final_observation"])[done], fins[b].numpy()[done]), f"final_observation, step {t} batch {b}"
            if t + 1 < T:
                enqueue(b, t + 1)
    assert finished >= B * n, "the case must exercise autoresets"
    for e in devs + hosts:
        assert e.status() == 0
        e.close()


@pytest.mark.parametrize("transport", ["full", "delta"])
def test_step_async_wait_two_batches_in_flight(transport, cuda_device):
    """step_async / step_wait (mg_step_host_async / _wait): two env batches alternated with both in flight return exactly what
    the blocking device path returns, also when device-path calls are mixed in between."""
    g = load_golden("collect_respawn_clustered")
    n = 3000
    ref = [_make(g, "collect_respawn_clustered", n, autoreset=True, seed=s) for s in (4, 5)]
    pip = [_make(g, "collect_respawn_clustered", n, autoreset=True, seed=s, host_transport=transport) for s in (4, 5)]
    for e in ref + pip:
        e.reset()
    rng = np.random.default_rng(8)
    acts = [[rng.integers(0, 4, size=(n, 2)).astype(np.int8) for _ in range(2)] for _ in range(40)]
    with pytest.raises(RuntimeError):
        pip[0].step_wait()
    pip[0].step_async(acts[0][0])
    with pytest.raises(RuntimeError):
        pip[0].step_async(acts[0][0])
    pip[1].step_async(acts[0][1])
    for t in range(40):
        for b in range(2):
            want = ref[b].step(torch.as_tensor(acts[t][b], device=cuda_device))
            got = pip[b].step_wait()
            for x, y in zip(want[:4], got[:4]):
                assert np.array_equal(_np(x), y), f"step {t} batch {b}"
            if t == 20 and b == 0:      # a device-path step between two host-path steps stays ordered
                a = torch.as_tensor(acts[t][1], device=cuda_device)
                x = ref[b].step(a)[0].clone(); y = pip[b].step(a)[0]
                assert torch.equal(x, y)
            if t + 1 < 40:
                pip[b].step_async(acts[t + 1][b])
    for e in ref + pip:
        assert e.status() == 0
        e.close()


def test_reseed_makes_everything_a_function_of_the_seed(cuda_device):
    """reseed(s) (mg_set_seed + zeroed block counters) then reset: the same as a fresh env constructed with seed s - same
    placements, same episode for the same actions; a different seed differs."""
    import gym_multigrid_b200 as mg
    n = 777
    fresh = mg.make_vec("multigrid-collect-respawn-clustered-v0", n, seed=123)
    used = mg.make_vec("multigrid-collect-respawn-clustered-v0", n, seed=9)
    gen = torch.Generator(device=cuda_device).manual_seed(0)
    used.reset()
    for _ in range(17):
        used.step(torch.randint(0, 4, (n, 2), generator=gen, device=cuda_device, dtype=torch.int8))
    other = _np(used.reset()[0]).copy()
    used.reseed(123)
    a, b = fresh.reset()[0], used.reset()[0]
    assert torch.equal(a, b) and not np.array_equal(_np(a), other)
    for _ in range(60):
        act = torch.randint(0, 4, (n, 2), generator=gen, device=cuda_device, dtype=torch.int8)
        x, y = fresh.step(act), used.step(act)
        assert all(torch.equal(u, v) for u, v in zip(x[:4], y[:4]))
    fresh.close(); used.close()


def test_properties_at_full_size(cuda_device):
    """Size-independent invariants at 65 536 envs (BASELINE config 2), Philox mode, autoreset."""
    import gym_multigrid_b200 as mg
    n = 65536
    env = mg.make_vec("multigrid-collect-respawn-clustered-v0", n, seed=1)
    obs, _ = env.reset()
    walls = torch.zeros((10, 10), dtype=torch.bool, device=cuda_device)
    walls[0, :] = walls[-1, :] = walls[:, 0] = walls[:, -1] = True
    gen = torch.Generator(device=cuda_device).manual_seed(0)
    total_reward = torch.zeros((n, 2), dtype=torch.float64, device=cuda_device)
    for t in range(1, 101):
        act = torch.randint(0, 4, (n, 2), generator=gen, device=cuda_device, dtype=torch.int8)
        prev_pick = env.pickups.sum(dim=(1, 2)).clone()
        obs, rew, term, trunc, info = env.step(act)
        typ = obs[..., 0]
        assert bool(((typ == 1) == walls).all()), "border walls intact, no wall elsewhere"
        assert bool(((typ == 3).sum(dim=(1, 2)) == 2).all()), "exactly two agents visible"   # quadrant layouts: agents never co-located
        assert bool(((typ == 2).sum(dim=(1, 2)) <= 15).all()), "balls never exceed 15 (respawn can only lose balls)"
        assert bool((obs[..., 2][typ == 3] == 3).all()) and bool((obs[..., 2][typ != 3] == 0).all()), "state plane = dir 3 on agents only"
        assert not bool(term.any()), "respawn variant never terminates"
        assert bool((trunc == (t % 50 == 0)).all()), "TimeLimit 50 truncation, in lockstep"
        if t % 50 != 0:
            assert bool((rew.sum(1) == (env.pickups.sum(dim=(1, 2)) - prev_pick)).all()), "reward == pickups this step"
            total_reward += rew
        else:
            assert int(env.step_count.max()) == 0 and int(env.pickups.abs().sum()) == 0, "autoreset cleared counters"
            total_reward.zero_()
    assert env.status() == 0
    env.close()


def test_single_env_adapter_matches_reference_trace(cuda_device):
    """config 1: the gymnasium.make-style single env with the reference's signatures."""
    import gym_multigrid_b200 as mg
    g = load_golden("collect_respawn_clustered")
    env = mg.make("multigrid-collect-respawn-clustered-v0")
    assert env.action_space.n == 4 and env.observation_space.shape == (10, 10, 3)
    e = 3
    env.vec.set_state_from_obs(g["init_obs"][e:e + 1], g["init_pos"][e:e + 1])
    for t in range(int(g["length"][e])):
        env.vec.set_trace(order=g["order"][e:e + 1, t], draws=g["draws"][e:e + 1, t], n_draws=g["n_draws"][e:e + 1, t])
        obs, rew, term, trunc, info = env.step([int(a) for a in g["actions"][e, t]])
        assert obs.dtype == np.uint8 and rew.dtype == np.float64 and isinstance(term, bool) and isinstance(trunc, bool)
        assert np.array_equal(obs, g["obs"][e, t]) and np.array_equal(rew, g["rewards"][e, t])
        assert term == bool(g["terminated"][e, t]) and trunc == bool(g["truncated"][e, t])
        assert [info[k] for k in env.keys] == g["info"][e, t].tolist()
    env.close()


def test_create_rejects_bad_configs(cuda_device):
    from gym_multigrid_b200.vector_env import CollectVecEnv
    with pytest.raises(ValueError):
        CollectVecEnv(4, size=2)
    with pytest.raises(ValueError):
        CollectVecEnv(4, size=10, num_balls=200)
    with pytest.raises(ValueError):
        CollectVecEnv(4, size=10, num_balls=16, layout="quadrants_respawn")
    with pytest.raises(ValueError):
        CollectVecEnv(4, device="cpu")


# (size, agents_index, balls_index, balls_reward, num_balls, respawn, layout, time_limit, n)
SWEEP = [
    (5, [4], [1], [2.5], 3, True, "even_dist", 7, 70),                      # smallest useful grid, one agent, one ball type
    (7, [3, 5, 8], [0, 2], [1, -1], 6, False, "even_dist", 0, 131),          # three agents, negative reward, terminates
    (9, [3, 5, 1, 2], [0, 1, 2, 4], [1, 2, 3, 4], 8, True, "even_dist", 20, 65),
    (12, [3, 5], [0, 1, 2], [1, 1, 1], 12, True, "quadrants_respawn", 30, 257),
    (16, [3, 5], [0, 1, 2, 3], [1, 1, 1, 1], 16, False, "quadrants", 40, 100),
    (11, [3, 5, 6], [0, 1, 2], [1, 1, 1], 9, True, "rooms", 25, 90),
    (20, [0, 9, 3, 5, 6, 7, 1, 2], [4], [0.5], 10, True, "even_dist", 15, 64),  # eight agents (MG_MAX_AGENTS), fractional reward
]


@pytest.mark.parametrize("cfg", SWEEP, ids=[f"{c[6]}-{c[0]}x{c[0]}-A{len(c[1])}-nb{len(c[2])}" for c in SWEEP])
def test_config_sweep_matches_oracle(cfg, cuda_device):
    """Grid sizes, agent counts, ball-type counts, rewards and layouts the registered ids never use: CUDA == oracle for the
    step (with autoreset and final observations), partial views, the toroid wrapper and Grid.encode."""
    import gym_multigrid_b200 as mg
    from gym_multigrid_b200.vector_env import CollectVecEnv
    size, agents, balls, rewards, num_balls, respawn, layout, tl, n = cfg
    kw = dict(size=size, agents_index=agents, balls_index=balls, balls_reward=rewards, num_balls=num_balls, respawn=respawn, layout=layout)
    env = CollectVecEnv(n, max_episode_steps=tl or None, seed=77, env_id_base=5, **kw)
    env.enable_final_observation()
    o = oc.CollectOracle(oc.make_collect_cfg(time_limit=tl, **kw), n)
    r = oc.PhiloxRng(seed=77, env_id_base=5)
    obs, _ = env.reset()
    assert np.array_equal(_np(obs), o.reset(r))
    rng = np.random.default_rng(size)
    A = len(agents)
    for t in range(60):
        act = rng.integers(0, 4, size=(n, A)).astype(np.int8)
        obs, rew, term, trunc, info = env.step(torch.as_tensor(act, device=cuda_device))
        oobs, orew, oterm, otrunc, ofin = o.step(act, r, autoreset=True, want_final_obs=True)
        assert np.array_equal(_np(obs), oobs) and np.array_equal(_np(rew), orew), f"step {t}"
        assert np.array_equal(_np(term), oterm) and np.array_equal(_np(trunc), otrunc), f"step {t}: flags"
        d = oterm | otrunc
        assert np.array_equal(_np(info["final_observation"])[d], ofin[d])
        assert np.array_equal(_np(env.grid), o.grid) and np.array_equal(_np(env.pickups).reshape(n, -1), o.info)
    dirs = rng.integers(0, 4, size=(n, A)).astype(np.uint8)
    for V in (3, 5, 7, 6):
        assert np.array_equal(_np(env.gen_obs(V, False, dirs=dirs)), oc.partial_view3(o.grid, o.agent_pos, size, size, V, False, dirs=dirs))
    assert np.array_equal(_np(env.toroid_obs()), oc.toroid(o.grid, o.agent_pos, size, len(balls)))
    assert np.array_equal(_np(env.encode()), oobs) and env.status() == 0
    env.close()
