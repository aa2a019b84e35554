"""GPU suite for the Wildfire EXTENSION (no reference code exists: the checker is the in-repo CPU oracle of OUR
specification, include/multigrid_b200.h "Wildfire"; parity vs the reference is unpinned by construction)."""
import numpy as np
import pytest
import torch

import oracle as oc

pytestmark = pytest.mark.gpu


def _np(t):
    return t.cpu().numpy()


@pytest.mark.parametrize("size,A,fires,n", [(16, 5, 3, 300), (64, 16, 4, 96), (32, 32, 6, 128), (8, 1, 1, 257), (4, 2, 2, 65),
                                            ((8, 10), 6, 3, 130), ((12, 20), 7, 5, 90), ((24, 6), 4, 2, 70), (128, 9, 5, 40)])
def test_wildfire_matches_oracle(size, A, fires, n, cuda_device):
    """Square grids and W x H grids; H % 4 == 0 takes the word-parallel kernel, other heights the generic one."""
    import gym_multigrid_b200 as mg
    kw = dict(num_agents=A, num_fires=fires, alpha=0.2, beta=0.08, max_steps=40, seed=13, env_id_base=9)
    kw.update(dict(width=size[0], height=size[1]) if isinstance(size, tuple) else dict(size=size))
    env = mg.make_wildfire_vec(n, **kw)
    env.enable_final_observation()
    o = oc.WildfireOracle(n, **kw)
    obs, _ = env.reset()
    assert np.array_equal(_np(obs), o.reset())
    assert np.array_equal(_np(env.agents), o.agents)
    rng = np.random.default_rng(0)
    ext = 0
    for t in range(90):
        act = rng.integers(0, 5, size=(n, A)).astype(np.int8)
        obs, rew, term, trunc, info = env.step(torch.as_tensor(act, device=cuda_device))
        oo, orew, oterm, otrunc, ofin = o.step(act, autoreset=True, want_final_obs=True)
        assert np.array_equal(_np(obs), oo), f"step {t}: obs"
        assert np.array_equal(_np(rew), orew), f"step {t}: rewards"
        assert np.array_equal(_np(term), oterm) and np.array_equal(_np(trunc), otrunc), f"step {t}: flags"
        d = oterm | otrunc
        assert np.array_equal(_np(info["final_observation"])[d], ofin[d])
        assert np.array_equal(_np(env.terrain), o.terrain) and np.array_equal(_np(env.agents), o.agents)
        assert np.array_equal(_np(env._planes["hdr"]), o.hdr)
        ext += int(orew.sum())
    assert ext > 0 and int(env.episode_count.min()) >= 2
    env.close()


def test_wildfire_replayed_order_and_invariants(cuda_device):
    import gym_multigrid_b200 as mg
    n, A, S = 512, 8, 32
    kw = dict(size=S, num_agents=A, num_fires=5, alpha=0.25, beta=0.05, max_steps=1000, seed=2)
    env = mg.make_wildfire_vec(n, autoreset=False, **kw)
    o = oc.WildfireOracle(n, **kw)
    env.reset(); o.reset()
    rng = np.random.default_rng(1)
    for t in range(40):
        order = np.stack([rng.permutation(A) for _ in range(n)]).astype(np.uint8)
        act = rng.integers(0, 5, size=(n, A)).astype(np.int8)
        env.set_order(order)
        obs, rew, term, trunc, _ = env.step(torch.as_tensor(act, device=cuda_device))
        oo, orew, oterm, otrunc = o.step(act, order=order)
        assert np.array_equal(_np(obs), oo) and np.array_equal(_np(rew), orew)
        typ = obs[..., 0]
        assert bool(((typ == 3).sum(dim=(1, 2)) == A).all()), "agents never share a cell"
        burnt_before = getattr(test_wildfire_replayed_order_and_invariants, "_b", None)
        burnt = (env.terrain == 2).sum(dim=1)
        if burnt_before is not None:
            assert bool((burnt >= burnt_before).all()), "burnt cells never recover"
        test_wildfire_replayed_order_and_invariants._b = burnt
    env.close()


def test_wildfire_rejects_bad_configs(cuda_device):
    import gym_multigrid_b200 as mg
    with pytest.raises(ValueError):
        mg.make_wildfire_vec(4, size=10)            # 100 cells: not a multiple of 16
    with pytest.raises(ValueError):
        mg.make_wildfire_vec(4, size=16, num_agents=33)


def test_wildfire_full_size_invariants(cuda_device):
    """BASELINE config 5 per-GPU share: 131 072 envs, 64x64 cells, 16 agents - size-independent properties."""
    import gym_multigrid_b200 as mg
    n, A, S = 131072, 16, 64
    env = mg.make_wildfire_vec(n, size=S, num_agents=A, num_fires=4, alpha=0.15, beta=0.05, max_steps=200, seed=5)
    obs, _ = env.reset()
    assert bool(((obs[..., 0] == 1).sum(dim=(1, 2)) + (obs[..., 0] == 3).sum(dim=(1, 2)) >= 4).all())
    gen = torch.Generator(device=cuda_device).manual_seed(1)
    prev_burnt = (env.terrain == 2).sum(dim=1)
    for t in range(8):
        obs, rew, term, trunc, _ = env.step(torch.randint(0, 5, (n, A), generator=gen, device=cuda_device, dtype=torch.int8))
        typ = obs[..., 0]
        assert bool(((typ == 3).sum(dim=(1, 2)) == A).all()), "agents never share a cell"
        assert bool((typ <= 3).all()) and bool(((obs[..., 2] == 0) | (typ == 3)).all())
        burnt = (env.terrain == 2).sum(dim=1)
        done = term | trunc
        assert bool((burnt[~done] >= prev_burnt[~done]).all()), "burnt cells never recover inside an episode"
        assert bool((rew.sum(dim=1) <= A).all()) and bool((rew >= 0).all())
        assert bool((term == ((env.terrain == 1).sum(dim=1) == 0))[~done].all())
        prev_burnt = burnt
    env.close()
