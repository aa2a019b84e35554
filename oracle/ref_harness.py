"""Run the UNMODIFIED reference under a shim and record RNG-tapped traces.

TEST INFRASTRUCTURE ONLY.  Usable only where /root/reference exists (the build
container); nothing in the product path, `-m gpu` tests, smoke() or bench.py imports
this module.  It is what `oracle/gen_golden.py` uses to produce `tests/golden/*.npz`.

How the reference is made to run here; no reference source is edited:
  * `oracle/refshim/` supplies stub `gymnasium` + `matplotlib` packages;
  * `np.float_` (removed in numpy 2) is aliased to `np.float64` before import
    (annotations at multigrid.py:399, collect_game.py:122,135,151, ctf.py:1365-1368);
  * `CollectGameQuadrantsRespawn.__init__(self)` (collect_game.py:372-374) accepts no
    kwargs although registration passes six (__init__.py:122-134): the instance is
    created with `object.__new__` and `CollectGameQuadrants.__init__` is called;
  * `env.num_balls = int(env.num_balls)` for the variants whose `_gen_grid` does
    `isinstance(self.num_balls, int)` on an `np.int64` (collect_game.py:37 vs 245,343).

RNG taps (the reference's hot path uses three generators, none seeded by reset()):
  * `random.randint`          <- MultiGridEnv._rand_int (multigrid.py:225-230)
  * `np.random.permutation`   <- CollectGameEnv.step agent order (collect_game.py:186)
  * `np.random.randint`       <- Maze start cell (maze.py:204)
  * proxy around `env.np_random` for CtF (`integers`, `shuffle`, `choice`).
"""
from __future__ import annotations

import os
import random
import sys
from contextlib import contextmanager

import numpy as np

REFERENCE_ROOT = os.environ.get("MG_REFERENCE_ROOT", "/root/reference")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "refshim")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "gym_multigrid"))


def import_reference():
    """Import `gym_multigrid` from /root/reference under the shim; returns the package."""
    if not reference_available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    if not hasattr(np, "float_"):
        np.float_ = np.float64  # numpy>=2 removed the alias the reference annotates with
    for p in (_SHIM, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import gym_multigrid  # noqa: F401  (registers the ids with the shim registry)
    import gym_multigrid.envs  # noqa: F401
    return gym_multigrid


# ----------------------------------------------------------------------------- taps
class Taps:
    """Records the outputs of the global-RNG call sites while installed."""

    def __init__(self):
        self.randint = []       # python random.randint outputs, in call order
        self.perm = []          # np.random.permutation outputs
        self.np_randint = []    # np.random.randint outputs

    def clear(self):
        self.randint.clear()
        self.perm.clear()
        self.np_randint.clear()


@contextmanager
def installed_taps():
    taps = Taps()
    o_randint, o_perm, o_nprandint = random.randint, np.random.permutation, np.random.randint

    def t_randint(a, b):
        v = o_randint(a, b)
        taps.randint.append(int(v))
        return v

    def t_perm(n):
        v = o_perm(n)
        taps.perm.append(np.asarray(v).copy())
        return v

    def t_nprandint(*a, **k):
        v = o_nprandint(*a, **k)
        taps.np_randint.append(int(v))
        return v

    random.randint, np.random.permutation, np.random.randint = t_randint, t_perm, t_nprandint
    try:
        yield taps
    finally:
        random.randint, np.random.permutation, np.random.randint = o_randint, o_perm, o_nprandint


class GeneratorProxy:
    """Wraps a numpy Generator and records integers/shuffle/choice outputs (CtF)."""

    def __init__(self, gen):
        self._gen = gen
        self.events = []  # (kind, payload)

    def integers(self, *a, **k):
        v = self._gen.integers(*a, **k)
        self.events.append(("integers", int(v)))
        return v

    def shuffle(self, x):
        self._gen.shuffle(x)
        self.events.append(("shuffle", list(int(i) for i in x)))

    def choice(self, a, size=None, replace=True, p=None):
        v = self._gen.choice(a, size=size, replace=replace, p=p)
        self.events.append(("choice", np.asarray(v).copy()))
        return v

    def __getattr__(self, name):
        return getattr(self._gen, name)


# ------------------------------------------------------------------- env construction
COLLECT_IDS = [
    "multigrid-collect-v0",
    "multigrid-collect-single-v0",
    "multigrid-collect-quadrants-v0",
    "multigrid-collect-rooms-v0",
    "multigrid-collect-rooms-fixed-horizon-v0",
    "multigrid-collect-rooms-respawn-v0",
    "multigrid-collect-respawn-v0",
    "multigrid-collect-respawn-clustered-v0",
    "multigrid-collect-quadrants15-v0",
]


def make_collect(env_id: str, **overrides):
    """Construct a reference Collect env for a registered id (`overrides` replace registered kwargs); returns (env, time_limit)."""
    import_reference()
    import importlib
    from gymnasium.envs.registration import registry

    spec = registry[env_id]
    mod, cls_name = spec["entry_point"].split(":")
    cls = getattr(importlib.import_module(mod), cls_name)
    kw = dict(spec["kwargs"])
    kw.update(overrides)
    if cls_name == "CollectGameQuadrantsRespawn":
        from gym_multigrid.envs.collect_game import CollectGameQuadrants
        env = object.__new__(cls)
        CollectGameQuadrants.__init__(env, **kw)
    else:
        env = cls(**kw)
    env.num_balls = int(env.num_balls)
    return env, spec["max_episode_steps"]


# ---------------------------------------------------------------- Collect trace record
def record_collect_episode(env, time_limit, seed, action_rng, oob_action_prob=0.0, max_len=None):
    """One episode of a reference Collect env -> dict of arrays.

    Seeds the generators the reference really uses (`random`, legacy `np.random`), the
    way the reference's own `set_seed` does (utils/misc.py:9-19).  The TimeLimit wrapper
    gymnasium.make would add (truncated |= elapsed >= max_episode_steps) is emulated here.
    """
    random.seed(seed)
    np.random.seed(seed)
    A = len(env.agents)
    with installed_taps() as taps:
        obs0, info0 = env.reset(seed=seed)
        reset_draws = list(taps.randint)
        taps.clear()
        W, H = env.width, env.height
        init_pos = np.array([np.asarray(a.pos) for a in env.agents], dtype=np.int16)
        init_dir = np.array([a.dir for a in env.agents], dtype=np.int8)
        T = max_len or (time_limit if time_limit is not None else env.max_steps)
        rec = dict(actions=[], order=[], draws=[], n_draws=[], obs=[], rewards=[], terminated=[],
                   truncated=[], info=[], pos=[], collected=[])
        t = 0
        while True:
            acts = action_rng.integers(0, 4, size=A)
            if oob_action_prob > 0:
                oob = action_rng.random(A) < oob_action_prob
                acts = np.where(oob, action_rng.choice([-1, 4, 7], size=A), acts)
            obs, rew, term, trunc, info = env.step([int(a) for a in acts])
            t += 1
            if time_limit is not None and t >= time_limit:
                trunc = True
            assert len(taps.perm) == 1
            rec["actions"].append(acts.astype(np.int8))
            rec["order"].append(taps.perm[0].astype(np.uint8))
            rec["draws"].append(np.array(taps.randint, dtype=np.uint8))
            rec["n_draws"].append(len(taps.randint))
            rec["obs"].append(obs.copy())
            rec["rewards"].append(np.asarray(rew, dtype=np.float64).copy())
            rec["terminated"].append(bool(term))
            rec["truncated"].append(bool(trunc))
            rec["info"].append(np.array([info[k] for k in env.keys], dtype=np.int32))
            rec["pos"].append(np.array([np.asarray(a.pos) for a in env.agents], dtype=np.int16))
            rec["collected"].append(int(env.collected_balls))
            taps.clear()
            if term or trunc or t >= T:
                break
    out = dict(
        init_obs=np.asarray(obs0, dtype=np.uint8).copy(), init_pos=init_pos, init_dir=init_dir,
        reset_draws=np.array(reset_draws, dtype=np.uint8), length=t,
        actions=np.stack(rec["actions"]), order=np.stack(rec["order"]),
        draws=rec["draws"], n_draws=np.array(rec["n_draws"], dtype=np.int32),
        obs=np.stack(rec["obs"]), rewards=np.stack(rec["rewards"]),
        terminated=np.array(rec["terminated"]), truncated=np.array(rec["truncated"]),
        info=np.stack(rec["info"]), pos=np.stack(rec["pos"]),
        collected=np.array(rec["collected"], dtype=np.int32),
    )
    return out


def pack_collect_episodes(eps, T, K):
    """Stack episodes into fixed-shape arrays: steps padded to T, draws padded to K ints."""
    E = len(eps)
    A = eps[0]["actions"].shape[1]
    W, H, D = eps[0]["init_obs"].shape
    R = max(len(e["reset_draws"]) for e in eps)
    out = dict(
        init_obs=np.stack([e["init_obs"] for e in eps]),
        init_pos=np.stack([e["init_pos"] for e in eps]),
        init_dir=np.stack([e["init_dir"] for e in eps]),
        length=np.array([e["length"] for e in eps], dtype=np.int32),
        n_reset_draws=np.array([len(e["reset_draws"]) for e in eps], dtype=np.int32),
        reset_draws=np.zeros((E, R), np.uint8),
        actions=np.zeros((E, T, A), np.int8), order=np.zeros((E, T, A), np.uint8),
        n_draws=np.zeros((E, T), np.int32), draws=np.zeros((E, T, K), np.uint8),
        obs=np.zeros((E, T, W, H, D), np.uint8), rewards=np.zeros((E, T, A), np.float64),
        terminated=np.zeros((E, T), bool), truncated=np.zeros((E, T), bool),
        info=np.zeros((E, T, eps[0]["info"].shape[1]), np.int32),
        pos=np.zeros((E, T, A, 2), np.int16), collected=np.zeros((E, T), np.int32),
    )
    for i, e in enumerate(eps):
        L = e["length"]
        out["reset_draws"][i, :len(e["reset_draws"])] = e["reset_draws"]
        for k in ("actions", "order", "n_draws", "obs", "rewards", "terminated", "truncated",
                  "info", "pos", "collected"):
            out[k][i, :L] = e[k]
        for t, d in enumerate(e["draws"]):
            if len(d) > K:
                raise ValueError(f"episode {i} step {t}: {len(d)} draws > K={K}")
            out["draws"][i, t, :len(d)] = d
    return out


# ------------------------------------------------------------------ generator proxy hook
@contextmanager
def tapped_generators(unseeded_entropy=None):
    """Every gymnasium np_random generator created while installed records into ONE ordered log.  `unseeded_entropy`
    makes the generators the reference never seeds (the scripted policies') reproducible."""
    import gymnasium
    log = []
    gymnasium.Env.unseeded_entropy = unseeded_entropy

    def wrap(g):
        p = GeneratorProxy(g)
        p.events = log
        return p

    old = gymnasium.Env.wrap_generator
    gymnasium.Env.wrap_generator = staticmethod(wrap)
    try:
        yield log
    finally:
        gymnasium.Env.wrap_generator = old
        gymnasium.Env.unseeded_entropy = None


# ----------------------------------------------------------------------- Maze recording
def record_maze_episode(map_path, seed, action_rng, obstacle_penalty_ratio=0.0, max_steps=100, invalid_prob=0.0):
    """One episode of the reference MazeSingleAgentEnv (maze.py:26-377), "map" observations."""
    import_reference()
    from gym_multigrid.envs.maze import MazeSingleAgentEnv
    env = MazeSingleAgentEnv(map_path, max_steps=max_steps, obstacle_penalty_ratio=obstacle_penalty_ratio,
                             observation_option="map")
    np.random.seed(seed)
    with installed_taps() as taps:
        obs0, info0 = env.reset(seed=seed)
        start_index = taps.np_randint[0]          # maze.py:204 np.random.randint(0, len(background))
    rec = dict(actions=[], obs=[], reward=[], terminated=[], truncated=[], pos=[], dir=[], info=[])
    while True:
        a = int(action_rng.integers(0, 5))
        obs, rew, term, trunc, info = env.step(a)
        rec["actions"].append(a)
        rec["obs"].append(np.asarray(obs).copy())
        rec["reward"].append(float(rew))
        rec["terminated"].append(bool(term))
        rec["truncated"].append(bool(trunc))
        rec["pos"].append(np.asarray(env.agents[0].pos, dtype=np.int16).copy())
        rec["dir"].append(int(env.agents[0].dir))
        rec["info"].append([info["d_a_f"], info["d_a_ob"]])
        if term or trunc:
            break
    L = len(rec["actions"])
    return dict(field_map=np.asarray(env._field_map).copy(), init_obs=np.asarray(obs0).copy(), start_index=start_index,
                init_info=np.array([info0["d_a_f"], info0["d_a_ob"]]), length=L,
                actions=np.array(rec["actions"], np.int8), obs=np.stack(rec["obs"]),
                reward=np.array(rec["reward"], np.float64), terminated=np.array(rec["terminated"]),
                truncated=np.array(rec["truncated"]), pos=np.stack(rec["pos"]), dir=np.array(rec["dir"], np.int8),
                info=np.array(rec["info"], np.float64))


def pack_episodes(eps, keys_per_step, keys_static, T=None):
    """Generic packer: per-step arrays padded to T along axis 1, static arrays stacked."""
    T = T or max(e["length"] for e in eps)
    out = {k: np.stack([e[k] for e in eps]) for k in keys_static}
    out["length"] = np.array([e["length"] for e in eps], np.int32)
    for k in keys_per_step:
        first = np.asarray(eps[0][k])
        arr = np.zeros((len(eps), T) + first.shape[1:], first.dtype)
        for i, e in enumerate(eps):
            arr[i, :e["length"]] = e[k]
        out[k] = arr
    return out


# ------------------------------------------------------------------------ CtF recording
# key order of CtFMvNEnv._get_info / Ctf1v1Env._get_info (ctf.py:1165-1182, 434-452)
CTF_INFO_KEYS = ("d_ba_ra", "d_ba_bf", "d_ba_rf", "d_ra_bf", "d_ra_rf", "d_bf_rf", "d_ba_bb", "d_ba_rb", "d_ra_bb", "d_ra_rb", "d_ba_ob")

def record_ctf_mvn_episode(map_path, seed, action_rng, num_blue=2, num_red=2, obstacle_penalty_ratio=0.0,
                           max_steps=100, observation_option="map", max_battles=16, render_every=0, tile_sizes=(32, 8)):
    """One episode of the reference CtFMvNEnv (ctf.py:657-1433) on a FRESH instance (agent.terminated is
    never cleared by reset in the reference), with the ordered RNG event log split per step."""
    import_reference()
    with tapped_generators(unseeded_entropy=770000 + seed) as log:
        from gym_multigrid.envs.ctf import CtFMvNEnv
        from gym_multigrid.policy.ctf.heuristic import RwPolicy
        env = CtFMvNEnv(map_path, num_blue_agents=num_blue, num_red_agents=num_red, enemy_policies=RwPolicy(),
                        obstacle_penalty_ratio=obstacle_penalty_ratio, max_steps=max_steps,
                        observation_option=observation_option)
        obs0, info0 = env.reset(seed=seed)
        # the red policies keep the generator they were given at construction (ctf.py:820-826); after reset(seed)
        # the env owns a new one.  Both record into `log`, in call order.
        place = [np.asarray(ev[1]).copy() for ev in log if ev[0] == "choice"]
        del log[:]
        n = num_blue + num_red
        rec = dict(actions=[], red_actions=[], order=[], n_battles=[], blue_win=[], obs=[], reward=[], terminated=[],
                   truncated=[], pos=[], dir=[], dead=[], info=[], stats_flags=[], stats_defeated=[])
        init_pos = np.array([np.asarray(a.pos) for a in env.agents], np.int16)
        init_dir = np.array([a.dir for a in env.agents], np.int8)
        # render recording (render_every > 0): the agents' sticky background colour after every step (1 light_blue, 2 light_red;
        # agent.py:197-200, ctf.py:1214-1230) and rgb_array frames after the reset and after every `render_every`-th step
        BG = {"light_blue": 1, "light_red": 2}
        frames = {ts: [] for ts in tile_sizes}
        frame_step = []

        def grab(step_index):
            from gym_multigrid.core.grid import Grid
            frame_step.append(step_index)
            for ts in tile_sizes:
                Grid.tile_cache.clear()
                frames[ts].append(env.render(tile_size=ts).copy())
        if render_every:
            rec["bg"] = []
            init_bg = np.array([BG[a.bg_color] for a in env.agents], np.uint8)
            grab(-1)
        while True:
            acts = action_rng.integers(0, 5, size=num_blue)
            obs, rew, term, trunc, info = env.step([int(a) for a in acts])
            if render_every:
                rec["bg"].append(np.array([BG[a.bg_color] for a in env.agents], np.uint8))
                assert all(a.color == (("blue_grey" if i < num_blue else "red_grey") if a.terminated else ("blue" if i < num_blue else "red"))
                           for i, a in enumerate(env.agents))     # grey <=> terminated (ctf.py:1316-1332, 1409-1418)
                if (len(rec["bg"]) % render_every) == 0 or term or trunc:
                    grab(len(rec["bg"]) - 1)
            ints = [ev[1] for ev in log if ev[0] == "integers"]
            shuf = [ev[1] for ev in log if ev[0] == "shuffle"]
            wins = [bool(ev[1]) for ev in log if ev[0] == "choice"]
            assert len(ints) == num_red and len(shuf) == 1 and len(wins) <= max_battles
            del log[:]
            rec["actions"].append(acts.astype(np.int8))
            rec["red_actions"].append(np.array(ints, np.int8))
            rec["order"].append(np.array(shuf[0], np.uint8))
            rec["n_battles"].append(len(wins))
            rec["blue_win"].append(np.array(wins + [False] * (max_battles - len(wins)), np.uint8))
            rec["obs"].append(np.asarray(obs).copy())
            rec["reward"].append(float(rew))
            rec["terminated"].append(bool(term))
            rec["truncated"].append(bool(trunc))
            rec["pos"].append(np.array([np.asarray(a.pos) for a in env.agents], np.int16))
            rec["dir"].append(np.array([a.dir for a in env.agents], np.int8))
            rec["dead"].append(np.array([a.terminated for a in env.agents], np.uint8))
            rec["info"].append(np.array([info[k] for k in CTF_INFO_KEYS], np.float64))
            gs = env.game_stats   # ctf.py:1068-1073
            rec["stats_flags"].append(np.array([gs["blue_flag_captured"], gs["red_flag_captured"]], np.uint8))
            rec["stats_defeated"].append(np.array(list(gs["blue_agent_defeated"]) + list(gs["red_agent_defeated"]), np.uint8))
            if term or trunc:
                break
    L = len(rec["actions"])
    out = dict(field_map=np.asarray(env._field_map).copy(), init_obs=np.asarray(obs0).copy(), init_pos=init_pos,
               init_dir=init_dir, blue_place=place[0].astype(np.int32), red_place=place[1].astype(np.int32), length=L,
               init_info=np.array([info0[k] for k in CTF_INFO_KEYS], np.float64))
    for k, v in rec.items():
        out[k] = np.stack(v) if isinstance(v[0], np.ndarray) else np.array(v)
    out["n_battles"] = out["n_battles"].astype(np.int32)
    out["reward"] = out["reward"].astype(np.float64)
    if render_every:
        out["init_bg"] = init_bg
        out["frame_step"] = np.array(frame_step, np.int32)
        for ts in tile_sizes:
            out[f"frames_{ts}"] = np.stack(frames[ts])
    return out


def record_ctf_mvn_session(map_path, seed, action_rng, episodes, num_blue=2, num_red=2, obstacle_penalty_ratio=0.0, max_steps=100,
                           max_battles=16):
    """`episodes` consecutive episodes of ONE reference CtFMvNEnv instance (reset(seed) once, then plain reset() calls): the
    reference's reset never clears Agent.terminated / collided / bg_color (agent.py:97-100; multigrid.py:114-153, ctf.py:1050-1075;
    SURVEY 3.3), so from the second episode on agents defeated earlier START defeated.  Returns one dict per episode in
    record_ctf_mvn_episode's format plus `init_dead` / `init_collided` (the flags the episode started with)."""
    import_reference()
    out = []
    with tapped_generators(unseeded_entropy=880000 + seed) as log:
        from gym_multigrid.envs.ctf import CtFMvNEnv
        from gym_multigrid.policy.ctf.heuristic import RwPolicy
        env = CtFMvNEnv(map_path, num_blue_agents=num_blue, num_red_agents=num_red, enemy_policies=RwPolicy(),
                        obstacle_penalty_ratio=obstacle_penalty_ratio, max_steps=max_steps, observation_option="map")
        for ep in range(episodes):
            del log[:]
            obs0, info0 = env.reset(seed=seed) if ep == 0 else env.reset()
            place = [np.asarray(ev[1]).copy() for ev in log if ev[0] == "choice"]
            assert len(place) == 2
            del log[:]
            rec = dict(actions=[], red_actions=[], order=[], n_battles=[], blue_win=[], obs=[], reward=[], terminated=[],
                       truncated=[], pos=[], dir=[], dead=[], collided=[], info=[], stats_flags=[], stats_defeated=[])
            d = dict(field_map=np.asarray(env._field_map).copy(), init_obs=np.asarray(obs0).copy(),
                     init_pos=np.array([np.asarray(a.pos) for a in env.agents], np.int16),
                     init_dir=np.array([a.dir for a in env.agents], np.int8),
                     init_dead=np.array([a.terminated for a in env.agents], np.uint8),
                     init_collided=np.array([a.collided for a in env.agents], np.uint8),
                     blue_place=place[0].astype(np.int32), red_place=place[1].astype(np.int32),
                     init_info=np.array([info0[k] for k in CTF_INFO_KEYS], np.float64))
            while True:
                acts = action_rng.integers(0, 5, size=num_blue)
                obs, rew, term, trunc, info = env.step([int(a) for a in acts])
                ints = [ev[1] for ev in log if ev[0] == "integers"]
                shuf = [ev[1] for ev in log if ev[0] == "shuffle"]
                wins = [bool(ev[1]) for ev in log if ev[0] == "choice"]
                assert len(ints) == num_red and len(shuf) == 1 and len(wins) <= max_battles
                del log[:]
                rec["actions"].append(acts.astype(np.int8))
                rec["red_actions"].append(np.array(ints, np.int8))
                rec["order"].append(np.array(shuf[0], np.uint8))
                rec["n_battles"].append(len(wins))
                rec["blue_win"].append(np.array(wins + [False] * (max_battles - len(wins)), np.uint8))
                rec["obs"].append(np.asarray(obs).copy())
                rec["reward"].append(float(rew))
                rec["terminated"].append(bool(term))
                rec["truncated"].append(bool(trunc))
                rec["pos"].append(np.array([np.asarray(a.pos) for a in env.agents], np.int16))
                rec["dir"].append(np.array([a.dir for a in env.agents], np.int8))
                rec["dead"].append(np.array([a.terminated for a in env.agents], np.uint8))
                rec["collided"].append(np.array([a.collided for a in env.agents], np.uint8))
                rec["info"].append(np.array([info[k] for k in CTF_INFO_KEYS], np.float64))
                gs = env.game_stats   # ctf.py:1068-1073
                rec["stats_flags"].append(np.array([gs["blue_flag_captured"], gs["red_flag_captured"]], np.uint8))
                rec["stats_defeated"].append(np.array(list(gs["blue_agent_defeated"]) + list(gs["red_agent_defeated"]), np.uint8))
                if term or trunc:
                    break
            d["length"] = len(rec["actions"])
            for k, v in rec.items():
                d[k] = np.stack(v) if isinstance(v[0], np.ndarray) else np.array(v)
            d["n_battles"] = d["n_battles"].astype(np.int32)
            d["reward"] = d["reward"].astype(np.float64)
            out.append(d)
    return out


# ------------------------------------------------------------------- partial-view recording
def record_partial_views(env_id, seed, n_samples, view_sizes=(3, 5, 7)):
    """Reference partial observations (MultiGridEnv.gen_obs_grid multigrid.py:485-515 + Grid.encode_for_agents
    grid.py:254-284) on states of a Collect env.  `gen_obs` itself passes one positional argument too many
    (multigrid.py:526-528 vs grid.py:254) and raises TypeError, so its two steps are called directly - the
    algorithm, not the crash.  Agent directions are set by hand (Collect never turns its agents)."""
    env, tl = make_collect(env_id)
    random.seed(seed); np.random.seed(seed)
    rng = np.random.default_rng(seed)
    env.reset(seed=seed)
    A = len(env.agents)
    out = dict(grid_obs=[], pos=[], dirs=[], V=[], see_through=[], views=[])
    for s in range(n_samples):
        for _ in range(int(rng.integers(1, 6))):
            _, _, term, trunc, _ = env.step([int(a) for a in rng.integers(0, 4, size=A)])
            if term or trunc or env.step_count >= 45:
                env.reset(seed=seed + s)
        V = int(rng.choice(view_sizes))
        st = bool(rng.integers(0, 2))
        for a in env.agents:
            a.dir = int(rng.integers(0, 4))
            a.view_size = V
        env.see_through_walls = st
        grids, masks = env.gen_obs_grid()
        views = [g.encode_for_agents(agent_pos=(V // 2, V - 1), vis_mask=m) for g, m in zip(grids, masks)]
        pad = np.zeros((A, max(view_sizes), max(view_sizes), 3), np.uint8)
        for k, v in enumerate(views):
            pad[k, :V, :V] = v
        out["grid_obs"].append(env.grid.encode().copy())
        out["pos"].append(np.array([np.asarray(a.pos) for a in env.agents], np.int16))
        out["dirs"].append(np.array([a.dir for a in env.agents], np.int8))
        out["V"].append(V); out["see_through"].append(st); out["views"].append(pad)
        for a in env.agents:   # Collect expects dir 3 forever
            a.dir = 3
    return {k: np.array(v) for k, v in out.items()}


# --------------------------------------------------------------------- rgb_array render recording
def record_render(env_id, seed, n_samples, tile_sizes=(32, 8)):
    """Reference `MultiGridEnv.render()` frames (multigrid.py:546-606 -> Grid.render grid.py:183-221 -> Grid.render_tile
    :132-181 -> utils/rendering.py) on states of a Collect env, highlight off (the default).  Agent directions are set by
    hand so that all four rotations of the agent triangle are covered (Collect itself never turns its agents)."""
    import_reference()
    from gym_multigrid.core.grid import Grid
    env, tl = make_collect(env_id)
    random.seed(seed); np.random.seed(seed)
    rng = np.random.default_rng(seed)
    env.reset(seed=seed)
    A = len(env.agents)
    out = dict(grid_obs=[], tile_size=[], frame=[], pos=[])
    for s in range(n_samples):
        for _ in range(int(rng.integers(1, 6))):
            _, _, term, trunc, _ = env.step([int(a) for a in rng.integers(0, 4, size=A)])
            if term or trunc or env.step_count >= 45:
                env.reset(seed=seed + s)
        for a in env.agents:
            a.dir = int(rng.integers(0, 4))
        for ts in tile_sizes:
            Grid.tile_cache.clear()            # every frame is rendered from scratch: no tile of an earlier size / test leaks in
            frame = env.render(tile_size=ts)
            assert frame.dtype == np.uint8 and frame.shape == (env.height * ts, env.width * ts, 3)
            out["grid_obs"].append(env.grid.encode().copy())
            out["tile_size"].append(ts)
            out["pos"].append(np.array([np.asarray(a.pos) for a in env.agents], np.int16))
            out["frame"].append(frame.copy())
        for a in env.agents:
            a.dir = 3
    r = {k: (np.array(v) if k != "frame" else v) for k, v in out.items()}
    # frames of the two tile sizes have different shapes: one array per size
    for ts in tile_sizes:
        idx = [i for i, t in enumerate(out["tile_size"]) if t == ts]
        r[f"frames_{ts}"] = np.stack([out["frame"][i] for i in idx])
        r[f"grid_obs_{ts}"] = np.stack([out["grid_obs"][i] for i in idx])
        r[f"pos_{ts}"] = np.stack([out["pos"][i] for i in idx])
    del r["frame"], r["grid_obs"], r["tile_size"], r["pos"]
    return r


def record_maze_render(map_path, seed, n_samples, tile_sizes=(32, 8)):
    """Reference `render()` frames of MazeSingleAgentEnv (maze.py:26-377; white Floor, grey Obstacle, red Flag on white, blue
    agent triangle on white) along a random walk; the state needed to redraw them: field_map, agent pos, agent dir."""
    import_reference()
    from gym_multigrid.core.grid import Grid
    from gym_multigrid.envs.maze import MazeSingleAgentEnv
    env = MazeSingleAgentEnv(map_path, max_steps=1000, observation_option="map")
    np.random.seed(seed)
    rng = np.random.default_rng(seed)
    env.reset(seed=seed)
    out = dict(pos=[], dir=[])
    frames = {ts: [] for ts in tile_sizes}
    for s in range(n_samples):
        for _ in range(int(rng.integers(0 if s == 0 else 1, 7))):   # sample 0 is the reset state (dir 3)
            _, _, term, trunc, _ = env.step(int(rng.integers(0, 5)))
            if term or trunc:
                env.reset(seed=seed + s)
        out["pos"].append(np.asarray(env.agents[0].pos, dtype=np.int16).copy())
        out["dir"].append(int(env.agents[0].dir))
        for ts in tile_sizes:
            Grid.tile_cache.clear()
            f = env.render(tile_size=ts)
            assert f.dtype == np.uint8 and f.shape == (env.height * ts, env.width * ts, 3)
            frames[ts].append(f.copy())
    r = dict(field_map=np.asarray(env._field_map).astype(np.uint8), pos=np.stack(out["pos"]), dir=np.array(out["dir"], np.int8))
    for ts in tile_sizes:
        r[f"frames_{ts}"] = np.stack(frames[ts])
    return r


# --------------------------------------------------------------------- Toroid wrapper recording
def record_toroid(env_id, seed, n_samples):
    """Reference ToroidObservation (wrappers/toroid.py:6-68) outputs on states of a Collect env."""
    import_reference()
    from gym_multigrid.wrappers.toroid import ToroidObservation
    env, tl = make_collect(env_id)
    random.seed(seed); np.random.seed(seed)
    rng = np.random.default_rng(seed)
    wrapped = ToroidObservation(env)
    env.reset(seed=seed)
    A = len(env.agents)
    out = dict(grid_obs=[], pos=[], toroid=[])
    for s in range(n_samples):
        for _ in range(int(rng.integers(1, 5))):
            _, _, term, trunc, _ = env.step([int(a) for a in rng.integers(0, 4, size=A)])
            if term or trunc or env.step_count >= 45:
                env.reset(seed=seed + s)
        tor = wrapped.observation(env.grid.encode())
        assert all(t.dtype == np.float32 for t in tor)
        out["grid_obs"].append(env.grid.encode().copy())
        out["pos"].append(np.array([np.asarray(a.pos) for a in env.agents], np.int16))
        out["toroid"].append(np.stack(tor))
    return {k: np.array(v) for k, v in out.items()}


def record_ctf_1v1_episode(map_path, seed, action_rng, max_steps=100, max_battles=4, observation_option="map"):
    """One episode of the reference Ctf1v1Env (ctf.py:50-654), "map" (or "flattened") observations, fresh instance."""
    import_reference()
    with tapped_generators(unseeded_entropy=880000 + seed) as log:
        from gym_multigrid.envs.ctf import Ctf1v1Env
        from gym_multigrid.policy.ctf.heuristic import RwPolicy
        env = Ctf1v1Env(map_path, enemy_policy=RwPolicy(), max_steps=max_steps, observation_option=observation_option)
        obs0, info0 = env.reset(seed=seed)
        place = [ev[1] for ev in log if ev[0] == "integers"]      # ctf.py:317, :322
        assert len(place) == 2
        del log[:]
        rec = dict(actions=[], red_actions=[], n_battles=[], blue_win=[], obs=[], reward=[], terminated=[], truncated=[],
                   pos=[], dir=[], dead=[], info=[], stats_flags=[], stats_defeated=[])
        init_pos = np.array([np.asarray(a.pos) for a in env.agents], np.int16)
        while True:
            a = int(action_rng.integers(0, 5))
            obs, rew, term, trunc, info = env.step(a)
            ints = [ev[1] for ev in log if ev[0] == "integers"]
            wins = [bool(ev[1]) for ev in log if ev[0] == "choice"]
            assert len(ints) == 1 and len(wins) <= 1 and not [ev for ev in log if ev[0] == "shuffle"]
            del log[:]
            rec["actions"].append(np.array([a], np.int8))
            rec["red_actions"].append(np.array(ints, np.int8))
            rec["n_battles"].append(len(wins))
            rec["blue_win"].append(np.array(wins + [False] * (max_battles - len(wins)), np.uint8))
            rec["obs"].append(np.asarray(obs).copy())
            rec["reward"].append(float(rew))
            rec["terminated"].append(bool(term))
            rec["truncated"].append(bool(trunc))
            rec["pos"].append(np.array([np.asarray(a_.pos) for a_ in env.agents], np.int16))
            rec["dir"].append(np.array([a_.dir for a_ in env.agents], np.int8))
            rec["dead"].append(np.array([0, int(env._is_red_agent_defeated)], np.uint8))
            rec["info"].append(np.array([info[k] for k in CTF_INFO_KEYS], np.float64))
            gs = env.game_stats   # ctf.py:340-345
            rec["stats_flags"].append(np.array([gs["blue_flag_captured"], gs["red_flag_captured"]], np.uint8))
            rec["stats_defeated"].append(np.array(list(gs["blue_agent_defeated"]) + list(gs["red_agent_defeated"]), np.uint8))
            if term or trunc:
                break
    L = len(rec["actions"])
    out = dict(field_map=np.asarray(env._field_map).copy(), init_obs=np.asarray(obs0).copy(), init_pos=init_pos,
               blue_place=np.array([place[0]], np.int32), red_place=np.array([place[1]], np.int32), length=L,
               init_info=np.array([info0[k] for k in CTF_INFO_KEYS], np.float64))
    for k, v in rec.items():
        out[k] = np.stack(v) if isinstance(v[0], np.ndarray) else np.array(v)
    out["n_battles"] = out["n_battles"].astype(np.int32)
    out["reward"] = out["reward"].astype(np.float64)
    return out


# ------------------------------------------------------------------ generic MultiGridEnv.step
def make_generic_env(size, n_agents, max_steps, layout_seed):
    """A DefaultWorld env built ONLY from reference classes, to exercise the base-class step
    (multigrid.py:397-483) and encode_dim-6 `encode_for_agents` (grid.py:254-284) that no shipped env reaches:
    border walls plus a random sprinkle of every DefaultWorld object type, SmallActions-compatible actions."""
    import_reference()
    from gym_multigrid.multigrid import MultiGridEnv
    from gym_multigrid.core.agent import Agent, DefaultActions
    from gym_multigrid.core.grid import Grid
    from gym_multigrid.core.object import Ball, Box, Door, Floor, Goal, Key, Lava, ObjectGoal, Switch, Wall
    from gym_multigrid.core.world import DefaultWorld

    class GenericEnv(MultiGridEnv):
        def __init__(self):
            agents = [Agent(DefaultWorld, i, actions=DefaultActions) for i in range(n_agents)]
            super().__init__(agents=agents, grid_size=size, max_steps=max_steps, world=DefaultWorld,
                             actions_set=DefaultActions, partial_obs=False)

        def _gen_grid(self, width, height):
            self.grid = Grid(width, height, self.world)
            self.grid.wall_rect(0, 0, width, height)
            w = self.world
            rs = random.Random(layout_seed)
            objs = [Goal(w, 1), Goal(w, 3), Switch(w), Floor(w, "blue"), Lava(w), Door(w, "yellow"), Door(w, "green", is_open=True),
                    Door(w, "purple", is_locked=True), Key(w, "yellow"), Ball(w, 2, 1), Box(w, "orange"), ObjectGoal(w, 4), Wall(w),
                    Goal(w, 0), Floor(w, "red"), Lava(w)]
            for o in objs:
                while True:
                    x, y = rs.randint(1, width - 2), rs.randint(1, height - 2)
                    if self.grid.get(x, y) is None:
                        self.put_obj(o, x, y)
                        break
            for a in self.agents:
                while True:
                    x, y = rs.randint(1, width - 2), rs.randint(1, height - 2)
                    if self.grid.get(x, y) is None:
                        self.place_agent(a, pos=(x, y))
                        a.dir = rs.randint(0, 3)
                        break

    return GenericEnv()


def record_generic_episode(size, n_agents, max_steps, seed, action_rng):
    env = make_generic_env(size, n_agents, max_steps, seed)
    np.random.seed(seed)
    with installed_taps() as taps:
        obs0, _ = env.reset(seed=seed)
        rec = dict(actions=[], order=[], obs=[], rewards=[], terminated=[], truncated=[], pos=[], dir=[])
        init_pos = np.array([np.asarray(a.pos) for a in env.agents], np.int16)
        init_dir = np.array([a.dir for a in env.agents], np.int8)
        taps.clear()
        while True:
            acts = action_rng.choice(4, size=n_agents, p=[0.1, 0.2, 0.2, 0.5])      # still / left / right / forward
            obs, rew, term, trunc, _ = env.step([int(a) for a in acts])
            rec["actions"].append(acts.astype(np.int8))
            rec["order"].append(np.asarray(taps.perm[0], np.uint8))
            taps.clear()
            rec["obs"].append(np.stack(obs).astype(np.uint8))
            rec["rewards"].append(np.asarray(rew, np.float64).copy())
            rec["terminated"].append(bool(term)); rec["truncated"].append(bool(trunc))
            rec["pos"].append(np.array([np.asarray(a.pos) for a in env.agents], np.int16))
            rec["dir"].append(np.array([a.dir for a in env.agents], np.int8))
            if term or trunc:
                break
    out = dict(init_obs=np.stack(obs0).astype(np.uint8), init_pos=init_pos, init_dir=init_dir, length=len(rec["actions"]))
    for k, v in rec.items():
        out[k] = np.stack(v) if isinstance(v[0], np.ndarray) else np.array(v)
    return out


def record_generic_partial(size, n_agents, seed, n_samples, view_sizes=(3, 5, 7)):
    """Reference partial observations with encode_dim 6 (DefaultWorld): gen_obs_grid (multigrid.py:485-515) +
    encode_for_agents (grid.py:254-284) on states of the DefaultWorld env of `make_generic_env` (doors open / closed /
    locked, walls, every other object type).  As in record_partial_views the two steps of `gen_obs` are called directly
    (its own call passes one argument too many, multigrid.py:526-528)."""
    env = make_generic_env(size, n_agents, 10 ** 6, seed)
    np.random.seed(seed)
    rng = np.random.default_rng(seed)
    env.reset(seed=seed)
    A = n_agents
    out = dict(obs6=[], pos=[], V=[], see_through=[], views=[])
    for s in range(n_samples):
        for _ in range(int(rng.integers(1, 6))):
            _, _, term, _, _ = env.step([int(a) for a in rng.choice(4, size=A, p=[0.1, 0.25, 0.25, 0.4])])
            if term:
                env.reset(seed=seed + s)
        V = int(rng.choice(view_sizes))
        st = bool(rng.integers(0, 2))
        for a in env.agents:
            a.view_size = V
        env.see_through_walls = st
        grids, masks = env.gen_obs_grid()
        views = [g.encode_for_agents(agent_pos=(V // 2, V - 1), vis_mask=m) for g, m in zip(grids, masks)]
        pad = np.zeros((A, max(view_sizes), max(view_sizes), 6), np.uint8)
        for k, v in enumerate(views):
            pad[k, :V, :V] = v
        out["obs6"].append(env.grid.encode_for_agents(agent_pos=env.agents[0].pos).copy())
        out["pos"].append(np.array([np.asarray(a.pos) for a in env.agents], np.int16))
        out["V"].append(V); out["see_through"].append(st); out["views"].append(pad)
    return {k: np.array(v) for k, v in out.items()}
