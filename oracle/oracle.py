"""ctypes binding of the CPU oracle (oracle/mg_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libmg_oracle.so")

MAX_AGENTS = 8
MAX_BALL_TYPES = 8
LAYOUTS = {"even_dist": 0, "quadrants": 1, "rooms": 2, "quadrants_respawn": 3}

ERR_TRACE_OVERFLOW, ERR_TRACE_RANGE, ERR_OOB = 1, 2, 4


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc, seconds)."""
    if force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB_PATH)
        for f in os.listdir(_HERE) if f.endswith((".c", ".h"))
    ):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _LIB_PATH


class CollectCfg(C.Structure):
    _fields_ = [
        ("width", C.c_int32), ("height", C.c_int32), ("num_agents", C.c_int32),
        ("num_ball_types", C.c_int32), ("agent_colour", C.c_int32 * MAX_AGENTS),
        ("ball_colour", C.c_int32 * MAX_BALL_TYPES), ("ball_reward", C.c_double * MAX_BALL_TYPES),
        ("num_balls", C.c_int32), ("respawn", C.c_int32), ("layout", C.c_int32),
        ("fixed_horizon", C.c_int32), ("max_steps", C.c_int32), ("time_limit", C.c_int32),
    ]


class CollectState(C.Structure):
    _fields_ = [("grid", C.c_void_p), ("agent_pos", C.c_void_p), ("step_count", C.c_void_p),
                ("collected", C.c_void_p), ("info", C.c_void_p), ("rng_ctr", C.c_void_p)]


class RngSrc(C.Structure):
    _fields_ = [("mode", C.c_int32), ("order", C.c_void_p), ("draws", C.c_void_p), ("n_draws", C.c_void_p),
                ("K", C.c_int32), ("draws_used", C.c_void_p), ("seed", C.c_uint64), ("env_id_base", C.c_uint64)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.oc_collect_reset.restype = C.c_int
        _lib.oc_collect_step.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def make_collect_cfg(size, agents_index, balls_index, balls_reward, num_balls, respawn, layout,
                     fixed_horizon=False, max_steps=100, time_limit=0, width=None, height=None) -> CollectCfg:
    c = CollectCfg()
    c.width = width or size
    c.height = height or size
    c.num_agents = len(agents_index)
    c.num_ball_types = len(balls_index)
    for i, v in enumerate(agents_index):
        c.agent_colour[i] = v
    for i, v in enumerate(balls_index):
        c.ball_colour[i] = v
    for i, v in enumerate(balls_reward):
        c.ball_reward[i] = float(v)
    c.num_balls = int(np.sum(np.array(num_balls)))
    c.respawn = int(bool(respawn))
    c.layout = LAYOUTS[layout] if isinstance(layout, str) else int(layout)
    c.fixed_horizon = int(bool(fixed_horizon))
    c.max_steps = max_steps
    c.time_limit = int(time_limit or 0)
    return c


class TraceRng:
    """Replay source: recorded np.random.permutation outputs + random.randint outputs."""

    def __init__(self, order=None, draws=None, n_draws=None):
        self.order = None if order is None else np.ascontiguousarray(order, np.uint8)
        self.draws = None if draws is None else np.ascontiguousarray(draws, np.uint8)
        self.n_draws = None if n_draws is None else np.ascontiguousarray(n_draws, np.int32)
        N = (self.draws if self.draws is not None else self.order).shape[0]
        self.draws_used = np.zeros(N, np.int32)

    def struct(self):
        s = RngSrc()
        s.mode = 0
        s.order, s.draws, s.n_draws = _p(self.order), _p(self.draws), _p(self.n_draws)
        s.K = 0 if self.draws is None else self.draws.shape[1]
        s.draws_used = _p(self.draws_used)
        return s


class PhiloxRng:
    def __init__(self, seed, env_id_base=0):
        self.seed, self.env_id_base = int(seed), int(env_id_base)

    def struct(self):
        s = RngSrc()
        s.mode = 1
        s.seed, s.env_id_base = self.seed, self.env_id_base
        return s


class CollectOracle:
    """Batched CPU Collect envs with the same state planes as the CUDA library."""

    def __init__(self, cfg: CollectCfg, num_envs: int, nthreads: int = 1):
        self.cfg, self.N, self.nthreads = cfg, int(num_envs), nthreads
        W, H, A, nb = cfg.width, cfg.height, cfg.num_agents, cfg.num_ball_types
        self.W, self.H, self.A, self.nb = W, H, A, nb
        N = self.N
        self.grid = np.zeros((N, W * H), np.uint8)
        self.agent_pos = np.zeros((N, A, 2), np.uint8)
        self.step_count = np.zeros(N, np.int32)
        self.collected = np.zeros(N, np.int32)
        self.info = np.zeros((N, A * nb), np.int32)
        self.rng_ctr = np.zeros(N, np.uint32)
        self.status = C.c_int32(0)

    def _state(self):
        s = CollectState()
        s.grid, s.agent_pos, s.step_count = _p(self.grid), _p(self.agent_pos), _p(self.step_count)
        s.collected, s.info, s.rng_ctr = _p(self.collected), _p(self.info), _p(self.rng_ctr)
        return s

    def reset(self, rng, mask=None, want_obs=True):
        obs = np.empty((self.N, self.W, self.H, 3), np.uint8) if want_obs else None
        if mask is not None:
            mask = np.ascontiguousarray(mask, np.uint8)
            if obs is not None:
                obs[:] = self.encode()
        st, rs = self._state(), rng.struct()
        rc = lib().oc_collect_reset(C.byref(self.cfg), C.c_int64(self.N), C.byref(st), _p(mask), C.byref(rs),
                                    _p(obs), C.byref(self.status), C.c_int(self.nthreads))
        if rc:
            raise RuntimeError("oc_collect_reset failed (invalid layout/config)")
        return obs

    def step(self, actions, rng, autoreset=False, reset_rng=None, want_final_obs=False, want_obs=True, reuse_buffers=False):
        actions = np.ascontiguousarray(actions, np.int8).reshape(self.N, self.A)
        if reuse_buffers:  # benchmark mode: no per-step allocation / page faults
            if not hasattr(self, "_out"):
                self._out = (np.zeros((self.N, self.W, self.H, 3), np.uint8), np.zeros((self.N, self.A), np.float64),
                             np.zeros(self.N, np.uint8), np.zeros(self.N, np.uint8))
            obs, rew, term, trunc = self._out
            obs = obs if want_obs else None
        else:
            obs = np.empty((self.N, self.W, self.H, 3), np.uint8) if want_obs else None
            rew = np.empty((self.N, self.A), np.float64)
            term = np.empty(self.N, np.uint8)
            trunc = np.empty(self.N, np.uint8)
        fin = np.zeros((self.N, self.W, self.H, 3), np.uint8) if want_final_obs else None
        st, rs = self._state(), rng.struct()
        rrs = reset_rng.struct() if reset_rng is not None else RngSrc()
        rc = lib().oc_collect_step(C.byref(self.cfg), C.c_int64(self.N), C.byref(st), _p(actions), C.byref(rs),
                                   _p(obs), _p(rew), _p(term), _p(trunc), C.c_int(int(autoreset)), C.byref(rrs),
                                   _p(fin), C.byref(self.status), C.c_int(self.nthreads))
        if rc:
            raise RuntimeError("oc_collect_step failed")
        out = (obs, rew, term.astype(bool), trunc.astype(bool))
        return out + (fin,) if want_final_obs else out

    def encode(self):
        return encode3(self.grid).reshape(self.N, self.W, self.H, 3)

    def set_state_from_obs(self, obs, agent_pos, step_count=0):
        """Inject a state given Grid.encode() arrays [N,W,H,3] (e.g. a reference reset)."""
        obs = np.asarray(obs, np.uint8).reshape(self.N, self.W * self.H, 3)
        self.grid[:] = obs[..., 0] | (obs[..., 1] << 2) | (obs[..., 2] << 6)
        self.agent_pos[:] = np.asarray(agent_pos).reshape(self.N, self.A, 2)
        self.step_count[:] = step_count
        self.collected[:] = 0
        self.info[:] = 0


def encode3(cells: np.ndarray) -> np.ndarray:
    cells = np.ascontiguousarray(cells, np.uint8)
    out = np.empty(cells.shape + (3,), np.uint8)
    lib().oc_encode3(_p(cells), C.c_int64(cells.size), _p(out))
    return out


def render_grid(obs: np.ndarray, tile_size: int = 32) -> np.ndarray:
    """Grid.render(tile_size) of Grid.encode() observations u8 [N, W, H, 3] -> frames u8 [N, H*ts, W*ts, 3]."""
    obs = np.ascontiguousarray(obs, np.uint8)
    N, W, H, _ = obs.shape
    out = np.empty((N, H * tile_size, W * tile_size, 3), np.uint8)
    rc = lib().oc_render_grid(_p(obs), C.c_int64(N), C.c_int(W), C.c_int(H), C.c_int(tile_size), _p(out))
    assert rc == 0, "oc_render_grid: cell outside the Collect world"
    return out


def render_maze(field_map: np.ndarray, pos: np.ndarray, dirs: np.ndarray, tile_size: int = 32) -> np.ndarray:
    """MazeSingleAgentEnv.render(tile_size) for agents at pos [N, 2] facing dirs [N] on one map -> u8 [N, S*ts, S*ts, 3]."""
    fm = np.ascontiguousarray(field_map, np.uint8)
    pos = np.ascontiguousarray(pos, np.int16); dirs = np.ascontiguousarray(dirs, np.int8)
    S, N = fm.shape[0], pos.shape[0]
    out = np.empty((N, S * tile_size, S * tile_size, 3), np.uint8)
    rc = lib().oc_render_maze(_p(fm), C.c_int(S), C.c_int64(N), _p(pos), _p(dirs), C.c_int(tile_size), _p(out))
    assert rc == 0, "oc_render_maze: cell outside the Maze world"
    return out


def render_ctf(field_map, pos, dirs, flags, num_blue, tile_size=32, variant_1v1=False) -> np.ndarray:
    """CtFMvNEnv.render(tile_size): agents at pos [N, n, 2] facing dirs [N, n] with oracle flags [N, n] -> u8 [N, S*ts, S*ts, 3]."""
    fm = np.ascontiguousarray(field_map, np.uint8)
    pos = np.ascontiguousarray(pos, np.uint8); dirs = np.ascontiguousarray(dirs, np.uint8); flags = np.ascontiguousarray(flags, np.uint8)
    S, (N, n) = fm.shape[0], dirs.shape
    out = np.empty((N, S * tile_size, S * tile_size, 3), np.uint8)
    rc = lib().oc_render_ctf(_p(fm), C.c_int(S), C.c_int64(N), C.c_int(n), C.c_int(num_blue), C.c_int(int(variant_1v1)), _p(pos), _p(dirs),
                             _p(flags), C.c_int(tile_size), _p(out))
    assert rc == 0, "oc_render_ctf: cell outside the CtF world"
    return out


def philox4x32_10(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().oc_philox4x32_10(c, k, o)
    return [int(v) for v in o]


# =========================================================================== Maze / CtF
class MapCfg(C.Structure):
    _fields_ = [("size", C.c_int32), ("field_map", C.c_void_p), ("flag_reward", C.c_double),
                ("obstacle_penalty", C.c_double), ("step_penalty", C.c_double), ("max_steps", C.c_int32),
                ("num_blue", C.c_int32), ("num_red", C.c_int32), ("battle_range", C.c_double),
                ("randomness", C.c_double), ("battle_reward", C.c_double), ("variant_1v1", C.c_int32),
                ("carry_agent_flags", C.c_int32)]


class MapState(C.Structure):
    _fields_ = [("pos", C.c_void_p), ("dir", C.c_void_p), ("flags", C.c_void_p), ("step_count", C.c_void_p),
                ("rng_ctr", C.c_void_p), ("stats", C.c_void_p)]


class MapRng(C.Structure):
    _fields_ = [("mode", C.c_int32), ("start_index", C.c_void_p), ("blue_place", C.c_void_p), ("red_place", C.c_void_p),
                ("red_actions", C.c_void_p), ("order", C.c_void_p), ("blue_win", C.c_void_p), ("KB", C.c_int32),
                ("battles_used", C.c_void_p), ("seed", C.c_uint64), ("env_id_base", C.c_uint64)]


def map_rng(mode=1, seed=0, env_id_base=0, **arrays):
    """Build a MapRng; keeps the numpy arrays alive on the returned struct."""
    r = MapRng()
    r.mode, r.seed, r.env_id_base = mode, int(seed), int(env_id_base)
    keep = {}
    dt = dict(start_index=np.int32, blue_place=np.int32, red_place=np.int32, red_actions=np.int8, order=np.uint8,
              blue_win=np.uint8, battles_used=np.int32)
    for k, v in arrays.items():
        if v is None:
            continue
        a = np.ascontiguousarray(v, dt[k])
        keep[k] = a
        setattr(r, k, a.ctypes.data)
        if k == "blue_win":
            r.KB = a.shape[1]
    r._keep = keep
    return r


class _MapOracle:
    def __init__(self, field_map, num_envs, n_agents, **cfg):
        self.fm = np.ascontiguousarray(np.asarray(field_map).astype(np.uint8))
        assert self.fm.ndim == 2 and self.fm.shape[0] == self.fm.shape[1], "square maps only"
        self.S, self.N, self.n = self.fm.shape[0], int(num_envs), n_agents
        c = MapCfg()
        c.size, c.field_map = self.S, self.fm.ctypes.data
        for k, v in cfg.items():
            setattr(c, k, v)
        self.cfg = c
        self.pos = np.zeros((self.N, n_agents, 2), np.uint8)
        self.dir = np.zeros((self.N, n_agents), np.uint8)
        self.flags = np.zeros((self.N, n_agents), np.uint8)
        self.step_count = np.zeros(self.N, np.int32)
        self.rng_ctr = np.zeros(self.N, np.uint32)
        self.stats = np.zeros(self.N, np.int32)    # CtF game_stats bits (mg_oracle.h)
        self.status = C.c_int32(0)

    def _state(self):
        s = MapState()
        s.pos, s.dir, s.flags = _p(self.pos), _p(self.dir), _p(self.flags)
        s.step_count, s.rng_ctr, s.stats = _p(self.step_count), _p(self.rng_ctr), _p(self.stats)
        return s

    def game_stats(self):
        """(flags [N, 2] = blue_flag_captured, red_flag_captured; defeated [N, n]) as the reference's game_stats dict holds them."""
        st = self.stats
        return np.stack([st & 1, (st >> 1) & 1], 1).astype(np.uint8), ((st[:, None] >> (8 + np.arange(self.n))) & 1).astype(np.uint8)

    def info(self):
        """`_get_info()` of every env: float64 [N, 2] (Maze) or [N, 11] (CtF), columns in the reference dict's key order."""
        maze = isinstance(self, MazeOracle)
        out = np.zeros((self.N, 2 if maze else 11), np.float64)
        st = self._state()
        lib().oc_map_info(C.byref(self.cfg), C.c_int(int(maze)), C.c_int64(self.N), C.byref(st), _p(out))
        return out

    def _call_reset(self, fn, rng, mask):
        obs = np.zeros((self.N, self.S, self.S), np.uint8)
        st = self._state()
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        fn(C.byref(self.cfg), C.c_int64(self.N), C.byref(st), _p(m), C.byref(rng), _p(obs), C.byref(self.status))
        return obs

    def _call_step(self, fn, actions, rng, autoreset, want_final_obs):
        actions = np.ascontiguousarray(actions, np.int8)
        obs = np.zeros((self.N, self.S, self.S), np.uint8)
        rew = np.zeros(self.N, np.float64)
        term, trunc = np.zeros(self.N, np.uint8), np.zeros(self.N, np.uint8)
        fin = np.zeros_like(obs) if want_final_obs else None
        st = self._state()
        fn(C.byref(self.cfg), C.c_int64(self.N), C.byref(st), _p(actions), C.byref(rng), _p(obs), _p(rew), _p(term),
           _p(trunc), C.c_int(int(autoreset)), _p(fin), C.byref(self.status))
        out = (obs, rew, term.astype(bool), trunc.astype(bool))
        return out + (fin,) if want_final_obs else out


MAZE_INFO_KEYS = ("d_a_f", "d_a_ob")                                                       # maze.py:262-269
CTF_INFO_KEYS = ("d_ba_ra", "d_ba_bf", "d_ba_rf", "d_ra_bf", "d_ra_rf", "d_bf_rf", "d_ba_bb", "d_ba_rb", "d_ra_bb", "d_ra_rb",
                 "d_ba_ob")                                                                 # ctf.py:1165-1182


class MazeOracle(_MapOracle):
    """MazeSingleAgentEnv (maze.py) batched on the CPU; obs = `_encode_map()` values as uint8 [N, W, H]."""

    def __init__(self, field_map, num_envs, flag_reward=1.0, obstacle_penalty_ratio=0.0, step_penalty_ratio=0.01, max_steps=100):
        super().__init__(field_map, num_envs, 1, flag_reward=float(flag_reward),
                         obstacle_penalty=float(flag_reward) * float(obstacle_penalty_ratio),   # maze.py:349
                         step_penalty=float(flag_reward) * float(step_penalty_ratio), max_steps=max_steps)  # :350

    def reset(self, rng, mask=None):
        return self._call_reset(lib().oc_maze_reset, rng, mask)

    def step(self, actions, rng, autoreset=False, want_final_obs=False):
        return self._call_step(lib().oc_maze_step, np.asarray(actions).reshape(self.N), rng, autoreset, want_final_obs)


class CtfOracle(_MapOracle):
    """CtFMvNEnv (ctf.py:657-1433) batched on the CPU; obs = `_encode_map()` (transposed) as uint8 [N, H, W]."""

    def __init__(self, field_map, num_envs, num_blue=2, num_red=2, battle_range=1.0, randomness=0.75, flag_reward=1.0,
                 battle_reward_ratio=0.25, obstacle_penalty_ratio=0.0, step_penalty_ratio=0.01, max_steps=100, variant_1v1=False,
                 carry_agent_flags=False):
        fr = float(flag_reward)
        super().__init__(field_map, num_envs, num_blue + num_red, flag_reward=fr, battle_reward=float(battle_reward_ratio) * fr,
                         obstacle_penalty=float(obstacle_penalty_ratio) * fr, step_penalty=float(step_penalty_ratio) * fr,
                         max_steps=max_steps, num_blue=num_blue, num_red=num_red, battle_range=float(battle_range),
                         randomness=float(randomness), variant_1v1=int(bool(variant_1v1)),    # ctf.py:724-727
                         carry_agent_flags=int(bool(carry_agent_flags)))
        self.nb, self.nr = num_blue, num_red

    def reset(self, rng, mask=None):
        return self._call_reset(lib().oc_ctf_reset, rng, mask)

    def step(self, blue_actions, rng, autoreset=False, want_final_obs=False):
        return self._call_step(lib().oc_ctf_step, np.asarray(blue_actions).reshape(self.N, self.nb), rng, autoreset, want_final_obs)

    def policy_actions(self, tables, seed, episode, env_id_base=0):
        """Red actions int8 [N, num_red] of the scripted opponents for the current state (oc_ctf_policy_actions); `tables` as
        `gym_multigrid_b200.policy.ctf.device.build_tables` returns them, `episode` = per-env episode counters."""
        out = np.zeros((self.N, self.nr), np.int8)
        st = self._state()
        keep = [np.ascontiguousarray(episode, np.int32), np.ascontiguousarray(tables["kind"], np.int32),
                np.ascontiguousarray(tables["randomness"], np.float64), np.ascontiguousarray(tables["first_move"], np.uint8),
                np.ascontiguousarray(tables["patrol_goal"], np.uint16), np.ascontiguousarray(tables["on_border"], np.uint8),
                np.ascontiguousarray(tables["along_border"], np.uint16)]
        lib().oc_ctf_policy_actions(C.byref(self.cfg), C.c_int64(self.N), C.byref(st), *[_p(a) for a in keep],
                                    C.c_int32(len(tables["along_border"])), C.c_uint64(int(seed)), C.c_uint64(int(env_id_base)), _p(out))
        return out

    def flattened(self):
        """observation_option="flattened" (ctf.py:1084-1104) of the current state: int64 [N, L]."""
        st = self._state()
        L = lib().oc_ctf_flattened(C.byref(self.cfg), C.c_int64(self.N), C.byref(st), None)
        out = np.zeros((self.N, L), np.int64)
        lib().oc_ctf_flattened(C.byref(self.cfg), C.c_int64(self.N), C.byref(st), _p(out))
        return out


def partial_view3(grid, pos, W, H, V, see_through_walls=False, dirs=None, oob_code=1 | 7 << 2, opaque_rule=0):
    """MultiGridEnv.gen_obs for packed Collect grids [N, W*H] and agent positions [N, A, 2] -> [N, A, V, V, 3]."""
    grid = np.ascontiguousarray(grid, np.uint8)
    pos = np.ascontiguousarray(pos, np.uint8)
    N, A = pos.shape[0], pos.shape[1]
    out = np.zeros((N, A, V, V, 3), np.uint8)
    dirs = None if dirs is None else np.ascontiguousarray(dirs, np.uint8)
    lib().oc_partial_view3(_p(grid), _p(pos), _p(dirs), C.c_int64(N), C.c_int(W), C.c_int(H), C.c_int(A), C.c_int(V),
                           C.c_int(int(see_through_walls)), C.c_int(int(oob_code)), C.c_int(int(opaque_rule)), _p(out))
    return out


def pack_obs(obs):
    """Grid.encode() arrays [..., 3] -> packed cells type | colour << 2 | state << 6."""
    o = np.asarray(obs, np.uint8)
    return (o[..., 0] | (o[..., 1] << 2) | (o[..., 2] << 6)).astype(np.uint8)


def toroid(grid, pos, W, num_ball_types):
    """ToroidObservation (wrappers/toroid.py) for packed Collect grids [N, W*W] -> float32 [N, A, W, W, nb + A]."""
    grid = np.ascontiguousarray(grid, np.uint8)
    pos = np.ascontiguousarray(pos, np.uint8)
    N, A = pos.shape[0], pos.shape[1]
    out = np.zeros((N, A, W, W, num_ball_types + A), np.float32)
    lib().oc_toroid(_p(grid), _p(pos), C.c_int64(N), C.c_int(W), C.c_int(A), C.c_int(num_ball_types), _p(out))
    return out


# ============================================================================ Wildfire (extension)
class WfCfg(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("num_agents", C.c_int32), ("agent_colour", C.c_int32 * 32),
                ("num_fires", C.c_int32), ("ignite_threshold", C.c_uint32 * 5), ("burnout_threshold", C.c_uint32),
                ("max_steps", C.c_int32)]


class WfState(C.Structure):
    _fields_ = [("terrain", C.c_void_p), ("agents", C.c_void_p), ("hdr", C.c_void_p)]


def wildfire_thresholds(alpha, beta):
    """The integer thresholds both the oracle and the kernels compare Philox words with."""
    ign = [min(2**32 - 1, int(np.floor((1.0 - (1.0 - alpha) ** k) * 2.0**32))) for k in range(5)]
    return ign, min(2**32 - 1, int(np.floor(beta * 2.0**32)))


class WildfireOracle:
    """CPU restatement of the Wildfire extension (spec: include/multigrid_b200.h).  No reference exists."""

    def __init__(self, num_envs, size=64, num_agents=16, agents_index=None, num_fires=4, alpha=0.15, beta=0.05, max_steps=200,
                 seed=0, env_id_base=0, width=None, height=None):
        self.N, self.W, self.H, self.A = int(num_envs), int(width or size), int(height or size), int(num_agents)
        c = WfCfg()
        c.width, c.height, c.num_agents, c.num_fires, c.max_steps = self.W, self.H, self.A, num_fires, max_steps
        for i, v in enumerate(agents_index or [4] * self.A):
            c.agent_colour[i] = v
        ign, bo = wildfire_thresholds(alpha, beta)
        for k in range(5):
            c.ignite_threshold[k] = ign[k]
        c.burnout_threshold = bo
        self.cfg, self.seed, self.base = c, int(seed), int(env_id_base)
        self.terrain = np.zeros((self.N, self.W * self.H), np.uint8)
        self.agents = np.zeros((self.N, self.A, 4), np.uint8)
        self.hdr = np.zeros((self.N, 4), np.int32)

    def _st(self):
        s = WfState()
        s.terrain, s.agents, s.hdr = _p(self.terrain), _p(self.agents), _p(self.hdr)
        return s

    def reset(self, mask=None):
        obs = np.zeros((self.N, self.W, self.H, 3), np.uint8)
        st = self._st()
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        lib().oc_wf_reset(C.byref(self.cfg), C.c_int64(self.N), C.byref(st), _p(m), C.c_uint64(self.seed), C.c_uint64(self.base), _p(obs))
        return obs

    def step(self, actions, order=None, autoreset=False, want_final_obs=False):
        actions = np.ascontiguousarray(actions, np.int8).reshape(self.N, self.A)
        order = None if order is None else np.ascontiguousarray(order, np.uint8)
        obs = np.zeros((self.N, self.W, self.H, 3), np.uint8)
        rew = np.zeros((self.N, self.A), np.float64)
        term, trunc = np.zeros(self.N, np.uint8), np.zeros(self.N, np.uint8)
        fin = np.zeros_like(obs) if want_final_obs else None
        st = self._st()
        lib().oc_wf_step(C.byref(self.cfg), C.c_int64(self.N), C.byref(st), _p(actions), _p(order), C.c_uint64(self.seed),
                         C.c_uint64(self.base), _p(obs), _p(rew), _p(term), _p(trunc), C.c_int(int(autoreset)), _p(fin))
        out = (obs, rew, term.astype(bool), trunc.astype(bool))
        return out + (fin,) if want_final_obs else out


# ========================================================================== generic MultiGridEnv.step
class GenericOracle:
    """Base-class MultiGridEnv.step with DefaultWorld on the CPU (see mg_oracle.h)."""

    def __init__(self, num_envs, width, height, num_agents, max_steps):
        self.N, self.W, self.H, self.A, self.max_steps = int(num_envs), width, height, num_agents, max_steps
        self.gcell = np.ones((self.N, width * height), np.uint8)
        self.gstate = np.zeros((self.N, width * height), np.uint8)
        self.pos = np.zeros((self.N, num_agents, 2), np.uint8)
        self.step_count = np.zeros(self.N, np.int32)
        self.status = C.c_int32(0)

    def set_state_from_obs(self, obs6, pos):
        """obs6: encode_for_agents arrays [N, W, H, 6] (any agent's view), pos [N, A, 2]."""
        o = np.asarray(obs6, np.uint8).reshape(self.N, -1, 6)
        self.gcell[:] = o[..., 0] | (o[..., 1] << 4)
        self.gstate[:] = np.where(o[..., 0] == 4, o[..., 2], np.where(o[..., 0] == 10, o[..., 4], 0))
        self.pos[:] = np.asarray(pos).reshape(self.N, self.A, 2)
        self.step_count[:] = 0

    def encode(self):
        obs = np.zeros((self.N, self.A, self.W, self.H, 6), np.uint8)
        lib().oc_generic_encode(C.c_int64(self.N), self.W, self.H, self.A, _p(self.gcell), _p(self.gstate), _p(self.pos), _p(obs))
        return obs

    def partial_views(self, V, see_through_walls=False, dirs=None):
        """MultiGridEnv.gen_obs (encode_dim 6): u8 [N, A, V, V, 6]."""
        out = np.zeros((self.N, self.A, V, V, 6), np.uint8)
        d = None if dirs is None else np.ascontiguousarray(dirs, np.uint8)
        lib().oc_partial_view6(C.c_int64(self.N), self.W, self.H, self.A, int(V), int(bool(see_through_walls)), _p(self.gcell),
                               _p(self.gstate), _p(self.pos), _p(d), _p(out))
        return out

    def step(self, actions, order):
        actions = np.ascontiguousarray(actions, np.int8).reshape(self.N, self.A)
        order = np.ascontiguousarray(order, np.uint8).reshape(self.N, self.A)
        obs = np.zeros((self.N, self.A, self.W, self.H, 6), np.uint8)
        rew = np.zeros((self.N, self.A), np.float64)
        term, trunc = np.zeros(self.N, np.uint8), np.zeros(self.N, np.uint8)
        lib().oc_generic_step(C.c_int64(self.N), self.W, self.H, self.A, self.max_steps, _p(self.gcell), _p(self.gstate), _p(self.pos),
                              _p(self.step_count), _p(actions), _p(order), _p(obs), _p(rew), _p(term), _p(trunc), C.byref(self.status))
        return obs, rew, term.astype(bool), trunc.astype(bool)
