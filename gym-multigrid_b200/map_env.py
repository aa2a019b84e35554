"""Batched Maze and Capture-the-Flag environments on one B200: host-side mirrors of the reference's
`MazeSingleAgentEnv` (envs/maze.py) and `CtFMvNEnv` (envs/ctf.py:657-1433).  The text map is shared by
all envs (device tables owned by the C handle); per-env state (agent positions, dirs, flags, step
counter) is one caller-owned torch uint8 CUDA tensor; `step` / `reset` are one kernel launch each."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .vector_base import VectorEnvSurface
from .spaces import Box, Discrete, MultiDiscrete
from .utils.map import load_text_map


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class _MapVecEnv(VectorEnvSurface):
    family = None
    ref_dtype = None

    def _create(self, num_envs, field_map, num_blue, num_red, flag_reward, battle_reward, obstacle_penalty, step_penalty,
                battle_range, randomness, max_steps, device, seed, autoreset, env_id_base, reference_dtypes, variant_1v1=False):
        self._lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("gym-multigrid_b200 runs on CUDA devices only (no CPU fallback)")
        if not torch.cuda.is_available():
            raise RuntimeError("no CUDA device available: gym-multigrid_b200 has no CPU fallback")
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", idx)
        fm = np.ascontiguousarray(np.asarray(field_map))
        if fm.ndim != 2 or fm.shape[0] != fm.shape[1]:
            raise ValueError("square maps only (the reference mixes width and height: maze.py:68-70, ctf.py:745-747)")
        if not np.array_equal(fm, fm.astype(np.uint8)):
            raise ValueError("map codes must be small non-negative integers")
        self.field_map = fm.astype(np.uint8)
        self.size = self.width = self.height = int(fm.shape[0])
        self.num_envs, self.num_blue, self.num_red = int(num_envs), int(num_blue), int(num_red)
        self.n_agents = self.num_blue + self.num_red
        self.max_steps, self.autoreset = int(max_steps), bool(autoreset)
        self.reference_dtypes = bool(reference_dtypes)
        cfg = _lib.MapConfig()
        cfg.struct_size = C.sizeof(_lib.MapConfig)
        cfg.family, cfg.num_envs, cfg.env_id_base = self.family, self.num_envs, int(env_id_base)
        cfg.size, cfg.field_map = self.size, self.field_map.ctypes.data
        cfg.num_blue, cfg.num_red = self.num_blue, self.num_red
        cfg.flag_reward, cfg.battle_reward = float(flag_reward), float(battle_reward)
        cfg.obstacle_penalty, cfg.step_penalty = float(obstacle_penalty), float(step_penalty)
        cfg.battle_range, cfg.randomness = float(battle_range), float(randomness)
        cfg.max_steps, cfg.autoreset = self.max_steps, int(self.autoreset)
        cfg.obs_dtype = _lib.OBS_REFERENCE if self.reference_dtypes else _lib.OBS_U8
        cfg.variant_1v1 = int(bool(variant_1v1))
        self._variant_1v1 = bool(variant_1v1)
        cfg.seed = int(seed) & (2**64 - 1)
        h = C.c_void_p()
        if self._lib.mg_create_map(C.byref(cfg), idx, C.byref(h)) != 0:
            raise ValueError(_lib.last_error(None))
        self._h = h
        N, S, n = self.num_envs, self.size, self.n_agents
        odt = self.ref_dtype if self.reference_dtypes else torch.uint8
        with torch.cuda.device(self.device):
            self.state = torch.zeros(self._lib.mg_state_bytes(self._h), dtype=torch.uint8, device=self.device)
            self._obs = torch.zeros((N, S, S), dtype=odt, device=self.device)
            self._rewards = torch.zeros(N, dtype=torch.float64, device=self.device)
            self._term = torch.zeros(N, dtype=torch.uint8, device=self.device)
            self._trunc = torch.zeros(N, dtype=torch.uint8, device=self.device)
        self._final_obs = None
        self._planes = {}
        slots = 1
        while slots < n:
            slots *= 2
        for name, pid, dt, cols in (("agents", _lib.MAP_PLANE_AGENTS, torch.uint8, slots * 4), ("hdr", _lib.MAP_PLANE_HDR, torch.int32, 4)):
            off, nbytes, row = C.c_size_t(), C.c_size_t(), C.c_size_t()
            self._lib.mg_state_plane(self._h, pid, C.byref(off), C.byref(nbytes), C.byref(row))
            assert row.value == cols * dt.itemsize, "state plane layout mismatch between the library and the host mirror"
            self._planes[name] = self.state[off.value: off.value + nbytes.value].view(dt).view(-1, cols)[:N]
        self._agents = self._planes["agents"].view(N, slots, 4)[:, :n]     # (x, y, dir, flags) per agent
        self._io = _lib.StepIO()
        self._bound = None
        self._red_actions = None
        self._skip_map_obs = False   # CtF "flattened" / "positional" modes: the step does not write the map observation
        self._host = None
        self.with_info = False      # True: step() / reset() also return the reference's info dict (one more launch)
        self._trace_keepalive = None
        self.closed = False

    # --- zero-copy state views
    @property
    def agent_pos(self):
        return self._agents[..., 0:2]

    @property
    def agent_dir(self):
        return self._agents[..., 2]

    @property
    def agent_flags(self):
        """u8 [N, n]: bit0 terminated (defeated), bit1 collided (agent.py:97-100)."""
        return self._agents[..., 3]

    @property
    def agent_terminated(self):
        return (self._agents[..., 3] & 1).bool()

    @property
    def step_count(self):
        return self._planes["hdr"][:, 0]

    @property
    def episode_count(self):
        return self._planes["hdr"][:, 3]

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(_lib.last_error(self._h))

    def reset(self, *, seed=None, options=None, mask=None):
        m = None if mask is None else torch.as_tensor(mask, device=self.device).to(torch.uint8).contiguous()
        self._check(self._lib.mg_reset(self._h, _ptr(self.state), _ptr(m), None if self._skip_map_obs else _ptr(self._obs), self._stream()))
        return self._obs, (self.get_info() if self.with_info else {})

    def _prep_actions(self, actions):
        a = actions
        if a.device != self.device:
            a = a.to(self.device, non_blocking=True)
        if a.dtype.is_floating_point:
            a = torch.round(a)  # "Just in case NN outputs are, for some reason, not discrete." (ctf.py:1303-1304)
        if a.dtype != torch.int8:
            a = a.to(torch.int8)
        return a.reshape(self.num_envs, self.num_blue).contiguous()

    def step(self, actions):
        if not isinstance(actions, torch.Tensor):
            return self.step_host(actions)
        a = actions
        if not (a.dtype is torch.int8 and a.is_cuda and a.is_contiguous() and a.numel() == self.num_envs * self.num_blue and a.device == self.device):
            a = self._prep_actions(actions)
        bound = self._bound
        if bound is None or bound[0] is not self._obs or bound[1] is not self._final_obs:   # (re)bind what does not change per call
            io = self._io
            io.obs, io.rewards = (None if self._skip_map_obs else self._obs.data_ptr()), self._rewards.data_ptr()
            io.terminated, io.truncated = self._term.data_ptr(), self._trunc.data_ptr()
            io.final_obs = self._final_obs.data_ptr() if self._final_obs is not None else None
            bound = self._bound = (self._obs, self._final_obs, C.byref(io), _ptr(self.state), self._term.view(torch.bool),
                                   self._trunc.view(torch.bool), getattr(torch._C, "_cuda_getCurrentRawStream", None))
        self._io.actions = a.data_ptr()
        stream = bound[6](self.device.index) if bound[6] else torch.cuda.current_stream(self.device).cuda_stream
        if self._lib.mg_step(self._h, bound[3], bound[2], stream):
            raise RuntimeError(_lib.last_error(self._h))
        return self._obs, self._rewards, bound[4], bound[5], (self._info() if (self.with_info or self._final_obs is not None) else {})

    def _host_io(self, actions):
        if self._host is None:
            N = self.num_envs
            blk, obs, rew, term, trunc = _lib.host_result_buffers(self._lib, self._h, tuple(self._obs.shape), self._obs.dtype, N, 1)
            self._host = dict(act=torch.zeros((N, self.num_blue), dtype=torch.int8, pin_memory=True), obs=obs, rew=rew, term=term,
                              trunc=trunc, block=blk)
            self._host_np = {k: v.numpy() for k, v in self._host.items()}
            h, io = self._host, _lib.StepIO()
            io.actions, io.obs, io.rewards = h["act"].data_ptr(), h["obs"].data_ptr(), h["rew"].data_ptr()
            io.terminated, io.truncated, io.final_obs = h["term"].data_ptr(), h["trunc"].data_ptr(), None
            self._host_io_struct = io
        self._host_np["act"][...] = np.round(np.asarray(actions)).astype(np.int64).reshape(self.num_envs, self.num_blue)
        return self._host_io_struct

    def _host_result(self):
        n = self._host_np
        return n["obs"], n["rew"], n["term"].view(np.bool_), n["trunc"].view(np.bool_), {}

    def step_host(self, actions):
        io = self._host_io(actions)
        self._check(self._lib.mg_step_host(self._h, _ptr(self.state), C.byref(io), self._stream()))
        return self._host_result()

    def get_info(self, out=None):
        """`_get_info()` of every env as a dict of float64 CUDA tensors [N], keys and values as the reference's dict
        (maze.py:262-269; ctf.py:1165-1182): one extra kernel launch over the state planes."""
        K = len(self.info_keys)
        if out is None:
            out = torch.empty((self.num_envs, K), dtype=torch.float64, device=self.device)
        self._check(self._lib.mg_map_info(self._h, _ptr(self.state), _ptr(out), self._stream()))
        return {k: out[:, i] for i, k in enumerate(self.info_keys)}

    def _info(self):
        info = self.get_info() if self.with_info else {}
        if self._final_obs is not None:
            info["final_observation"] = self._final_obs
            info["_final_observation"] = (self._term | self._trunc).view(torch.bool)
        return info

    def enable_final_observation(self, enable=True):
        self._final_obs = torch.zeros_like(self._obs) if enable else None

    def set_trace(self, **arrays):
        """Validation mode: recorded reference RNG outputs (see include/multigrid_b200.h: mg_map_trace).
        No arguments = back to Philox mode."""
        arrays = {k: v for k, v in arrays.items() if v is not None}
        if not arrays:
            self._lib.mg_set_map_trace(self._h, None)
            self._trace_keepalive = None
            return None
        dt = dict(start_index=torch.int32, blue_place=torch.int32, red_place=torch.int32, red_actions=torch.int8,
                  order=torch.uint8, blue_win=torch.uint8)
        t = {k: torch.as_tensor(np.ascontiguousarray(v), device=self.device).to(dt[k]).contiguous() for k, v in arrays.items()}
        t["battles_used"] = torch.zeros(self.num_envs, dtype=torch.int32, device=self.device)
        tr = _lib.MapTrace()
        for k, v in t.items():
            setattr(tr, k, v.data_ptr())
        tr.KB = t["blue_win"].reshape(self.num_envs, -1).shape[1] if "blue_win" in t else 0
        self._lib.mg_set_map_trace(self._h, C.byref(tr))
        self._trace_keepalive = t
        return t

    def status(self) -> int:
        s = C.c_int32(0)
        self._check(self._lib.mg_status(self._h, self._stream(), C.byref(s)))
        return s.value

    @property
    def launch_count(self) -> int:
        return int(self._lib.mg_launch_count(self._h))

    def close(self):
        if not self.closed and getattr(self, "_h", None):
            torch.cuda.synchronize(self.device)
            self._lib.mg_destroy(self._h)
            self._h, self.closed = None, True

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


class MazeVecEnv(_MapVecEnv):
    """`num_envs` x MazeSingleAgentEnv (maze.py:26-377), constructor kwargs as in maze.py:31-40.
    Observations: the "map" option - `_encode_map()` values [N, W, H] indexed [x][y] (uint8, or float64 with
    `reference_dtypes=True` as the reference returns, maze.py:246); actions MazeActions Discrete(5)."""
    _render_family = "maze"
    metadata = {"render_modes": ["rgb_array"], "autoreset_mode": "same_step"}
    render_mode = "rgb_array"
    family = _lib.FAMILY_MAZE
    ref_dtype = torch.float64
    info_keys = ("d_a_f", "d_a_ob")

    def __init__(self, num_envs, map_path, max_steps=100, flag_reward=1.0, obstacle_penalty_ratio=0.0, step_penalty_ratio=0.01,
                 observation_option="map", device="cuda:0", seed=0, autoreset=True, env_id_base=0, reference_dtypes=False,
                 view_size=7, see_through_walls=False):
        if observation_option not in ("map", "partial"):
            raise NotImplementedError('observation_option "map" (the reference\'s) or "partial" (gen_obs views, BASELINE config 4); '
                                      '"positional" is agent_pos + the static lists')
        fm = load_text_map(map_path)
        fr = float(flag_reward)
        self._create(num_envs, fm, 1, 0, fr, 0.0, fr * float(obstacle_penalty_ratio), fr * float(step_penalty_ratio), 0.0, 0.0,
                     max_steps, device, seed, autoreset, env_id_base, reference_dtypes)
        self.single_action_space = Discrete(5)
        self.action_space = MultiDiscrete(np.full((self.num_envs,), 5))
        hi = 3  # len(MazeWorld.OBJECT_TO_IDX) - 1 (maze.py:150-155)
        odt = np.float64 if reference_dtypes else np.uint8
        self.single_observation_space = Box(0, hi, (self.size, self.size), odt)
        self.observation_space = Box(0, hi, (self.num_envs, self.size, self.size), odt)
        self.background = [tuple(int(v) for v in c) for c in zip(*np.where(self.field_map == 0))]
        self.obstacle = [tuple(int(v) for v in c) for c in zip(*np.where(self.field_map == 3))]
        self.flag = [tuple(int(v) for v in c) for c in zip(*np.where(self.field_map == 2))]
        if observation_option == "partial":
            self.set_partial_obs(view_size, see_through_walls)

    def positional_obs(self):
        """observation_option="positional" (maze.py:224-231): `agent` int64 [N, 2] plus the map's background / flag / obstacle cell
        lists (np.where order, flattened) broadcast over the envs - the lists are static, so they are expanded views, not copies."""
        N = self.num_envs
        st = lambda cells: torch.as_tensor(np.array(cells, np.int64).reshape(-1), device=self.device).expand(N, -1)  # noqa: E731
        return {"agent": self.agent_pos[:, 0].to(torch.int64), "background": st(self.background), "flag": st(self.flag),
                "obstacle": st(self.obstacle)}

    def set_partial_obs(self, view_size=7, see_through_walls=False):
        """Switch `reset` / `step` observations to MultiGridEnv.gen_obs partial views u8 [N, 1, V, V, 3] (V in 3, 5, 7), computed
        by the same kernel launch that steps the envs; view_size 0 switches back to the "map" observation."""
        V = int(view_size)
        self._check(self._lib.mg_set_partial_obs(self._h, V, int(bool(see_through_walls))))
        self._final_obs = None
        if V:
            self._obs = torch.zeros((self.num_envs, 1, V, V, 3), dtype=torch.uint8, device=self.device)
            self.single_observation_space = Box(0, 255, (V, V, 3), np.uint8)
            self.observation_space = Box(0, 255, (self.num_envs, 1, V, V, 3), np.uint8)
        else:
            odt = self.ref_dtype if self.reference_dtypes else torch.uint8
            self._obs = torch.zeros((self.num_envs, self.size, self.size), dtype=odt, device=self.device)
        self._host = None


    def gen_obs(self, view_size=7, see_through_walls=False, out=None):
        """Partial view of the maze (BASELINE config 4): reference dynamics x the reference's gen_obs algorithm
        (multigrid.py:485-532) with MazeWorld encodings; u8 [N, 1, V, V, 3].  Cells outside the map show the filler
        (3, 7, 1) - MazeWorld has no wall type, so this one code is an extension (see include/multigrid_b200.h)."""
        V = int(view_size)
        if out is None:
            out = torch.empty((self.num_envs, 1, V, V, 3), dtype=torch.uint8, device=self.device)
        self._check(self._lib.mg_gen_obs(self._h, _ptr(self.state), None, V, int(bool(see_through_walls)), _ptr(out), self._stream()))
        return out


class CtfVecEnv(_MapVecEnv):
    """`num_envs` x CtFMvNEnv (ctf.py:657-1433) with RwPolicy red agents (policy/ctf/heuristic.py:69-72) drawn on
    the device.  Observations: the "map" option - `_encode_map()` (transposed, [N, H, W]; uint8, or int64 with
    `reference_dtypes=True`); `positional_obs()` / `flattened_obs()` assemble the other two options from the
    state planes.  Actions MultiDiscrete([5] * num_blue_agents); reward = scalar team reward (float64)."""
    _render_family = "ctf"
    metadata = {"render_modes": ["rgb_array"], "autoreset_mode": "same_step"}
    render_mode = "rgb_array"
    family = _lib.FAMILY_CTF
    ref_dtype = torch.int64
    info_keys = ("d_ba_ra", "d_ba_bf", "d_ba_rf", "d_ra_bf", "d_ra_rf", "d_bf_rf", "d_ba_bb", "d_ba_rb", "d_ra_bb", "d_ra_rb", "d_ba_ob")

    def __init__(self, num_envs, map_path, num_blue_agents=2, num_red_agents=2, battle_range=1, randomness=0.75, flag_reward=1,
                 battle_reward_ratio=0.25, obstacle_penalty_ratio=0, step_penalty_ratio=0.01, max_steps=100,
                 observation_option="map", observation_scaling=1, device="cuda:0", seed=0, autoreset=True, env_id_base=0,
                 reference_dtypes=False, variant_1v1=False, carry_agent_flags=False):
        if observation_option not in ("map", "flattened", "positional"):
            raise ValueError(f"Invalid observation_option: {observation_option}")     # ctf.py:1105-1108
        self.observation_option = observation_option
        fm = load_text_map(map_path)
        fr = flag_reward
        self._create(num_envs, fm, num_blue_agents, num_red_agents, float(fr), float(battle_reward_ratio * fr),
                     float(obstacle_penalty_ratio * fr), float(step_penalty_ratio * fr), float(battle_range), float(randomness),
                     max_steps, device, seed, autoreset, env_id_base, reference_dtypes, variant_1v1)
        self.num_blue_agents, self.num_red_agents = self.num_blue, self.num_red
        if carry_agent_flags:
            self.set_carry_agent_flags(True)
        self.single_action_space = MultiDiscrete([5] * self.num_blue)
        self.action_space = MultiDiscrete(np.full((self.num_envs, self.num_blue), 5))
        odt = np.int64 if reference_dtypes else np.uint8
        self.single_observation_space = Box(0, 6, (self.size, self.size), odt)
        self.observation_space = Box(0, 6, (self.num_envs, self.size, self.size), odt)
        f = self.field_map
        cells = lambda code: [tuple(int(v) for v in c) for c in zip(*np.where(f == code))]  # noqa: E731
        self.obstacle, self.blue_flag, self.red_flag = cells(6), cells(4)[0], cells(5)[0]
        self.blue_territory = cells(0) + [self.blue_flag]
        self.red_territory = cells(1) + [self.red_flag]
        if observation_option != "map":   # reset / step return the flattened vector (or its dict of views): step launch + ctf_flat_kernel
            L = self._lib.mg_ctf_flat_len(self._h)
            fdt = torch.int64 if reference_dtypes else torch.uint8      # as for the map observation: the reference's dtype on request
            self._flat = torch.zeros((self.num_envs, L), dtype=fdt, device=self.device)
            self._skip_map_obs = True
            if observation_option == "flattened":                                # ctf.py:940-948
                odt = np.int64 if reference_dtypes else np.uint8
                self.single_observation_space = Box(0, max(self.size - 1, 1), (L,), odt)
                self.observation_space = Box(0, max(self.size - 1, 1), (self.num_envs, L), odt)

    def _option_obs(self):
        if self.observation_option == "map":
            return self._obs
        self.flattened_obs(out=self._flat)
        return self._flat if self.observation_option == "flattened" else self._positional_views(self._flat)

    def reset(self, *, seed=None, options=None, mask=None):
        """`seed` (an int) re-keys the env's generator as `super().reset(seed=seed)` does in the reference (ctf.py:1056,
        multigrid.py:114-119): the same seed gives the same placements and the same episode for the same actions."""
        if seed is not None:
            if mask is not None:
                raise ValueError("reset(seed=..., mask=...): re-keying the generator restarts the RNG streams of EVERY env of the batch; "
                                 "reseed with a full reset, or reset the masked envs without a seed")
            self.reseed(seed)
            self._planes["hdr"][:, 3] = 0     # episode counters too: the device opponents' draws are keyed by (step, episode)
        _, info = super().reset(seed=seed, options=options, mask=mask)
        return self._option_obs(), info

    def _check_host_option(self):
        if self.observation_option != "map":     # before anything touches the state
            raise NotImplementedError('host-array steps return observation_option="map" only; pass CUDA tensors')

    def step(self, actions):
        if not isinstance(actions, torch.Tensor):
            self._check_host_option()
        if self._device_policies and not self._fused_policies:      # one small launch ahead of the step's: red actions from the current state
            self._check(self._lib.mg_red_policy_actions(self._h, _ptr(self.state), _ptr(self._red_buf), self._stream()))
        elif self._enemy_policies is not None:
            self._decide_red_actions()
        out = super().step(actions)
        if self.observation_option == "map":
            return out
        return (self._option_obs(),) + tuple(out[1:])

    def step_async(self, actions):
        """`step` with host arrays, first half: the red team decides exactly as in `step` - the device opponents' kernel is
        enqueued on the host-path stream ahead of the step, host policies decide now - then H2D actions -> step -> D2H results."""
        self._check_host_option()
        if self._host_pending:
            raise RuntimeError("step_async called again before step_wait")
        if self._host_stream is None:
            self._host_stream = torch.cuda.Stream(device=self.device)
        if self._device_policies and not self._fused_policies:
            self._host_stream.wait_stream(torch.cuda.current_stream(self.device))
            self._check(self._lib.mg_red_policy_actions(self._h, _ptr(self.state), _ptr(self._red_buf), C.c_void_p(self._host_stream.cuda_stream)))
        elif self._enemy_policies is not None:
            self._host_stream.synchronize()      # the previous host-path step must have landed before the state is read back
            self._decide_red_actions()
        super().step_async(actions)

    def set_carry_agent_flags(self, on=True):
        """One reference env INSTANCE per slot, stepped through several episodes: the reference's `reset()` never clears
        `Agent.terminated` / `collided` / `bg_color` (core/agent.py:97-100 are their only assignments outside `step`),
        so an agent defeated in one episode starts the next one defeated.  Off (default): every reset - explicit, masked, or the
        same-step autoreset - starts a fresh instance, which is how the reference's own tests and training script use the env."""
        self._check(self._lib.mg_set_carry_agent_flags(self._h, int(bool(on))))
        self.carry_agent_flags = bool(on)

    def set_red_actions(self, red_actions=None):
        """Drive the red agents from outside (the reference's `enemy_policies`, ctf.py:666): `red_actions` int8 CUDA tensor
        [N, num_red] that every following `step` reads - overwrite it in place between steps (a learned opponent, self-play, a
        host-side A* policy).  None = back to the built-in RwPolicy drawn on the device."""
        if self._device_policies or self._enemy_policies is not None:     # the caller takes the red team over: scripted opponents off
            if self._device_policies:
                self._check(self._lib.mg_set_red_policies(self._h, None))     # (switches the fusion off too)
            self._device_policies = self._fused_policies = False
            self._enemy_policies = None
        return self._bind_red_actions(red_actions)

    def _bind_red_actions(self, red_actions):
        if red_actions is None:
            self._red_actions = None
            self._check(self._lib.mg_set_red_actions(self._h, None))
            return None
        t = torch.as_tensor(red_actions, device=self.device).to(torch.int8).reshape(self.num_envs, self.num_red).contiguous()
        self._red_actions = t
        self._check(self._lib.mg_set_red_actions(self._h, _ptr(t)))
        return t

    _enemy_policies = None

    _device_policies = False

    _fused_policies = False

    def set_enemy_policies(self, enemy_policies=None, random_generator=None, device=False, fused=True):
        """The reference's `enemy_policies` argument (ctf.py:666, 775-826) for every env of the batch: one policy for all red
        agents or a list of `num_red_agents` of them - any object with `act(observation_dict, curr_pos) -> int` (the reference's
        CtfPolicy interface; `policy/ctf/heuristic.py` has Fight / Capture / Patrol / PatrolFight).  `random_generator`,
        `field_map` and `action_set` attributes are filled in as the reference's constructor does (`random_generator` and
        `action_set` always, `field_map` when None).  The policies decide on the HOST, as in the reference: before each `step`
        the positional observation is read back once, `act` is called per env and red agent in index order (env 0 first), and
        the actions go to the kernel through `set_red_actions` - a per-step device sync, meant for SB3-sized batches.
        None, or only None / RwPolicy entries = the built-in opponent drawn on the device.  Returns the list in use (or None).

        `device=True` keeps the whole loop on the GPU for any batch size: the policies of this package (exact types, ego red)
        are turned into tables - the first move of the reference's A* route per (cell, target), the patrol targets - and a
        small kernel decides for every env before each step (`mg_set_red_policies` / `mg_red_policy_actions`).  Targets and
        routes are the reference's; the random draws (follow the route or not, random action, patrol target) come from the
        env's Philox generator instead of numpy's, as RwPolicy's do.  `fused=True` (default) makes the decision part of the step
        itself (`mg_set_red_policy_fusion`: one launch for 2v2 with the u8 map observation, the policy kernel ahead of the step
        kernel otherwise); `fused=False` keeps the explicit `mg_red_policy_actions` call before every step.  Same results."""
        from .actions import CtfActions
        from .policy.ctf.heuristic import RwPolicy
        nr = self.num_red
        pols = list(enemy_policies) if isinstance(enemy_policies, (list, tuple)) else [enemy_policies] * nr
        if len(pols) != nr:
            raise AssertionError("len(enemy_policies) must equal num_red_agents")       # ctf.py:779
        if self._device_policies:
            self._check(self._lib.mg_set_red_policies(self._h, None))     # (switches the fusion off too)
            self._device_policies = self._fused_policies = False
        if all(p is None or type(p) is RwPolicy for p in pols):
            self._enemy_policies = None
            self._bind_red_actions(None)
            return None
        if device:
            from .policy.ctf.device import build_tables
            t = build_tables(pols, self.field_map)
            rp = _lib.RedPolicies()
            rp.struct_size, rp.num_red, rp.n_along = C.sizeof(_lib.RedPolicies), nr, len(t["along_border"])
            for k in range(nr):
                rp.kind[k], rp.randomness[k] = int(t["kind"][k]), float(t["randomness"][k])
            keep = [np.ascontiguousarray(t[name]) for name in ("first_move", "patrol_goal", "on_border", "along_border")]
            rp.first_move, rp.patrol_goal, rp.on_border, rp.along_border = (a.ctypes.data for a in keep)
            self._check(self._lib.mg_set_red_policies(self._h, C.byref(rp)))      # copies the tables during the call
            self._enemy_policies, self._device_policies, self._policy_tables = None, True, t
            self._red_buf = self._bind_red_actions(torch.zeros((self.num_envs, nr), dtype=torch.int8, device=self.device))
            if fused:
                self._check(self._lib.mg_set_red_policy_fusion(self._h, _ptr(self._red_buf)))
                self._fused_policies = True
            return pols
        gen = random_generator if random_generator is not None else np.random.default_rng()
        pols = [RwPolicy() if p is None else p for p in pols]
        for p in pols:
            if hasattr(p, "random_generator"):
                p.random_generator = gen
            if getattr(p, "field_map", 0) is None:
                p.field_map = np.asarray(self.field_map)
            if hasattr(p, "action_set"):
                p.action_set = CtfActions
        self._enemy_policies = pols
        self._red_host = np.zeros((self.num_envs, nr), np.int8)
        self._red_buf = self._bind_red_actions(self._red_host)
        return pols

    def set_policy_trace(self, patrol_target=None, follow=None, action=None):
        """Validation mode of the device-decided opponents (mg_set_policy_trace): recorded outputs of the reference's generator
        per (env, red agent) - the border cell PatrolPolicy drew (cell index x * size + y), follow-the-route booleans, uniform
        actions - replace the Philox draws of `mg_red_policy_actions`.  No arguments = back to Philox."""
        if patrol_target is None and follow is None and action is None:
            self._check(self._lib.mg_set_policy_trace(self._h, None, None, None))
            self._policy_trace = None
            return None
        shape = (self.num_envs, self.num_red)
        t = (torch.as_tensor(np.ascontiguousarray(patrol_target, dtype=np.uint16).view(np.int16), device=self.device).reshape(shape).contiguous(),   # same 16 bits
             torch.as_tensor(np.asarray(follow), device=self.device).to(torch.uint8).reshape(shape).contiguous(),
             torch.as_tensor(np.asarray(action), device=self.device).to(torch.int8).reshape(shape).contiguous())
        self._check(self._lib.mg_set_policy_trace(self._h, *(_ptr(x) for x in t)))
        self._policy_trace = t
        return t

    def _decide_red_actions(self):
        d = {k: v.cpu().numpy() for k, v in self.positional_obs().items()}
        pos = d["red_agent"].reshape(self.num_envs, self.num_red, 2)
        for e in range(self.num_envs):
            obs = {k: v[e] for k, v in d.items()}
            if "is_red_agent_defeated" in obs:      # Ctf1v1Env hands a plain int there (ctf.py:395)
                obs["is_red_agent_defeated"] = int(obs["is_red_agent_defeated"][0])
            for k, p in enumerate(self._enemy_policies):
                self._red_host[e, k] = int(p.act(obs, (int(pos[e, k, 0]), int(pos[e, k, 1]))))     # ctf.py:1297-1301
        self._red_buf.copy_(torch.from_numpy(self._red_host))

    def game_stats(self):
        """The reference's `env.game_stats` (ctf.py:1068-1073) for every env, as bool CUDA tensors.  Cleared by reset - with
        same-step autoreset the finished episode's stats are gone when `step` returns, so read them with `autoreset=False`."""
        st = self._planes["hdr"][:, 1]
        bits = (st[:, None] >> (8 + torch.arange(self.n_agents, device=self.device))) & 1
        return {"blue_agent_defeated": bits[:, :self.num_blue].bool(), "red_agent_defeated": bits[:, self.num_blue:].bool(),
                "blue_flag_captured": (st & 1).bool(), "red_flag_captured": ((st >> 1) & 1).bool()}

    def flattened_obs(self, out=None, dtype=None):
        """observation_option="flattened" (ctf.py:1084-1104) of every env: CUDA tensor [N, L], one kernel (mg_ctf_flat_obs).
        dtype int64 (default: the reference's) or uint8 (1/8 of the bytes; every entry is a coordinate < 256 or a flag)."""
        L = self._lib.mg_ctf_flat_len(self._h)
        if out is None:
            out = torch.empty((self.num_envs, L), dtype=dtype or torch.int64, device=self.device)
        if out.dtype not in (torch.int64, torch.uint8) or tuple(out.shape) != (self.num_envs, L) or not out.is_contiguous():
            raise ValueError(f"flattened_obs: out must be a contiguous int64 or uint8 tensor of shape {(self.num_envs, L)}")
        fn = self._lib.mg_ctf_flat_obs if out.dtype is torch.int64 else self._lib.mg_ctf_flat_obs_u8
        self._check(fn(self._h, _ptr(self.state), _ptr(out), self._stream()))
        return out

    def positional_obs(self):
        """observation_option="positional" (ctf.py:1112-1135) as batched int64 CUDA tensors: the flattened vector cut at the
        key boundaries (views of one buffer)."""
        return self._positional_views(self.flattened_obs())

    def _positional_views(self, f):
        n, nb = self.num_blue + self.num_red, self.num_blue
        sizes = [("blue_agent", 2 * nb), ("red_agent", 2 * (n - nb)), ("blue_flag", 2), ("red_flag", 2),
                 ("blue_territory", 2 * len(self.blue_territory)), ("red_territory", 2 * len(self.red_territory)),
                 ("obstacle", 2 * len(self.obstacle)),
                 ("is_red_agent_defeated", 1) if self._variant_1v1 else ("terminated_agents", n)]   # ctf.py:378-396 vs :1112-1135
        out, k = {}, 0
        for key, w in sizes:
            out[key] = f[:, k:k + w]
            k += w
        assert k == f.shape[1]
        return out


class Ctf1v1VecEnv(CtfVecEnv):
    """`num_envs` x Ctf1v1Env (ctf.py:50-654): one blue agent vs one RwPolicy red agent, fixed move order blue then
    red, a lost battle ends the episode.  `obstacle_penalty_ratio` must be 0: with a non-zero ratio the reference's
    own step raises (`blue_agent_loc in self.obstacle`, ctf.py:639, is an ambiguous ndarray test), so no behaviour
    is defined there.  Action space Discrete(5); constructor kwargs as ctf.py:55-70."""

    def __init__(self, num_envs, map_path, battle_range=1.0, randomness=0.75, flag_reward=1.0, battle_reward_ratio=0.25,
                 obstacle_penalty_ratio=0.0, step_penalty_ratio=0.01, max_steps=100, observation_option="map", **kwargs):
        if obstacle_penalty_ratio != 0:
            raise ValueError("Ctf1v1Env is only defined for obstacle_penalty_ratio == 0 (the reference raises otherwise, ctf.py:639)")
        super().__init__(num_envs, map_path, num_blue_agents=1, num_red_agents=1, battle_range=battle_range, randomness=randomness,
                         flag_reward=flag_reward, battle_reward_ratio=battle_reward_ratio, obstacle_penalty_ratio=0,
                         step_penalty_ratio=step_penalty_ratio, max_steps=max_steps, observation_option=observation_option,
                         variant_1v1=True, **kwargs)
        self.single_action_space = Discrete(5)
