"""Batched base-class `MultiGridEnv.step` with `DefaultWorld` (multigrid.py:397-483, world.py:33-52): still / left /
right / forward, goal termination with `_reward` (multigrid.py:218-223), per-agent full-grid observations
`Grid.encode_for_agents` (encode_dim 6, grid.py:254-284).  No shipped reference env reaches this path (every env
overrides `step`); `_gen_grid` is abstract in the base class (multigrid.py:199-201), so the layout is injected:
`set_layout` takes the encoded grid(s) a reference `_gen_grid` would produce and `reset` / autoreset restore it."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .vector_base import VectorEnvSurface
from .spaces import Box, Discrete, MultiDiscrete

# DefaultWorld.OBJECT_TO_IDX (world.py:37-51)
DEFAULT_OBJECT_TO_IDX = dict(unseen=0, empty=1, wall=2, floor=3, door=4, key=5, ball=6, box=7, goal=8, lava=9, agent=10,
                             objgoal=11, switch=12)


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class GenericVecEnv(VectorEnvSurface):
    """`step(actions[N, A]) -> (obs u8 [N, A, W, H, 6], rewards f64 [N, A], terminated [N], truncated [N], info)`.
    The reference returns a list of A arrays per env (multigrid.py:476-481); axis 1 is that list."""

    def __init__(self, num_envs, width, height=None, num_agents=1, max_steps=100, device="cuda:0", seed=0, autoreset=True,
                 env_id_base=0):
        self._lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("gym-multigrid_b200 runs on CUDA devices only (no CPU fallback)")
        if not torch.cuda.is_available():
            raise RuntimeError("no CUDA device available: gym-multigrid_b200 has no CPU fallback")
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", idx)
        self.num_envs, self.width, self.height = int(num_envs), int(width), int(height or width)
        self.num_agents, self.max_steps, self.autoreset = int(num_agents), int(max_steps), bool(autoreset)
        cfg = _lib.GenericConfig()
        cfg.struct_size = C.sizeof(_lib.GenericConfig)
        cfg.family, cfg.num_envs, cfg.env_id_base = _lib.FAMILY_GENERIC, self.num_envs, int(env_id_base)
        cfg.width, cfg.height, cfg.num_agents = self.width, self.height, self.num_agents
        cfg.max_steps, cfg.autoreset, cfg.seed = self.max_steps, int(self.autoreset), int(seed) & (2**64 - 1)
        h = C.c_void_p()
        if self._lib.mg_create_generic(C.byref(cfg), idx, C.byref(h)) != 0:
            raise ValueError(_lib.last_error(None))
        self._h = h
        N, W, H, A = self.num_envs, self.width, self.height, self.num_agents
        with torch.cuda.device(self.device):
            self.state = torch.zeros(self._lib.mg_state_bytes(self._h), dtype=torch.uint8, device=self.device)
            self._obs = torch.zeros((N, A, W, H, 6), dtype=torch.uint8, device=self.device)
            self._rewards = torch.zeros((N, A), dtype=torch.float64, device=self.device)
            self._term = torch.zeros(N, dtype=torch.uint8, device=self.device)
            self._trunc = torch.zeros(N, dtype=torch.uint8, device=self.device)
        self._final_obs = None
        self._planes = {}
        cols = dict(cell=W * H, state=W * H, pos=A * 2, hdr=4, init_cell=W * H, init_state=W * H, init_pos=A * 2)
        for name, pid in _lib.GEN_PLANES.items():
            off, nbytes, row = C.c_size_t(), C.c_size_t(), C.c_size_t()
            self._lib.mg_state_plane(self._h, pid, C.byref(off), C.byref(nbytes), C.byref(row))
            dt = torch.int32 if name == "hdr" else torch.uint8
            self._planes[name] = self.state[off.value: off.value + nbytes.value].view(dt).view(-1, cols[name])[:N]
        self.single_action_space = MultiDiscrete([4] * A)      # the four actions the base-class step defines
        self.action_space = MultiDiscrete(np.full((N, A), 4))
        self.single_observation_space = Box(0, 255, (W, H, 6), np.uint8)   # multigrid.py:105-110, encode_dim 6
        self.observation_space = Box(0, 255, (N, A, W, H, 6), np.uint8)
        self._io = _lib.StepIO()
        self._order = None
        self.closed = False

    # ------------------------------------------------------------------ layout
    def set_layout(self, obs6, agent_pos):
        """obs6: `encode_for_agents` output of the episode-start grid, u8 [W, H, 6] (shared) or [N, W, H, 6];
        agent_pos [A, 2] or [N, A, 2].  Stored in the INIT_* planes; `reset` copies them into the live state."""
        N, W, H, A = self.num_envs, self.width, self.height, self.num_agents
        o = torch.as_tensor(np.asarray(obs6), device=self.device).to(torch.uint8)
        o = o.reshape(-1, W * H, 6)
        p = torch.as_tensor(np.asarray(agent_pos), device=self.device).to(torch.uint8).reshape(-1, A * 2)
        cell = o[..., 0] | (o[..., 1] << 4)
        is_door, is_agent = o[..., 0] == DEFAULT_OBJECT_TO_IDX["door"], o[..., 0] == DEFAULT_OBJECT_TO_IDX["agent"]
        state = torch.where(is_door, o[..., 2], torch.where(is_agent, o[..., 4], torch.zeros_like(o[..., 2])))
        self._planes["init_cell"].copy_(cell.expand(N, -1))
        self._planes["init_state"].copy_(state.expand(N, -1))
        self._planes["init_pos"].copy_(p.expand(N, -1))

    @property
    def agent_pos(self):
        return self._planes["pos"].view(self.num_envs, self.num_agents, 2)

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
        self._check(self._lib.mg_reset(self._h, _ptr(self.state), _ptr(m), _ptr(self._obs), self._stream()))
        return self._obs, {}

    def set_order(self, order=None):
        """Replay recorded `np.random.permutation` outputs [N, A] (multigrid.py:402); None = Philox Fisher-Yates."""
        if order is None:
            self._lib.mg_set_trace(self._h, None)
            self._order = None
            return
        self._order = torch.as_tensor(np.ascontiguousarray(order), device=self.device).to(torch.uint8).contiguous()
        tr = _lib.Trace()
        tr.order = self._order.data_ptr()
        self._lib.mg_set_trace(self._h, C.byref(tr))

    def step(self, actions):
        a = actions if isinstance(actions, torch.Tensor) else torch.as_tensor(np.asarray(actions))
        a = a.to(self.device, non_blocking=True)
        if a.dtype != torch.int8:
            a = a.to(torch.int8)
        a = a.reshape(self.num_envs, self.num_agents).contiguous()
        io = self._io
        io.actions, io.obs, io.rewards = a.data_ptr(), self._obs.data_ptr(), self._rewards.data_ptr()
        io.terminated, io.truncated = self._term.data_ptr(), self._trunc.data_ptr()
        io.final_obs = self._final_obs.data_ptr() if self._final_obs is not None else None
        self._check(self._lib.mg_step(self._h, _ptr(self.state), C.byref(io), self._stream()))
        info = {} if self._final_obs is None else {"final_observation": self._final_obs,
                                                   "_final_observation": (self._term | self._trunc).view(torch.bool)}
        return self._obs, self._rewards, self._term.view(torch.bool), self._trunc.view(torch.bool), info

    def gen_obs(self, view_size=7, see_through_walls=False, dirs=None, out=None):
        """Partial observations (MultiGridEnv.gen_obs, multigrid.py:485-532) with encode_dim 6: u8 [N, A, V, V, 6], the
        egocentric V x V window in front of each agent (walls and closed / locked doors occlude)."""
        V = int(view_size)
        if out is None:
            out = torch.empty((self.num_envs, self.num_agents, V, V, 6), dtype=torch.uint8, device=self.device)
        d = None if dirs is None else torch.as_tensor(dirs, device=self.device).to(torch.uint8).contiguous()
        self._check(self._lib.mg_gen_obs(self._h, _ptr(self.state), _ptr(d), V, int(bool(see_through_walls)), _ptr(out), self._stream()))
        return out

    def set_state_from_obs(self, obs6, agent_pos):
        """Load a live state (not the reset snapshot) from `encode_for_agents` arrays [N, W, H, 6] + positions (validation)."""
        keep = {k: self._planes[k].clone() for k in ("init_cell", "init_state", "init_pos")}
        self.set_layout(obs6, agent_pos)
        for live, init in (("cell", "init_cell"), ("state", "init_state"), ("pos", "init_pos")):
            self._planes[live].copy_(self._planes[init])
            self._planes[init].copy_(keep[init])

    def enable_final_observation(self, enable=True):
        self._final_obs = torch.zeros_like(self._obs) if enable else None

    def status(self) -> int:
        """Sticky device error bits (`MG_ERR_*`); synchronises.  MG_ERR_BAD_ACTION: an action the base-class step
        raises for in the reference (pickup / drop / toggle / done, multigrid.py:447)."""
        st = C.c_int32(0)
        self._check(self._lib.mg_status(self._h, self._stream(), C.byref(st)))
        return st.value

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
