"""Batched Wildfire environments - an EXTENSION: the reference ships no Wildfire code (only a README heading,
README.md:43), BASELINE.json config 5 asks for one.  The rules are specified in include/multigrid_b200.h
("Wildfire") and DESIGN.md section 10, restated by the CPU oracle (oracle/mg_oracle_wildfire.c) and implemented
by csrc/wildfire_kernels.cu: one CTA per env, shared-memory fire-spread stencil, warp-vote move resolution."""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from .vector_base import VectorEnvSurface
from .spaces import Box, Discrete, MultiDiscrete


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def thresholds(alpha: float, beta: float):
    """Integer thresholds compared with Philox words: ignite iff u < floor(2^32 (1 - (1 - alpha)^k)), burn out iff u < floor(2^32 beta)."""
    ign = [min(2**32 - 1, int(math.floor((1.0 - (1.0 - alpha) ** k) * 2.0**32))) for k in range(5)]
    return ign, min(2**32 - 1, int(math.floor(beta * 2.0**32)))


class WildfireVecEnv(VectorEnvSurface):
    def __init__(self, num_envs, size=64, num_agents=16, agents_index=None, num_fires=4, alpha=0.15, beta=0.05, max_steps=200,
                 device="cuda:0", seed=0, autoreset=True, env_id_base=0, width=None, height=None):
        self._lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("gym-multigrid_b200 runs on CUDA devices only (no CPU fallback)")
        if not torch.cuda.is_available():
            raise RuntimeError("no CUDA device available: gym-multigrid_b200 has no CPU fallback")
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", idx)
        self.num_envs, self.width, self.height = int(num_envs), int(width or size), int(height or size)
        self.num_agents, self.max_steps, self.autoreset = int(num_agents), int(max_steps), bool(autoreset)
        self.agents_index = list(agents_index or [4] * self.num_agents)
        cfg = _lib.WildfireConfig()
        cfg.struct_size = C.sizeof(_lib.WildfireConfig)
        cfg.family, cfg.num_envs, cfg.env_id_base = _lib.FAMILY_WILDFIRE, self.num_envs, int(env_id_base)
        cfg.width, cfg.height, cfg.num_agents, cfg.num_fires = self.width, self.height, self.num_agents, int(num_fires)
        if self.num_agents > _lib.MAX_WILDFIRE_AGENTS or len(self.agents_index) != self.num_agents:
            raise ValueError("1..32 agents, one colour index per agent")
        for i, v in enumerate(self.agents_index):
            cfg.agent_colour[i] = int(v)
        ign, bo = thresholds(alpha, beta)
        for k in range(5):
            cfg.ignite_threshold[k] = ign[k]
        cfg.burnout_threshold = bo
        cfg.max_steps, cfg.autoreset, cfg.seed = self.max_steps, int(self.autoreset), int(seed) & (2**64 - 1)
        h = C.c_void_p()
        if self._lib.mg_create_wildfire(C.byref(cfg), idx, C.byref(h)) != 0:
            raise ValueError(_lib.last_error(None))
        self._h = h
        N, W, H, A = self.num_envs, self.width, self.height, self.num_agents
        with torch.cuda.device(self.device):
            self.state = torch.zeros(self._lib.mg_state_bytes(self._h), dtype=torch.uint8, device=self.device)
            self._obs = torch.zeros((N, W, H, 3), dtype=torch.uint8, device=self.device)
            self._rewards = torch.zeros((N, A), dtype=torch.float64, device=self.device)
            self._term = torch.zeros(N, dtype=torch.uint8, device=self.device)
            self._trunc = torch.zeros(N, dtype=torch.uint8, device=self.device)
        self._final_obs = None
        self._planes = {}
        for name, pid, dt, cols in (("terrain", _lib.WF_PLANE_TERRAIN, torch.uint8, W * H), ("agents", _lib.WF_PLANE_AGENTS, torch.uint8, A * 4),
                                    ("hdr", _lib.WF_PLANE_HDR, torch.int32, 4)):
            off, nbytes, row = C.c_size_t(), C.c_size_t(), C.c_size_t()
            self._lib.mg_state_plane(self._h, pid, C.byref(off), C.byref(nbytes), C.byref(row))
            self._planes[name] = self.state[off.value: off.value + nbytes.value].view(dt).view(-1, cols)[:N]
        self.single_action_space = MultiDiscrete([5] * A)
        self.action_space = MultiDiscrete(np.full((N, A), 5))
        self.single_observation_space = Box(0, 255, (W, H, 3), np.uint8)
        self.observation_space = Box(0, 255, (N, W, H, 3), np.uint8)
        self._io = _lib.StepIO()
        self._order = None
        self.closed = False

    @property
    def terrain(self):
        """u8 [N, W*H]: 0 healthy, 1 burning, 2 burnt (index x*H + y)."""
        return self._planes["terrain"]

    @property
    def agents(self):
        """u8 [N, A, 4]: x, y, dir, 0."""
        return self._planes["agents"].view(self.num_envs, self.num_agents, 4)

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
        """Replay a given per-step agent order [N, A] (validation); None = Philox Fisher-Yates."""
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

    def enable_final_observation(self, enable=True):
        self._final_obs = torch.zeros_like(self._obs) if enable else None

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
