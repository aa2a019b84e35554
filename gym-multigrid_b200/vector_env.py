"""Batched Collect environments on one B200: the host-side mirror of the reference's
`CollectGameEnv` family (gym_multigrid/envs/collect_game.py) behind a gymnasium
VectorEnv-style surface.  All env state lives in ONE caller-owned torch uint8 tensor on the
GPU (struct-of-arrays planes, see include/multigrid_b200.h); `step` / `reset` enqueue one
hand-written sm_100a kernel each through the C ABI and never synchronise with the host.
"""
from __future__ import annotations

import ctypes as C
import enum

import numpy as np
import torch

from . import _lib
from .vector_base import VectorEnvSurface
from .spaces import Box, Discrete, MultiDiscrete

# CollectGameEnv.keys (collect_game.py:48-55)
INFO_KEYS = ["agent1ball1", "agent1ball2", "agent1ball3", "agent2ball1", "agent2ball2", "agent2ball3"]


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class CollectVecEnv(VectorEnvSurface):
    """`num_envs` independent Collect games stepped in lockstep on `device`.

    Constructor kwargs are the reference's (collect_game.py:17-72: size, num_balls, agents_index,
    balls_index, balls_reward, respawn) plus `layout` / `fixed_horizon` (which reference subclass),
    `max_episode_steps` (the registration's TimeLimit) and the vector-env additions.

    RNG.  The reference draws agent order from the global legacy numpy RNG and ball placement
    from python `random` (collect_game.py:186, multigrid.py:225-230) - `reset(seed=)` does not
    seed either.  Here: production mode = counter-based Philox4x32-10 keyed by (`seed`, global env
    id); validation mode (`set_trace`) = replay of recorded reference RNG outputs, bit-exact.
    """
    _render_family = "collect"
    metadata = {"render_modes": ["rgb_array"], "autoreset_mode": "same_step"}
    render_mode = "rgb_array"

    def __init__(self, num_envs, size=10, num_balls=15, agents_index=(3, 5), balls_index=(0, 1, 2),
                 balls_reward=(1, 1, 1), respawn=False, layout="even_dist", fixed_horizon=False,
                 max_steps=100, max_episode_steps=None, device="cuda:0", seed=0, autoreset=True,
                 env_id_base=0, width=None, height=None, host_transport="delta", host_threads=None):
        self._lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("gym-multigrid_b200 runs on CUDA devices only (no CPU fallback)")
        if not torch.cuda.is_available():
            raise RuntimeError("no CUDA device available: gym-multigrid_b200 has no CPU fallback")
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", dev_index)

        self.num_envs = int(num_envs)
        self.width = int(width or size)
        self.height = int(height or size)
        self.size = size
        self.agents_index = list(agents_index)
        self.balls_index = list(balls_index)
        self.balls_reward = [float(r) for r in balls_reward]
        self.num_agents = len(self.agents_index)
        self.num_ball_types = len(self.balls_index)
        self.num_balls = int(np.sum(np.array(num_balls)))  # collect_game.py:37
        self.respawn = bool(respawn)
        self.layout = layout
        self.max_steps = int(max_steps)
        self.max_episode_steps = max_episode_steps
        self.autoreset = bool(autoreset)
        self.keys = INFO_KEYS

        cfg = _lib.Config()
        cfg.struct_size = C.sizeof(_lib.Config)
        cfg.family = _lib.FAMILY_COLLECT
        cfg.num_envs = self.num_envs
        cfg.env_id_base = int(env_id_base)
        cfg.width, cfg.height = self.width, self.height
        cfg.num_agents, cfg.num_ball_types = self.num_agents, self.num_ball_types
        if self.num_agents > _lib.MAX_AGENTS or self.num_ball_types > _lib.MAX_BALL_TYPES:
            raise ValueError("at most 8 agents and 8 ball types")
        for i, v in enumerate(self.agents_index):
            cfg.agent_colour[i] = int(v)
        for i, v in enumerate(self.balls_index):
            cfg.ball_colour[i] = int(v)
        for i, v in enumerate(self.balls_reward):
            cfg.ball_reward[i] = v
        cfg.num_balls = self.num_balls
        cfg.respawn = int(self.respawn)
        cfg.layout = _lib.LAYOUTS[layout]
        cfg.fixed_horizon = int(bool(fixed_horizon))
        cfg.max_steps = self.max_steps
        cfg.time_limit = int(max_episode_steps or 0)
        cfg.autoreset = int(self.autoreset)
        cfg.seed = int(seed) & (2**64 - 1)
        self._cfg = cfg
        handle = C.c_void_p()
        if self._lib.mg_create(C.byref(cfg), dev_index, C.byref(handle)) != 0:
            raise ValueError(_lib.last_error(None))
        self._h = handle

        N, W, H, A, nb = self.num_envs, self.width, self.height, self.num_agents, self.num_ball_types
        with torch.cuda.device(self.device):
            self.state = torch.zeros(self._lib.mg_state_bytes(self._h), dtype=torch.uint8, device=self.device)
            self._obs = torch.zeros((N, W, H, 3), dtype=torch.uint8, device=self.device)
            self._rewards = torch.zeros((N, A), dtype=torch.float64, device=self.device)
            self._term = torch.zeros(N, dtype=torch.uint8, device=self.device)
            self._trunc = torch.zeros(N, dtype=torch.uint8, device=self.device)
            self._final_obs = None
        self._planes = {}
        for name, pid, dt, cols in (("grid", _lib.PLANE_GRID, torch.uint8, W * H),
                                    ("agent_pos", _lib.PLANE_AGENT_POS, torch.uint8, A * 2),
                                    ("hdr", _lib.PLANE_HDR, torch.int32, 4),
                                    ("info", _lib.PLANE_INFO, torch.int32, A * nb)):
            off, nbytes, row = C.c_size_t(), C.c_size_t(), C.c_size_t()
            self._lib.mg_state_plane(self._h, pid, C.byref(off), C.byref(nbytes), C.byref(row))
            flat = self.state[off.value: off.value + nbytes.value].view(dt)
            self._planes[name] = flat.view(-1, cols)[:N]

        self.single_action_space = Discrete(4)  # CollectActions (agent.py:32-36; multigrid.py:66)
        self.single_observation_space = Box(0, 255, (W, H, 3), np.uint8)  # multigrid.py:105-110
        self.action_space = MultiDiscrete(np.full((N, A), 4))
        self.observation_space = Box(0, 255, (N, W, H, 3), np.uint8)
        self._host = None
        self._trace_keepalive = None
        self._io = _lib.StepIO()
        self.closed = False
        self._bind_io()
        self.set_host_transport(host_transport, host_threads)

    def set_host_transport(self, mode="delta", host_threads=None):
        """What `step(numpy)` / `step_async` move over PCIe for the observation (mg_set_host_transport): "full" = the expanded
        (W, H, 3) array (300 B per 10x10 env), "packed" = the 1-byte-per-cell grid plane expanded on the host, "delta" (default) =
        a 16-byte record of the cells the step changed, patched into the page-locked observation buffer this env returns (plus
        the packed rows of the envs that autoreset).  The results are identical; only the bytes on the wire differ.
        `host_threads`: host threads of the decoder (default: this process's cores / LOCAL_WORLD_SIZE)."""
        import os
        if host_threads is None:
            try:
                cores = len(os.sched_getaffinity(0))
            except AttributeError:
                cores = os.cpu_count() or 1
            host_threads = max(1, min(32, cores // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))))
        if getattr(self, "_host_pending", False):
            raise RuntimeError("set_host_transport while a step_async is in flight")
        self._check(self._lib.mg_set_host_transport(self._h, _lib.TRANSPORTS[mode], int(host_threads)))
        self.host_transport, self.host_threads = mode, int(host_threads)

    def _bind_io(self):
        """Everything `step` needs that does not change from call to call: the output pointers of the io block, the bool views
        of the flag buffers, the info dict.  (A step call costs ~5 us on the host; at small batches that is the bottleneck.)"""
        io = self._io
        io.obs, io.rewards = self._obs.data_ptr(), self._rewards.data_ptr()
        io.terminated, io.truncated = self._term.data_ptr(), self._trunc.data_ptr()
        io.final_obs = self._final_obs.data_ptr() if self._final_obs is not None else None
        self._io_ref, self._state_ptr = C.byref(io), _ptr(self.state)
        self._term_b, self._trunc_b = self._term.view(torch.bool), self._trunc.view(torch.bool)
        self._info_static = {"pickups": self.pickups}
        self._shape = (self.num_envs, self.num_agents)
        self._dev_index = self.device.index
        self._raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)

    # ------------------------------------------------------------------ state views (zero copy)
    @property
    def grid(self):
        """Packed cells u8 [N, W*H] (index x*H + y): type | colour << 2 | state << 6."""
        return self._planes["grid"]

    @property
    def agent_pos(self):
        return self._planes["agent_pos"].view(self.num_envs, self.num_agents, 2)

    @property
    def step_count(self):
        return self._planes["hdr"][:, 0]

    @property
    def collected_balls(self):
        return self._planes["hdr"][:, 1]

    @property
    def episode_count(self):
        return self._planes["hdr"][:, 3]

    @property
    def pickups(self):
        """info counters i32 [N, A, num_ball_types] (reference: env.info[keys[nb*i + colour]])."""
        return self._planes["info"].view(self.num_envs, self.num_agents, self.num_ball_types)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(_lib.last_error(self._h))

    # ------------------------------------------------------------------------------ gym API
    def reset(self, *, seed=None, options=None, mask=None):
        """Re-generate every env (or those selected by `mask`); returns (obs[N,W,H,3] u8 cuda, info).

        `seed` re-keys nothing in the reference either (its Collect RNGs are global); it is accepted
        for signature parity.  Use the constructor's `seed` to key the Philox streams."""
        m = None
        if mask is not None:
            m = torch.as_tensor(mask, device=self.device).to(torch.uint8).contiguous()
        self._check(self._lib.mg_reset(self._h, _ptr(self.state), _ptr(m), _ptr(self._obs), self._stream()))
        return self._obs, self._info()

    def step(self, actions):
        """actions: int8 CUDA tensor [N, A] (other integer dtypes are converted), values 0..3 =
        north/east/south/west, anything else is a no-op exactly as in the reference.
        Returns (obs u8 [N,W,H,3], rewards f64 [N,A], terminated bool [N], truncated bool [N], info),
        all CUDA tensors, valid until the next step/reset (they are reused output buffers)."""
        if not isinstance(actions, torch.Tensor):
            return self.step_host(actions)
        a = actions
        if not (a.dtype is torch.int8 and a.is_cuda and a.shape == self._shape and a.is_contiguous() and a.device == self.device):
            if a.device != self.device:
                a = a.to(self.device, non_blocking=True)
            if a.dtype != torch.int8:
                a = a.to(torch.int8)
            a = a.reshape(self._shape).contiguous()
        self._io.actions = a.data_ptr()
        stream = self._raw_stream(self._dev_index) if self._raw_stream else torch.cuda.current_stream(self.device).cuda_stream
        if self._lib.mg_step(self._h, self._state_ptr, self._io_ref, stream):
            raise RuntimeError(_lib.last_error(self._h))
        return self._obs, self._rewards, self._term_b, self._trunc_b, (self._info_static if self._final_obs is None else self._info())

    def rollout(self, actions=None, steps=None, obs=True, final_observation=False, out=None):
        """T env steps in ONE kernel launch with the env state held in shared memory (mg_rollout): the loop
        `for t in range(T): env.step(actions[t])` of the reference (collect_game.py:183-214), bit-identical to T `step` calls
        incl. same-step autoresets.  `actions`: int8 CUDA tensor [T, N, A]; or None with `steps=T` for the uniform random policy
        drawn on the device (the actions taken come back as info["actions"]).  `obs=False` skips the observations (planners
        that score action sequences by reward).  Returns (obs u8 [T,N,W,H,3] | None, rewards f64 [T,N,A], terminated bool [T,N],
        truncated bool [T,N], info); the tensors are reused by the next rollout of the same length unless `out` (a dict with
        the keys obs / rewards / terminated / truncated [/ final_obs / actions]) supplies them."""
        N, A, W, H = self.num_envs, self.num_agents, self.width, self.height
        if actions is not None:
            a = actions
            if not (isinstance(a, torch.Tensor) and a.dtype is torch.int8 and a.is_cuda and a.is_contiguous() and a.device == self.device):
                a = torch.as_tensor(a).to(self.device, non_blocking=True).to(torch.int8).contiguous()
            a = a.reshape(-1, N, A)
            T = a.shape[0]
        else:
            if steps is None:
                raise ValueError("rollout: pass actions [T, N, A] or steps=T (uniform random policy on the device)")
            a, T = None, int(steps)
        key = (T, bool(obs), bool(final_observation), a is None)
        b = out if out is not None else getattr(self, "_roll_bufs", {}).get(key)
        if b is None:
            with torch.cuda.device(self.device):
                b = dict(rewards=torch.empty((T, N, A), dtype=torch.float64, device=self.device),
                         terminated=torch.empty((T, N), dtype=torch.uint8, device=self.device),
                         truncated=torch.empty((T, N), dtype=torch.uint8, device=self.device))
                if obs:
                    b["obs"] = torch.empty((T, N, W, H, 3), dtype=torch.uint8, device=self.device)
                if final_observation:
                    b["final_obs"] = torch.zeros((T, N, W, H, 3), dtype=torch.uint8, device=self.device)
                if a is None:
                    b["actions"] = torch.empty((T, N, A), dtype=torch.int8, device=self.device)
            self._roll_bufs = {key: b}      # one cached set: the last shape used
        io = _lib.RolloutIO()
        io.steps, io.actions = T, (a.data_ptr() if a is not None else None)
        io.obs = b["obs"].data_ptr() if obs else None
        io.rewards, io.terminated, io.truncated = b["rewards"].data_ptr(), b["terminated"].data_ptr(), b["truncated"].data_ptr()
        io.final_obs = b["final_obs"].data_ptr() if final_observation else None
        io.actions_out = b["actions"].data_ptr() if (a is None and "actions" in b) else None
        self._check(self._lib.mg_rollout(self._h, self._state_ptr, C.byref(io), self._stream()))
        info = {"pickups": self.pickups}
        if a is None:
            info["actions"] = b.get("actions")
        if final_observation:
            info["final_observation"] = b["final_obs"]
            info["_final_observation"] = (b["terminated"] | b["truncated"]).view(torch.bool)
        return (b.get("obs") if obs else None), b["rewards"], b["terminated"].view(torch.bool), b["truncated"].view(torch.bool), info

    def _host_io(self, actions):
        if self._host is None:
            N, W, H, A = self.num_envs, self.width, self.height, self.num_agents
            blk, obs, rew, term, trunc = _lib.host_result_buffers(self._lib, self._h, (N, W, H, 3), torch.uint8, N, A)
            self._host = dict(act=torch.zeros((N, A), dtype=torch.int8, pin_memory=True), obs=obs, rew=rew.view(N, A), term=term,
                              trunc=trunc, block=blk)
            self._host_np = {k: v.numpy() for k, v in self._host.items()}
            h, io = self._host, _lib.StepIO()
            io.actions, io.obs, io.rewards = h["act"].data_ptr(), h["obs"].data_ptr(), h["rew"].data_ptr()
            io.terminated, io.truncated, io.final_obs = h["term"].data_ptr(), h["trunc"].data_ptr(), None
            self._host_io_struct = io
        self._host_np["act"][...] = np.asarray(actions).reshape(self.num_envs, self.num_agents)
        return self._host_io_struct

    def _host_result(self):
        n = self._host_np
        if "term_b" not in n:
            n["term_b"], n["trunc_b"] = n["term"].view(np.bool_), n["trunc"].view(np.bool_)
        return n["obs"], n["rew"], n["term_b"], n["trunc_b"], (self._info_static if self._final_obs is None else self._info())

    def step_host(self, actions):
        """gymnasium-style call with HOST arrays: numpy in, numpy out (views of page-locked buffers,
        overwritten by the next call).  Host<->device copies and the wait are inside the call.
        `step_async` / `step_wait` (vector_base.py) are its two halves."""
        io = self._host_io(actions)
        self._check(self._lib.mg_step_host(self._h, _ptr(self.state), C.byref(io), self._stream()))
        return self._host_result()

    def encode(self, out=None):
        """Grid.encode() of the current state (grid.py:223-252) -> u8 [N, W, H, 3]."""
        out = self._obs if out is None else out
        self._check(self._lib.mg_encode(self._h, _ptr(self.state), _ptr(out), self._stream()))
        return out

    def gen_obs(self, view_size=7, see_through_walls=False, dirs=None, out=None):
        """Partial observations (MultiGridEnv.gen_obs, multigrid.py:485-532): u8 [N, A, V, V, 3], the egocentric
        V x V window in front of each agent with MiniGrid-style occlusion.  `dirs` [N, A] overrides the agents'
        directions (Collect agents always face 3 = up)."""
        V = int(view_size)
        if out is None:
            out = torch.empty((self.num_envs, self.num_agents, V, V, 3), dtype=torch.uint8, device=self.device)
        d = None if dirs is None else torch.as_tensor(dirs, device=self.device).to(torch.uint8).contiguous()
        self._check(self._lib.mg_gen_obs(self._h, _ptr(self.state), _ptr(d), V, int(bool(see_through_walls)), _ptr(out), self._stream()))
        return out

    def toroid_obs(self, out=None):
        """ToroidObservation (wrappers/toroid.py:28-68) of the current state: float32 [N, A, W, H, nb + A]."""
        if out is None:
            out = torch.empty((self.num_envs, self.num_agents, self.width, self.height, self.num_ball_types + self.num_agents),
                              dtype=torch.float32, device=self.device)
        self._check(self._lib.mg_toroid_obs(self._h, _ptr(self.state), _ptr(out), self._stream()))
        return out

    def _info(self):
        info = {"pickups": self.pickups}
        if self._final_obs is not None:
            info["final_observation"] = self._final_obs
            info["_final_observation"] = (self._term | self._trunc).view(torch.bool)
        return info

    def enable_final_observation(self, enable=True):
        """gymnasium 0.29.1 same-step autoreset hands the terminal obs out as info['final_observation']."""
        if enable and self._final_obs is None:
            self._final_obs = torch.zeros_like(self._obs)
        elif not enable:
            self._final_obs = None
        self._bind_io()

    # ------------------------------------------------------------------- validation / state
    def set_trace(self, order=None, draws=None, n_draws=None, reset_draws=None, n_reset_draws=None):
        """Validation mode: replay RNG outputs recorded from the reference (see include: mg_trace).
        Call with no arguments to return to Philox mode."""
        if order is None and draws is None and reset_draws is None:
            self._lib.mg_set_trace(self._h, None)
            self._trace_keepalive = None
            return None
        N = self.num_envs

        def dev(x, dt):
            return None if x is None else torch.as_tensor(np.ascontiguousarray(x), device=self.device).to(dt).contiguous()

        t = dict(order=dev(order, torch.uint8), draws=dev(draws, torch.uint8), n_draws=dev(n_draws, torch.int32),
                 reset_draws=dev(reset_draws, torch.uint8), n_reset_draws=dev(n_reset_draws, torch.int32),
                 draws_used=torch.zeros(N, dtype=torch.int32, device=self.device),
                 reset_draws_used=torch.zeros(N, dtype=torch.int32, device=self.device))
        tr = _lib.Trace()
        tr.order, tr.draws, tr.n_draws = (x.data_ptr() if x is not None else None for x in (t["order"], t["draws"], t["n_draws"]))
        tr.K = 0 if t["draws"] is None else t["draws"].reshape(N, -1).shape[1]
        tr.reset_draws = t["reset_draws"].data_ptr() if t["reset_draws"] is not None else None
        tr.n_reset_draws = t["n_reset_draws"].data_ptr() if t["n_reset_draws"] is not None else None
        tr.R = 0 if t["reset_draws"] is None else t["reset_draws"].reshape(N, -1).shape[1]
        tr.draws_used, tr.reset_draws_used = t["draws_used"].data_ptr(), t["reset_draws_used"].data_ptr()
        self._lib.mg_set_trace(self._h, C.byref(tr))
        self._trace_keepalive = t
        return t

    def set_state_from_obs(self, obs, agent_pos, step_count=0):
        """Inject states given Grid.encode() arrays [N,W,H,3] and agent positions [N,A,2]
        (e.g. recorded reference resets)."""
        o = torch.as_tensor(np.asarray(obs), device=self.device).to(torch.uint8).reshape(self.num_envs, -1, 3)
        self.grid.copy_(o[..., 0] | (o[..., 1] << 2) | (o[..., 2] << 6))
        self.agent_pos.copy_(torch.as_tensor(np.asarray(agent_pos), device=self.device).to(torch.uint8).reshape(self.agent_pos.shape))
        self._planes["hdr"][:, 0] = step_count
        self._planes["hdr"][:, 1] = 0
        self._planes["info"].zero_()
        self._lib.mg_host_invalidate(self._h)

    def get_state(self):
        return self.state.clone()

    def set_state(self, state):
        self.state.copy_(state)
        self._lib.mg_host_invalidate(self._h)      # the delta transport's host mirror no longer matches

    def status(self) -> int:
        """Read-and-clear the device error word (MG_ERR_* bits); synchronises the stream."""
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
            self._h = None
            self.closed = True

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


class CollectActions(enum.IntEnum):   # core/agent.py:32-36
    north = 0
    east = 1
    south = 2
    west = 3


class _AgentView:
    """What callers read from `env.agents[i]` (core/agent.py:92-100): index, pos, dir."""

    def __init__(self, env, i):
        self._env, self.i = env, i
        self.index, self.dir = env.vec.agents_index[i], 3        # Collect agents never turn (multigrid.py:371-374)

    @property
    def pos(self):
        return self._env.vec.agent_pos[0, self.i].cpu().numpy()


class CollectEnv:
    """Single-env adaptor with the reference's own signatures (collect_game.py:107-119,183-214):
    `reset(seed=, options=) -> (obs ndarray (W,H,3) uint8, info dict)`,
    `step(actions) -> (obs, rewards float64 (A,), terminated bool, truncated bool, info dict)`.
    It is one env of a CollectVecEnv (num_envs=1, no autoreset); truncation includes the
    registration's TimeLimit exactly as `gymnasium.make` would apply it."""

    def __init__(self, **kwargs):
        kwargs.setdefault("autoreset", False)
        self.vec = CollectVecEnv(1, **kwargs)
        v = self.vec
        self.width, self.height, self.num_ball_types = v.width, v.height, v.num_ball_types
        self.action_space, self.observation_space = v.single_action_space, v.single_observation_space
        self.keys, self.max_steps, self.respawn, self.num_balls = v.keys, v.max_steps, v.respawn, v.num_balls
        self.agents_index = v.agents_index
        self.actions_set = CollectActions
        self.agents = [_AgentView(self, i) for i in range(v.num_agents)]      # `for a in env.agents` (tests/test_collect.py:16)

    @property
    def step_count(self):
        return int(self.vec.step_count[0])

    @property
    def collected_balls(self):
        return int(self.vec.collected_balls[0])

    @property
    def agent_positions(self):
        return self.vec.agent_pos[0].cpu().numpy()

    def _info(self):
        c = self.vec.pickups[0].reshape(-1).cpu().numpy()
        return {k: int(c[i]) if i < len(c) else 0 for i, k in enumerate(self.keys)}

    def reset(self, *, seed=None, options=None):
        obs, _ = self.vec.reset(seed=seed, options=options)
        return obs[0].cpu().numpy(), self._info()

    def step(self, actions):
        a = torch.as_tensor(np.asarray(actions, dtype=np.int64).clip(-128, 127).astype(np.int8)).reshape(1, -1)
        obs, rew, term, trunc, _ = self.vec.step(a.to(self.vec.device))
        return (obs[0].cpu().numpy(), rew[0].cpu().numpy(), bool(term[0]), bool(trunc[0]), self._info())

    def render(self, close=False, highlight=False, tile_size=32):
        """MultiGridEnv.render (multigrid.py:546-606): the frame as ndarray (H * tile_size, W * tile_size, 3) uint8."""
        if close:
            return None
        if highlight:
            raise NotImplementedError("render(highlight=True) is not available (the default is off)")
        return self.vec.render(tile_size=tile_size)[0].cpu().numpy()

    def close(self):
        self.vec.close()
