"""ctypes binding of include/multigrid_b200.h (the C ABI of libmultigrid_b200.so).

There is no CPU fallback: loading fails loudly when the library has not been built, and
`mg_create` fails when no sm_100 CUDA device is usable.
"""
from __future__ import annotations

import ctypes as C
import os

from ._build import LIB_PATH

MAX_AGENTS = 8
MAX_BALL_TYPES = 8
FAMILY_COLLECT, FAMILY_MAZE, FAMILY_CTF, FAMILY_WILDFIRE, FAMILY_GENERIC = 0, 1, 2, 3, 4
GEN_PLANES = dict(cell=0, state=1, pos=2, hdr=3, init_cell=4, init_state=5, init_pos=6)
WF_PLANE_TERRAIN, WF_PLANE_AGENTS, WF_PLANE_HDR = 0, 1, 2
MAX_WILDFIRE_AGENTS = 32
OBS_U8, OBS_REFERENCE = 0, 1
MAP_PLANE_AGENTS, MAP_PLANE_HDR = 0, 1
ERR_BAD_ACTION = 8
MAX_MAP_AGENTS = 16
LAYOUTS = {"even_dist": 0, "quadrants": 1, "rooms": 2, "quadrants_respawn": 3}
PLANE_GRID, PLANE_AGENT_POS, PLANE_HDR, PLANE_INFO = 0, 1, 2, 3
ERR_TRACE_OVERFLOW, ERR_TRACE_RANGE, ERR_OOB = 1, 2, 4

EXPORTS = [
    "mg_abi_version", "mg_create", "mg_destroy", "mg_last_error", "mg_state_bytes", "mg_obs_bytes",
    "mg_state_plane", "mg_reset", "mg_step", "mg_encode", "mg_step_host", "mg_set_trace", "mg_status",
    "mg_launch_count", "mg_debug_set_timeline", "mg_tile_envs", "mg_create_map", "mg_set_map_trace", "mg_gen_obs", "mg_toroid_obs", "mg_create_wildfire", "mg_create_generic", "mg_map_info", "mg_set_partial_obs", "mg_host_layout", "mg_set_red_actions", "mg_step_host_async", "mg_step_host_wait", "mg_render", "mg_ctf_flat_len", "mg_ctf_flat_obs", "mg_ctf_flat_obs_u8", "mg_set_seed",
    "mg_set_red_policies", "mg_red_policy_actions", "mg_set_red_policy_fusion", "mg_set_carry_agent_flags", "mg_astar_first_moves",
    "mg_set_policy_trace", "mg_rollout", "mg_set_host_transport", "mg_host_invalidate", "mg_host_expand_plane", "mg_delta_record_bytes", "mg_host_apply_delta", "mg_stream_idle",
]
TRANSPORTS = {"full": 0, "packed": 1, "delta": 2}


class Config(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("family", C.c_int32), ("num_envs", C.c_int64), ("env_id_base", C.c_int64),
        ("width", C.c_int32), ("height", C.c_int32), ("num_agents", C.c_int32), ("num_ball_types", C.c_int32),
        ("agent_colour", C.c_int32 * MAX_AGENTS), ("ball_colour", C.c_int32 * MAX_BALL_TYPES),
        ("ball_reward", C.c_double * MAX_BALL_TYPES), ("num_balls", C.c_int32), ("respawn", C.c_int32),
        ("layout", C.c_int32), ("fixed_horizon", C.c_int32), ("max_steps", C.c_int32), ("time_limit", C.c_int32),
        ("autoreset", C.c_int32), ("seed", C.c_uint64),
    ]


class StepIO(C.Structure):
    _fields_ = [("actions", C.c_void_p), ("obs", C.c_void_p), ("rewards", C.c_void_p), ("terminated", C.c_void_p),
                ("truncated", C.c_void_p), ("final_obs", C.c_void_p)]


class RolloutIO(C.Structure):
    _fields_ = [("steps", C.c_int32), ("actions", C.c_void_p), ("obs", C.c_void_p), ("rewards", C.c_void_p), ("terminated", C.c_void_p),
                ("truncated", C.c_void_p), ("final_obs", C.c_void_p), ("actions_out", C.c_void_p)]


class Trace(C.Structure):
    _fields_ = [("order", C.c_void_p), ("draws", C.c_void_p), ("n_draws", C.c_void_p), ("K", C.c_int32),
                ("reset_draws", C.c_void_p), ("n_reset_draws", C.c_void_p), ("R", C.c_int32),
                ("draws_used", C.c_void_p), ("reset_draws_used", C.c_void_p)]


class MapConfig(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("family", C.c_int32), ("num_envs", C.c_int64), ("env_id_base", C.c_int64),
        ("size", C.c_int32), ("field_map", C.c_void_p), ("num_blue", C.c_int32), ("num_red", C.c_int32),
        ("flag_reward", C.c_double), ("battle_reward", C.c_double), ("obstacle_penalty", C.c_double),
        ("step_penalty", C.c_double), ("battle_range", C.c_double), ("randomness", C.c_double),
        ("max_steps", C.c_int32), ("autoreset", C.c_int32), ("obs_dtype", C.c_int32), ("variant_1v1", C.c_int32),
        ("seed", C.c_uint64),
    ]


class WildfireConfig(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("family", C.c_int32), ("num_envs", C.c_int64), ("env_id_base", C.c_int64),
        ("width", C.c_int32), ("height", C.c_int32), ("num_agents", C.c_int32), ("agent_colour", C.c_int32 * 32),
        ("num_fires", C.c_int32), ("ignite_threshold", C.c_uint32 * 5), ("burnout_threshold", C.c_uint32),
        ("max_steps", C.c_int32), ("autoreset", C.c_int32), ("seed", C.c_uint64),
    ]


class GenericConfig(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("family", C.c_int32), ("num_envs", C.c_int64), ("env_id_base", C.c_int64),
                ("width", C.c_int32), ("height", C.c_int32), ("num_agents", C.c_int32), ("max_steps", C.c_int32),
                ("autoreset", C.c_int32), ("seed", C.c_uint64)]


class MapTrace(C.Structure):
    _fields_ = [("start_index", C.c_void_p), ("blue_place", C.c_void_p), ("red_place", C.c_void_p),
                ("red_actions", C.c_void_p), ("order", C.c_void_p), ("blue_win", C.c_void_p), ("KB", C.c_int32),
                ("battles_used", C.c_void_p)]


class RedPolicies(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("num_red", C.c_int32), ("kind", C.c_int32 * 16), ("randomness", C.c_double * 16),
                ("first_move", C.c_void_p), ("patrol_goal", C.c_void_p), ("on_border", C.c_void_p), ("along_border", C.c_void_p),
                ("n_along", C.c_int32)]


_lib = None


class LibraryMissing(RuntimeError):
    pass


def load():
    """Load libmultigrid_b200.so (built in-tree by __graft_entry__.build / _build.build_library)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LibraryMissing(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  gym-multigrid_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.mg_abi_version.restype = C.c_int
    lib.mg_create.restype = C.c_int
    lib.mg_create.argtypes = [C.POINTER(Config), C.c_int, C.POINTER(C.c_void_p)]
    lib.mg_destroy.argtypes = [C.c_void_p]
    lib.mg_last_error.restype = C.c_char_p
    lib.mg_last_error.argtypes = [C.c_void_p]
    lib.mg_state_bytes.restype = C.c_size_t
    lib.mg_state_bytes.argtypes = [C.c_void_p]
    lib.mg_obs_bytes.restype = C.c_size_t
    lib.mg_obs_bytes.argtypes = [C.c_void_p]
    lib.mg_state_plane.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    lib.mg_reset.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.mg_step.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(StepIO), C.c_void_p]
    lib.mg_step_host.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(StepIO), C.c_void_p]
    lib.mg_step_host_async.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(StepIO), C.c_void_p]
    lib.mg_step_host_wait.argtypes = [C.c_void_p, C.c_void_p]
    lib.mg_ctf_flat_len.argtypes = [C.c_void_p]
    lib.mg_ctf_flat_obs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.mg_ctf_flat_obs_u8.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.mg_set_seed.argtypes = [C.c_void_p, C.c_uint64]
    lib.mg_render.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.mg_encode.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.mg_set_trace.argtypes = [C.c_void_p, C.POINTER(Trace)]
    lib.mg_status.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int32)]
    lib.mg_debug_set_timeline.argtypes = [C.c_void_p, C.c_void_p]
    lib.mg_tile_envs.argtypes = [C.c_void_p]
    lib.mg_create_map.restype = C.c_int
    lib.mg_create_map.argtypes = [C.POINTER(MapConfig), C.c_int, C.POINTER(C.c_void_p)]
    lib.mg_set_map_trace.argtypes = [C.c_void_p, C.POINTER(MapTrace)]
    lib.mg_gen_obs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.mg_toroid_obs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.mg_create_wildfire.restype = C.c_int
    lib.mg_create_wildfire.argtypes = [C.POINTER(WildfireConfig), C.c_int, C.POINTER(C.c_void_p)]
    lib.mg_create_generic.restype = C.c_int
    lib.mg_create_generic.argtypes = [C.POINTER(GenericConfig), C.c_int, C.POINTER(C.c_void_p)]
    lib.mg_set_red_actions.restype = C.c_int
    lib.mg_set_red_actions.argtypes = [C.c_void_p, C.c_void_p]
    lib.mg_set_red_policies.restype = C.c_int
    lib.mg_set_red_policies.argtypes = [C.c_void_p, C.POINTER(RedPolicies)]
    lib.mg_set_carry_agent_flags.restype = C.c_int
    lib.mg_set_carry_agent_flags.argtypes = [C.c_void_p, C.c_int]
    lib.mg_set_red_policy_fusion.restype = C.c_int
    lib.mg_set_red_policy_fusion.argtypes = [C.c_void_p, C.c_void_p]
    lib.mg_red_policy_actions.restype = C.c_int
    lib.mg_red_policy_actions.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.mg_astar_first_moves.restype = C.c_int
    lib.mg_astar_first_moves.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
    lib.mg_host_layout.restype = C.c_int
    lib.mg_host_layout.argtypes = [C.c_void_p] + [C.POINTER(C.c_size_t)] * 4
    lib.mg_set_partial_obs.restype = C.c_int
    lib.mg_set_partial_obs.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.mg_map_info.restype = C.c_int
    lib.mg_map_info.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.mg_set_policy_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.mg_rollout.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(RolloutIO), C.c_void_p]
    lib.mg_set_host_transport.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.mg_host_invalidate.argtypes = [C.c_void_p]
    lib.mg_host_expand_plane.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    lib.mg_delta_record_bytes.argtypes = [C.c_int, C.c_int]
    lib.mg_host_apply_delta.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_int]
    lib.mg_stream_idle.argtypes = [C.c_void_p]
    lib.mg_launch_count.restype = C.c_int64
    lib.mg_launch_count.argtypes = [C.c_void_p]
    if lib.mg_abi_version() != 1:
        raise RuntimeError("libmultigrid_b200.so ABI version mismatch")
    _lib = lib
    return lib


def last_error(handle=None) -> str:
    msg = load().mg_last_error(handle)
    return msg.decode() if msg else ""


def host_result_buffers(lib, handle, obs_shape, obs_dtype, n, rew_cols):
    """Page-locked host buffers for mg_step_host laid out as mg_host_layout asks (one block -> one D2H copy per step):
    returns (block, obs, rewards, terminated, truncated) torch tensors sharing the block's memory."""
    import torch
    offs = [C.c_size_t() for _ in range(4)]
    if lib.mg_host_layout(handle, *[C.byref(o) for o in offs]) != 0:
        raise RuntimeError("mg_host_layout failed")
    off_r, off_t, off_u, total = (o.value for o in offs)
    block = torch.zeros(total, dtype=torch.uint8, pin_memory=True)
    nobs = int(torch.tensor(obs_shape).prod()) * torch.empty((), dtype=obs_dtype).element_size()
    obs = block[:nobs].view(obs_dtype).view(*obs_shape)
    rew = block[off_r: off_r + n * rew_cols * 8].view(torch.float64)
    rew = rew.view(n, rew_cols) if rew_cols > 1 else rew.view(n)
    return block, obs, rew, block[off_t: off_t + n], block[off_u: off_u + n]
