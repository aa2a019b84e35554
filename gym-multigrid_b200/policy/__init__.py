"""Host-side agent policies that call into / are called by the CtF step (reference: `gym_multigrid/policy/`).

These are CALLERS of the hot path, not part of it ("A*-based policies stay host-side inputs"):
the CtF adaptor calls `policy.act(observation_dict, curr_pos)` where the reference does (ctf.py:1297-1301) and hands
the resulting red actions to the CUDA step through `mg_set_red_actions`.  Pure numpy; no torch, no CUDA."""
from .base import BaseAgentPolicy  # noqa: F401
