"""`from gym_multigrid_b200.envs.ctf import Ctf1v1Env, CtFMvNEnv` (reference: envs/ctf.py:50-654, 657-1433; imported that way by
tests/test_ctf.py:8 and scripts/main_mvn_ctf_rl.py:7)."""
from ..map_env import Ctf1v1VecEnv, CtfVecEnv  # noqa: F401
from ..single_env import Ctf1v1Env, CtFMvNEnv  # noqa: F401
