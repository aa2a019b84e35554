"""The Collect family (reference: envs/collect_game.py).  The reference's classes differ only in `_gen_grid`; here one class takes
the layout as an argument and `gym_multigrid_b200.make(id)` / `make_vec(id, n)` pick it from the registered id
(registration.COLLECT_CLASSES maps the reference's class names to layouts)."""
from ..registration import COLLECT_CLASSES  # noqa: F401
from ..vector_env import CollectEnv, CollectVecEnv  # noqa: F401
