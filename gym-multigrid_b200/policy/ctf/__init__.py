from .heuristic import (CapturePolicy, CtfPolicy, DestinationPolicy, FightPolicy, PatrolFightPolicy,  # noqa: F401
                        PatrolPolicy, RwPolicy)
from .utils import a_star, manhattan_distance  # noqa: F401
