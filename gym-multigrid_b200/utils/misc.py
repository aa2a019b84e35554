"""`gym_multigrid.utils.misc.set_seed` (utils/misc.py:9-19): seeds numpy's legacy generator, `random` and torch - what makes
the reference's Collect envs reproducible.  The batched envs here take `seed=` (Philox, keyed by env id) instead; this is for
user code that seeds its own sampling the way the reference's scripts do.  The GIF writer (utils/misc.py:22-34) needs
matplotlib + imagemagick and is not carried over: `render()` returns the frames as arrays."""
import os
import random

import numpy as np


def set_seed(seed: int = 42) -> None:
    np.random.seed(seed)
    random.seed(seed)
    os.environ["PYTHONHASHSEED"] = str(seed)
    try:
        import torch
    except ImportError:
        return
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)
