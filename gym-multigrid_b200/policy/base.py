"""policy/base.py:13-42 of the reference: the attributes every policy carries (`name`, `action_set`,
`random_generator`); the CtF env overwrites the last two with its own (ctf.py:821-826)."""
from __future__ import annotations

import numpy as np


class BaseAgentPolicy:
    name = "base"

    def __init__(self, action_set=None, random_generator=None):
        self.action_set = action_set
        # base.py:35-39: a private default_rng() unless the caller (or, later, the env) supplies one
        self.random_generator = np.random.default_rng() if random_generator is None else random_generator

    def act(self, observation):
        raise NotImplementedError
