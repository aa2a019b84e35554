"""Minimal stand-in for the `gymnasium` package (TEST INFRASTRUCTURE ONLY).

The reference (Tran-Research-Group/gym-multigrid) imports gymnasium 0.29.1, which is
not installed in this image and cannot be fetched (no network).  This stub provides
just the surface the reference touches (multigrid.py:5-6,21,66,91-112,114-128;
envs/ctf.py:4; envs/maze.py:3; wrappers/toroid.py:2-3; __init__.py:1) so the
UNMODIFIED reference sources under /root/reference can be imported and executed by
oracle/gen_golden.py to record golden traces.  Nothing in the product path imports it.

Seeding follows gymnasium 0.29.1 `utils.seeding.np_random`:
Generator(PCG64(SeedSequence(seed))).
"""
import numpy as np

from . import spaces  # noqa: F401


class Env:
    metadata = {}
    _np_random = None
    # test hook: oracle/ref_harness.py sets this to wrap every generator in a recording proxy
    wrap_generator = staticmethod(lambda g: g)
    # test hook: entropy of lazily created (unseeded) generators; None = OS entropy as in gymnasium.  The recorders set it
    # so that the committed fixtures are reproducible (the reference's scripted-policy generators are never seeded).
    unseeded_entropy = None

    @property
    def np_random(self):
        if self._np_random is None:
            ss = np.random.SeedSequence() if Env.unseeded_entropy is None else np.random.SeedSequence(Env.unseeded_entropy)
            self._np_random = Env.wrap_generator(np.random.Generator(np.random.PCG64(ss)))
        return self._np_random

    @np_random.setter
    def np_random(self, value):
        self._np_random = value

    def reset(self, *, seed=None, options=None):
        if seed is not None:
            self._np_random = Env.wrap_generator(np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed))))

    def close(self):
        pass

    @property
    def unwrapped(self):
        return self


class Wrapper(Env):
    def __init__(self, env):
        self.env = env

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.env, name)

    @property
    def unwrapped(self):
        return self.env.unwrapped


class ObservationWrapper(Wrapper):
    def reset(self, *, seed=None, options=None):
        obs, info = self.env.reset(seed=seed, options=options)
        return self.observation(obs), info

    def step(self, action):
        obs, reward, terminated, truncated, info = self.env.step(action)
        return self.observation(obs), reward, terminated, truncated, info

    def observation(self, observation):
        raise NotImplementedError


from .envs.registration import make, register, registry  # noqa: E402,F401
