"""`register` / `make` stubs: a dict registry + lazy entry-point import.  `make` does
NOT add gymnasium's wrappers; TimeLimit is emulated by the golden harness itself."""
import importlib

registry = {}


def register(id, entry_point=None, max_episode_steps=None, kwargs=None, **_):
    registry[id] = dict(entry_point=entry_point, max_episode_steps=max_episode_steps,
                        kwargs=dict(kwargs or {}))


def make(id, **overrides):
    spec = registry[id.split(":")[-1]]
    mod, cls = spec["entry_point"].split(":")
    kw = dict(spec["kwargs"])
    kw.update(overrides)
    return getattr(importlib.import_module(mod), cls)(**kw)
