"""The parts of `gymnasium.vector.VectorEnv`'s surface that carry no computation, shared by every batched env class of this
package (num_envs / spaces / reset / step / close live in the classes themselves)."""
from __future__ import annotations


class VectorEnvSurface:
    metadata: dict = {"render_modes": [], "autoreset_mode": "same_step"}   # gymnasium 0.29.1 semantics: reset inside the finishing step
    spec = None
    render_mode = None
    closed = False

    @property
    def unwrapped(self):
        return self

    def render(self):
        """Rendering is out of scope: observations are the interface."""
        return None

    def close_extras(self, **kwargs):
        return None

    def get_attr(self, name: str):
        """VectorEnv.get_attr: per-env values come back as one tensor / value for the whole batch (state is struct-of-arrays)."""
        return getattr(self, name)

    def set_attr(self, name: str, values):
        setattr(self, name, values)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False
