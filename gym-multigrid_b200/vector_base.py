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

    def render(self, env_ids=None, tile_size=32, out=None):
        """`MultiGridEnv.render()` frames (rgb_array, highlight off; multigrid.py:546-606, Grid.render grid.py:183-221) of the
        envs listed in `env_ids` (default: all) -> u8 CUDA tensor [n, H * tile_size, W * tile_size, 3]; one blit kernel over a
        per-code tile atlas (mg_render).  Collect, Maze and CtF families; the other classes return None (no renderer)."""
        import ctypes as C

        import torch
        if getattr(self, "_render_family", None) is None:
            return None
        ids = None
        n = self.num_envs
        if env_ids is not None:
            ids = torch.as_tensor(env_ids, device=self.device).to(torch.int32).reshape(-1).contiguous()
            n = ids.numel()
        ts = int(tile_size)
        shape = (n, self.height * ts, self.width * ts, 3)
        if out is None:
            out = torch.empty(shape, dtype=torch.uint8, device=self.device)
        elif tuple(out.shape) != shape or out.dtype != torch.uint8 or not out.is_contiguous():
            raise ValueError(f"render: out must be a contiguous uint8 tensor of shape {shape}")
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        rc = self._lib.mg_render(self._h, C.c_void_p(self.state.data_ptr()), None if ids is None else C.c_void_p(ids.data_ptr()), n, ts,
                                 C.c_void_p(out.data_ptr()), stream)
        self._check(rc)
        return out

    def close_extras(self, **kwargs):
        return None

    def get_attr(self, name: str):
        """VectorEnv.get_attr: per-env values come back as one tensor / value for the whole batch (state is struct-of-arrays)."""
        return getattr(self, name)

    def set_attr(self, name: str, values):
        setattr(self, name, values)

    def reseed(self, seed: int):
        """Re-key the Philox streams (mg_set_seed) and zero the per-env block counters: everything that follows - resets,
        respawns, shuffles, battles - is then a function of (`seed`, global env id) alone, as after construction with that seed."""
        import ctypes as C
        hdr = getattr(self, "_planes", {}).get("hdr")
        if hdr is None:
            raise NotImplementedError(f"{type(self).__name__} keeps its RNG counters elsewhere; construct it with the seed instead")
        self._check(self._lib.mg_set_seed(self._h, C.c_uint64(int(seed) & 0xFFFFFFFFFFFFFFFF)))
        hdr[:, 2] = 0

    # ---- host-buffer path split in two (gymnasium VectorEnv.step_async / step_wait); the classes provide _host_io / _host_result
    _host_stream = None
    _host_pending = False

    def step_async(self, actions):
        """Enqueue one step with HOST arrays (H2D actions -> step kernel -> D2H results) on this env's private stream and
        return at once.  Two env batches driven alternately keep the PCIe link busy: the result copy of one overlaps the host
        work and the step of the other (the way gymnasium's AsyncVectorEnv / EnvPool's async mode are used)."""
        import ctypes as C

        import torch
        if self._host_pending:
            raise RuntimeError("step_async called again before step_wait")
        if not hasattr(self, "_host_io"):
            raise NotImplementedError(f"{type(self).__name__} has no host-buffer step; pass CUDA tensors to step()")
        io = self._host_io(actions)
        if self._host_stream is None:
            self._host_stream = torch.cuda.Stream(device=self.device)
        # ordered after device-path calls on the caller's stream - nothing to order after when that stream is idle (the usual case
        # in a host-buffer loop; asking costs a microsecond, the event record + wait it spares ~10)
        raw = getattr(torch._C, "_cuda_getCurrentRawStream", None)
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        if raw is None or self._lib.mg_stream_idle(C.c_void_p(raw(dev_index))) != 1:
            self._host_stream.wait_stream(torch.cuda.current_stream(self.device))
        if self._lib.mg_step_host_async(self._h, C.c_void_p(self.state.data_ptr()), C.byref(io), C.c_void_p(self._host_stream.cuda_stream)):
            self._check(-1)
        self._host_pending = True

    def step_wait(self):
        """Block until the step enqueued by `step_async` has landed in the page-locked host buffers; returns
        (obs, rewards, terminated, truncated, info) as numpy views of them (overwritten by this env's next host-path step)."""
        import ctypes as C
        if not self._host_pending:
            raise RuntimeError("step_wait without step_async")
        self._host_pending = False
        self._check(self._lib.mg_step_host_wait(self._h, C.c_void_p(self._host_stream.cuda_stream)))
        return self._host_result()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False
