"""Host halves of the compact result transports (csrc/host_transport.cpp), checked WITHOUT a device against the oracle:
the packed-plane expansion must equal `Grid.encode` (oracle oc_encode3, pinned to the reference's goldens), and delta
records built from consecutive oracle states must patch an observation mirror into exactly the next observation."""
import ctypes as C

import numpy as np
import pytest

import oracle as oc


def _lib():
    from gym_multigrid_b200 import _lib
    return _lib.load()


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("threads", [1, 3, 8])
def test_expand_plane_matches_oracle_encode(threads):
    lib = _lib()
    rng = np.random.default_rng(threads)
    for n in (0, 1, 31, 32, 33, 100, 4095, 4097, 100 * 1037, 256 * 257 + 5):
        cells = rng.integers(0, 256, size=n, dtype=np.uint8)
        if n >= 256:
            cells[:256] = np.arange(256, dtype=np.uint8)     # every code at least once, incl. marked balls (type 2, bit 6)
        out = np.full(3 * n + 64, 0xAB, np.uint8)
        assert lib.mg_host_expand_plane(_p(cells), _p(out), n, threads) == 0
        assert np.array_equal(out[:3 * n].reshape(-1, 3), oc.encode3(cells).reshape(-1, 3))
        assert np.all(out[3 * n:] == 0xAB), "wrote past the end"


def _records_from_oracle(lib, o, prev_grid, rew_codes, term, trunc, done, cells, A):
    """Delta records as the step kernel writes them, derived here from the oracle's state before / after the step."""
    R = lib.mg_delta_record_bytes(cells, A)
    wide = cells > 256
    rec = np.zeros((o.N, R), np.uint8)
    for e in range(o.N):
        idx = np.nonzero(prev_grid[e] != o._post_grid[e])[0]
        assert len(idx) <= 3 * A
        rec[e, 0] = len(idx) | (int(term[e]) << 5) | (int(trunc[e]) << 6) | (int(done[e]) << 7)
        rec[e, 1:1 + A] = rew_codes[e]
        for q, i in enumerate(idx):
            if wide:
                rec[e, 1 + A + 3 * q: 4 + A + 3 * q] = (i & 255, i >> 8, o._post_grid[e, i])
            else:
                rec[e, 1 + A + 2 * q: 3 + A + 2 * q] = (i, o._post_grid[e, i])
    return rec


@pytest.mark.parametrize("size,threads", [(10, 4), (17, 2)])
def test_delta_records_patch_the_mirror(size, threads):
    lib = _lib()
    N, A = 777, 2
    cfg = oc.make_collect_cfg(size=size, agents_index=[3, 5], balls_index=[0, 1, 2], balls_reward=[1, 1, 1], num_balls=15, respawn=True,
                              layout="quadrants_respawn", time_limit=0)
    o = oc.CollectOracle(cfg, N)
    r = oc.PhiloxRng(seed=5)
    mirror = np.ascontiguousarray(o.reset(r))
    rng = np.random.default_rng(1)
    table = np.zeros(33, np.float64)
    table[1:17] = 1.0
    table[17:33] = 1.0
    for t in range(40):
        prev = o.grid.copy()
        act = rng.integers(0, 4, size=(N, A)).astype(np.int8)
        obs, rew, term, trunc = o.step(act, r, autoreset=False)
        o._post_grid = o.grid
        codes = np.where(rew > 0, 1, 0).astype(np.uint8)     # colour is irrelevant with an all-ones table: any non-zero code
        rec = _records_from_oracle(lib, o, prev, codes, term, trunc, np.zeros(N, bool), size * size, A)
        got_rew = np.full((N, A), -1.0)
        got_t, got_u = np.full(N, 9, np.uint8), np.full(N, 9, np.uint8)
        assert lib.mg_host_apply_delta(_p(rec), N, size * size, A, _p(table), _p(mirror), _p(got_rew), _p(got_t), _p(got_u), None, threads) == 0
        assert np.array_equal(mirror, obs), f"step {t}"
        assert np.array_equal(got_rew, rew) and np.array_equal(got_t.astype(bool), term) and np.array_equal(got_u.astype(bool), trunc)


def test_abi_exports_transport_symbols():
    from gym_multigrid_b200 import _lib
    lib = _lib.load()
    for name in ("mg_set_host_transport", "mg_host_invalidate", "mg_host_expand_plane", "mg_delta_record_bytes", "mg_host_apply_delta"):
        assert hasattr(lib, name)
    assert lib.mg_delta_record_bytes(100, 2) == 16 and lib.mg_delta_record_bytes(225, 2) == 16 and lib.mg_delta_record_bytes(17 * 17, 2) == 24
