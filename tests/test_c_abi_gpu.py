"""The C ABI from plain C (tests/abi/c_driver.c: no Python, no torch in the process) against the oracle."""
import os
import subprocess

import numpy as np
import pytest

import oracle as oc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_plain_c_program_drives_the_library(tmp_path, cuda_device):
    import gym_multigrid_b200 as mg
    from gym_multigrid_b200._build import LIB_PATH
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    exe, out = str(tmp_path / "c_driver"), str(tmp_path / "out.bin")
    pkg = os.path.dirname(LIB_PATH)
    subprocess.run(["gcc", "-O1", "-std=c11", os.path.join(ROOT, "tests", "abi", "c_driver.c"), "-I", os.path.join(ROOT, "include"),
                    "-I", os.path.join(cuda, "include"), "-L", pkg, "-lmultigrid_b200", "-L", os.path.join(cuda, "lib64"), "-lcudart",
                    f"-Wl,-rpath,{pkg}", f"-Wl,-rpath,{os.path.join(cuda, 'lib64')}", "-o", exe], check=True)
    n, steps, seed = 777, 120, 31
    r = subprocess.run([exe, str(n), str(steps), str(seed), out], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert f"{steps + 2} launches" in r.stdout      # reset + steps + render
    raw = np.fromfile(out, dtype=np.uint8)
    per_step = raw[: steps * 24].view(np.int64).reshape(steps, 3)
    sums = raw[: steps * 24].view(np.float64).reshape(steps, 3)[:, 0]
    final_obs = raw[steps * 24: steps * 24 + n * 300].reshape(n, 10, 10, 3)
    frames = raw[steps * 24 + n * 300:].reshape(3, 320, 320, 3)

    s = mg.spec("multigrid-collect-respawn-clustered-v0")
    o = oc.CollectOracle(oc.make_collect_cfg(layout="quadrants_respawn", time_limit=s.max_episode_steps, **s.kwargs), n)
    rng = oc.PhiloxRng(seed=seed)
    o.reset(rng)
    e, i = np.meshgrid(np.arange(n), np.arange(2), indexing="ij")
    for t in range(steps):
        act = ((e * 7 + i * 3 + t * 5 + (e >> 3)) % 4).astype(np.int8)
        obs, rew, term, trunc = o.step(act, rng, autoreset=True)
        assert rew.sum() == sums[t] and int(term.sum()) == per_step[t, 1] and int(trunc.sum()) == per_step[t, 2], f"step {t}"
    assert np.array_equal(final_obs, obs)
    assert np.array_equal(frames, oc.render_grid(obs[[0, n // 2, n - 1]], 32))
