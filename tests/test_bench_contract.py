"""CPU suite: the driver's contract on `bench.py --impl reference` (the CPU arm: oracle/mg_oracle.c on the host cores, plus the
unmodified Python reference when a copy is on the box) - one JSON line with the keys the driver reads, and the same `config` dict the
GPU arm prints for the same flags."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "1",
                          "--budget-s", "2", "--python-seconds", "0.2"], capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1, "exactly one JSON line"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "env_steps_per_sec" and d["unit"] == "env-steps/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] == 1 and d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert "workload" in d["config"] and "65536 envs per step per GPU" in d["config"]["workload"]
    assert 0.05 < d["pickups_per_env_step"] < 0.4, "uniform random actions: ~0.18 pickups per env-step"
    # the GPU arm prints the same config for the same flags (the driver compares them: same_config)
    sys.path.insert(0, ROOT)
    import argparse
    import bench
    ns = argparse.Namespace(num_envs=65536, batches=16, action_ring=64, repeats=5)
    try:
        cfg = bench.workload_config(ns)
    except AttributeError:      # workload_config reads more flags than listed here: fall back to the printed line alone
        cfg = d["config"]
    assert cfg["workload"] == d["config"]["workload"]


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=120, env=env, cwd=ROOT)
    assert res.returncode == 0 and res.stdout.strip() == ""
