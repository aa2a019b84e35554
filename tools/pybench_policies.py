#!/usr/bin/env python
"""Host-side cost of one scripted-opponent decision: this package's policies (`gym_multigrid_b200.policy.ctf`) next to the
unmodified reference classes, same inputs, same seeds, decisions asserted equal while timing.  Build container only (needs
/root/reference for the reference arm); pure CPU, one thread.  Prints one JSON line (kept in profiles/)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import ref_harness as rh  # noqa: E402
from replay import POLICY_NAMES, policy_maps, policy_observation  # noqa: E402

from gym_multigrid_b200.policy.ctf import heuristic as ours  # noqa: E402


def main():
    rh.import_reference()
    from gym_multigrid.policy.ctf import heuristic as ref
    maps = policy_maps()
    big = np.where(np.arange(32)[:, None] < 16, 1.0, 0.0) * np.ones((32, 32))
    big[3, 3], big[28, 28] = 5.0, 4.0
    maps = {"board 10x10": maps["board"], "wide 14x11": maps["wide"], "open 32x32": big}
    out = {"what": "microseconds per policy.act(observation_dict, curr_pos), one CPU thread, build container", "cases": []}
    for mname, fm in maps.items():
        rows, cols = fm.shape
        rng = np.random.default_rng(1)
        K = 300 if rows < 30 else 60
        inputs = [(policy_observation(fm, np.stack([rng.integers(0, rows, 2), rng.integers(0, cols, 2)], 1),
                                      np.stack([rng.integers(0, rows, 2), rng.integers(0, cols, 2)], 1)),
                   (int(rng.integers(0, rows)), int(rng.integers(0, cols)))) for _ in range(K)]
        for pname in POLICY_NAMES:
            res = {}
            acts = {}
            for label, mod, passes in (("reference", ref, 1), ("ours_cold", ours, 1), ("ours_warm", ours, 2)):
                pol = getattr(mod, pname)(field_map=fm, random_generator=np.random.Generator(np.random.PCG64(3)))
                for k in range(passes):          # the second pass of `ours_warm` runs on memoised routes
                    pol.random_generator = np.random.Generator(np.random.PCG64(3))
                    t0 = time.perf_counter()
                    a = [int(pol.act(o, c)) for o, c in inputs]
                    dt = time.perf_counter() - t0
                res[label], acts[label] = round(dt / K * 1e6, 1), a
            assert acts["reference"] == acts["ours_cold"] == acts["ours_warm"], (mname, pname)
            out["cases"].append(dict(map=mname, policy=pname, decisions=K, us_per_decision=res,
                                     speedup_cold=round(res["reference"] / res["ours_cold"], 1),
                                     speedup_warm=round(res["reference"] / res["ours_warm"], 1)))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
