import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import gym_multigrid_b200 as mg
import oracle as oc
for early in ("1", "0"):
    os.environ["MG_EARLY_OBS"] = early
    n = 65536
    env = mg.make_vec("multigrid-collect-respawn-clustered-v0", n, device="cuda:0", seed=0)
    s = mg.spec("multigrid-collect-respawn-clustered-v0")
    o = oc.CollectOracle(oc.make_collect_cfg(layout="quadrants_respawn", time_limit=s.max_episode_steps, **s.kwargs), n, nthreads=16)
    r = oc.PhiloxRng(seed=0)
    env.reset(); o.reset(r)
    act = np.random.default_rng(0).integers(0, 4, size=(n, 2)).astype(np.int8)
    a = torch.as_tensor(act, device="cuda:0")
    bad = 0
    for t in range(230):
        obs, rew, term, trunc, _ = env.step(a)
        oobs, orew, oterm, otrunc = o.step(act, r, autoreset=True)
        if not np.array_equal(obs.cpu().numpy(), oobs) or not np.array_equal(rew.cpu().numpy(), orew):
            bad += 1
            if bad < 3:
                d = np.nonzero((obs.cpu().numpy() != oobs).reshape(n, -1).any(1))[0]
                print("early", early, "step", t, "obs differ in envs", d[:10], len(d))
    print("early", early, "status", env.status(), "mismatching steps", bad, flush=True)
    env.close()
