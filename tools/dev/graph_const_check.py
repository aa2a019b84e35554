import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import gym_multigrid_b200 as mg
import oracle as oc
early, const = sys.argv[1], sys.argv[2] == "1"
os.environ["MG_EARLY_OBS"] = early
n, B, RING, launches = 65536, 16, 64, 1024
dev = torch.device("cuda:0")
gen = torch.Generator(device=dev).manual_seed(1)
envs = [mg.make_vec("multigrid-collect-respawn-clustered-v0", n, device=dev, seed=0, autoreset=True, env_id_base=b * n) for b in range(B)]
rings = [torch.randint(0, 4, (RING, n, 2), generator=gen, device=dev, dtype=torch.int8) for _ in range(B)]
for e in envs:
    e.reset()
s = mg.spec("multigrid-collect-respawn-clustered-v0")
main = torch.cuda.Stream(device=dev)
with torch.cuda.stream(main):
    for i in range(B):
        envs[i].step(rings[i][0])
    main.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=main):
        for gi in range(launches):
            b = gi % B
            envs[b].step(rings[b][0 if const else (gi // B) % RING])
    REPLAYS = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    for _ in range(REPLAYS):
        g.replay()   # after this: 1 + 64 steps per batch
    main.synchronize()
print("status per batch", [e.status() for e in envs])
for b in (0, 7, 15):
    o = oc.CollectOracle(oc.make_collect_cfg(layout="quadrants_respawn", time_limit=s.max_episode_steps, **s.kwargs), n, nthreads=16)
    r = oc.PhiloxRng(seed=0, env_id_base=b * n)
    o.reset(r)
    ring = rings[b].cpu().numpy()
    o.step(ring[0], r, autoreset=True)
    for k in range(64 * REPLAYS):
        oobs = o.step(ring[0 if const else k % RING], r, autoreset=True)[0]
    same = np.array_equal(envs[b]._obs.cpu().numpy(), oobs)
    gs = np.array_equal(envs[b].grid.cpu().numpy(), o.grid)
    print("batch", b, "obs equal", same, "grid equal", gs, "pos equal", np.array_equal(envs[b].agent_pos.cpu().numpy(), o.agent_pos))
